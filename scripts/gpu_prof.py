"""Per-phase cycle breakdown of sim_kernel (instrumented build, libabides_b200_prof.so).
usage: ABIDES_B200_LIB=abides_b200/csrc/libabides_b200_prof.so python scripts/gpu_prof.py <n> <config> [args...]"""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("ABIDES_B200_LIB", os.path.join(ROOT, "abides_b200", "csrc", "libabides_b200_prof.so"))
import numpy as np
from abides_b200 import runtime, lowering
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = sys.argv[2] if len(sys.argv) > 2 else "sparse_zi_1000"
extra = sys.argv[3:]
seeds = runtime.parallel_seeds(7, n)
specs = runtime.build_specs_parallel(cfg, [["-s", s] + extra for s in seeds])
hb = lowering.HostBatch(specs)
from abides_b200.engine import Engine
eng = Engine(0)
print(eng._L.abides_version().decode())
eng.load(hb); eng.run()
res, _ = eng.results_arrays()
ms = eng.last_run_ms()[0]
msgs = int(res["ttl_messages"].sum())
print("status", np.unique(res["status"]), "msgs", msgs, "kernel ms %.1f" % ms, "M msg/s %.1f" % (msgs / ms / 1e3))
out = (C.c_uint64 * 16)()
eng._ck(eng._L.abides_read_profile(eng._h, out, 16), "abides_read_profile")
names = ["pop", "refill", "rec_load", "exch_order", "exch_cancel", "exch_query", "exch_other", "ag_wake", "ag_spread",
         "ag_other", "requeue", "n_refill", "n_bpull", "n_requeue", "n_recload", "total_cycles"]
v = dict(zip(names, [int(x) for x in out]))
tot = v["total_cycles"]
print("cycles/event (avg over warps): %.0f" % (tot / msgs))
for k in names[:11]:
    print("  %-12s %6.1f%%  %8.0f cyc/event" % (k, 100.0 * v[k] / tot, v[k] / msgs))
acc = sum(v[k] for k in names[:11])
print("  accounted %.1f%%" % (100.0 * acc / tot))
for k in names[11:15]:
    print("  %-12s %10d  (%.4f per event)" % (k, v[k], v[k] / msgs))
# utilisation: a warp's busy cycles against the kernel's duration (tail = instances that finish early)
clk = 1.965e9
print("mean busy share of a warp: %.1f%% of the kernel's %.1f ms (the rest is the tail after its instance ended)"
      % (100.0 * (tot / n) / (ms * 1e-3 * clk), ms))
