"""Bring-up: zi_1000 batch on device vs oracle (per-instance ttl / finals)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from abides_b200 import runtime, lowering
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = sys.argv[2] if len(sys.argv) > 2 else "sparse_zi_1000"
extra = sys.argv[3:]
seeds = runtime.parallel_seeds(7, n)
t = time.time(); specs = runtime.build_specs_parallel(cfg, [["-s", s] + extra for s in seeds]); print("build s", time.time() - t, flush=True)
hb = lowering.HostBatch(specs)
import oracle
from abides_b200.engine import Engine
eng = Engine(0)
t = time.time(); eng.load(hb); print("load s", time.time() - t, "h2d MB", hb.h2d_bytes() / 1e6, flush=True)
t = time.time(); eng.run(); print("run s", time.time() - t, "kernel ms", eng.last_run_ms(), flush=True)
res, ares = eng.results_arrays()
print("status", np.unique(res["status"]), "msgs", int(res["ttl_messages"].sum()), "msg/s", res["ttl_messages"].sum() / (eng.last_run_ms()[0] / 1e3))
print("queue_peak", res["queue_peak"].max(), "book_peak", res["book_peak_orders"].max())
t = time.time(); ores, oares = oracle.run_batch(hb); dt = time.time() - t
ottl = np.array([ores[i].ttl_messages for i in range(n)])
print("oracle s", dt, "oracle msg/s", ottl.sum() / dt, "threads", os.cpu_count())
print("ttl mismatches", int((ottl != res["ttl_messages"]).sum()))
o = np.frombuffer(oares, dtype=ares.dtype)
for f in ("cash", "holdings", "marked_to_market"):
    print(f, "mismatches", int((o[f] != ares[f]).sum()))
fv, ofv = ares["final_valuation"], o["final_valuation"]
print("fv mismatches", int(((fv != ofv) & ~(np.isnan(fv) & np.isnan(ofv))).sum()))
