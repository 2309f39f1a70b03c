"""Run a batch and print the device status histogram and capacity peaks (diagnosis of ABIDES_EOVERFLOW)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from abides_b200 import runtime, lowering
from abides_b200.engine import Engine
n = int(sys.argv[1]); cfg = sys.argv[2]; extra = sys.argv[3:]
seeds = runtime.parallel_seeds(20261019, n)
specs = runtime.build_specs_parallel(cfg, [["-s", s] + extra for s in seeds])
hb = lowering.HostBatch(specs)
eng = Engine(0); eng.load(hb); eng.run()
res, _ = eng.results_arrays()
print("status", np.unique(res["status"], return_counts=True))
bad = np.nonzero(res["status"] != 0)[0]
print("bad instances", bad[:20])
for f in ("ttl_messages", "queue_peak", "book_peak_orders", "n_orders", "n_fills", "final_time"):
    print(f, "ok max", res[f][res["status"] == 0].max(), "bad", res[f][bad][:8])
