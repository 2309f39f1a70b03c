"""How throughput moves with the number of resident instances: build `n_unique` instances of a config once, then run
batches made of 1x, 2x, ... copies of them (identical copies cost the same as distinct seeds; only the build is saved).

    python scripts/gpu_occupancy.py rmsc03 148 1,2,4,8 -t ABM -d 20200603 --end-time 16:00:00
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                                   # noqa: E402
from abides_b200 import lowering, runtime                           # noqa: E402
from abides_b200.engine import Engine                               # noqa: E402

cfg, n_unique, reps = sys.argv[1], int(sys.argv[2]), [int(x) for x in sys.argv[3].split(",")]
extra = sys.argv[4:]
seeds = runtime.parallel_seeds(7, n_unique)
t = time.time()
specs = runtime.build_specs_parallel(cfg, [["-s", s] + extra for s in seeds])
print("built %d instances in %.1f s" % (n_unique, time.time() - t), flush=True)
eng = Engine(0)
for r in reps:
    hb = lowering.HostBatch(specs * r)
    eng.load(hb)
    eng.run()
    res, _ = eng.results_arrays()
    ms = eng.last_run_ms()[0]
    msgs = int(res["ttl_messages"].sum())
    print("%5d instances: %8.1f ms, %7.1f M msg/s, %6.1f instance-runs/s, status %s"
          % (len(specs) * r, ms, msgs / ms / 1e3, len(specs) * r / ms * 1e3, np.unique(res["status"]).tolist()), flush=True)
