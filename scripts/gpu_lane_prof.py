"""GPU: per-phase cycle breakdown of the lane engine's CTA round (profiling build of the library).

    ABIDES_B200_LIB=abides_b200/csrc/libabides_b200_prof.so python scripts/gpu_lane_prof.py <workload> <instances>
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("ABIDES_B200_LIB", os.path.join(ROOT, "abides_b200", "csrc", "libabides_b200_prof.so"))
import ctypes as C                                     # noqa: E402
import numpy as np                                     # noqa: E402
from abides_b200 import runtime, sweep                 # noqa: E402
from abides_b200.engine import Engine                  # noqa: E402
import bench                                           # noqa: E402


def main():
    workload, n = sys.argv[1], int(sys.argv[2])
    cfg, extra = bench.WORKLOADS[workload]
    hb = sweep.build_sweep(cfg, runtime.parallel_seeds(20261019, n), extra)
    eng = Engine(0)
    eng.set_engine("lane")
    eng.load(hb)
    eng.run()
    ms = eng.last_run_ms()[0]
    out = (C.c_uint64 * 16)()
    eng._ck(eng._L.abides_read_profile(eng._h, out, 16), "abides_read_profile")
    rounds, pop, sort, handle, wait, events = [int(out[i]) for i in range(6)]
    res, _ = eng.results_arrays()
    cta = int(os.environ.get("ABIDES_LANE_CTA", "128"))
    n_cta = (n + cta - 1) // cta
    print(json.dumps({"lib": os.path.basename(os.environ["ABIDES_B200_LIB"]), "workload": workload, "instances": n, "run_ms": ms, "messages": int(res["ttl_messages"].sum()),
                      "delivered": int(res["delivered"].sum()), "ctas": n_cta, "rounds_per_cta": rounds / n_cta,
                      "events_per_round": events / max(rounds, 1),
                      "cycles_per_round": {"pop": pop / max(rounds, 1), "sort": sort / max(rounds, 1),
                                           "handle_thread0": handle / max(rounds, 1), "wait_for_slowest_warp": wait / max(rounds, 1)}}))


if __name__ == "__main__":
    main()
