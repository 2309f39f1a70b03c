"""GPU: throughput of the lane engine against the warp engine on one workload, with a cross-check of the results.

    python scripts/gpu_lane_perf.py <workload> <instances> [engine ...]      (engines: lane warp; default both)

Prints one JSON line per engine (messages/s from the CUDA-event time of abides_run) and whether the two engines agree
on every per-instance and per-agent result."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                    # noqa: E402
from abides_b200 import _abi, lowering, runtime, sweep   # noqa: E402
from abides_b200.engine import Engine                 # noqa: E402
import bench                                          # noqa: E402


def mem_available_gb():
    for line in open("/proc/meminfo"):
        if line.startswith("MemAvailable"):
            return int(line.split()[1]) / 1e6
    return 0.0


def main():
    workload, n = sys.argv[1], int(sys.argv[2])
    engines = sys.argv[3:] or ["lane", "warp"]
    cfg, extra = bench.WORKLOADS[workload]
    seeds = runtime.parallel_seeds(20261019, n)
    t0 = time.time()
    if sweep.supported(cfg) and not os.environ.get("ABIDES_FULL_STATE_UPLOAD"):
        rng = _abi.RNG_PHILOX if os.environ.get("ABIDES_PERF_RNG") == "philox" else _abi.RNG_MT19937
        hb = sweep.build_sweep(cfg, seeds, extra, rng_mode=rng)    # streams by seed, vectorised config-time draws
    else:
        specs = runtime.build_specs_parallel(cfg, [list(extra) + ["-s", s] for s in seeds])
        need = 2 * sum(s.mt_key.nbytes for s in specs[:1]) * n / 1e9
        if need > 0.6 * mem_available_gb():
            print(json.dumps({"workload": workload, "instances": n, "skipped": "needs %.0f GB of host memory" % need}))
            return
        hb = lowering.HostBatch(specs)
        del specs
    build_s = time.time() - t0
    digests = {}
    for name in engines:
        eng = Engine(0)
        eng.set_engine(name)
        t1 = time.time()
        eng.load(hb)
        load_s = time.time() - t1
        eng.run()
        ms = ms2 = eng.last_run_ms()[0]
        if os.environ.get("ABIDES_PERF_RUNS", "2") != "1":
            eng.reset()
            eng.run()
            ms2 = eng.last_run_ms()[0]
        res, ares = eng.results_arrays()
        msgs = int(res["ttl_messages"].sum())
        bad = int((res["status"] != 0).sum())
        d = hashlib.sha256(res["ttl_messages"].tobytes() + res["n_fills"].tobytes() + res["last_trade"].tobytes()
                           + ares["cash"].tobytes() + ares["holdings"].tobytes()).hexdigest()[:16]
        digests[name] = d
        print(json.dumps({"workload": workload, "instances": n, "engine": name, "variant": eng.engine_info()[1],
                          "messages": msgs, "run_ms": [round(ms, 2), round(ms2, 2)], "msg_per_s": msgs / (min(ms, ms2) / 1e3),
                          "days_per_s": n / (min(ms, ms2) / 1e3), "load_s": round(load_s, 2), "host_build_s": round(build_s, 1),
                          "nonzero_status": bad, "digest": d, "h2d_gb": round(hb.h2d_bytes() / 1e9, 2)}), flush=True)
        eng.close()
    if len(digests) > 1:
        print(json.dumps({"engines_agree": len(set(digests.values())) == 1}))


if __name__ == "__main__":
    main()
