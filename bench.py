#!/usr/bin/env python
"""bench.py -- simulated messages/sec (and market-days/sec) of the batched ABIDES hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload rmsc03]
                    [--instances 16384] [--scaling weak|strong] [--rng mt|philox] [--engine auto|lane|warp]

One "step" = one pass of the hot path over one batch: `--instances` independent market days of the workload (one seed
each, seeds drawn as config/parallel.py:11 does) simulated from midnight to the kernel stop time.  `value` =
ttl_messages (Kernel.py:204) of all instances on all ranks / device time, inputs resident in HBM; `e2e` = the same
through abides_load -> abides_run -> abides_read_results with HOST buffers (pinned), copies inside the timed region.
Under torchrun each rank owns one GPU and its own `--instances` seeds (weak scaling; `--scaling strong` splits a fixed
total of `--instances` over the ranks); the only collective is the final gather of per-instance summaries.

The headline workload is rmsc03 (BASELINE.json configs[3], the one the north star is quoted on); `also` carries
sparse_zi_1000 (configs[1]) measured the same way with fewer steps.  `--workload rmsc03_grid` is configs[4]: a
heterogeneous parameter grid (population sizes x market-maker parameters x seeds) in one batch.

`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) on a bounded sample of the
same workload; the Python reference itself cannot travel to the GPU box.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_MESSAGE = 160          # pop 16 + payload 16 + agent state r/w 80 + emit 32 + book 16
REFERENCE_NOTE = ("the reference arm / cpu_baseline is the C restatement of the reference (oracle/abides_oracle.c) on all "
                  "host threads, one instance per thread; the Python reference itself cannot travel to the GPU box and "
                  "runs 14.7 k (sparse_zi_1000) and 4-5 k (rmsc03) messages/s per core in the build container "
                  "(BASELINE.md section 2), i.e. the C port is ~400x the reference per core")
RMSC03 = ["-t", "ABM", "-d", "20200603"]
WORKLOADS = {
    "sparse_zi_100": ("sparse_zi_100", []),
    "sparse_zi_1000": ("sparse_zi_1000", []),
    "rmsc03": ("rmsc03", RMSC03),                                  # config/rmsc03.py defaults: 09:30-11:30
    "rmsc03_2h": ("rmsc03", RMSC03),
    "rmsc03_day": ("rmsc03", RMSC03 + ["--end-time", "16:00:00"]),
    # BASELINE configs[2] "rmsc01: value + noise agents + market maker", from the supported classes
    "rmsc01": ("rmsc03", RMSC03 + ["--end-time", "16:00:00", "--num-noise", "50", "--num-value", "25", "--num-momentum", "24"]),
    # the reference's literal config/rmsc01.py population: 1 MarketMaker + 50 ZI + 25 HBL + 24 momentum (warp engine)
    "rmsc01_literal": ("rmsc01", []),
    "rmsc02": ("rmsc02", []),
    "execution": ("execution", RMSC03 + ["-e"]),
}
# BASELINE configs[4]: util/make_grid.py-style sweep of agent counts / market-maker parameters x seeds, one batch
GRID = {"--num-noise": [2500, 5000], "--num-value": [50, 100], "--mm-pov": [0.025, 0.05], "--mm-num-ticks": [10, 20]}
HOURS = {"rmsc03": "09:30-11:30, kernel 00:00-11:31", "rmsc03_2h": "09:30-11:30, kernel 00:00-11:31",
         "rmsc03_grid": "09:30-11:30, kernel 00:00-11:31", "execution": "09:30-11:30, kernel 00:00-11:31",
         "sparse_zi_100": "09:30-16:00, kernel 00:00-17:00", "sparse_zi_1000": "09:30-16:00, kernel 00:00-17:00"}
# instances per GPU: the headline runs on the warp engine -- 2368 = 148 SMs x 16 resident warps, one full wave, a step of
# 6.7 s so that the driver's 25 steps fit -- and the
# `also` workload on the lane engine, whose throughput needs tens of thousands of resident instances and whose step is
# the ~15 s one instance's event chain takes
DEFAULT_INSTANCES = {"rmsc03": 2368, "rmsc03_2h": 2368, "rmsc03_grid": 2048, "sparse_zi_1000": 45056,
                     "sparse_zi_100": 65536}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 2 + j and r[2 + j] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def ncu_traffic(workload, instances, engine):
    """DRAM bytes per message of the dominant kernel from the committed ncu capture of the workload (profiles/), or None."""
    for name in ("r2_roofline.json", "r1_roofline.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
        except Exception:
            continue
        for key in ("%s_%s_x%d" % (workload, engine, instances), "%s_%s" % (workload, engine), "%s_x%d" % (workload, instances)):
            e = t.get(key)
            if e:
                return e
    return None


def pin_arrays(hb):
    """Page-lock the big host tables so the e2e leg copies from pinned memory (done once, outside the timed region)."""
    import torch
    rt = torch.cuda.cudart()
    n = 0
    for name in ("mt_key", "lat_to", "lat_from", "size", "wakeup_time", "agent_type", "agent_theta", "theta", "mt_seed",
                 "mt_pos", "mt_has_gauss", "mt_gauss", "mt_key_row", "cfg"):
        a = getattr(hb, name, None)
        if a is not None and getattr(a, "nbytes", 0) >= (1 << 20):
            if int(rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)) == 0:
                n += a.nbytes
    return n


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def grid_groups(seeds_per_point, seed):
    """(extra flags, seeds) per grid point, heaviest populations first (a CTA then holds instances of one point)."""
    import itertools
    from abides_b200 import runtime
    keys = list(GRID)
    points = list(itertools.product(*[GRID[k] for k in keys]))
    all_seeds = runtime.parallel_seeds(seed, seeds_per_point * len(points))
    out = []
    for j, combo in enumerate(points):
        extra = list(RMSC03)
        for k, v in zip(keys, combo):
            extra += [k, str(v)]
        out.append((extra, all_seeds[j * seeds_per_point:(j + 1) * seeds_per_point]))
    out.sort(key=lambda g: -(int(g[0][g[0].index("--num-noise") + 1]) + int(g[0][g[0].index("--num-value") + 1])))
    return out


def build_batch(workload, seeds, rng_mode, seed0=0):
    """Host tables of one rank's batch.  The sweep builder (abides_b200/sweep.py) where one exists: config-time draws
    vectorised over the seeds, streams handed over as seeds; else the config module once per seed."""
    from abides_b200 import runtime, lowering, sweep
    t0 = time.perf_counter()
    if workload == "rmsc03_grid":
        n_points = 1
        for v in GRID.values():
            n_points *= len(v)
        per = max(1, len(seeds) // n_points)
        groups = [sweep.build_group("rmsc03", s, extra) for extra, s in grid_groups(per, seed0)]
        hb = sweep.SeededBatch(groups, rng_mode=rng_mode)
    else:
        cfg, extra = WORKLOADS[workload]
        if sweep.supported(cfg):
            hb = sweep.build_sweep(cfg, seeds, extra, rng_mode=rng_mode)
        else:
            specs = runtime.build_specs_parallel(cfg, [["-s", s] + extra for s in seeds])
            hb = lowering.HostBatch(specs, rng_mode=rng_mode)
    return hb, time.perf_counter() - t0


def build_cpu_sample(workload, seeds, n_sample):
    """The first n_sample instances of the batch as config-built, fully uploaded instances: what the oracle reads."""
    from abides_b200 import runtime, lowering
    if workload == "rmsc03_grid":
        extra, gseeds = grid_groups(max(1, n_sample), 0)[0]
        cfg, seeds = "rmsc03", gseeds[:n_sample]
    else:
        cfg, extra = WORKLOADS[workload]
    specs = runtime.build_specs_parallel(cfg, [["-s", s] + list(extra) for s in seeds[:n_sample]])
    return lowering.HostBatch(specs)


def cpu_run(hb, n_sample, threads):
    """The oracle port on the host cores over the first n_sample instances; returns (messages, seconds, ttl list)."""
    import oracle
    t0 = time.perf_counter()
    res, _ = oracle.run_batch(hb, n_threads=threads, first=0, count=n_sample)
    dt = time.perf_counter() - t0
    ttl = [int(res[i].ttl_messages) for i in range(n_sample)]
    return sum(ttl), dt, ttl


def cpu_sample_size(workload, cores):
    # about 10-30 s of CPU: one rmsc03-sized instance takes ~0.3 s (2 h) per thread, a sparse_zi_1000 day ~0.05 s
    per_thread = {"sparse_zi_100": 64, "sparse_zi_1000": 32, "rmsc01": 8, "rmsc01_literal": 4}.get(workload, 4)
    return cores * per_thread


def measure(eng, hb, steps, warmup, barrier, gather):
    """W warm-up and K timed steps on the resident batch; returns (device ms, kernel ms, wall ms, launches, res)."""
    for _ in range(warmup):
        eng.reset()
        eng.run()
    res, _ = eng.results_arrays()
    barrier()
    t0 = time.perf_counter()
    l0 = eng.launch_count()
    dev_ms = kern_ms = 0.0
    for _ in range(steps):
        eng.reset()
        eng.run()
        r, _ = eng.results_arrays()
        gather(r)
        dev_ms += eng.last_reset_ms() + eng.last_run_ms()[0]
        kern_ms += eng.last_run_ms()[0]
    barrier()
    return dev_ms, kern_ms, 1e3 * (time.perf_counter() - t0), eng.launch_count() - l0, res.copy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rmsc03", choices=sorted(WORKLOADS) + ["rmsc03_grid"])
    ap.add_argument("--instances", type=int, default=0, help="instances per GPU (weak) or in total (strong); 0 = default")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--rng", default="mt", choices=["mt", "philox"])
    ap.add_argument("--engine", default="auto", choices=["auto", "lane", "warp"])
    ap.add_argument("--seed", type=int, default=20261019)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances in the CPU baseline sample (0 = auto)")
    ap.add_argument("--also", default="sparse_zi_1000", help="second workload reported under `also` ('' = none)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from abides_b200 import _abi, runtime
    rng_mode = _abi.RNG_MT19937 if args.rng == "mt" else _abi.RNG_PHILOX
    cores = host_cores()
    n_default = DEFAULT_INSTANCES.get(args.workload, 2048)
    if args.scaling == "strong":
        total = args.instances or n_default
        lo, hi = runtime.shard_bounds(total, world, rank)
        per_rank = hi - lo
    else:
        per_rank = args.instances or n_default
        total = per_rank * world
        lo, hi = rank * per_rank, (rank + 1) * per_rank
    hours = HOURS.get(args.workload, "09:30-16:00, kernel 00:00-16:01")
    config = {"workload": "%s x %d seeds %s, one market day each (%s)"
                          % (args.workload, per_rank if args.scaling == "weak" else total,
                             "per GPU" if args.scaling == "weak" else "in total, split over the GPUs", hours),
              "instances_per_gpu": per_rank, "instances_total": total, "rng": args.rng,
              "l2": "per-step working set (agent records + RNG state + queues, several GB) exceeds the 126 MB L2",
              "seeds": "np.random.RandomState(%d).randint(0, 2**32, n) as config/parallel.py:11" % args.seed}

    if args.impl == "reference":
        if rank != 0:
            return 0
        n_sample = args.cpu_sample or cpu_sample_size(args.workload, cores)
        seeds = runtime.parallel_seeds(args.seed, max(n_sample, 1))
        hb = build_cpu_sample(args.workload, seeds, n_sample)
        cpu_run(hb, min(n_sample, cores), cores)                      # warm-up
        msgs, t = 0, 0.0
        for _ in range(args.steps):
            m, dt, _ = cpu_run(hb, n_sample, cores)
            msgs += m
            t += dt
        v = msgs / t
        line = {"impl": "reference", "metric": "simulated order messages/sec", "value": v, "unit": "messages/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "int64+f64", "data": "synthetic", "config": config,
                "market_days_per_s": n_sample * args.steps / t,
                "cpu_baseline": {"value": v, "unit": "messages/s", "cores": cores, "kind": "port",
                                 "sample": "%d instances of %s per step on %d host threads (C restatement of the "
                                           "reference, oracle/abides_oracle.c)" % (n_sample, args.workload, cores)},
                "e2e": {"value": v, "unit": "messages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "reference_note": REFERENCE_NOTE}
        print(json.dumps(line), flush=True)
        return 0

    # ---- our arm: build the host batch BEFORE touching CUDA (the config-built path forks a process pool) ----
    seeds_all = runtime.parallel_seeds(args.seed, total)
    hb, build_s = build_batch(args.workload, seeds_all[lo:hi], rng_mode, args.seed)
    cpu_hb = None
    if world == 1:
        n_sample = args.cpu_sample or cpu_sample_size(args.workload, cores)
        n_sample = min(n_sample, per_rank)
        cpu_hb = build_cpu_sample(args.workload, seeds_all[lo:hi], n_sample)

    import torch
    import numpy as np
    from abides_b200.engine import Engine
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = Engine(local_rank)
    eng.set_engine(args.engine)
    pinned = pin_arrays(hb)
    eng.load(hb)
    engine_id, variant = eng.engine_info()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(res):
        """K6: the final gather of per-instance summary records (the path's only collective)."""
        return runtime.gather_summaries(res, dist, device=torch.device("cuda", local_rank))

    sampler = ClockSampler(local_rank)
    sampler.start()
    dev_ms, kern_ms, wall_ms, launches, res = measure(eng, hb, args.steps, args.warmup, barrier, gather)
    clocks = sampler.stop()
    # ABIDES_ESTATE (-4) marks an instance in which the REFERENCE itself would raise (e.g. AdaptiveMarketMakerAgent
    # placing orders before it has ever seen a two-sided book: int(None), AdaptiveMarketMakerAgent.py:226-231);
    # such instances stop there and are reported, anything else is a device error
    n_estate = int((res["status"] == -4).sum())
    if int(((res["status"] != 0) & (res["status"] != -4)).sum()):
        raise RuntimeError("device status != 0 in %d instances: %s" % (int((res["status"] != 0).sum()),
                                                                      np.unique(res["status"]).tolist()))
    msgs_rank = int(res["ttl_messages"].sum())

    # end to end through the public API with host buffers: H2D of every table, run, D2H of the results, gather
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.e2e_steps):
        eng.load(hb)
        eng.run()
        r, a = eng.results_arrays()
        gather(r)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t1)
    h2d = hb.h2d_bytes()
    d2h = int(r.nbytes + a.nbytes)

    tt = torch.tensor([dev_ms, wall_ms, e2e_ms, kern_ms], dtype=torch.float64, device="cuda")
    mm = torch.tensor([float(msgs_rank), float(n_estate), float(per_rank)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(mm, op=dist.ReduceOp.SUM)
    dev_ms, wall_ms, e2e_ms, kern_ms = [float(x) for x in tt.tolist()]
    msgs_total, n_estate_total, inst_total = [float(x) for x in mm.tolist()]

    also = None
    if rank == 0 and world == 1 and args.also and args.also != args.workload:
        # the second BASELINE config measured the same way (fewer steps): resident-input throughput only
        w2 = args.also
        n2 = DEFAULT_INSTANCES.get(w2, 4096)
        hb2, _ = build_batch(w2, runtime.parallel_seeds(args.seed, n2), rng_mode, args.seed)
        eng.load(hb2)
        d2, k2, _, _, res2 = measure(eng, hb2, 1, 1, barrier, lambda r_: r_)
        m2 = float(res2["ttl_messages"].sum())
        also = {w2: {"value": m2 / (d2 / 1e3), "unit": "messages/s", "instances": n2, "steps": 1, "warmup": 1,
                     "market_days_per_s": n2 / (d2 / 1e3), "engine": _abi.VARIANT_NAMES.get(eng.engine_info()[1]),
                     "roofline_frac": m2 * ALGO_BYTES_PER_MESSAGE / (k2 / 1e3) / 1e9 / measured_hbm_peak()[0]}}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_hbm_peak()
    value = msgs_total * args.steps / (dev_ms / 1e3)
    achieved = msgs_rank * args.steps * ALGO_BYTES_PER_MESSAGE / (kern_ms / 1e3) / 1e9
    traffic = ncu_traffic(args.workload, per_rank, "lane" if engine_id == _abi.ENGINE_LANE else "warp")
    line = {
        "metric": "simulated order messages/sec", "value": value, "unit": "messages/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None,
        "dtype": "int64+f64", "data": "synthetic", "config": config,
        "engine": _abi.VARIANT_NAMES.get(variant, str(variant)),
        "market_days_per_s": inst_total * args.steps / (dev_ms / 1e3),
        "messages_per_step": msgs_total, "wall_ms_per_step": wall_ms / args.steps, "host_build_s": build_s,
        "clocks": clocks, "instances_stopped_where_reference_raises": int(n_estate_total),
        "e2e": {"value": msgs_total * args.e2e_steps / (e2e_ms / 1e3), "unit": "messages/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
                "pinned_host_bytes": pinned,
                "path": "Engine.load (abides_load: H2D of every table from pinned host memory, device-side seeding of "
                        "the RNG streams) -> Engine.run (abides_run) -> Engine.results_arrays (abides_read_results, "
                        "D2H) -> summary gather"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None if not traffic else traffic.get("dram_bytes_per_message", 0) * msgs_rank,
                     "traffic_source": None if not traffic else traffic.get("source"),
                     "peak_source": peak_src, "kernel_ms_per_launch": kern_ms / args.steps,
                     "note": "dominant kernel lane_sim_kernel / sim_kernel: %d B algorithmic per message x messages per "
                             "launch / CUDA-event time of the launch (events on the library's launch stream); `traffic` = "
                             "ncu DRAM bytes per message of the committed capture x messages per launch. The warp engine's "
                             "kernel is bound by instruction fetch and dependent latency, the lane engine's at its bench size "
                             "by the L2 / DRAM traffic of thread-local memory: DESIGN.md section 4 and profiles/" % ALGO_BYTES_PER_MESSAGE},
        "reference_note": REFERENCE_NOTE,
        "scaling_note": ("weak scaling: every rank runs its own %d instances, so the job at N GPUs is N times the N=1 job; "
                         "the reference arm always times one bounded CPU sample on one host, whatever N is"
                         % per_rank) if args.scaling == "weak" else
                        "strong scaling: %d instances in total, split evenly over the ranks" % total,
    }
    if also:
        line["also"] = also
    if cpu_hb is not None:
        m = dt = 0.0
        passes = 0
        while dt < 10.0 and passes < 40:                      # about 10 s of CPU work on the sample
            m1, dt1, ttl = cpu_run(cpu_hb, n_sample, cores)
            m, dt, passes = m + m1, dt + dt1, passes + 1
        line["cpu_sample_ttl_messages_match"] = bool(ttl == [int(x) for x in res["ttl_messages"][:n_sample]]) \
            if args.workload != "rmsc03_grid" and args.rng == "mt" else None
        line["cpu_baseline"] = {"value": m / dt, "unit": "messages/s", "cores": cores, "kind": "port",
                                "sample": "first %d instances of the same batch on %d host threads, %d passes, %.1f s "
                                          "(oracle/abides_oracle.c)" % (n_sample, cores, passes, dt)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
