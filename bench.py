#!/usr/bin/env python
"""bench.py -- simulated messages/sec (and market-days/sec) of the batched ABIDES hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload sparse_zi_1000] [--instances 1024] [--rng mt|philox]

One "step" = one pass of the hot path over one batch: `--instances` independent market days of the
workload (one seed each, seeds drawn as config/parallel.py:11 does) simulated from midnight to the
kernel stop time.  `value` = ttl_messages (Kernel.py:204) of all instances on all ranks / device time,
inputs resident in HBM; `e2e` = the same through abides_load -> abides_run -> abides_read_results with
host buffers.  Under torchrun each rank owns one GPU and its own `--instances` seeds (weak scaling);
the only collective is the final gather of per-instance summaries.

`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) on a bounded
sample of the same workload; the Python reference itself cannot travel to the GPU box.
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

CPU_TTL = []
ALGO_BYTES_PER_MESSAGE = 160          # pop 16 + payload 16 + agent state r/w 80 + emit 32 + book 16
BASELINE_PUBLISHED = {"sparse_zi_1000": 3100.4, "sparse_zi_100": 3585.4}   # BASELINE.md section 1 (1 core i7)
REFERENCE_NOTE = ("the reference arm / cpu_baseline is the C restatement of the reference (oracle/abides_oracle.c) on all "
                  "host threads; the Python reference itself cannot travel to the GPU box and runs 14.7 k (sparse_zi_1000) "
                  "and 4-5 k (rmsc03) messages/s per core in the build container (BASELINE.md section 2), i.e. the C port "
                  "is ~400x the reference per core")
WORKLOADS = {
    "sparse_zi_100": ("sparse_zi_100", []),
    "sparse_zi_1000": ("sparse_zi_1000", []),
    "rmsc03": ("rmsc03", ["-t", "ABM", "-d", "20200603", "--end-time", "16:00:00"]),
    "rmsc03_2h": ("rmsc03", ["-t", "ABM", "-d", "20200603"]),
    # BASELINE configs[2] "rmsc01: value + noise agents + market maker", from the supported classes
    "rmsc01": ("rmsc03", ["-t", "ABM", "-d", "20200603", "--end-time", "16:00:00", "--num-noise", "50",
                          "--num-value", "25", "--num-momentum", "24"]),
    # the reference's literal config/rmsc01.py population: 1 MarketMaker + 50 ZI + 25 HBL + 24 momentum
    "rmsc01_literal": ("rmsc01", []),
    # config/rmsc02.py: the same population with the market maker and momentum agents on market-data subscriptions
    "rmsc02": ("rmsc02", []),
    # config/execution.py -e: the rmsc03 market 09:30-11:30 with a TWAP agent buying 1.2 M shares
    "execution": ("execution", ["-t", "ABM", "-d", "20200603", "-e"]),
}


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


def ncu_traffic(workload, instances):
    """DRAM bytes per sim_kernel launch from the committed ncu capture of the same workload (profiles/), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_roofline.json")))
        e = t.get("%s_x%d" % (workload, instances))
        return float(e["dram_bytes_per_launch"]) if e else None
    except Exception:
        return None


def pin_host_batch(hb):
    """Page-lock the big host tables so the e2e leg copies from pinned memory (done once, outside the timed region)."""
    import torch
    rt = torch.cuda.cudart()
    n = 0
    for a in (hb.mt_key, hb.lat_to, hb.lat_from, hb.size, hb.wakeup_time, hb.agent_type, hb.agent_theta, hb.theta):
        if a is not None and a.nbytes >= (1 << 20):
            if int(rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)) == 0:
                n += a.nbytes
    return n


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def build_batch(workload, seeds, rng_mode):
    from abides_b200 import runtime, lowering
    cfg, extra = WORKLOADS[workload]
    t0 = time.perf_counter()
    specs = runtime.build_specs_parallel(cfg, [["-s", s] + extra for s in seeds])
    hb = lowering.HostBatch(specs, rng_mode=rng_mode)
    return hb, time.perf_counter() - t0


def cpu_baseline(hb, n_sample, threads):
    """The oracle port on the host cores over the first n_sample instances (bounded sample)."""
    import oracle
    t0 = time.perf_counter()
    res, _ = oracle.run_batch(hb, n_threads=threads, first=0, count=n_sample)
    dt = time.perf_counter() - t0
    global CPU_TTL
    CPU_TTL = [int(res[i].ttl_messages) for i in range(n_sample)]      # also a parity spot check against the device
    msgs = sum(CPU_TTL)
    return msgs, dt


def bounded_sample(hb, limit, threads, budget_s=20.0):
    """How many instances the CPU leg takes: one instance per host thread is timed first, and the sample grows from
    there only as far as about `budget_s` seconds of wall time allow (the whole batch for the ZI workloads, one
    round of `threads` instances for rmsc03-sized ones).  Returns (n_sample, messages, seconds) of the probe if the
    probe already is the sample, else (n_sample, None, None)."""
    probe = min(limit, threads)
    m, dt = cpu_baseline(hb, probe, threads)
    n = min(limit, max(probe, int(budget_s / max(dt, 1e-3)) * probe))
    return (n, m, dt) if n == probe else (n, None, None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sparse_zi_1000", choices=sorted(WORKLOADS))
    ap.add_argument("--instances", type=int, default=1024, help="instances per GPU")
    ap.add_argument("--rng", default="mt", choices=["mt", "philox"])
    ap.add_argument("--seed", type=int, default=20261019)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances in the CPU baseline sample (0 = auto)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from abides_b200 import _abi, runtime
    rng_mode = _abi.RNG_MT19937 if args.rng == "mt" else _abi.RNG_PHILOX
    cores = host_cores()
    hours = {"rmsc03_2h": "09:30-11:30, kernel 00:00-11:31", "execution": "09:30-11:30, kernel 00:00-11:31",
             "sparse_zi_100": "09:30-16:00, kernel 00:00-17:00", "sparse_zi_1000": "09:30-16:00, kernel 00:00-17:00"
             }.get(args.workload, "09:30-16:00, kernel 00:00-16:01")
    config = {"workload": "%s x %d seeds per GPU, one market day each (%s)" % (args.workload, args.instances, hours),
              "instances_per_gpu": args.instances, "rng": args.rng,
              "l2": "per-step working set (agent records + RNG state + queues) exceeds the 126 MB L2",
              "seeds": "np.random.RandomState(%d).randint(0, 2**32, n) as config/parallel.py:11" % args.seed}

    if args.impl == "reference":
        if rank != 0:
            return 0
        seeds = runtime.parallel_seeds(args.seed, args.cpu_sample or args.instances)
        hb, build_s = build_batch(args.workload, seeds, _abi.RNG_MT19937)
        # one batch per step where that is a few seconds of CPU (the ZI workloads), else one instance per thread;
        # the sizing probe doubles as the warm-up
        n_sample = args.cpu_sample or bounded_sample(hb, len(seeds), cores)[0]
        msgs = 0
        t = 0.0
        for _ in range(args.steps):
            m, dt = cpu_baseline(hb, n_sample, cores)
            msgs += m
            t += dt
        v = msgs / t
        line = {"impl": "reference", "metric": "simulated order messages/sec", "value": v, "unit": "messages/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": v / BASELINE_PUBLISHED[args.workload] if args.workload in BASELINE_PUBLISHED else None,
                "dtype": "int64+f64", "data": "synthetic", "config": config,
                "market_days_per_s": n_sample * args.steps / t,
                "cpu_baseline": {"value": v, "unit": "messages/s", "cores": cores, "kind": "port",
                                 "sample": "%d instances of %s per step on %d host threads (C restatement of the "
                                           "reference, oracle/abides_oracle.c)" % (n_sample, args.workload, cores)},
                "e2e": {"value": v, "unit": "messages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "reference_note": REFERENCE_NOTE}
        print(json.dumps(line), flush=True)
        return 0

    # ---- our arm: build the host batch BEFORE touching CUDA (the builder forks a process pool) ----
    seeds_all = runtime.parallel_seeds(args.seed, args.instances * world)
    seeds = seeds_all[rank * args.instances:(rank + 1) * args.instances]
    hb, build_s = build_batch(args.workload, seeds, rng_mode)

    import torch
    import numpy as np
    from abides_b200.engine import Engine
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = Engine(local_rank)
    eng.load(hb)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_summaries(res):
        """K6: the final gather of per-instance summary records (the path's only collective)."""
        return runtime.gather_summaries(res, dist, device=torch.device("cuda", local_rank))

    for _ in range(args.warmup):
        eng.reset()
        eng.run()
    res, _ = eng.results_arrays()
    # ABIDES_ESTATE (-4) marks an instance in which the REFERENCE itself would raise (e.g. AdaptiveMarketMakerAgent
    # placing orders before it has ever seen a two-sided book: int(None), AdaptiveMarketMakerAgent.py:226-231);
    # such instances stop there and are reported, anything else is a device error
    n_estate = int((res["status"] == -4).sum())
    if int(((res["status"] != 0) & (res["status"] != -4)).sum()):
        raise RuntimeError("device status != 0 in %d instances: %s" % (int((res["status"] != 0).sum()),
                                                                      np.unique(res["status"]).tolist()))
    msgs_rank = int(res["ttl_messages"].sum())

    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    kern_ms = 0.0
    for _ in range(args.steps):
        eng.reset()
        eng.run()
        r, _ = eng.results_arrays()
        gather_summaries(r)
        dev_ms += eng.last_reset_ms() + eng.last_run_ms()[0]
        kern_ms += eng.last_run_ms()[0]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()

    # e2e: host buffers -> abides_load (H2D) -> abides_run -> abides_read_results (D2H)
    pinned = pin_host_batch(hb)
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.e2e_steps):
        eng.load(hb)
        eng.run()
        r, a = eng.results_arrays()
        gather_summaries(r)
    barrier()
    e2e_wall = time.perf_counter() - t1
    h2d = hb.h2d_bytes()
    d2h = r.nbytes + a.nbytes

    tt = torch.tensor([dev_ms, wall * 1e3, e2e_wall * 1e3, kern_ms], dtype=torch.float64, device="cuda")
    mm = torch.tensor([float(msgs_rank)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(mm, op=dist.ReduceOp.SUM)
    dev_ms, wall_ms, e2e_ms, kern_ms = [float(x) for x in tt.tolist()]
    msgs_total = float(mm.item())
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_hbm_peak()
    value = msgs_total * args.steps / (dev_ms / 1e3)
    achieved = msgs_rank * args.steps * ALGO_BYTES_PER_MESSAGE / (kern_ms / 1e3) / 1e9
    line = {
        "metric": "simulated order messages/sec", "value": value, "unit": "messages/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": value / BASELINE_PUBLISHED[args.workload] if args.workload in BASELINE_PUBLISHED else None,
        "dtype": "int64+f64", "data": "synthetic", "config": config,
        "market_days_per_s": args.instances * world * args.steps / (dev_ms / 1e3),
        "messages_per_step": msgs_total, "wall_ms_per_step": wall_ms / args.steps, "host_build_s": build_s,
        "clocks": clocks, "instances_stopped_where_reference_raises": n_estate,
        "e2e": {"value": msgs_total * args.e2e_steps / (e2e_ms / 1e3), "unit": "messages/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
                "pinned_host_bytes": pinned,
                "path": "Engine.load (abides_load, H2D of every table) -> Engine.run (abides_run) -> "
                        "Engine.results_arrays (abides_read_results, D2H) -> summary gather"},
        "gpu_launches": 3 * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(args.workload, args.instances), "peak_source": peak_src,
                     "kernel_ms_per_launch": kern_ms / args.steps,
                     "note": "dominant kernel sim_kernel: %d B algorithmic per message x messages per launch / "
                             "CUDA-event time of the launch (events on the library's launch stream). The kernel is "
                             "instruction-fetch bound, not HBM bound: see DESIGN.md and profiles/"
                             % ALGO_BYTES_PER_MESSAGE},
    }
    line["reference_note"] = REFERENCE_NOTE
    if world == 1:
        if args.cpu_sample:
            n_sample, m, dt = args.cpu_sample, None, None
        else:
            n_sample, m, dt = bounded_sample(hb, args.instances, cores)    # the whole batch when that takes seconds
        if m is None:
            m, dt = cpu_baseline(hb, n_sample, cores)
        line["cpu_sample_ttl_messages_match"] = bool(CPU_TTL == [int(x) for x in res["ttl_messages"][:n_sample]])
        line["cpu_baseline"] = {"value": m / dt, "unit": "messages/s", "cores": cores, "kind": "port",
                                "sample": "first %d instances of the same batch on %d host threads, %.1f s "
                                          "(oracle/abides_oracle.c)" % (n_sample, cores, dt)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
