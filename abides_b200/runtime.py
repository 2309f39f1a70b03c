"""Host runtime above the C ABI: run lowered instances on the GPU(s), batch seed sweeps, write
results back into the descriptor objects, print the reference's end-of-run lines.

Seed sweeps (reference: config/parallel.py:9-22 spawns one process per seed) become ONE batch: every
seed's config module body is executed with `capture()` active, so `Kernel.runner` records the
lowered instance instead of running it, and all instances then go to the device in one launch."""
import contextlib
import os
import runpy
import sys
import time

import numpy as np

from . import _abi, lowering
from .engine import Engine

COMPAT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "compat")
_capture = None
_engines = {}
_BUILD_LOCK = __import__("threading").RLock()


def capturing():
    return _capture is not None


def captured(spec, kernel):
    _capture.append((spec, kernel))


@contextlib.contextmanager
def capture():
    global _capture
    prev, _capture = _capture, []
    try:
        yield _capture
    finally:
        _capture = prev


def engine(device=0):
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]


class RunOutput:
    """Results of a run: `res` per instance, `ares` per agent (instances concatenated in input order)."""

    def __init__(self, hb, res, ares, kernel_ms, wall_s, agent_offsets=None, placement=None):
        self.hb, self.res, self.ares, self.kernel_ms, self.wall_s = hb, res, ares, kernel_ms, wall_s
        if agent_offsets is None:
            agent_offsets = [(hb.instances[i].agent_offset, hb.instances[i].n_agents) for i in range(hb.n_instances)]
        self.agent_offsets = agent_offsets
        self.placement = placement              # multi-device runs: [(device, first instance, count, kernel ms)] per chunk

    def agents(self, inst):
        off, n = self.agent_offsets[inst]
        return self.ares[off:off + n]


def _check_status(res, allow_status):
    bad = np.nonzero((res["status"] != 0) & ~np.isin(res["status"], list(allow_status)))[0]
    if len(bad):
        raise RuntimeError("device reported status %d for instance %d (see enum abides_error)"
                           % (int(res["status"][bad[0]]), int(bad[0])))


def run_specs(specs, device=0, rng_mode=_abi.RNG_MT19937, trace_capacity=0, allow_status=(), devices=None,
              chunk_instances=None, engine_name=None, trace_flags=0):
    """Kernel.runner for a list of lowered instances: H2D, run to completion, D2H.  An instance that ends with a
    status outside `allow_status` raises (ABIDES_ESTATE marks a run in which the reference itself raises).
    `devices` (a list of CUDA device indices) spreads the instances over several GPUs: see run_chunks."""
    if devices is not None and len(devices) > 1:
        n = len(specs)
        chunk = chunk_instances or max(1, -(-n // (2 * len(devices))))
        bounds = [(lo, min(n, lo + chunk)) for lo in range(0, n, chunk)]
        return run_chunks([(lambda lo=lo, hi=hi: lowering.HostBatch(specs[lo:hi], rng_mode=rng_mode)) for lo, hi in bounds],
                          [hi - lo for lo, hi in bounds], devices, allow_status=allow_status, engine_name=engine_name)
    if devices:
        device = devices[0]
    hb = lowering.HostBatch(specs, rng_mode=rng_mode, trace_capacity=trace_capacity, trace_flags=trace_flags)
    eng = engine(device)
    if engine_name is not None:
        eng.set_engine(engine_name)
    t0 = time.perf_counter()
    eng.load(hb)
    eng.run()
    res, ares = eng.results_arrays()
    wall = time.perf_counter() - t0
    _check_status(res, allow_status)
    return RunOutput(hb, res.copy(), ares.copy(), eng.last_run_ms()[0], wall)


def run_chunks(builders, counts, devices, allow_status=(), engine_name=None):
    """The box-wide sweep driver (reference: config/parallel.py:14-18 keeps `num_parallel` processes busy with the
    next seed).  `builders[i]()` returns the host batch of chunk i (`counts[i]` instances); one worker thread per
    device takes the next chunk as soon as its GPU is free -- a ticket counter, so faster GPUs and shorter chunks even
    out -- loads it, runs it and reads the results back.  ctypes releases the GIL during the C calls, so the GPUs
    run concurrently from this one process.  Results come back in chunk (= input) order."""
    import threading
    n_chunks = len(builders)
    out = [None] * n_chunks
    errors = []
    ticket = [0]
    lock = threading.Lock()
    t0 = time.perf_counter()

    def worker(dev):
        try:
            eng = engine(dev)
            if engine_name is not None:
                eng.set_engine(engine_name)
            while True:
                with lock:
                    i = ticket[0]
                    ticket[0] += 1
                if i >= n_chunks or errors:
                    return
                hb = builders[i]()
                eng.load(hb)
                eng.run()
                res, ares = eng.results_arrays()
                out[i] = (dev, res.copy(), ares.copy(), eng.last_run_ms()[0],
                          [hb.instances[k].n_agents for k in range(hb.n_instances)])
        except Exception as e:                                 # surfaced by the caller
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(d,)) for d in devices]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    res = np.concatenate([o[1] for o in out])
    ares = np.concatenate([o[2] for o in out])
    _check_status(res, allow_status)
    offsets, off = [], 0
    for o in out:
        for n in o[4]:
            offsets.append((off, n))
            off += n
    placement, first = [], 0
    for o, c in zip(out, counts):
        placement.append((o[0], first, c, o[3]))
        first += c
    return RunOutput(None, res, ares, max(sum(p[3] for p in placement if p[0] == d) for d in devices),
                     time.perf_counter() - t0, agent_offsets=offsets, placement=placement)


def run_sweep(config, seeds, extra_args=(), devices=(0,), chunk_instances=None, allow_status=(_abi.ESTATE,),
              engine_name=None):
    """Seed sweep of `config` over the GPUs in `devices` from one process.  Chunks are built with the vectorised sweep
    builder where one exists (abides_b200/sweep.py: streams travel as seeds) and by executing the config module per
    seed otherwise.  Returns a RunOutput whose rows follow `seeds`."""
    from . import sweep as sweep_mod
    seeds = [int(s) for s in seeds]
    n = len(seeds)
    chunk = chunk_instances or max(1, -(-n // (2 * len(devices)))) if len(devices) > 1 else (chunk_instances or n)
    bounds = [(lo, min(n, lo + chunk)) for lo in range(0, n, chunk)]

    def builder(lo, hi):
        if sweep_mod.supported(config):
            return lambda: sweep_mod.build_sweep(config, seeds[lo:hi], extra_args)
        return lambda: lowering.HostBatch([s for s, _ in build_instances(
            config, [list(extra_args) + ["-s", str(x)] for x in seeds[lo:hi]])])

    return run_chunks([builder(lo, hi) for lo, hi in bounds], [hi - lo for lo, hi in bounds], list(devices),
                      allow_status=allow_status, engine_name=engine_name)


def write_back(kernel, out, inst):
    """Copy the end-of-run portfolio of instance `inst` into kernel.agents (TradingAgent.py:122-147)."""
    ar = out.agents(inst)
    r = out.res[inst]
    kernel.meanResultByAgentType, kernel.agentCountByType = {}, {}
    ex = kernel.agents[0]
    symbol = next(iter(ex.order_books))
    ex.order_books[symbol].last_trade = int(r["last_trade"])
    for a in kernel.agents[1:]:
        rec = ar[a.id]
        a.holdings = {'CASH': int(rec["cash"])}
        if rec["holdings"] != 0:
            a.holdings[symbol] = int(rec["holdings"])
        a.marked_to_market = int(rec["marked_to_market"])
        fv = float(rec["final_valuation"])
        a.final_valuation = None if np.isnan(fv) else fv
        if hasattr(a, "execution_time_horizon"):              # TWAP / VWAP: the slot carries the arrival mid price
            a.arrival_price, a.final_valuation = a.final_valuation, None
            executed = abs(int(rec["holdings"]))
            a.rem_quantity = a.quantity - executed
            a.average_transaction_price = (round(abs(int(rec["cash"]) - a.starting_cash) / executed, 2)
                                           if executed else None)
        gain = a.marked_to_market - a.starting_cash
        kernel.meanResultByAgentType[a.type] = kernel.meanResultByAgentType.get(a.type, 0) + gain
        kernel.agentCountByType[a.type] = kernel.agentCountByType.get(a.type, 0) + 1
    kernel.ttl_messages = int(r["ttl_messages"])
    kernel.custom_state['kernel_event_queue_elapsed_wallclock'] = out.wall_s
    kernel.custom_state['kernel_slowest_agent_finish_time'] = int(r["slowest_agent_finish_time"])
    kernel.custom_state['ttl_messages'] = int(r["ttl_messages"])


def summary_log_rows(kernel):
    """The rows the reference's agents append to the kernel summary log (Kernel.py:514-520): STARTING_CASH from
    every TradingAgent.kernelStarting (TradingAgent.py:108), then per agent FINAL_CASH_POSITION / ENDING_CASH
    (TradingAgent.py:127-132) and, for the classes that report it, FINAL_VALUATION (ZeroIntelligenceAgent.py:93,
    ValueAgent.py:75, NoiseAgent.py:68).  Needs write_back() to have run."""
    from abides_b200 import exchange_log
    rows = []

    def row(a, ev, val):
        rows.append({'AgentID': a.id, 'AgentStrategy': a.type, 'EventType': ev, 'Event': val})

    for a in kernel.agents[1:]:
        row(a, 'STARTING_CASH', a.starting_cash)
    for a in kernel.agents[1:]:
        row(a, 'FINAL_CASH_POSITION', a.holdings['CASH'])
        row(a, 'ENDING_CASH', a.marked_to_market)
        for ev, val in exchange_log.summary_tail(a):
            row(a, ev, val)
    return rows


def fundamental_frame(kernel, trace):
    """The oracle's f_log as the frame ExchangeAgent.kernelTerminating archives (ExchangeAgent.py:96-101,
    SparseMeanRevertingOracle.py:66,:121): the opening (mkt_open, r_bar) row, then one row per oracle step, taken
    from the device trace's ABIDES_TRACE_FUNDAMENTAL records."""
    import pandas as pd
    sym = next(iter(kernel.agents[0].order_books))
    fund = trace[trace[:, 2] == _abi.TRACE_FUNDAMENTAL]
    t = np.concatenate([[pd.Timestamp(kernel.oracle.mkt_open).value], fund[:, 0]]).astype(np.int64)
    v = np.concatenate([[float(kernel.oracle.symbols[sym]['r_bar'])], fund[:, 3].astype(np.float64)])
    df = pd.DataFrame({'FundamentalTime': pd.to_datetime(t), 'FundamentalValue': v})
    return sym, df.set_index('FundamentalTime')


def write_logs(kernel, trace=None, root=".", exchange_log=None, spec=None):
    """Kernel.writeSummaryLog (Kernel.py:523-532) and, when a trace was taken, what the agents archive at
    kernelTerminating: the exchange's fundamental series (ExchangeAgent.py:91-101) and -- when the trace carries the
    exchange-log records (run_spec_logged) -- the exchange's event log, its ORDERBOOK_<symbol>_FULL depth archive and
    the trading agents' logs (abides_b200/exchange_log.py).  bz2-pickled DataFrames under ./log/<log_dir>/ with the
    reference's file names, columns and dtypes.  `exchange_log`: None = write those archives iff the trace holds
    exchange-log records, True = the trace was taken with ABIDES_TRACE_EXCHANGE_LOG (a run without a single order
    has none and still gets its archives), False = never; `spec`: the lowered instance (the logs of agents with log_orders=True need its latencies)."""
    import pandas as pd
    path = os.path.join(root, "log", kernel.log_dir)
    os.makedirs(path, exist_ok=True)
    kernel.summaryLog = summary_log_rows(kernel)
    pd.DataFrame(kernel.summaryLog).to_pickle(os.path.join(path, "summary_log.bz2"), compression='bz2')
    if trace is not None and getattr(kernel.oracle, "symbols", None):
        sym, df = fundamental_frame(kernel, trace)
        df.to_pickle(os.path.join(path, "fundamental_{}.bz2".format(sym)), compression='bz2')
    if exchange_log is None:
        exchange_log = trace is not None and len(trace) > 0 and bool((trace[:, 2] >= _abi.TRACE_ORDER_PLACED).any())
    if exchange_log and trace is not None:
        from abides_b200 import exchange_log
        try:
            kernel.log_archive = exchange_log.write_archives(kernel, trace, path, spec=spec)
        except NotImplementedError as e:                 # the documented gaps (MODIFY_ORDER in the archives): say so, keep the run
            kernel.log_archive = None
            print("abides_b200: exchange / agent log archives not written: %s" % e, file=sys.stderr)
    return path


def run_spec_logged(spec, device=0, engine_name=None, trace_flags=_abi.TRACE_EXCHANGE_LOG):
    """One instance with its full device trace: a first run sizes the trace (every dispatched event plus at most
    one oracle step per event, plus the exchange-log records), the second records it."""
    out = run_specs([spec], device=device, engine_name=engine_name)
    cap = 2 * int(out.res["delivered"][0]) + 2 * spec.n_agents + 64
    if trace_flags & _abi.TRACE_EXCHANGE_LOG:
        # every exchange-log record belongs to a dispatched event: a placed order or a modification to the message the
        # exchange receives, a notification to its delivery, top of book / last trade to the order message handled
        cap += 3 * int(out.res["delivered"][0]) + 5 * int(out.res["n_orders"][0]) + 2 * int(out.res["n_fills"][0])
    out = run_specs([spec], device=device, trace_capacity=cap, engine_name=engine_name, trace_flags=trace_flags)
    tr, n = engine(device).trace(0)
    if n > cap:
        raise RuntimeError("trace overflow: %d records, capacity %d" % (n, cap))
    return out, tr


def mean_ending_values(kernel):
    """"Mean ending value by agent type" (Kernel.py:318-322): int(round(sum / count))."""
    return {t: int(round(v / kernel.agentCountByType[t])) for t, v in kernel.meanResultByAgentType.items()}


def report(kernel, out, inst, per_agent=True, file=None):
    file = file or sys.stdout
    if per_agent:
        for a in kernel.agents[1:]:
            print("Final holdings for {}: {}.  Marked to market: {}".format(
                a.name, a.fmtHoldings(a.holdings), a.marked_to_market), file=file)
    secs = max(out.wall_s, 1e-9)
    print("Event Queue elapsed: {:.6f}s, messages: {}, messages per second: {:0.1f}".format(
        secs, kernel.ttl_messages, kernel.ttl_messages / secs), file=file)
    print("Mean ending value by agent type:", file=file)
    for t, v in mean_ending_values(kernel).items():
        print("{}: {:d}".format(t, v), file=file)
    print("Simulation ending!", file=file)


def find_config(name):
    for base in (os.path.join(COMPAT_DIR, "config"), os.getcwd(), os.path.join(os.getcwd(), "config")):
        p = os.path.join(base, name + ".py")
        if os.path.exists(p):
            return p
    if os.path.exists(name):
        return name
    raise FileNotFoundError("config %r not found" % name)


@contextlib.contextmanager
def _compat_path():
    sys.path.insert(0, COMPAT_DIR)
    try:
        yield
    finally:
        sys.path.remove(COMPAT_DIR)


def _reset_process_globals():
    """The reference keeps process-global counters; each captured instance starts from a clean process."""
    for mod, attr in (("message.Message", "Message"), ("util.order.Order", "Order")):
        m = sys.modules.get(mod)
        if m is not None:
            setattr(getattr(m, attr), "uniq" if attr == "Message" else "order_id", 0)


def build_instances(config, arg_lists, quiet=True):
    """Execute the config module body once per argument list (e.g. one per seed) under capture();
    returns [(InstanceSpec, Kernel)]."""
    path = find_config(config)
    out = []
    # a config body is process-global state (sys.argv, sys.path, the np.random stream): one at a time
    with _BUILD_LOCK, _compat_path(), capture() as cap:
        for args in arg_lists:
            _reset_process_globals()
            argv, sys.argv = sys.argv, ["abides.py", "-c", config] + [str(a) for a in args]
            try:
                if quiet:
                    with open(os.devnull, "w") as dn, contextlib.redirect_stdout(dn):
                        runpy.run_path(path, run_name="__abides_config__")
                else:
                    runpy.run_path(path, run_name="__abides_config__")
            finally:
                sys.argv = argv
        out = list(cap)
    return out


def _build_chunk(job):
    config, arg_lists = job
    return [s for s, _ in build_instances(config, arg_lists)]


def build_specs_parallel(config, arg_lists, workers=None):
    """build_instances over a process pool (config-time numpy work is host-bound); returns specs only."""
    import multiprocessing as mp
    workers = workers or os.cpu_count() or 1
    workers = max(1, min(workers, len(arg_lists)))
    if workers == 1:
        return _build_chunk((config, arg_lists))
    chunks = [arg_lists[i::workers] for i in range(workers)]
    with mp.get_context("fork").Pool(workers) as pool:
        parts = pool.map(_build_chunk, [(config, c) for c in chunks])
    specs = [None] * len(arg_lists)
    for w, part in enumerate(parts):
        for j, s in enumerate(part):
            specs[w + j * workers] = s
    return specs


def parallel_seeds(seed, num_simulations):
    """The seed stream of the reference's sweep driver (config/parallel.py:10-11)."""
    rs = np.random.RandomState(seed)
    return [int(s) for s in rs.randint(low=0, high=2 ** 32, size=num_simulations)]


def run_config_batch(config, arg_lists, device=0, write=True, allow_status=(), devices=None):
    """Seed sweep / parameter grid in one launch (or, with `devices`, spread over several GPUs of the box).
    Returns (RunOutput, [Kernel])."""
    built = build_instances(config, arg_lists)
    specs = [s for s, _ in built]
    kernels = [k for _, k in built]
    out = run_specs(specs, device=device, allow_status=allow_status, devices=devices)
    if write:
        for i, k in enumerate(kernels):
            write_back(k, out, i)
    return out, kernels


# ---- multi-GPU: instances shard by contiguous ranges, one final gather of summaries ----
def shard_bounds(n_instances, world, rank):
    """[lo, hi) of the instances rank `rank` owns: n // world each, the remainder to the low ranks."""
    base, rem = divmod(int(n_instances), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_summaries(res, dist=None, device=None):
    """All-gather the per-instance result records (numpy structured array of engine.INSTANCE_RESULT_DTYPE) of
    every rank, in instance order.  `dist` is torch.distributed (NCCL on the GPUs, gloo in the CPU tests) or
    None for a single process.  This is the path's only collective."""
    import torch
    from .engine import INSTANCE_RESULT_DTYPE
    res = np.ascontiguousarray(res)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return res.copy()
    world = dist.get_world_size()
    words = INSTANCE_RESULT_DTYPE.itemsize // 8
    n_local = torch.tensor([len(res)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    pad = max(counts)
    buf = np.zeros(pad * words, np.int64)
    buf[:len(res) * words] = np.frombuffer(res.tobytes(), dtype=np.int64)
    t = torch.from_numpy(buf).to(device) if device is not None else torch.from_numpy(buf)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    parts = [o.cpu().numpy()[:c * words] for o, c in zip(out, counts)]
    return np.frombuffer(np.concatenate(parts).tobytes(), dtype=INSTANCE_RESULT_DTYPE).copy()


def grid_arg_lists(base_args, grid, seeds):
    """Cartesian parameter grid x seeds as config argument lists (the reference's util/make_grid.py and
    scripts/market_maker sweeps launch one process per point; here every point is one instance of ONE batch).
    `grid` maps a config flag (e.g. "--mm-pov") to its list of values."""
    import itertools
    keys = list(grid)
    out = []
    for combo in itertools.product(*[grid[k] for k in keys]):
        for s in seeds:
            args = list(base_args) + ["-s", s]
            for k, v in zip(keys, combo):
                args += [k, v]
            out.append(args)
    return out
