"""Lower the reference's operator surface (agent / oracle / latency objects handed to
`Kernel.runner`, Kernel.py:51-55) to the plain tables of include/abides_b200.h.

The functions here are duck-typed on the reference's attribute names, so they lower both
this package's descriptor classes (abides_b200/compat) and -- in the golden-vector
generator only -- the reference's own objects.
"""
import ctypes as C

import numpy as np

from . import _abi

_CLASS_BY_NAME = {
    "ExchangeAgent": _abi.EXCHANGE,
    "ZeroIntelligenceAgent": _abi.ZI,
    "ValueAgent": _abi.VALUE,
    "NoiseAgent": _abi.NOISE,
    "MomentumAgent": _abi.MOMENTUM,
    "AdaptiveMarketMakerAgent": _abi.ADAPTIVE_MM,
    "POVExecutionAgent": _abi.POV_EXEC,
    "MarketMakerAgent": _abi.MARKET_MAKER,
    "HeuristicBeliefLearningAgent": _abi.HBL,
    "TWAPExecutionAgent": _abi.SCHEDULE_EXEC,
    "VWAPExecutionAgent": _abi.SCHEDULE_EXEC,
    "SubscriptionAgent": _abi.SUBSCRIPTION,
    "MarketReplayAgent": _abi.MARKET_REPLAY,
}


def _ns(t):
    """pandas Timestamp / Timedelta / numpy datetime / int -> integer nanoseconds."""
    if t is None:
        raise ValueError("time value is None")
    if hasattr(t, "value"):
        return int(t.value)
    if isinstance(t, str):
        import pandas as pd
        return int(pd.Timedelta(t).value)
    if isinstance(t, (np.datetime64, np.timedelta64)):
        return int(t.astype("int64"))
    return int(t)


def agent_class(agent):
    for k in type(agent).__mro__:
        if k.__name__ in _CLASS_BY_NAME:
            if k is not type(agent) and k.__name__ != "ExchangeAgent":
                # a subclass of a supported agent that is not itself supported: different strategy
                raise NotImplementedError(
                    "agent class %s (subclass of %s) is not supported by the B200 engine"
                    % (type(agent).__name__, k.__name__))
            return _CLASS_BY_NAME[k.__name__]
    raise NotImplementedError("agent class %s is not supported by the B200 engine" % type(agent).__name__)


def _mt_state(rs):
    st = rs.get_state(legacy=True) if hasattr(rs, "get_state") else rs
    name, key, pos, has_gauss, gauss = st
    if name != "MT19937":
        raise ValueError("only MT19937 RandomState streams are supported")
    return np.asarray(key, dtype=np.uint32), int(pos), int(has_gauss), float(gauss)


def _regular_horizon(agent):
    """(first point, step, number of points) of a TWAP / VWAP agent's execution_time_horizon, all in ns."""
    pts = [_ns(t) for t in agent.execution_time_horizon]
    if len(pts) < 3:
        raise NotImplementedError("execution_time_horizon needs at least three points")
    step = pts[1] - pts[0]
    if step <= 0 or any(b - a != step for a, b in zip(pts, pts[1:])):
        raise NotImplementedError("execution_time_horizon must be a regular grid")
    return pts[0], step, len(pts)


def _schedule_rows(agent):
    """schedule[slot] for the slots that place limit orders.  ExecutionAgent.placeOrders (:96-97) looks the slot up
    as pd.Interval(t, t + 1 minute), so only a one-minute grid with one-minute bins can work in the reference."""
    import pandas as pd
    h0, step, n = _regular_horizon(agent)
    if not getattr(agent, "trade", False):
        return []
    out = []
    for i in range(n - 2):
        t = pd.Timestamp(h0 + i * step)
        key = pd.Interval(t, t + pd.Timedelta(minutes=1))
        if key not in agent.schedule:
            raise NotImplementedError("schedule has no bin %s (the reference raises KeyError there)" % (key,))
        q = agent.schedule[key]
        if int(q) != q or not (0 <= int(q) < 2 ** 31):
            raise NotImplementedError("schedule quantity %r" % (q,))
        out.append(int(q))
    return out


def _replay_rows(agent, mkt_open_ns):
    """The tape of a MarketReplayAgent as the device reads it: six int32 per row in time order -- time lo, time hi (ns),
    ORDER_ID, PRICE, SIZE, BUY -- from `historical_orders.orders_dict` ({timestamp: [row dicts]}, MarketReplayAgent.py:
    121); returns (rows, first wake-up time).  A row that re-prices or flips an order the agent still holds open ends
    the instance on the device with ABIDES_EINVAL: the reference's modifyOrder then leaves a book whose level prices are
    inconsistent (util/OrderBook.py:359 overwrites the first order of the OLD price level with the new order), which no
    sorted book can mirror."""
    od = agent.historical_orders.orders_dict
    out = []
    times = [_ns(t) for t in od]
    if not times or any(b <= a for a, b in zip(times, times[1:])):
        raise NotImplementedError("MarketReplayAgent: order timestamps must be non-empty and ascending")
    if times[0] < mkt_open_ns:
        raise NotImplementedError("MarketReplayAgent: first order before the market opens")
    for t, group in zip(times, od.values()):
        for r in group:
            oid, price, size = int(r['ORDER_ID']), int(r['PRICE']), int(r['SIZE'])
            buy = 1 if r['BUY_SELL_FLAG'] == 'BUY' else 0
            if not (0 < oid < 2 ** 31 and 0 <= price < 2 ** 31 and 0 <= size < 2 ** 31):
                raise NotImplementedError("MarketReplayAgent: order row out of range: %r" % (r,))
            out.extend([np.int32(np.uint32(t & 0xFFFFFFFF)), np.int32(t >> 32), oid, price, size, buy])
    return [int(x) for x in out], times[0]


_MM_LEVELS_QUOTE_DICT = {1: [1, 0, 0, 0, 0], 2: [.5, .5, 0, 0, 0], 3: [.34, .33, .33, 0, 0], 4: [.25, .25, .25, .25, 0],
                         5: [.20, .20, .20, .20, .20]}


def _subscription(levels, freq):
    if int(levels) != levels or not (1 <= int(levels) <= _abi.MD_MAX_LEVELS):
        raise NotImplementedError("market-data subscription with %r levels (1..%d supported)" % (levels, _abi.MD_MAX_LEVELS))
    if not (float(freq) >= 0):
        raise ValueError("subscription freq %r" % (freq,))
    return dict(subscribe=1, sub_levels=int(levels), subscribe_freq=float(freq))


def _type_row(agent, klass):
    g = lambda n, d=0: getattr(agent, n, d)
    row = dict(klass=klass, q_max=0, starting_cash=int(g("starting_cash", 0)),
               sigma_n=0.0, r_bar=0.0, kappa=0.0, sigma_s=0.0, eta=0.0, lambda_a=0.0,
               R_min=0, R_max=0, wake_up_freq=0, pov=0.0, level_spacing=0.0, spread_alpha=0.0,
               skew_beta=0.0, min_order_size=0, num_ticks=0, cancel_limit_delay=0,
               backstop_quantity=-1, window_size=-1, anchor=0,
               log_orders=1 if g("log_orders", False) else 0, percent_aggr=0.0, depth_spread=0,
               exec_start_time=0, exec_end_time=0, exec_lookback=0, exec_quantity=0.0, exec_is_buy=0, exec_trade=0,
               mm_min_size=0, mm_max_size=0, mm_num_levels=0, hbl_L=0, subscribe=0, sub_levels=0, subscribe_freq=0.0)
    if klass in (_abi.ZI, _abi.VALUE, _abi.HBL):
        row.update(sigma_n=float(agent.sigma_n), r_bar=float(agent.r_bar), kappa=float(agent.kappa),
                   sigma_s=float(agent.sigma_s), lambda_a=float(agent.lambda_a))
    if klass in (_abi.ZI, _abi.HBL):
        row.update(q_max=int(agent.q_max), R_min=int(agent.R_min), R_max=int(agent.R_max),
                   eta=float(agent.eta))
    if klass == _abi.HBL:
        row.update(hbl_L=int(agent.L))
    if klass == _abi.MARKET_MAKER:
        if g("subscribe", False):
            row.update(_subscription(agent.subscribe_num_levels, agent.subscribe_freq))
            if getattr(agent, "levels_quote_dict", None) not in (None, _MM_LEVELS_QUOTE_DICT):
                raise NotImplementedError("MarketMakerAgent with a non-default levels_quote_dict")
        row.update(wake_up_freq=_ns(agent.wake_up_freq), mm_min_size=int(agent.min_size),
                   mm_max_size=int(agent.max_size), mm_num_levels=int(agent.subscribe_num_levels))
    if klass == _abi.VALUE:
        row.update(percent_aggr=float(agent.percent_aggr), depth_spread=int(agent.depth_spread))
    if klass == _abi.MOMENTUM:
        if g("subscribe", False):
            row.update(_subscription(1, 10e9))      # MomentumAgent.wakeup (:37): levels=1, freq=10e9, hard-coded
        row.update(wake_up_freq=_ns(agent.wake_up_freq))
    if klass == _abi.SUBSCRIPTION:
        row.update(_subscription(agent.levels, agent.freq), wake_up_freq=10 ** 9)   # getWakeFrequency: '1s' 
    if klass == _abi.POV_EXEC:
        row.update(wake_up_freq=_ns(agent.freq))
        if g("trade", False):                      # trade=False is a pure heartbeat: nothing else is read
            row.update(pov=float(agent.pov), exec_start_time=_ns(agent.start_time),
                       exec_end_time=_ns(agent.end_time), exec_lookback=_ns(agent.look_back_period),
                       exec_quantity=float(agent.quantity), exec_is_buy=1 if agent.direction == "BUY" else 0,
                       exec_trade=1)
    if klass == _abi.SCHEDULE_EXEC:
        h0, step, n = _regular_horizon(agent)
        # getWakeFrequency needs the first horizon point even for a non-trading agent; the rest is only read when trading
        row.update(exec_start_time=h0)
        if g("trade", False):
            if float(agent.quantity) != int(agent.quantity):
                raise NotImplementedError("non-integral execution quantity")
            row.update(wake_up_freq=step, exec_end_time=h0 + (n - 1) * step, exec_quantity=float(agent.quantity),
                       exec_is_buy=1 if agent.direction == "BUY" else 0, exec_trade=1)
    if klass == _abi.ADAPTIVE_MM:
        if g("subscribe", False):
            raise NotImplementedError("AdaptiveMarketMakerAgent(subscribe=True) is not supported")
        anchor = {"middle": 0, "bottom": 1, "top": 2}[agent.anchor]
        bq = agent.backstop_quantity
        if bq is not None and int(bq) != bq:
            raise NotImplementedError("non-integral backstop_quantity")
        row.update(wake_up_freq=_ns(agent.wake_up_freq), pov=float(agent.pov),
                   level_spacing=float(agent.level_spacing), spread_alpha=float(agent.spread_alpha),
                   skew_beta=float(agent.skew_beta), min_order_size=int(agent.min_order_size),
                   num_ticks=int(agent.num_ticks), cancel_limit_delay=int(agent.cancel_limit_delay),
                   backstop_quantity=-1 if bq is None else int(bq),
                   window_size=-1 if agent.is_adaptive else int(agent.window_size), anchor=anchor)
    return row


class InstanceSpec:
    """One lowered simulation instance (everything `Kernel.runner` was handed)."""

    def __init__(self):
        self.cfg = {}
        self.types = []            # list of dict rows
        self.agent_type = None     # int32[N]
        self.lat_to = None
        self.lat_from = None
        self.size = None
        self.wakeup_time = None
        self.agent_theta = None    # int32[N] offsets, -1 none
        self.theta = None          # int32[]
        self.mt_key = None         # uint32[N+4, 624]
        self.mt_pos = None
        self.mt_has_gauss = None
        self.mt_gauss = None
        self.agent_names = []
        self.agent_type_names = []
        self.device_draws = {}     # constructor arguments the device needs when IT performs the agents' own-stream
                                   # constructor draws (sweep.SeededBatch): sigma_pv, momentum_size_range

    @property
    def n_agents(self):
        return len(self.agent_type)


def lower_instance(agents, startTime, stopTime, defaultComputationDelay=1, defaultLatency=1,
                   agentLatency=None, latencyNoise=(1.0,), agentLatencyModel=None, oracle=None,
                   kernel_random_state=None, global_state=None, seed=0):
    """Lower one `Kernel.runner(...)` call.  `global_state` is np.random.get_state() at call time."""
    n = len(agents)
    spec = InstanceSpec()
    if n == 0 or agent_class(agents[0]) != _abi.EXCHANGE:
        raise NotImplementedError("agent 0 must be the ExchangeAgent")
    ex = agents[0]
    if len(ex.order_books) != 1:
        raise NotImplementedError("exactly one symbol is supported")
    symbol = next(iter(ex.order_books))

    types, type_index = [], {}
    agent_type = np.zeros(n, np.int32)
    size = np.zeros(n, np.int64)
    wake = np.zeros(n, np.int64)
    agent_theta = np.full(n, -1, np.int32)
    theta = []
    mt_key = np.zeros((n + _abi.EXTRA_STREAMS, _abi.MT_N), np.uint32)
    mt_pos = np.full(n + _abi.EXTRA_STREAMS, _abi.MT_N, np.int32)
    mt_hg = np.zeros(n + _abi.EXTRA_STREAMS, np.int32)
    mt_g = np.zeros(n + _abi.EXTRA_STREAMS, np.float64)

    for i, a in enumerate(agents):
        if a.id != i:
            raise ValueError("agents[%d].id == %r: list index must equal agent id" % (i, a.id))
        k = agent_class(a)
        if i > 0 and k == _abi.EXCHANGE:
            raise NotImplementedError("only one ExchangeAgent is supported")
        if k != _abi.EXCHANGE and getattr(a, "symbol", symbol) != symbol:
            raise NotImplementedError("agent %d trades %r, the exchange lists %r" % (i, a.symbol, symbol))
        row = _type_row(a, k)
        if k == _abi.MARKET_REPLAY:
            replay_rows, first = _replay_rows(a, _ns(ex.mkt_open))
            row["wake_up_freq"] = first - _ns(ex.mkt_open)                       # getWakeFrequency (:64-66)
        key = tuple(sorted(row.items()))
        if key not in type_index:
            type_index[key] = len(types)
            types.append(row)
        agent_type[i] = type_index[key]
        if k in (_abi.NOISE, _abi.VALUE, _abi.MOMENTUM):
            size[i] = int(a.size)
        if k == _abi.NOISE:
            wt = a.wakeup_time
            wake[i] = _ns(wt[0] if isinstance(wt, tuple) else wt)
        if k in (_abi.ZI, _abi.HBL):
            spec.device_draws["sigma_pv"] = float(a.sigma_pv)
        if k == _abi.MOMENTUM:
            spec.device_draws["momentum_size_range"] = (int(a.min_size), int(a.max_size))
        if k in (_abi.ZI, _abi.HBL):
            th = [int(x) for x in a.theta]
            if len(th) != 2 * int(a.q_max):
                raise ValueError("ZI theta must have 2*q_max entries")
            agent_theta[i] = len(theta)
            theta.extend(th)
        if k == _abi.SCHEDULE_EXEC:
            agent_theta[i] = len(theta)
            theta.extend(_schedule_rows(a))
        if k == _abi.MARKET_REPLAY:
            agent_theta[i] = len(theta)
            theta.extend(replay_rows)
            size[i] = len(replay_rows) // 6
        mt_key[i], mt_pos[i], mt_hg[i], mt_g[i] = _mt_state(a.random_state)
        spec.agent_names.append(getattr(a, "name", str(i)))
        spec.agent_type_names.append(getattr(a, "type", ""))

    # latency: only agent<->exchange traffic exists on this path, i.e. row 0 / column 0.
    cfg = spec.cfg
    lat_to = np.zeros(n, np.float64)
    lat_from = np.zeros(n, np.float64)
    cfg.update(jitter=0.0, jitter_clip=0.0, jitter_unit=1.0, n_latency_noise=0)
    lat_rs = None
    if agentLatencyModel is not None:
        m = agentLatencyModel
        kw = m.kwargs
        ml = kw["min_latency"]
        if hasattr(ml, "to_exchange"):          # util.ExchangeStarLatency: row 0 / column 0 only
            lat_to[:] = ml.to_exchange
            lat_from[:] = ml.from_exchange
        elif np.isscalar(ml):
            lat_to[:] = ml
            lat_from[:] = ml
        else:
            ml = np.asarray(ml)
            if ml.ndim == 1:
                lat_to[:] = ml
                lat_from[:] = ml[0]
            else:
                lat_to[:] = ml[:, 0]
                lat_from[:] = ml[0, :]
        if m.latency_model == "deterministic":
            cfg["latency_kind"] = _abi.LAT_DETERMINISTIC
        elif m.latency_model == "cubic":
            cfg["latency_kind"] = _abi.LAT_CUBIC
            for name in ("jitter", "jitter_clip", "jitter_unit", "connected"):
                if not np.isscalar(kw[name]):
                    raise NotImplementedError("per-pair %s is not supported" % name)
            if not kw["connected"]:
                raise NotImplementedError("connected=False")
            cfg.update(jitter=float(kw["jitter"]), jitter_clip=float(kw["jitter_clip"]),
                       jitter_unit=float(kw["jitter_unit"]))
        else:
            raise ValueError("unknown latency model %r" % m.latency_model)
        lat_rs = m.random_state
    else:
        cfg["latency_kind"] = _abi.LAT_LEGACY_NOISE
        cfg["n_latency_noise"] = len(latencyNoise)
        if agentLatency is None:
            lat_to[:] = defaultLatency
            lat_from[:] = defaultLatency
        elif hasattr(agentLatency, "to_exchange"):
            lat_to[:] = agentLatency.to_exchange
            lat_from[:] = agentLatency.from_exchange
        else:
            lat_from[:] = np.asarray(agentLatency[0], dtype=np.float64)
            lat_to[:] = np.asarray([agentLatency[i][0] for i in range(n)], dtype=np.float64)

    # oracle (util/oracle/SparseMeanRevertingOracle.py); none at all for a pure replay (config/marketreplay.py:138):
    # the exchange then opens without a price (ExchangeAgent.py:84-88)
    e = _abi.EXTRA_STREAMS
    if oracle is None and all(agent_class(a) == _abi.MARKET_REPLAY for a in agents[1:]):
        cfg.update(r_bar=0.0, kappa=0.0, fund_vol=0.0, megashock_lambda_a=1.0, megashock_mean=0.0, megashock_var=0.0,
                   megashock_time0=2 ** 62, megashock_value0=0.0, flags=_abi.INST_NO_ORACLE)
    else:
        if oracle is None or type(oracle).__name__ != "SparseMeanRevertingOracle":
            raise NotImplementedError("oracle must be a SparseMeanRevertingOracle")
        if list(oracle.symbols) != [symbol]:
            raise NotImplementedError("oracle symbols must match the exchange's single symbol")
        s = oracle.symbols[symbol]
        ms = oracle.megashocks[symbol][-1]
        if len(oracle.megashocks[symbol]) != 1 or _ns(oracle.r[symbol][0]) != _ns(oracle.mkt_open):
            raise ValueError("oracle has already been advanced")
        cfg.update(r_bar=float(s["r_bar"]), kappa=float(s["kappa"]), fund_vol=float(s["fund_vol"]),
                   megashock_lambda_a=float(s["megashock_lambda_a"]),
                   megashock_mean=float(s["megashock_mean"]), megashock_var=float(s["megashock_var"]),
                   megashock_time0=_ns(ms["MegashockTime"]), megashock_value0=float(ms["MegashockValue"]))
        if _ns(oracle.mkt_open) != _ns(ex.mkt_open) or _ns(oracle.mkt_close) != _ns(ex.mkt_close):
            raise NotImplementedError("oracle and exchange market hours differ")
        mt_key[n + _abi.STREAM_ORACLE], mt_pos[n + _abi.STREAM_ORACLE], mt_hg[n + _abi.STREAM_ORACLE], \
            mt_g[n + _abi.STREAM_ORACLE] = _mt_state(s["random_state"])
    if kernel_random_state is not None:
        mt_key[n + 1], mt_pos[n + 1], mt_hg[n + 1], mt_g[n + 1] = _mt_state(kernel_random_state)
    if lat_rs is not None:
        mt_key[n + 2], mt_pos[n + 2], mt_hg[n + 2], mt_g[n + 2] = _mt_state(lat_rs)
    if global_state is None:
        global_state = np.random.get_state()
    mt_key[n + 3], mt_pos[n + 3], mt_hg[n + 3], mt_g[n + 3] = _mt_state(global_state)
    assert e == 4

    cfg.update(start_time=_ns(startTime), stop_time=_ns(stopTime), mkt_open=_ns(ex.mkt_open),
               mkt_close=_ns(ex.mkt_close), default_computation_delay=int(defaultComputationDelay),
               exch_pipeline_delay=int(ex.pipeline_delay), exch_computation_delay=int(ex.computation_delay),
               stream_history=int(ex.stream_history), exch_log_orders=1 if ex.log_orders else 0,
               seed=int(seed) & 0xFFFFFFFFFFFFFFFF)
    spec.types = types
    spec.agent_type = agent_type
    spec.lat_to, spec.lat_from = lat_to, lat_from
    spec.size, spec.wakeup_time = size, wake
    spec.agent_theta = agent_theta
    spec.theta = np.asarray(theta, np.int32)
    spec.mt_key, spec.mt_pos, spec.mt_has_gauss, spec.mt_gauss = mt_key, mt_pos, mt_hg, mt_g
    return spec


class HostBatch:
    """Concatenated host tables of a list of InstanceSpec plus the ctypes `abides_batch` view."""

    def __init__(self, specs, rng_mode=_abi.RNG_MT19937, trace_capacity=0, trace_flags=0):
        self.specs = specs
        ni = len(specs)
        self.instances = (_abi.InstanceCfg * ni)()
        type_rows = []
        a_off = t_off = 0
        for i, s in enumerate(specs):
            ic = self.instances[i]
            for k, v in s.cfg.items():
                setattr(ic, k, v)
            ic.agent_offset, ic.n_agents = a_off, s.n_agents
            ic.type_base, ic.theta_offset = len(type_rows), t_off
            type_rows.extend(s.types)
            a_off += s.n_agents
            t_off += len(s.theta)
        self.types = (_abi.AgentType * max(1, len(type_rows)))()
        for j, row in enumerate(type_rows):
            for k, v in row.items():
                setattr(self.types[j], k, v)
        cat = lambda xs, dt: np.ascontiguousarray(np.concatenate(xs) if xs else np.zeros(0, dt), dtype=dt)
        self.agent_type = cat([s.agent_type for s in specs], np.int32)
        self.lat_to = cat([s.lat_to for s in specs], np.float64)
        self.lat_from = cat([s.lat_from for s in specs], np.float64)
        self.size = cat([s.size for s in specs], np.int64)
        self.wakeup_time = cat([s.wakeup_time for s in specs], np.int64)
        self.agent_theta = cat([s.agent_theta for s in specs], np.int32)
        self.theta = cat([s.theta for s in specs] + [np.zeros(1, np.int32)], np.int32)
        self.n_agents_total, self.n_theta_total = a_off, t_off
        self.rng_mode = rng_mode
        if rng_mode == _abi.RNG_MT19937:
            self.mt_key = cat([s.mt_key.reshape(-1) for s in specs], np.uint32)
            self.mt_pos = cat([s.mt_pos for s in specs], np.int32)
            self.mt_has_gauss = cat([s.mt_has_gauss for s in specs], np.int32)
            self.mt_gauss = cat([s.mt_gauss for s in specs], np.float64)
        else:
            self.mt_key = self.mt_pos = self.mt_has_gauss = self.mt_gauss = None
        b = _abi.Batch()
        b.n_instances, b.n_types = ni, len(type_rows)
        b.n_agents_total, b.n_theta_total = a_off, t_off
        b.rng_mode, b.trace_capacity, b.trace_flags = rng_mode, trace_capacity, trace_flags
        b.instances = C.cast(self.instances, C.POINTER(_abi.InstanceCfg))
        b.types = C.cast(self.types, C.POINTER(_abi.AgentType))
        p = lambda arr, ct: None if arr is None else arr.ctypes.data_as(C.POINTER(ct))
        b.agent_type = p(self.agent_type, C.c_int32)
        b.lat_to_exch = p(self.lat_to, C.c_double)
        b.lat_from_exch = p(self.lat_from, C.c_double)
        b.agent_size = p(self.size, C.c_int64)
        b.agent_wakeup_time = p(self.wakeup_time, C.c_int64)
        b.agent_theta = p(self.agent_theta, C.c_int32)
        b.theta = p(self.theta, C.c_int32)
        b.mt_key = p(self.mt_key, C.c_uint32)
        b.mt_pos = p(self.mt_pos, C.c_int32)
        b.mt_has_gauss = p(self.mt_has_gauss, C.c_int32)
        b.mt_gauss = p(self.mt_gauss, C.c_double)
        self.c = b

    @property
    def n_instances(self):
        return len(self.specs)

    def h2d_bytes(self):
        n = C.sizeof(self.instances) + C.sizeof(self.types)
        for a in (self.agent_type, self.lat_to, self.lat_from, self.size, self.wakeup_time,
                  self.agent_theta, self.theta, self.mt_key, self.mt_pos, self.mt_has_gauss, self.mt_gauss):
            if a is not None:
                n += a.nbytes
        return n


def save_spec(path, spec):
    """Persist an InstanceSpec as a compressed .npz (golden fixtures)."""
    import json
    np.savez_compressed(
        path, cfg=json.dumps(spec.cfg), types=json.dumps(spec.types), agent_type=spec.agent_type,
        lat_to=spec.lat_to, lat_from=spec.lat_from, size=spec.size, wakeup_time=spec.wakeup_time,
        agent_theta=spec.agent_theta, theta=spec.theta, mt_key=spec.mt_key, mt_pos=spec.mt_pos,
        mt_has_gauss=spec.mt_has_gauss, mt_gauss=spec.mt_gauss,
        agent_type_names=json.dumps(spec.agent_type_names))


def load_spec(path):
    import json
    z = np.load(path, allow_pickle=False)
    s = InstanceSpec()
    s.cfg = json.loads(str(z["cfg"]))
    s.types = json.loads(str(z["types"]))
    s.agent_type = z["agent_type"]
    s.lat_to, s.lat_from = z["lat_to"], z["lat_from"]
    s.size, s.wakeup_time = z["size"], z["wakeup_time"]
    s.agent_theta, s.theta = z["agent_theta"], z["theta"]
    s.mt_key, s.mt_pos = z["mt_key"], z["mt_pos"]
    s.mt_has_gauss, s.mt_gauss = z["mt_has_gauss"], z["mt_gauss"]
    s.agent_type_names = json.loads(str(z["agent_type_names"]))
    s.agent_names = [str(i) for i in range(len(s.agent_type))]
    return s
