"""Vectorised seed sweeps: build thousands of instances of one config without executing the config body per seed.

The reference's sweep driver starts one process per seed (config/parallel.py:9-22); `runtime.build_instances` does the
same thing in-process -- it runs the config module once per seed and lowers the agents it constructed, ~17 ms (1000
agents) to ~80 ms (5000 agents) of Python per instance, and every stream travels as its 2.5 KB MT19937 state.  For the
two config families the BASELINE sweeps use (sparse_zi_*, rmsc03 with its population / market-maker flags) this module
replays the config's draw order for ALL seeds at once on `vecrng.VecRandomState` and describes every stream by its seed
(abides_batch.mt_seed): the device seeds the generators and performs the constructor draws the agents make from their
own streams (ZI private values, momentum order sizes) itself.

Everything that does not depend on the seed -- agent types, class layout, market hours, delays -- is taken from ONE
instance built the ordinary way (the template), and the vectorised draws are checked against that instance before the
batch is returned, so a config change that this module does not mirror fails loudly instead of drifting.
"""
import ctypes as C

import numpy as np

from . import _abi, runtime
from .vecrng import VecRandomState

E = _abi.EXTRA_STREAMS
_CFG_DTYPE = np.dtype(_abi.InstanceCfg)
_LIGHT_NS_PER_M = 299792458e-9


class SweepGroup:
    """n instances of one population (same flags, different seeds): per-instance tables as [n, n_agents] arrays."""

    def __init__(self, template, seeds):
        self.template = template                   # lowering.InstanceSpec of seeds[0], built by the config module
        self.seeds = [int(s) for s in seeds]
        self.n = len(self.seeds)
        self.n_agents = template.n_agents
        na, n = self.n_agents, self.n
        self.lat_to = np.zeros((n, na), np.float64)
        self.lat_from = np.zeros((n, na), np.float64)
        self.size = np.zeros((n, na), np.int64)
        self.wakeup_time = np.zeros((n, na), np.int64)
        self.stream_seed = np.zeros((n, na + E), np.uint32)
        self.stream_pos = np.zeros((n, na + E), np.int32)
        self.stream_has_gauss = np.zeros((n, na + E), np.int32)
        self.stream_gauss = np.zeros((n, na + E), np.float64)
        self.megashock_time0 = np.zeros(n, np.int64)
        self.megashock_value0 = np.zeros(n, np.float64)
        self.draw_flags = 0


def _oracle_ctor(grp, g, oracle_rs, mkt_open_ns, s):
    """SparseMeanRevertingOracle.__init__ (:58-72): the gap to the first megashock from the global stream, its size
    from the symbol's stream."""
    import pandas as pd
    gap = g.exponential(1.0 / s["megashock_lambda_a"])
    open_ts = pd.Timestamp(mkt_open_ns)
    grp.megashock_time0[:] = [(open_ts + pd.Timedelta(float(x), unit='ns')).value for x in gap.tolist()]
    size = oracle_rs.normal(s["megashock_mean"], np.sqrt(s["megashock_var"]))
    flip = oracle_rs.randint(0, 2) != 0
    grp.megashock_value0[:] = np.where(flip, -size, size)


def _set_stream(grp, idx, rs):
    """Describe stream column idx by the seeds / words drawn / cached gaussian of a VecRandomState."""
    grp.stream_seed[:, idx] = rs.seeds
    grp.stream_pos[:, idx] = rs.consumed
    grp.stream_has_gauss[:, idx] = rs.has_gauss
    grp.stream_gauss[:, idx] = rs.gauss


def _fresh(grp, idx, g):
    """random_state = np.random.RandomState(seed=np.random.randint(low=0, high=2**32)): one global word, unused stream."""
    grp.stream_seed[:, idx] = g.randint(0, 2 ** 32).astype(np.uint32)


def _check_against_template(grp, names):
    """Row 0 of the vectorised tables must equal what the config module itself produced for seeds[0]."""
    t = grp.template
    for name in names:
        got = getattr(grp, name)[0]
        want = getattr(t, name) if name != "wakeup_time" else t.wakeup_time
        if not np.array_equal(got, want):
            bad = int(np.nonzero(got != want)[0][0])
            raise RuntimeError("sweep builder out of step with the config module: %s[%d] = %r, config built %r"
                               % (name, bad, got[bad], want[bad]))
    if grp.megashock_time0[0] != t.cfg["megashock_time0"] or grp.megashock_value0[0] != t.cfg["megashock_value0"]:
        raise RuntimeError("sweep builder out of step with the config module: first megashock")
    # streams: expand the seeded description of a few streams and compare with the state the config lowered
    na = grp.n_agents
    for idx in sorted(set([0, 1, na // 2, na - 1, na + _abi.STREAM_ORACLE, na + _abi.STREAM_KERNEL, na + _abi.STREAM_LATENCY])):
        if idx < na and (grp.draw_flags and t.mt_pos[idx] != _abi.MT_N):
            continue                                         # constructor draws left to the device: cannot compare here
        rs = np.random.RandomState(seed=int(grp.stream_seed[0, idx]))
        k = int(grp.stream_pos[0, idx])
        if k:
            rs.randint(0, 2 ** 32, size=k, dtype='uint64')   # one 32-bit word each
        st = rs.get_state()
        if idx == na + _abi.STREAM_LATENCY and t.cfg["latency_kind"] == _abi.LAT_LEGACY_NOISE:
            continue                                         # no latency stream in the legacy model
        if not (np.array_equal(st[1], t.mt_key[idx]) and st[2] == t.mt_pos[idx]):
            raise RuntimeError("sweep builder out of step with the config module: stream %d" % idx)


# ----------------------------------------------------------------------------------------------
# sparse_zi_100 / sparse_zi_1000  (compat/config/_sparse_zi.py; reference config/sparse_zi_100.py:69-194)
# ----------------------------------------------------------------------------------------------
def sweep_sparse_zi(config, seeds, extra_args=()):
    (template, _), = runtime.build_instances(config, [list(extra_args) + ["-s", str(seeds[0])]])
    grp = SweepGroup(template, seeds)
    n, na = grp.n, grp.n_agents
    new_model = template.cfg["latency_kind"] == _abi.LAT_CUBIC
    g = VecRandomState(np.asarray(grp.seeds, np.uint64))                 # np.random.seed(seed)
    oracle_seed = g.randint(0, 2 ** 32)
    _fresh(grp, na + _abi.STREAM_KERNEL, g)
    if new_model:
        _fresh(grp, na + _abi.STREAM_LATENCY, g)
    oracle_rs = VecRandomState(oracle_seed)
    _oracle_ctor(grp, g, oracle_rs, template.cfg["mkt_open"],
                 dict(megashock_lambda_a=template.cfg["megashock_lambda_a"], megashock_mean=template.cfg["megashock_mean"],
                      megashock_var=template.cfg["megashock_var"]))
    _set_stream(grp, na + _abi.STREAM_ORACLE, oracle_rs)
    for a in range(na):                                                  # exchange, then the ZI agents: one seed each
        _fresh(grp, a, g)
    if new_model:
        # np.random.uniform(21000, 100000, (n, n)): column 0 is agent -> exchange, row 0 exchange -> agent
        for i in range(na):
            row = np.stack([g.uniform(21000, 100000) for _ in range(na)], axis=1) if i == 0 else None
            if i == 0:
                grp.lat_from[:] = row
                grp.lat_to[:, 0] = row[:, 0]
            else:
                grp.lat_to[:, i] = g.uniform(21000, 100000)
                for _ in range(na - 1):
                    g.random_sample()
    else:
        # np.random.uniform(21000, 13000000, (n, n)): only row 0 is read, with [0, 0] = 20000
        row0 = np.stack([g.uniform(21000, 13000000) for _ in range(na)], axis=1)
        row0[:, 0] = 20000
        grp.lat_to[:] = row0
        grp.lat_from[:] = row0
        g.skip_words(2 * (na * na - na))
    grp.stream_seed[:, na + _abi.STREAM_GLOBAL] = np.asarray(grp.seeds, np.uint64).astype(np.uint32)
    grp.stream_pos[:, na + _abi.STREAM_GLOBAL] = g.consumed
    grp.draw_flags = _abi.DRAW_ZI_THETA
    _check_against_template(grp, ["lat_to", "lat_from"])
    return grp


# ----------------------------------------------------------------------------------------------
# rmsc03 and its population / market-maker flags  (compat/config/rmsc03.py; reference config/rmsc03.py:134-367)
# ----------------------------------------------------------------------------------------------
def _u_quadratic_offsets(y, span_value, span_unit_ns):
    """util.get_wake_time (util/util.py:35-57) for a = 0, b = 1: cbrt((3 / 12) y - 1 / 8) + 1 / 2 of the open-close span,
    as pandas scales a Timedelta by a float: truncated in the Timedelta's own unit.  numpy's vector pow may differ
    from the scalar libm pow the config calls in the last bit, so every element that lands within 1e-3 of a
    truncation boundary is recomputed with scalar arithmetic."""
    alpha, beta = 12.0, 0.5
    arg = (3 / alpha) * y - (beta - 0) ** 3
    mult = np.where(arg < 0, -np.power(np.abs(arg), 1.0 / 3.0), np.power(np.abs(arg), 1.0 / 3.0)) + beta
    v = mult * span_value
    frac = v - np.floor(v)
    risky = np.nonzero((frac < 1e-3) | (frac > 1 - 1e-3))
    if len(risky[0]):
        def one(a):
            c = -((-a) ** (1.0 / 3.0)) if a < 0 else a ** (1.0 / 3.0)
            return (c + beta) * span_value
        v[risky] = [one(float(a)) for a in arg[risky].tolist()]
    return np.trunc(v).astype(np.int64) * span_unit_ns


def sweep_rmsc03(seeds, extra_args=("-t", "ABM", "-d", "20200603")):
    import argparse
    import pandas as pd
    from dateutil.parser import parse
    ap = argparse.ArgumentParser()                          # the two flags whose values enter the vectorised draws
    ap.add_argument('-d', '--historical-date', required=True, type=parse)
    ap.add_argument('--latency-scale', type=float, default=1.0)
    args, _ = ap.parse_known_args([str(a) for a in extra_args])
    (template, kernel), = runtime.build_instances("rmsc03", [list(extra_args) + ["-s", str(seeds[0])]])
    grp = SweepGroup(template, seeds)
    n, na = grp.n, grp.n_agents
    klass = np.array([template.types[t]["klass"] for t in template.agent_type])
    g = VecRandomState(np.asarray(grp.seeds, np.uint64))
    oracle_rs = VecRandomState(g.randint(0, 2 ** 32))
    _oracle_ctor(grp, g, oracle_rs, template.cfg["mkt_open"],
                 dict(megashock_lambda_a=template.cfg["megashock_lambda_a"], megashock_mean=template.cfg["megashock_mean"],
                      megashock_var=template.cfg["megashock_var"]))
    _set_stream(grp, na + _abi.STREAM_ORACLE, oracle_rs)
    _fresh(grp, 0, g)                                                    # exchange
    day = pd.to_datetime(args.historical_date)
    noise_open, noise_close = day + pd.to_timedelta("09:00:00"), day + pd.to_timedelta("16:00:00")
    span = noise_close - noise_open
    unit_ns = {"ns": 1, "us": 10 ** 3, "ms": 10 ** 6, "s": 10 ** 9}[getattr(span, "unit", "ns")]
    span_value = span.value // unit_ns
    ys = np.empty((n, na), np.float64)
    for a in range(1, na):
        k = klass[a]
        if k == _abi.NOISE:
            ys[:, a] = g.rand()                                          # util.get_wake_time, then the agent's seed,
            _fresh(grp, a, g)
            grp.size[:, a] = g.randint(20, 50)                           # then NoiseAgent.__init__ (:34)
        elif k == _abi.VALUE:
            _fresh(grp, a, g)
            grp.size[:, a] = g.randint(20, 50)                           # ValueAgent.__init__
        else:
            _fresh(grp, a, g)                                            # market makers, momentum, POV: seed only
    noise = klass == _abi.NOISE
    if noise.any():
        grp.wakeup_time[:, noise] = noise_open.value + _u_quadratic_offsets(ys[:, noise], span_value, unit_ns)
        # spot-check the truncation rule against pandas itself
        for a in np.nonzero(noise)[0][:4]:
            y = float(ys[0, a])
            arg = (3 / 12.0) * y - 0.125
            c = -((-arg) ** (1.0 / 3.0)) if arg < 0 else arg ** (1.0 / 3.0)
            if (noise_open + (c + 0.5) * span).value != grp.wakeup_time[0, a]:
                raise RuntimeError("sweep builder: pandas scales Timedeltas differently from what this module assumes")
    _fresh(grp, na + _abi.STREAM_KERNEL, g)
    lat_rs = VecRandomState(g.randint(0, 2 ** 32))
    # util.exchange_distances_on_line + meters_to_light_ns (util/util.py:106-133)
    right = 3866660 * args.latency_scale
    x = np.stack([lat_rs.uniform(0.0, right) for _ in range(na)], axis=1)
    d = (np.abs(x - x[:, :1]) / _LIGHT_NS_PER_M).astype(int).astype(np.float64)
    grp.lat_to[:] = d
    grp.lat_from[:] = d
    _set_stream(grp, na + _abi.STREAM_LATENCY, lat_rs)
    grp.stream_seed[:, na + _abi.STREAM_GLOBAL] = np.asarray(grp.seeds, np.uint64).astype(np.uint32)
    grp.stream_pos[:, na + _abi.STREAM_GLOBAL] = g.consumed
    grp.draw_flags = _abi.DRAW_MOMENTUM_SIZE
    mom = klass == _abi.MOMENTUM
    grp.size[0, mom] = template.size[mom]                    # drawn on the device; keep row 0 comparable
    _check_against_template(grp, ["lat_to", "lat_from", "size", "wakeup_time"])
    grp.size[:, mom] = 0
    return grp


# ----------------------------------------------------------------------------------------------
# groups -> one abides_batch
# ----------------------------------------------------------------------------------------------
class SeededBatch:
    """Host tables of a sweep (the same interface as lowering.HostBatch: .c, .instances, .n_instances, ...)."""

    def __init__(self, groups, trace_capacity=0, rng_mode=_abi.RNG_MT19937):
        self.groups = groups
        ni = sum(g.n for g in groups)
        na_total = sum(g.n * g.n_agents for g in groups)
        self.cfg = np.zeros(ni, dtype=_CFG_DTYPE)
        type_rows = []
        parts = {k: [] for k in ("agent_type", "lat_to", "lat_from", "size", "wakeup_time", "agent_theta",
                                 "seed", "pos", "hg", "gauss")}
        i0 = a_off = t_off = 0
        draw_flags = 0
        momentum_types = []
        for g in groups:
            t = g.template
            type_base = len(type_rows)
            for row in t.types:
                row = dict(row)
                type_rows.append(row)
            sl = slice(i0, i0 + g.n)
            for k, v in t.cfg.items():
                self.cfg[k][sl] = v
            self.cfg["megashock_time0"][sl] = g.megashock_time0
            self.cfg["megashock_value0"][sl] = g.megashock_value0
            self.cfg["seed"][sl] = np.asarray(g.seeds, np.uint64)
            self.cfg["n_agents"][sl] = g.n_agents
            self.cfg["type_base"][sl] = type_base
            self.cfg["agent_offset"][sl] = a_off + np.arange(g.n, dtype=np.int64) * g.n_agents
            n_theta = len(t.theta)
            self.cfg["theta_offset"][sl] = t_off + np.arange(g.n, dtype=np.int64) * n_theta
            parts["agent_type"].append(np.tile(t.agent_type, g.n))
            parts["agent_theta"].append(np.tile(t.agent_theta, g.n))
            parts["lat_to"].append(g.lat_to.reshape(-1))
            parts["lat_from"].append(g.lat_from.reshape(-1))
            parts["size"].append(g.size.reshape(-1))
            parts["wakeup_time"].append(g.wakeup_time.reshape(-1))
            parts["seed"].append(g.stream_seed.reshape(-1))
            parts["pos"].append(g.stream_pos.reshape(-1))
            parts["hg"].append(g.stream_has_gauss.reshape(-1))
            parts["gauss"].append(g.stream_gauss.reshape(-1))
            draw_flags |= g.draw_flags
            i0 += g.n
            a_off += g.n * g.n_agents
            t_off += g.n * n_theta
        cat = lambda k, dt: np.ascontiguousarray(np.concatenate(parts[k]), dtype=dt)
        self.agent_type = cat("agent_type", np.int32)
        self.agent_theta = cat("agent_theta", np.int32)
        self.lat_to, self.lat_from = cat("lat_to", np.float64), cat("lat_from", np.float64)
        self.size, self.wakeup_time = cat("size", np.int64), cat("wakeup_time", np.int64)
        self.mt_seed, self.mt_pos = cat("seed", np.uint32), cat("pos", np.int32)
        self.mt_has_gauss, self.mt_gauss = cat("hg", np.int32), cat("gauss", np.float64)
        # streams are laid out per instance as [agents..., extras...]: exactly the order of the group arrays
        self.mt_key_row = np.full(len(self.mt_seed), -1, np.int32)
        self.theta = np.zeros(1, np.int32)                     # values are drawn on the device (layout only)
        self.n_agents_total, self.n_theta_total = na_total, t_off
        self.types = (_abi.AgentType * max(1, len(type_rows)))()
        for j, row in enumerate(type_rows):
            for k, v in row.items():
                setattr(self.types[j], k, v)
        self._fill_device_draw_parameters(groups)
        self.instances = (_abi.InstanceCfg * ni).from_buffer(self.cfg)
        b = _abi.Batch()
        b.n_instances, b.n_types = ni, len(type_rows)
        b.n_agents_total, b.n_theta_total = na_total, t_off
        b.rng_mode, b.trace_capacity = rng_mode, trace_capacity
        b.instances = C.cast(self.instances, C.POINTER(_abi.InstanceCfg))
        b.types = C.cast(self.types, C.POINTER(_abi.AgentType))
        p = lambda arr, ct: arr.ctypes.data_as(C.POINTER(ct))
        b.agent_type, b.agent_theta = p(self.agent_type, C.c_int32), p(self.agent_theta, C.c_int32)
        b.lat_to_exch, b.lat_from_exch = p(self.lat_to, C.c_double), p(self.lat_from, C.c_double)
        b.agent_size, b.agent_wakeup_time = p(self.size, C.c_int64), p(self.wakeup_time, C.c_int64)
        b.theta = p(self.theta, C.c_int32)
        b.mt_key = None
        b.mt_pos, b.mt_has_gauss, b.mt_gauss = p(self.mt_pos, C.c_int32), p(self.mt_has_gauss, C.c_int32), p(self.mt_gauss, C.c_double)
        b.mt_key_row, b.mt_seed = p(self.mt_key_row, C.c_int32), p(self.mt_seed, C.c_uint32)
        b.n_mt_rows = 0
        b.draw_flags = draw_flags
        self.c = b
        self.rng_mode = rng_mode

    def _fill_device_draw_parameters(self, groups):
        """The constructor arguments the device needs for its own draws: sigma_pv of the ZI agents, min / max order
        size of the momentum agents (in the type row's mm_min_size / mm_max_size)."""
        j = 0
        for g in groups:
            for row in g.template.types:
                ty = self.types[j]
                if row["klass"] in (_abi.ZI, _abi.HBL):
                    ty.sigma_pv = g.template.device_draws["sigma_pv"]
                if row["klass"] == _abi.MOMENTUM and (g.draw_flags & _abi.DRAW_MOMENTUM_SIZE):
                    ty.mm_min_size, ty.mm_max_size = g.template.device_draws["momentum_size_range"]
                j += 1

    @property
    def n_instances(self):
        return len(self.cfg)

    def h2d_bytes(self):
        n = self.cfg.nbytes + C.sizeof(self.types)
        for a in (self.agent_type, self.agent_theta, self.lat_to, self.lat_from, self.size, self.wakeup_time,
                  self.mt_seed, self.mt_pos, self.mt_has_gauss, self.mt_gauss, self.mt_key_row):
            n += a.nbytes
        return n


SPARSE_ZI = ("sparse_zi_100", "sparse_zi_1000")


def supported(config):
    return config in SPARSE_ZI or config == "rmsc03"


def build_sweep(config, seeds, extra_args=(), trace_capacity=0, rng_mode=_abi.RNG_MT19937):
    """SeededBatch of `config` for every seed (same flags for all); raises NotImplementedError for other configs.
    rng_mode=_abi.RNG_PHILOX keeps the seeds for the constructor draws only: the run itself uses counter-based streams."""
    return SeededBatch([build_group(config, seeds, extra_args)], trace_capacity=trace_capacity, rng_mode=rng_mode)


def build_group(config, seeds, extra_args=()):
    if config in SPARSE_ZI:
        return sweep_sparse_zi(config, seeds, extra_args)
    if config == "rmsc03":
        return sweep_rmsc03(seeds, extra_args)
    raise NotImplementedError("no vectorised sweep builder for config %r (use runtime.build_instances)" % config)
