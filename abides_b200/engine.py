"""Thin host wrapper over the C ABI (include/abides_b200.h).  No compute happens in Python and
there is no CPU fallback: every method ends in a call into libabides_b200.so."""
import ctypes as C

import numpy as np

from . import _abi

INT64_MAX = (1 << 63) - 1


class EngineError(RuntimeError):
    pass


class Engine:
    """One handle per CUDA device.  load() -> run() -> results()."""

    def __init__(self, device=0):
        self._L = _abi.lib()
        self._h = C.c_void_p()
        rc = self._L.abides_create(int(device), C.byref(self._h))
        if rc != 0:
            raise EngineError("abides_create(%d) failed (%d): %s" % (device, rc, self.last_error()))
        self.device = device
        self.batch = None

    def last_error(self):
        s = self._L.abides_last_error()
        return s.decode() if s else ""

    def _ck(self, rc, what):
        if rc != 0:
            raise EngineError("%s failed (%d): %s" % (what, rc, self.last_error()))

    def close(self):
        if self._h:
            self._L.abides_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load(self, host_batch):
        """Host tables -> device (H2D) and build the initial device state."""
        self._ck(self._L.abides_load(self._h, C.byref(host_batch.c)), "abides_load")
        self.batch = host_batch

    def set_engine(self, engine):
        """Force the device engine of the following loads: _abi.ENGINE_AUTO / ENGINE_WARP / ENGINE_LANE, or its name."""
        engine = {"auto": _abi.ENGINE_AUTO, "warp": _abi.ENGINE_WARP, "lane": _abi.ENGINE_LANE}.get(engine, engine)
        self._ck(self._L.abides_set_engine(self._h, int(engine)), "abides_set_engine")

    def engine_info(self):
        """(engine, variant) the loaded batch runs on; names in _abi.VARIANT_NAMES."""
        e, v = C.c_int32(0), C.c_int32(0)
        self._ck(self._L.abides_engine_info(self._h, C.byref(e), C.byref(v)), "abides_engine_info")
        return e.value, v.value

    def dense_fundamental(self, states, r_bar, kappa, sigma_s, n_steps):
        """MeanRevertingOracle.generate_fundamental_value_series (util/oracle/MeanRevertingOracle.py:49-81) for many
        series at once.  `states` is a list of numpy legacy state tuples (np.random.get_state()) or of integer seeds;
        returns (int64[n_series, n_steps], advanced states)."""
        from .vecrng import init_genrand
        n = len(states)
        key = np.zeros((n, _abi.MT_N), np.uint32)
        pos = np.full(n, _abi.MT_N, np.int32)
        hg = np.zeros(n, np.int32)
        gs = np.zeros(n, np.float64)
        for i, st in enumerate(states):
            if isinstance(st, (int, np.integer)):
                key[i] = init_genrand([int(st)])[0]
            else:
                key[i], pos[i], hg[i], gs[i] = np.asarray(st[1], np.uint32), int(st[2]), int(st[3]), float(st[4])
        out = np.zeros((n, int(n_steps)), np.int64)
        p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        self._ck(self._L.abides_dense_fundamental(self._h, n, p(key, C.c_uint32), p(pos, C.c_int32), p(hg, C.c_int32),
                                                  p(gs, C.c_double), float(r_bar), float(kappa), float(sigma_s),
                                                  int(n_steps), p(out, C.c_int64)), "abides_dense_fundamental")
        return out, [("MT19937", key[i].copy(), int(pos[i]), int(hg[i]), float(gs[i])) for i in range(n)]

    def launch_count(self):
        """Kernels launched through this handle so far."""
        n = C.c_int64(0)
        self._ck(self._L.abides_launch_count(self._h, C.byref(n)), "abides_launch_count")
        return n.value

    def reset(self):
        """Re-arm the resident batch (no host->device traffic)."""
        self._ck(self._L.abides_reset(self._h), "abides_reset")

    def run(self, until_time=INT64_MAX):
        self._ck(self._L.abides_run(self._h, int(until_time)), "abides_run")

    def last_run_ms(self):
        ms, n = C.c_float(0), C.c_int32(0)
        self._ck(self._L.abides_last_run_ms(self._h, C.byref(ms), C.byref(n)), "abides_last_run_ms")
        return ms.value, n.value

    def last_reset_ms(self):
        ms = C.c_float(0)
        self._ck(self._L.abides_last_reset_ms(self._h, C.byref(ms)), "abides_last_reset_ms")
        return ms.value

    def results(self, per_agent=True):
        hb = self.batch
        res = (_abi.InstanceResult * hb.n_instances)()
        ares = (_abi.AgentResult * hb.n_agents_total)() if per_agent else None
        self._ck(self._L.abides_read_results(self._h, res, ares), "abides_read_results")
        return res, ares

    def results_arrays(self):
        """Per-instance / per-agent results as numpy structured arrays (zero-copy views)."""
        res, ares = self.results(True)
        r = np.frombuffer(res, dtype=INSTANCE_RESULT_DTYPE)
        a = np.frombuffer(ares, dtype=AGENT_RESULT_DTYPE)
        return r, a

    def trace(self, instance=0):
        hb = self.batch
        cap = hb.c.trace_capacity
        buf = (_abi.TraceRec * max(1, cap))()
        n = C.c_int64(0)
        self._ck(self._L.abides_read_trace(self._h, instance, buf, cap, C.byref(n)), "abides_read_trace")
        k = min(n.value, cap)
        a = np.frombuffer(buf, dtype=TRACE_DTYPE, count=k)
        out = np.empty((k, 7), np.int64)
        for j, f in enumerate(("time", "agent", "code", "p0", "p1", "p2", "p3")):
            out[:, j] = a[f]
        return out, n.value


TRACE_DTYPE = np.dtype([("time", "<i8"), ("agent", "<i4"), ("code", "<i4"),
                        ("p0", "<i8"), ("p1", "<i8"), ("p2", "<i8"), ("p3", "<i8")])
INSTANCE_RESULT_DTYPE = np.dtype([
    ("ttl_messages", "<i8"), ("delivered", "<i8"), ("final_time", "<i8"),
    ("slowest_agent_finish_time", "<i8"), ("last_trade", "<i8"), ("final_fundamental", "<i8"),
    ("n_orders", "<i8"), ("n_fills", "<i8"), ("traded_volume", "<i8"),
    ("status", "<i4"), ("queue_peak", "<i4"), ("book_peak_levels", "<i4"), ("book_peak_orders", "<i4"),
    ("reserved", "<i8", (2,))])
AGENT_RESULT_DTYPE = np.dtype([("cash", "<i8"), ("holdings", "<i8"), ("marked_to_market", "<i8"),
                               ("final_valuation", "<f8")])
assert INSTANCE_RESULT_DTYPE.itemsize == C.sizeof(_abi.InstanceResult)
assert AGENT_RESULT_DTYPE.itemsize == C.sizeof(_abi.AgentResult)
assert TRACE_DTYPE.itemsize == C.sizeof(_abi.TraceRec)


def make_book_ops(rows):
    """int64[n,8] rows (kind, agent, order_id, price, qty, is_buy, new_qty, time) -> ctypes abides_book_op array."""
    n = len(rows)
    ops = (_abi.BookOp * max(1, n))()
    for i, r in enumerate(rows):
        o = ops[i]
        o.kind, o.agent, o.order_id, o.price, o.qty, o.is_buy, o.new_qty, o.time = [int(x) for x in r]
    return ops


def replay_book(engine, op_offset, rows, stream_history, max_events_per_op=64):
    """util/OrderBook.py replay on the device (abides_replay_book): returns (events int64[m,8], counts, states int64[n,7])."""
    n_books = len(op_offset) - 1
    n = int(op_offset[-1])
    ops = make_book_ops(rows)
    ev = (_abi.BookEvent * max(1, n * max_events_per_op))()
    cnt = np.zeros(max(1, n), np.int32)
    st = (_abi.BookState * max(1, n))()
    off = np.ascontiguousarray(op_offset, np.int64)
    engine._ck(engine._L.abides_replay_book(engine._h, n_books, off.ctypes.data_as(C.POINTER(C.c_int64)), ops,
                                            int(stream_history), int(max_events_per_op), ev,
                                            cnt.ctypes.data_as(C.POINTER(C.c_int32)), st), "abides_replay_book")
    return unpack_replay(ev, cnt[:n], st, n, max_events_per_op)


def unpack_replay(ev, cnt, st, n, max_events_per_op):
    events = []
    for i in range(n):
        for k in range(min(int(cnt[i]), max_events_per_op)):
            e = ev[i * max_events_per_op + k]
            events.append((i, e.kind, e.agent, e.is_buy, e.order_id, e.price, e.qty, e.fill_price))
    states = np.array([(s.bid, s.bid_vol, s.ask, s.ask_vol, s.last_trade, s.n_bid_levels, s.n_ask_levels)
                       for s in st[:n]], np.int64).reshape(-1, 7)
    return np.array(events, np.int64).reshape(-1, 8), np.asarray(cnt), states
