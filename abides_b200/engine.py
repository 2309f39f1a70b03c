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
