"""ctypes wrapper of oracle/_build/libabides_oracle.so (test infrastructure only)."""
import ctypes as C
import os
import subprocess
import threading

import numpy as np

from abides_b200 import _abi

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "_build", "libabides_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_DIR, "abides_oracle.c")
    hdrs = [os.path.join(_DIR, "..", "include", h) for h in ("abides_b200.h", "abides_dd_math.h", "abides_math_tables.h")]
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(LIB_PATH) for p in [src] + hdrs)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _DIR] + (["-B"] if force else []))
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if os.path.exists(os.path.join(_DIR, "abides_oracle.c")):
            try:
                build()
            except Exception:
                if not os.path.exists(LIB_PATH):
                    raise
        L = C.CDLL(LIB_PATH)
        L.oracle_run_instance.argtypes = [
            C.POINTER(_abi.Batch), C.c_int, C.POINTER(_abi.InstanceResult), C.POINTER(_abi.AgentResult),
            C.POINTER(_abi.TraceRec), C.c_int64, C.POINTER(C.c_int64),
            C.POINTER(C.c_int64), C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_int64)]
        L.oracle_run_instance.restype = C.c_int
        L.oracle_run_range.argtypes = [C.POINTER(_abi.Batch), C.c_int, C.c_int,
                                       C.POINTER(_abi.InstanceResult), C.POINTER(_abi.AgentResult)]
        L.oracle_run_range.restype = C.c_int
        L.oracle_replay_book.argtypes = [C.c_int32, C.POINTER(C.c_int64), C.POINTER(_abi.BookOp), C.c_int32,
                                         C.c_int32, C.POINTER(_abi.BookEvent), C.POINTER(C.c_int32),
                                         C.POINTER(_abi.BookState)]
        L.oracle_replay_book.restype = C.c_int
        L.oracle_rng_probe.argtypes = [C.POINTER(C.c_uint32), C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                       C.c_double, C.c_double, C.POINTER(C.c_double)]
        L.oracle_rng_probe.restype = None
        L.oracle_dense_fundamental.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                               C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double, C.c_int64,
                                               C.POINTER(C.c_int64)]
        L.oracle_dense_fundamental.restype = None
        L.oracle_math_probe.argtypes = [C.c_int, C.c_double, C.c_double]
        L.oracle_math_probe.restype = C.c_double
        L.oracle_math_probe_n.argtypes = [C.c_int, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                          C.POINTER(C.c_double)]
        L.oracle_math_probe_n.restype = None
        _lib = L
    return _lib


def trace_to_array(buf, n):
    """TraceRec ctypes array -> int64[n, 7] (time, agent, code, p0..p3)."""
    a = np.frombuffer(buf, dtype=np.dtype([("time", "<i8"), ("agent", "<i4"), ("code", "<i4"),
                                           ("p0", "<i8"), ("p1", "<i8"), ("p2", "<i8"), ("p3", "<i8")]), count=n)
    out = np.empty((n, 7), np.int64)
    for j, k in enumerate(("time", "agent", "code", "p0", "p1", "p2", "p3")):
        out[:, j] = a[k]
    return out


def results_to_dict(hb, inst, res, ares):
    ic = hb.instances[inst]
    o, n = ic.agent_offset, ic.n_agents
    d = {f: getattr(res, f) for f, _ in _abi.InstanceResult._fields_ if f != "reserved"}
    d["cash"] = np.array([ares[o + a].cash for a in range(n)], np.int64)
    d["holdings"] = np.array([ares[o + a].holdings for a in range(n)], np.int64)
    d["mtm"] = np.array([ares[o + a].marked_to_market for a in range(n)], np.int64)
    d["final_valuation"] = np.array([ares[o + a].final_valuation for a in range(n)], np.float64)
    return d


def math_probe(what, x, y=None):
    """Vector probe of include/abides_dd_math.h: what = 0 log, 1 exp, 2 pow; +10 = abm_slow_* versions."""
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(np.zeros_like(x) if y is None else y, np.float64)
    out = np.empty_like(x)
    dp = C.POINTER(C.c_double)
    lib().oracle_math_probe_n(int(what), x.size, x.ctypes.data_as(dp), y.ctypes.data_as(dp), out.ctypes.data_as(dp))
    return out


def run_instance(hb, inst=0, trace_cap=0, flog_cap=0):
    """Run one instance of a lowering.HostBatch through the CPU oracle."""
    L = lib()
    res = _abi.InstanceResult()
    ares = (_abi.AgentResult * hb.n_agents_total)()
    tr = (_abi.TraceRec * max(1, trace_cap))()
    ntr = C.c_int64(0)
    ft = np.zeros(max(1, flog_cap), np.int64)
    fv = np.zeros(max(1, flog_cap), np.float64)
    nf = C.c_int64(0)
    rc = L.oracle_run_instance(C.byref(hb.c), inst, C.byref(res), ares, tr, trace_cap, C.byref(ntr),
                               ft.ctypes.data_as(C.POINTER(C.c_int64)), fv.ctypes.data_as(C.POINTER(C.c_double)),
                               flog_cap, C.byref(nf))
    if rc != 0:
        raise RuntimeError("oracle_run_instance failed: %d" % rc)
    d = results_to_dict(hb, inst, res, ares)
    d["n_trace"] = ntr.value
    if trace_cap:
        d["trace"] = trace_to_array(tr, min(ntr.value, trace_cap))
    if flog_cap:
        k = min(nf.value, flog_cap)
        d["f_time"], d["f_value"] = ft[:k].copy(), fv[:k].copy()
    return d


def run_batch(hb, n_threads=None, first=0, count=None):
    """Run instances [first, first+count) on `n_threads` host threads (one instance per thread at a
    time; ctypes releases the GIL).  Returns (InstanceResult array, AgentResult array)."""
    L = lib()
    ni = hb.n_instances
    count = ni - first if count is None else count
    n_threads = n_threads or os.cpu_count() or 1
    res = (_abi.InstanceResult * ni)()
    ares = (_abi.AgentResult * hb.n_agents_total)()
    nxt = [first]
    lock = threading.Lock()

    def work():
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= first + count:
                return
            L.oracle_run_range(C.byref(hb.c), i, 1, res, ares)

    ts = [threading.Thread(target=work) for _ in range(min(n_threads, count))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return res, ares


def replay_book(op_offset, ops, stream_history, max_events_per_op=64):
    L = lib()
    n_books = len(op_offset) - 1
    n_ops = int(op_offset[-1])
    ev = (_abi.BookEvent * max(1, n_ops * max_events_per_op))()
    cnt = np.zeros(max(1, n_ops), np.int32)
    st = (_abi.BookState * max(1, n_ops))()
    off = np.ascontiguousarray(op_offset, np.int64)
    rc = L.oracle_replay_book(n_books, off.ctypes.data_as(C.POINTER(C.c_int64)), ops, stream_history,
                              max_events_per_op, ev, cnt.ctypes.data_as(C.POINTER(C.c_int32)), st)
    if rc != 0:
        raise RuntimeError("oracle_replay_book failed")
    return ev, cnt, st


def dense_fundamental(state, r_bar, kappa, sigma_s, n_steps):
    """MeanRevertingOracle series from a numpy legacy state tuple; returns (int64[n_steps], advanced state)."""
    key = np.ascontiguousarray(state[1], np.uint32).copy()
    pos, hg, gs = C.c_int32(int(state[2])), C.c_int32(int(state[3])), C.c_double(float(state[4]))
    out = np.zeros(int(n_steps), np.int64)
    lib().oracle_dense_fundamental(key.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(pos), C.byref(hg), C.byref(gs),
                                   float(r_bar), float(kappa), float(sigma_s), int(n_steps),
                                   out.ctypes.data_as(C.POINTER(C.c_int64)))
    return out, ("MT19937", key, pos.value, hg.value, gs.value)
