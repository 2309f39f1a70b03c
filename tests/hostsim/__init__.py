"""TEST INFRASTRUCTURE ONLY: the lane engine's device code compiled for the host (tests/hostsim/lane_host.cpp), so the
CPU suite can check the engine's logic against the reference goldens without a GPU.  Never imported by abides_b200."""
import ctypes as C
import os
import subprocess

import numpy as np

from abides_b200 import _abi
from abides_b200.engine import INSTANCE_RESULT_DTYPE, AGENT_RESULT_DTYPE, TRACE_DTYPE

_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(_DIR))
CSRC = os.path.join(ROOT, "abides_b200", "csrc")
# ABIDES_HOSTSIM_DEFINES="LANE_STAGE=0 LANE_ONE_CTX=1": build options of the lane engine under test (csrc/lane_types.h)
DEFINES = os.environ.get("ABIDES_HOSTSIM_DEFINES", "").split()
LIB_PATH = os.path.join(_DIR, "_build", "liblane_host%s.so" % "".join("_" + d.replace("=", "") for d in DEFINES))
LV_ZI_LEGACY, LV_ZI_CUBIC, LV_RMSC, LV_FULL = range(4)
_lib = None


def build(force=False):
    src = os.path.join(_DIR, "lane_host.cpp")
    deps = [src] + [os.path.join(CSRC, f) for f in ("lane_device.inc", "lane_types.h", "lane_variants.inc",
                                                     "lane_host_common.h")]
    deps += [os.path.join(ROOT, "include", h) for h in ("abides_b200.h", "abides_dd_math.h", "abides_math_tables.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps):
        os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-g", "-ffp-contract=off", "-fPIC", "-shared"] +
                              ["-D" + d for d in DEFINES] +
                              ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB_PATH, src])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.hostsim_run.argtypes = [C.POINTER(_abi.Batch), C.c_int, C.POINTER(_abi.InstanceResult),
                                  C.POINTER(_abi.AgentResult), C.c_int32, C.POINTER(_abi.TraceRec), C.c_int64,
                                  C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        L.hostsim_run.restype = C.c_int
        L.hostsim_replay_book.argtypes = [C.c_int32, C.POINTER(C.c_int64), C.POINTER(_abi.BookOp), C.c_int32, C.c_int32,
                                          C.POINTER(_abi.BookEvent), C.POINTER(C.c_int32), C.POINTER(_abi.BookState)]
        L.hostsim_replay_book.restype = C.c_int
        _lib = L
    return _lib


def run(hb, variant=-1, trace_instance=-1, trace_cap=0):
    """Run every instance of a lowering.HostBatch through the host build of the lane engine.
    Returns (res, ares, trace int64[n,7] or None, n_trace, variant)."""
    L = lib()
    res = (_abi.InstanceResult * hb.n_instances)()
    ares = (_abi.AgentResult * hb.n_agents_total)()
    tr = (_abi.TraceRec * max(1, trace_cap))()
    ntr, var = C.c_int64(0), C.c_int32(-1)
    rc = L.hostsim_run(C.byref(hb.c), variant, res, ares, trace_instance, tr, trace_cap, C.byref(ntr), C.byref(var))
    if rc != 0:
        raise RuntimeError("hostsim_run failed: %d" % rc)
    r = np.frombuffer(res, dtype=INSTANCE_RESULT_DTYPE).copy()
    a = np.frombuffer(ares, dtype=AGENT_RESULT_DTYPE).copy()
    trace = None
    if trace_cap:
        k = min(ntr.value, trace_cap)
        t = np.frombuffer(tr, dtype=TRACE_DTYPE, count=k)
        trace = np.empty((k, 7), np.int64)
        for j, f in enumerate(("time", "agent", "code", "p0", "p1", "p2", "p3")):
            trace[:, j] = t[f]
    return r, a, trace, ntr.value, var.value


def replay_book(op_offset, ops, stream_history, max_events_per_op=64):
    L = lib()
    n_books = len(op_offset) - 1
    n_ops = int(op_offset[-1])
    ev = (_abi.BookEvent * max(1, n_ops * max_events_per_op))()
    cnt = np.zeros(max(1, n_ops), np.int32)
    st = (_abi.BookState * max(1, n_ops))()
    off = np.ascontiguousarray(op_offset, np.int64)
    rc = L.hostsim_replay_book(n_books, off.ctypes.data_as(C.POINTER(C.c_int64)), ops, stream_history,
                               max_events_per_op, ev, cnt.ctypes.data_as(C.POINTER(C.c_int32)), st)
    if rc != 0:
        raise RuntimeError("hostsim_replay_book failed")
    return ev, cnt, st
