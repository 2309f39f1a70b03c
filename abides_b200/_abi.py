"""ctypes mirror of include/abides_b200.h and the loader of the CUDA library.

There is deliberately no fallback here: if the shared library cannot be loaded the
import error propagates, and if no CUDA device is present `abides_create` fails.
"""
import ctypes as C
import os

MT_N = 624
EXTRA_STREAMS = 4
STREAM_ORACLE, STREAM_KERNEL, STREAM_LATENCY, STREAM_GLOBAL = 0, 1, 2, 3

# enum abides_agent_class
EXCHANGE, ZI, VALUE, NOISE, MOMENTUM, ADAPTIVE_MM, POV_EXEC, MARKET_MAKER, HBL, SCHEDULE_EXEC, SUBSCRIPTION, MARKET_REPLAY = range(12)
# enum abides_latency_kind
LAT_DETERMINISTIC, LAT_CUBIC, LAT_LEGACY_NOISE = range(3)
# enum abides_rng_mode
RNG_MT19937, RNG_PHILOX = 0, 1
OK, EINVAL, ENOMEM, ECUDA, ESTATE, EOVERFLOW = 0, -1, -2, -3, -4, -5     # enum abides_error

MSG_NAMES = {
    0: "WAKEUP", 1: "WHEN_MKT_OPEN", 2: "WHEN_MKT_CLOSE", 3: "QUERY_LAST_TRADE", 4: "QUERY_SPREAD",
    5: "QUERY_ORDER_STREAM", 6: "QUERY_TRANSACTED_VOLUME", 7: "LIMIT_ORDER", 8: "MARKET_ORDER",
    9: "CANCEL_ORDER", 10: "MODIFY_ORDER", 11: "ORDER_ACCEPTED", 12: "ORDER_EXECUTED",
    13: "ORDER_CANCELLED", 14: "ORDER_MODIFIED", 15: "MKT_CLOSED",
    16: "MARKET_DATA_SUBSCRIPTION_REQUEST", 17: "MARKET_DATA_SUBSCRIPTION_CANCELLATION", 18: "MARKET_DATA",
}
MD_MAX_LEVELS = 8
MSG_CODES = {v: k for k, v in MSG_NAMES.items()}
TRACE_FUNDAMENTAL = 64          # trace code of an oracle step (include/abides_b200.h: enum abides_trace_code)
TRACE_ORDER_PLACED, TRACE_BOOK_TOP, TRACE_LAST_TRADE, TRACE_EXCH_SENT = 65, 66, 67, 128
TRACE_EXCHANGE_LOG = 1          # enum abides_trace_flags
INST_NO_ORACLE = 1              # enum abides_instance_flags
MSG_REPLY = 32


class AgentType(C.Structure):
    _fields_ = [
        ("klass", C.c_int32), ("q_max", C.c_int32),
        ("starting_cash", C.c_int64),
        ("sigma_n", C.c_double), ("r_bar", C.c_double), ("kappa", C.c_double),
        ("sigma_s", C.c_double), ("eta", C.c_double), ("lambda_a", C.c_double),
        ("R_min", C.c_int64), ("R_max", C.c_int64),
        ("wake_up_freq", C.c_int64),
        ("pov", C.c_double), ("level_spacing", C.c_double), ("spread_alpha", C.c_double),
        ("skew_beta", C.c_double),
        ("min_order_size", C.c_int64), ("num_ticks", C.c_int64), ("cancel_limit_delay", C.c_int64),
        ("backstop_quantity", C.c_int64), ("window_size", C.c_int64),
        ("anchor", C.c_int32), ("log_orders", C.c_int32),
        ("percent_aggr", C.c_double), ("depth_spread", C.c_int64),
        ("exec_start_time", C.c_int64), ("exec_end_time", C.c_int64), ("exec_lookback", C.c_int64),
        ("exec_quantity", C.c_double), ("exec_is_buy", C.c_int32), ("exec_trade", C.c_int32),
        ("mm_min_size", C.c_int64), ("mm_max_size", C.c_int64), ("mm_num_levels", C.c_int32), ("hbl_L", C.c_int32),
        ("subscribe", C.c_int32), ("sub_levels", C.c_int32), ("subscribe_freq", C.c_double),
        ("sigma_pv", C.c_double),
    ]


class InstanceCfg(C.Structure):
    _fields_ = [
        ("start_time", C.c_int64), ("stop_time", C.c_int64),
        ("mkt_open", C.c_int64), ("mkt_close", C.c_int64),
        ("default_computation_delay", C.c_int64),
        ("exch_pipeline_delay", C.c_int64), ("exch_computation_delay", C.c_int64),
        ("stream_history", C.c_int32), ("exch_log_orders", C.c_int32),
        ("latency_kind", C.c_int32), ("n_latency_noise", C.c_int32),
        ("jitter", C.c_double), ("jitter_clip", C.c_double), ("jitter_unit", C.c_double),
        ("r_bar", C.c_double), ("kappa", C.c_double), ("fund_vol", C.c_double),
        ("megashock_lambda_a", C.c_double), ("megashock_mean", C.c_double), ("megashock_var", C.c_double),
        ("megashock_time0", C.c_int64), ("megashock_value0", C.c_double),
        ("agent_offset", C.c_int64),
        ("n_agents", C.c_int32), ("type_base", C.c_int32),
        ("theta_offset", C.c_int64),
        ("seed", C.c_uint64),
        ("flags", C.c_int64), ("reserved", C.c_int64),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("n_instances", C.c_int32), ("n_types", C.c_int32),
        ("n_agents_total", C.c_int64), ("n_theta_total", C.c_int64),
        ("rng_mode", C.c_int32), ("trace_capacity", C.c_int32),
        ("instances", C.POINTER(InstanceCfg)),
        ("types", C.POINTER(AgentType)),
        ("agent_type", C.POINTER(C.c_int32)),
        ("lat_to_exch", C.POINTER(C.c_double)),
        ("lat_from_exch", C.POINTER(C.c_double)),
        ("agent_size", C.POINTER(C.c_int64)),
        ("agent_wakeup_time", C.POINTER(C.c_int64)),
        ("agent_theta", C.POINTER(C.c_int32)),
        ("theta", C.POINTER(C.c_int32)),
        ("mt_key", C.POINTER(C.c_uint32)),
        ("mt_pos", C.POINTER(C.c_int32)),
        ("mt_has_gauss", C.POINTER(C.c_int32)),
        ("mt_gauss", C.POINTER(C.c_double)),
        ("mt_key_row", C.POINTER(C.c_int32)),
        ("mt_seed", C.POINTER(C.c_uint32)),
        ("n_mt_rows", C.c_int64),
        ("draw_flags", C.c_int32), ("trace_flags", C.c_int32),
    ]


DRAW_ZI_THETA, DRAW_MOMENTUM_SIZE = 1, 2               # enum abides_draw_flags


class InstanceResult(C.Structure):
    _fields_ = [
        ("ttl_messages", C.c_int64), ("delivered", C.c_int64), ("final_time", C.c_int64),
        ("slowest_agent_finish_time", C.c_int64), ("last_trade", C.c_int64),
        ("final_fundamental", C.c_int64), ("n_orders", C.c_int64), ("n_fills", C.c_int64),
        ("traded_volume", C.c_int64),
        ("status", C.c_int32), ("queue_peak", C.c_int32),
        ("book_peak_levels", C.c_int32), ("book_peak_orders", C.c_int32),
        ("reserved", C.c_int64 * 2),
    ]


class AgentResult(C.Structure):
    _fields_ = [
        ("cash", C.c_int64), ("holdings", C.c_int64), ("marked_to_market", C.c_int64),
        ("final_valuation", C.c_double),
    ]


class TraceRec(C.Structure):
    _fields_ = [
        ("time", C.c_int64), ("agent", C.c_int32), ("code", C.c_int32),
        ("p0", C.c_int64), ("p1", C.c_int64), ("p2", C.c_int64), ("p3", C.c_int64),
    ]


class BookOp(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("agent", C.c_int32), ("order_id", C.c_int64), ("price", C.c_int64),
        ("qty", C.c_int64), ("is_buy", C.c_int32), ("reserved", C.c_int32), ("new_qty", C.c_int64),
        ("time", C.c_int64),
    ]


class BookEvent(C.Structure):
    _fields_ = [
        ("op_index", C.c_int32), ("kind", C.c_int32), ("agent", C.c_int32), ("is_buy", C.c_int32),
        ("order_id", C.c_int64), ("price", C.c_int64), ("qty", C.c_int64), ("fill_price", C.c_int64),
    ]


class BookState(C.Structure):
    _fields_ = [
        ("bid", C.c_int64), ("bid_vol", C.c_int64), ("ask", C.c_int64), ("ask_vol", C.c_int64),
        ("last_trade", C.c_int64), ("n_bid_levels", C.c_int32), ("n_ask_levels", C.c_int32),
    ]


_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ABIDES_B200_LIB") or os.path.join(_PKG_DIR, "csrc", "libabides_b200.so")

EXPORTS = [
    "abides_create", "abides_destroy", "abides_last_error", "abides_version", "abides_load",
    "abides_reset", "abides_run", "abides_read_results", "abides_read_trace", "abides_last_run_ms", "abides_last_reset_ms", "abides_read_profile",
    "abides_replay_book", "abides_set_engine", "abides_engine_info", "abides_launch_count", "abides_dense_fundamental",
]
ENGINE_AUTO, ENGINE_WARP, ENGINE_LANE = 0, 1, 2        # enum abides_engine
LANE_AUTO_MIN_INSTANCES_ZI, LANE_AUTO_MIN_INSTANCES = 8192, 24576
VARIANT_NAMES = {0: "warp/full", 1: "warp/zi_legacy", 2: "warp/zi_cubic", 3: "warp/rmsc", 4: "warp/rmsc01",
                 5: "warp/rmsc02", 6: "warp/exec", 16: "lane/zi_legacy", 17: "lane/zi_cubic", 18: "lane/rmsc",
                 19: "lane/full"}

_lib = None


def lib():
    """Load the CUDA engine (abides_b200/csrc/libabides_b200.so).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "abides_b200: CUDA library %s not built -- run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    H = C.c_void_p
    L.abides_create.argtypes = [C.c_int, C.POINTER(H)]
    L.abides_create.restype = C.c_int
    L.abides_destroy.argtypes = [H]
    L.abides_destroy.restype = None
    L.abides_last_error.argtypes = []
    L.abides_last_error.restype = C.c_char_p
    L.abides_version.argtypes = []
    L.abides_version.restype = C.c_char_p
    L.abides_load.argtypes = [H, C.POINTER(Batch)]
    L.abides_load.restype = C.c_int
    L.abides_reset.argtypes = [H]
    L.abides_reset.restype = C.c_int
    L.abides_run.argtypes = [H, C.c_int64]
    L.abides_run.restype = C.c_int
    L.abides_read_results.argtypes = [H, C.POINTER(InstanceResult), C.POINTER(AgentResult)]
    L.abides_read_results.restype = C.c_int
    L.abides_read_trace.argtypes = [H, C.c_int32, C.POINTER(TraceRec), C.c_int64, C.POINTER(C.c_int64)]
    L.abides_read_trace.restype = C.c_int
    L.abides_last_run_ms.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    L.abides_last_run_ms.restype = C.c_int
    L.abides_last_reset_ms.argtypes = [H, C.POINTER(C.c_float)]
    L.abides_last_reset_ms.restype = C.c_int
    L.abides_read_profile.argtypes = [H, C.POINTER(C.c_uint64), C.c_int32]
    L.abides_read_profile.restype = C.c_int
    L.abides_replay_book.argtypes = [H, C.c_int32, C.POINTER(C.c_int64), C.POINTER(BookOp), C.c_int32,
                                     C.c_int32, C.POINTER(BookEvent), C.POINTER(C.c_int32),
                                     C.POINTER(BookState)]
    L.abides_replay_book.restype = C.c_int
    L.abides_set_engine.argtypes = [H, C.c_int32]
    L.abides_set_engine.restype = C.c_int
    L.abides_engine_info.argtypes = [H, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.abides_engine_info.restype = C.c_int
    L.abides_launch_count.argtypes = [H, C.POINTER(C.c_int64)]
    L.abides_launch_count.restype = C.c_int
    L.abides_dense_fundamental.argtypes = [H, C.c_int32, C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                           C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double, C.c_int64,
                                           C.POINTER(C.c_int64)]
    L.abides_dense_fundamental.restype = C.c_int
    _lib = L
    return L
