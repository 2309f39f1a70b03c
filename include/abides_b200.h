/*
 * abides_b200.h -- C ABI of the B200-native batched ABIDES engine.
 *
 * The reference (abides-sim/abides) is pure Python and has no FFI; the surface this
 * library sits under is the call a config module makes,
 *     Kernel(...).runner(agents, startTime, stopTime, ..., agentLatencyModel,
 *                        defaultComputationDelay, oracle, log_dir)   Kernel.py:51-55
 * `abides_b200.Kernel.runner` lowers the agent / oracle / latency objects it is
 * handed to the plain tables below and calls these entry points through ctypes.
 *
 * Conventions: every function returns 0 on success or a negative ABIDES_E* code and
 * leaves a message for abides_last_error(); the caller owns every host buffer, the
 * library owns every device buffer; one handle per device; no internal threads; there
 * is NO CPU fallback -- without a CUDA device abides_create() fails.
 *
 * Times are integer nanoseconds (pandas Timestamp.value), prices integer cents.
 */
#ifndef ABIDES_B200_H
#define ABIDES_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABIDES_MT_N 624
#define ABIDES_EXTRA_STREAMS 4          /* per instance, after the per-agent streams */
#define ABIDES_STREAM_ORACLE  0         /* symbols[sym]['random_state']   config/sparse_zi_100.py:130 */
#define ABIDES_STREAM_KERNEL  1         /* Kernel(random_state=...)       Kernel.py:12,:380 */
#define ABIDES_STREAM_LATENCY 2         /* LatencyModel(random_state=...) model/LatencyModel.py:137 */
#define ABIDES_STREAM_GLOBAL  3         /* the process-global np.random   e.g. agent/ValueAgent.py:199 */

enum abides_error {
  ABIDES_OK = 0, ABIDES_EINVAL = -1, ABIDES_ENOMEM = -2, ABIDES_ECUDA = -3,
  ABIDES_ESTATE = -4, ABIDES_EOVERFLOW = -5
};

/* agent classes */
enum abides_agent_class {
  ABIDES_EXCHANGE = 0,      /* agent/ExchangeAgent.py */
  ABIDES_ZI = 1,            /* agent/ZeroIntelligenceAgent.py */
  ABIDES_VALUE = 2,         /* agent/ValueAgent.py */
  ABIDES_NOISE = 3,         /* agent/NoiseAgent.py */
  ABIDES_MOMENTUM = 4,      /* agent/examples/MomentumAgent.py */
  ABIDES_ADAPTIVE_MM = 5,   /* agent/market_makers/AdaptiveMarketMakerAgent.py */
  ABIDES_POV_EXEC = 6,      /* agent/execution/POVExecutionAgent.py */
  ABIDES_MARKET_MAKER = 7,  /* agent/market_makers/MarketMakerAgent.py (polling mode, subscribe=False) */
  ABIDES_HBL = 8,           /* agent/HeuristicBeliefLearningAgent.py (a ZeroIntelligenceAgent subclass) */
  ABIDES_SCHEDULE_EXEC = 9, /* agent/execution/ExecutionAgent.py with a schedule: TWAPExecutionAgent.py, VWAPExecutionAgent.py */
  ABIDES_SUBSCRIPTION = 10, /* agent/examples/SubscriptionAgent.py: subscribes to market data, never trades */
  /* agent/examples/MarketReplayAgent.py: replays an order tape (lane engine only).  The agent's `theta` rows hold the
   * tape, six ints per row in time order -- time lo, time hi (ns), ORDER_ID, PRICE, SIZE, BUY (1 / 0) --, agent_size the
   * number of rows, the type row's wake_up_freq = first tape time - mkt_open (getWakeFrequency :64-66) */
  ABIDES_MARKET_REPLAY = 11
};

enum abides_latency_kind {
  ABIDES_LAT_DETERMINISTIC = 0,   /* model/LatencyModel.py:142-143 */
  ABIDES_LAT_CUBIC = 1,           /* model/LatencyModel.py:131-140 */
  ABIDES_LAT_LEGACY_NOISE = 2     /* Kernel.py:378-381 (agentLatency matrix + latencyNoise) */
};

enum abides_rng_mode {
  ABIDES_RNG_MT19937 = 0,   /* the reference's np.random.RandomState streams, bit-exact */
  ABIDES_RNG_PHILOX = 1     /* counter-based Philox4x32-10 keyed by (instance seed, stream) */
};

/* Strategy parameters shared by all agents of one "type" (one row per distinct ctor kwargs set). */
typedef struct abides_agent_type {
  int32_t klass;              /* enum abides_agent_class */
  int32_t q_max;              /* ZI */
  int64_t starting_cash;
  double  sigma_n, r_bar, kappa, sigma_s, eta, lambda_a;     /* ZI / Value estimator */
  int64_t R_min, R_max;       /* ZI */
  int64_t wake_up_freq;       /* Momentum / AdaptiveMM / POV period, ns */
  double  pov, level_spacing, spread_alpha, skew_beta;       /* AdaptiveMM */
  int64_t min_order_size, num_ticks, cancel_limit_delay;
  int64_t backstop_quantity;  /* < 0: None */
  int64_t window_size;        /* < 0: 'adaptive' */
  int32_t anchor;             /* 0 middle, 1 bottom, 2 top */
  int32_t log_orders;         /* TradingAgent.log_orders: only affects order-id consumption of order 0 */
  double  percent_aggr;       /* Value: 0.1 */
  int64_t depth_spread;       /* Value: 2 */
  /* POVExecutionAgent (agent/execution/POVExecutionAgent.py:14-33); also uses pov and wake_up_freq (= freq) */
  int64_t exec_start_time, exec_end_time;   /* trades while start < now < end */
  int64_t exec_lookback;      /* lookback_period of its QUERY_TRANSACTED_VOLUME, ns */
  double  exec_quantity;      /* total quantity to execute */
  int32_t exec_is_buy;        /* direction == "BUY" */
  int32_t exec_trade;         /* trade flag: 0 = heartbeat only */
  /* TWAP / VWAP ExecutionAgent (agent/execution/ExecutionAgent.py:10-30): execution_time_horizon is the regular
   * grid exec_start_time, + wake_up_freq, ..., exec_end_time; exec_quantity / exec_is_buy / exec_trade as above; the
   * agent's `theta` rows hold schedule[slot] for the slots that place limit orders (all but the last two);
   * abides_agent_result.final_valuation returns its arrival mid price (NaN if never set) */
  /* MarketMakerAgent (agent/market_makers/MarketMakerAgent.py:27-50); also uses wake_up_freq */
  int64_t mm_min_size, mm_max_size;
  int32_t mm_num_levels;      /* subscribe_num_levels: depth of its spread query, 2x = price levels quoted per side */
  int32_t hbl_L;              /* HeuristicBeliefLearningAgent: order-stream buckets it looks at */
  /* market-data subscriptions (TradingAgent.requestDataSubscription :170-173, ExchangeAgent.py:285-317):
   * MarketMakerAgent / MomentumAgent with subscribe=True and the SubscriptionAgent ask once, at their first wake,
   * for `sub_levels` book levels per side (<= ABIDES_MD_MAX_LEVELS) no more often than every `subscribe_freq` ns */
  int32_t subscribe;
  int32_t sub_levels;
  double  subscribe_freq;
  double  sigma_pv;           /* ZI / HBL private-value variance (ZeroIntelligenceAgent.py:46-50); only read when the
                                 device draws theta itself (ABIDES_DRAW_ZI_THETA) */
} abides_agent_type;
#define ABIDES_MD_MAX_LEVELS 8

/* One simulation instance (one reference process). */
typedef struct abides_instance_cfg {
  int64_t start_time, stop_time;          /* Kernel.runner startTime / stopTime */
  int64_t mkt_open, mkt_close;            /* ExchangeAgent */
  int64_t default_computation_delay;      /* Kernel.runner defaultComputationDelay */
  int64_t exch_pipeline_delay, exch_computation_delay;
  int32_t stream_history;                 /* ExchangeAgent.stream_history */
  int32_t exch_log_orders;
  int32_t latency_kind;                   /* enum abides_latency_kind */
  int32_t n_latency_noise;                /* len(latencyNoise) for the legacy model */
  double  jitter, jitter_clip, jitter_unit;                 /* cubic model scalars */
  double  r_bar, kappa, fund_vol;         /* SparseMeanRevertingOracle symbol parameters */
  double  megashock_lambda_a, megashock_mean, megashock_var;
  int64_t megashock_time0;                /* first megashock, drawn in the oracle ctor (:67-72) */
  double  megashock_value0;
  int64_t agent_offset;                   /* first row of this instance in the per-agent tables */
  int32_t n_agents;                       /* rows (agent ids 0..n_agents-1; id 0 is the exchange) */
  int32_t type_base;                      /* first row of this instance's types in `types` */
  int64_t theta_offset;                   /* first int of this instance in `theta` */
  uint64_t seed;                          /* Philox mode: instance key */
  int64_t flags;                          /* enum abides_instance_flags */
  int64_t reserved;
} abides_instance_cfg;
enum abides_instance_flags {
  ABIDES_INST_NO_ORACLE = 1               /* Kernel.runner(oracle=None) (config/marketreplay.py): the exchange has no opening
                                             price (ExchangeAgent.py:84-88), no agent may ask the oracle */
};

/* A batch of instances: tables are concatenated over instances in instance order. */
typedef struct abides_batch {
  int32_t n_instances;
  int32_t n_types;                        /* rows in `types` */
  int64_t n_agents_total;                 /* sum of n_agents */
  int64_t n_theta_total;                  /* ints in `theta` */
  int32_t rng_mode;                       /* enum abides_rng_mode */
  int32_t trace_capacity;                 /* per-instance trace records to keep (0 = off) */
  const abides_instance_cfg* instances;   /* [n_instances] */
  const abides_agent_type*   types;       /* [n_types] */
  const int32_t* agent_type;              /* [n_agents_total] index relative to type_base */
  const double*  lat_to_exch;             /* [n_agents_total] min latency agent -> exchange (ns) */
  const double*  lat_from_exch;           /* [n_agents_total] min latency exchange -> agent (ns) */
  const int64_t* agent_size;              /* [n_agents_total] Noise/Value/Momentum order size */
  const int64_t* agent_wakeup_time;       /* [n_agents_total] Noise: wakeup_time[0] */
  const int32_t* agent_theta;             /* [n_agents_total] offset into theta rel. to theta_offset, -1 none */
  const int32_t* theta;                   /* [n_theta_total] ZI private values, 2*q_max each */
  /* MT19937 mode: one stream per agent row plus ABIDES_EXTRA_STREAMS per instance, laid out as
   * stream(inst, s) = agent_offset + ABIDES_EXTRA_STREAMS*inst + s, s < n_agents for agent streams,
   * s = n_agents + ABIDES_STREAM_* for the extra ones.  NULL in Philox mode. */
  const uint32_t* mt_key;                 /* [n_streams][624] ([n_mt_rows][624] when mt_key_row is given) */
  const int32_t*  mt_pos;                 /* [n_streams] */
  const int32_t*  mt_has_gauss;           /* [n_streams] */
  const double*   mt_gauss;               /* [n_streams] */
  /* Compact stream description (optional; NULL / 0 = every stream carries its full state, row = stream).  A seed sweep
   * of n instances of a 1000-agent config has 10^3 n streams of 2.5 KB each, but every one of them is
   * np.random.RandomState(seed) after a known number of draws -- so the host may hand over the seed and the device
   * runs init_genrand and the draws itself:
   *   mt_key_row[s] >= 0 : row of stream s's 624-word state in mt_key (mt_pos = numpy's position in it);
   *   mt_key_row[s] == -1: stream s is RandomState(seed = mt_seed[s]) from which mt_pos[s] 32-bit words have been
   *                        drawn so far; mt_has_gauss / mt_gauss are numpy's cached gaussian at that point. */
  const int32_t*  mt_key_row;             /* [n_streams] or NULL */
  const uint32_t* mt_seed;                /* [n_streams] or NULL */
  int64_t n_mt_rows;                      /* rows of mt_key when mt_key_row is given */
  /* constructor-time draws the device performs from the agents' own (seeded) streams, enum abides_draw_flags */
  int32_t draw_flags;
  int32_t trace_flags;                    /* enum abides_trace_flags: extra trace records (only with trace_capacity > 0) */
} abides_batch;

enum abides_trace_flags {
  /* Everything the reference's EXCHANGE_AGENT.bz2 and ORDERBOOK_<symbol>_FULL.bz2 archives hold beyond the dispatched
   * events (agent/ExchangeAgent.py:147-150,:404-408, util/OrderBook.py:108-156), as extra trace records in program order:
   *   ABIDES_TRACE_ORDER_PLACED  agent = the placing agent, time = its currentTime (Order.time_placed):
   *                              p0 = order id in the message, p1 = limit price, p2 = signed quantity, p3 = 1 for a
   *                              market order (TradingAgent.placeLimitOrder / placeMarketOrder :295-365), 2 for the
   *                              new order of a modification (MarketReplayAgent.py:56-59)
   *   ABIDES_TRACE_EXCH_SENT | ORDER_ACCEPTED / _EXECUTED / _CANCELLED / _MODIFIED   agent = recipient, time = the
   *                              exchange's currentTime: p0 = order id, p1 = limit price, p2 = signed quantity,
   *                              p3 = fill price (-1 none) -- one per notification, in the order the book sends them
   *   ABIDES_TRACE_BOOK_TOP      after every handleLimitOrder: p0 / p1 = best bid and its volume, p2 / p3 = best ask and
   *                              its volume (price -1 = empty side): the BEST_BID / BEST_ASK rows and the snapshot marker
   *   ABIDES_TRACE_LAST_TRADE    after a handleLimitOrder that traded: p0 = traded quantity, p1 = average price */
  ABIDES_TRACE_EXCHANGE_LOG = 1
};
enum abides_trace_code {
  ABIDES_TRACE_FUNDAMENTAL = 64,          /* an oracle step: time = its timestamp, agent = -1, p0 = the new value */
  ABIDES_TRACE_ORDER_PLACED = 65, ABIDES_TRACE_BOOK_TOP = 66, ABIDES_TRACE_LAST_TRADE = 67,
  ABIDES_TRACE_EXCH_SENT = 128
};

enum abides_draw_flags {
  /* `theta` is not given (agent_theta holds the layout only): every ZI / HBL agent draws
   * sorted(round(random_state.normal(0, sqrt(sigma_pv), 2 * q_max)), reverse=True)   ZeroIntelligenceAgent.py:46-50 */
  ABIDES_DRAW_ZI_THETA = 1,
  /* agent_size of momentum agents is not given: size = random_state.randint(mm_min_size, mm_max_size)
   * (agent/examples/MomentumAgent.py:21; the type row's mm_min_size / mm_max_size carry min_size / max_size) */
  ABIDES_DRAW_MOMENTUM_SIZE = 2
};

/* Per-instance result (the record gathered at the end). */
typedef struct abides_instance_result {
  int64_t ttl_messages;                   /* Kernel.py:204 -- every queue pop incl. re-queues */
  int64_t delivered;                      /* pops that reached wakeup()/receiveMessage() */
  int64_t final_time;                     /* kernel currentTime at loop exit */
  int64_t slowest_agent_finish_time;      /* Kernel.py:309 */
  int64_t last_trade;                     /* OrderBook.last_trade */
  int64_t final_fundamental;              /* oracle r at exit */
  int64_t n_orders;                       /* order ids consumed */
  int64_t n_fills;                        /* executions (each produces two ORDER_EXECUTED) */
  int64_t traded_volume;                  /* sum of fill quantities */
  int32_t status;                         /* 0 ok; else enum abides_error raised on device */
  int32_t queue_peak;
  int32_t book_peak_levels, book_peak_orders;
  int64_t reserved[2];
} abides_instance_result;

/* Per-agent result (TradingAgent.kernelStopping, TradingAgent.py:122-147 and subclasses). */
typedef struct abides_agent_result {
  int64_t cash;                           /* holdings['CASH'] */
  int64_t holdings;                       /* holdings[symbol] (0 if absent) */
  int64_t marked_to_market;               /* markToMarket(holdings) */
  double  final_valuation;                /* FINAL_VALUATION summary-log value (class specific) */
} abides_agent_result;

/* Dispatch trace record (parity tests): one per delivered event. */
typedef struct abides_trace_rec {
  int64_t time;
  int32_t agent;                          /* recipient */
  int32_t code;                           /* 0 = WAKEUP, else enum abides_msg */
  int64_t p0, p1, p2, p3;                 /* message payload, see abides_msg: to the exchange p0 = sender and, for
                                             orders, id / limit price / signed quantity (CANCEL: side only), for
                                             QUERY_SPREAD the depth, for a subscription request the levels; to an agent
                                             order id / price / signed quantity / fill price for ORDER_*, best bid /
                                             size / best ask / size for QUERY_SPREAD and MARKET_DATA, the value for
                                             WHEN_MKT_*, QUERY_TRANSACTED_VOLUME and QUERY_LAST_TRADE */
} abides_trace_rec;

/* Message vocabulary.  Payload words for the trace:
 *   to exchange  : p0 = sender; orders: p1 = order id, p2 = price, p3 = qty (negative = sell;
 *                  CANCEL_ORDER: +-1, side only)
 *   ORDER_*      : p0 = order id, p1 = price (limit), p2 = qty (negative = sell), p3 = fill price
 *   QUERY_SPREAD reply : p0 = bid, p1 = bid volume, p2 = ask, p3 = ask volume (price -1 = none)
 *   WHEN_MKT_* reply   : p0 = time ; QUERY_TRANSACTED_VOLUME reply: p0 = volume */
enum abides_msg {
  ABIDES_MSG_WAKEUP = 0,
  ABIDES_MSG_WHEN_MKT_OPEN = 1, ABIDES_MSG_WHEN_MKT_CLOSE = 2,
  ABIDES_MSG_QUERY_LAST_TRADE = 3, ABIDES_MSG_QUERY_SPREAD = 4,
  ABIDES_MSG_QUERY_ORDER_STREAM = 5, ABIDES_MSG_QUERY_TRANSACTED_VOLUME = 6,
  ABIDES_MSG_LIMIT_ORDER = 7, ABIDES_MSG_MARKET_ORDER = 8,
  ABIDES_MSG_CANCEL_ORDER = 9, ABIDES_MSG_MODIFY_ORDER = 10,
  ABIDES_MSG_ORDER_ACCEPTED = 11, ABIDES_MSG_ORDER_EXECUTED = 12,
  ABIDES_MSG_ORDER_CANCELLED = 13, ABIDES_MSG_ORDER_MODIFIED = 14,
  ABIDES_MSG_MKT_CLOSED = 15,
  ABIDES_MSG_MARKET_DATA_SUBSCRIPTION_REQUEST = 16, ABIDES_MSG_MARKET_DATA_SUBSCRIPTION_CANCELLATION = 17,
  ABIDES_MSG_MARKET_DATA = 18,
  ABIDES_MSG_REPLY = 32          /* OR-ed in for exchange -> agent echoes of a query type */
};

/* ---- order-book replay harness (K2) -------------------------------------- */
typedef struct abides_book_op {           /* one message handed to util/OrderBook.py */
  int32_t kind;                           /* ABIDES_MSG_LIMIT_ORDER / MARKET_ORDER / CANCEL_ORDER / MODIFY_ORDER */
  int32_t agent;
  int64_t order_id;                       /* MARKET_ORDER children take ids from next_order_id */
  int64_t price;                          /* limit price */
  int64_t qty;                            /* >0 */
  int32_t is_buy;
  int32_t reserved;
  int64_t new_qty;                        /* MODIFY_ORDER: quantity of new_order */
  int64_t time;                           /* owner.currentTime */
} abides_book_op;

typedef struct abides_book_event {        /* outbound notification, in the order the book sends them */
  int32_t op_index;                       /* which abides_book_op caused it */
  int32_t kind;                           /* ORDER_ACCEPTED / ORDER_EXECUTED / ORDER_CANCELLED / ORDER_MODIFIED */
  int32_t agent;                          /* recipient */
  int32_t is_buy;
  int64_t order_id, price, qty, fill_price;
} abides_book_event;

typedef struct abides_book_state {        /* after every op */
  int64_t bid, bid_vol, ask, ask_vol;     /* getInsideBids(1) / getInsideAsks(1); price -1 = empty */
  int64_t last_trade;                     /* -1 = None */
  int32_t n_bid_levels, n_ask_levels;
} abides_book_state;

typedef struct abides_handle abides_handle;

/* lifecycle */
int  abides_create(int device, abides_handle** out);
void abides_destroy(abides_handle* h);
const char* abides_last_error(void);
const char* abides_version(void);

/* Two device engines run the same path.  The WARP engine gives every instance a warp: a step of a few
 * thousand instances takes seconds, it carries every agent class and market-data subscriptions.  The LANE engine gives
 * every instance a thread and regroups the popped events of a CTA by handler, so that a warp runs one handler for 32
 * instances: it needs tens of thousands of resident instances and then sustains 2-3x the warp engine's messages/s on the
 * sparse_zi populations (a step is the 10-20 s a single instance's event chain takes).  abides_load picks one (AUTO:
 * the lane engine when it carries the batch and the batch has at least ABIDES_LANE_AUTO_MIN_INSTANCES_ZI instances of a
 * ZI-only population, or ABIDES_LANE_AUTO_MIN_INSTANCES of any other; the environment variable ABIDES_ENGINE=warp|lane
 * overrides AUTO); abides_set_engine forces the choice for the following loads and abides_engine_info reports what the
 * loaded batch runs on: engine, and the kernel variant (warp engine: 0 full, 1 zi_legacy, 2 zi_cubic, 3 rmsc,
 * 4 rmsc01, 5 rmsc02, 6 exec; lane engine: 16 zi_legacy, 17 zi_cubic, 18 rmsc, 19 full). */
enum abides_engine { ABIDES_ENGINE_AUTO = 0, ABIDES_ENGINE_WARP = 1, ABIDES_ENGINE_LANE = 2 };
#define ABIDES_LANE_AUTO_MIN_INSTANCES_ZI 8192
#define ABIDES_LANE_AUTO_MIN_INSTANCES 24576
int  abides_set_engine(abides_handle* h, int32_t engine);
int  abides_engine_info(abides_handle* h, int32_t* engine, int32_t* variant);

/* Kernel.runner: load a batch (host -> device), run it, read results (device -> host). */
int  abides_load(abides_handle* h, const abides_batch* batch);
int  abides_reset(abides_handle* h);                        /* re-arm the resident batch for another run */
int  abides_run(abides_handle* h, int64_t until_time);      /* INT64_MAX: to completion incl. kernelStopping */
int  abides_read_results(abides_handle* h, abides_instance_result* inst_out /*[n_instances]*/,
                         abides_agent_result* agents_out /*[n_agents_total] or NULL*/);
int  abides_read_trace(abides_handle* h, int32_t instance, abides_trace_rec* out, int64_t capacity,
                       int64_t* n_out);
/* device-side timing of the last abides_run (CUDA events on the launching stream), ms */
int  abides_last_run_ms(abides_handle* h, float* ms, int32_t* n_launches);
int  abides_last_reset_ms(abides_handle* h, float* ms);
/* kernels this handle has launched since abides_create (loads, resets and runs) */
int  abides_launch_count(abides_handle* h, int64_t* n);
/* per-phase cycle / event counters summed over instances since the last load / reset (all zero unless the
 * library was built with -DABIDES_PROFILE; slots: see enum P_* in csrc/engine.cu) */
int  abides_read_profile(abides_handle* h, uint64_t* out, int32_t n);

/* util/OrderBook.py replay: n_books independent books, book b replays ops[op_offset[b] .. op_offset[b+1]).
 * events_out is [n_ops_total * max_events_per_op]; event_count_out [n_ops_total]; states_out [n_ops_total]. */
int  abides_replay_book(abides_handle* h, int32_t n_books, const int64_t* op_offset,
                        const abides_book_op* ops, int32_t stream_history, int32_t max_events_per_op,
                        abides_book_event* events_out, int32_t* event_count_out,
                        abides_book_state* states_out);

/* util/oracle/MeanRevertingOracle.py:49-81 -- the dense mean-reverting fundamental (one value per time step):
 *   r[0] = r_bar,  r[t] = max(0, kappa r_bar + (1 - kappa) r[t-1] + shock[t]),  out[s][t] = int(round(r[t])),
 * shock = normal(scale = sqrt(sigma_s), size = n_steps) drawn from series s's MT19937 stream (the reference draws from
 * the process-global np.random), whose state is given in and handed back advanced -- n_series independent series
 * (seeds x symbols) in one launch, one thread each.  Floats follow the north star's 1e-9 relative tolerance: the
 * normal deviates use the engine's table-driven log (include/abides_dd_math.h), correctly rounded except ~1e-4 of
 * arguments. */
int  abides_dense_fundamental(abides_handle* h, int32_t n_series, uint32_t* mt_key /*[n][624] in/out*/,
                              int32_t* mt_pos /*in/out*/, int32_t* mt_has_gauss /*in/out*/, double* mt_gauss /*in/out*/,
                              double r_bar, double kappa, double sigma_s, int64_t n_steps,
                              int64_t* out /*[n_series][n_steps]*/);

#ifdef __cplusplus
}
#endif
#endif /* ABIDES_B200_H */
