// engine.cu -- B200-native batched ABIDES engine (sm_100a).  One simulation instance per warp.
//
// What runs here: the kernel event loop (Kernel.py:190-271), sendMessage/setWakeup
// (:329-418), ExchangeAgent.receiveMessage (agent/ExchangeAgent.py:113-283), the limit order
// book (util/OrderBook.py:46-399), TradingAgent bookkeeping (agent/TradingAgent.py:152-524),
// the ZeroIntelligence / Value / Noise / Momentum / AdaptiveMarketMaker / POV-heartbeat
// strategies, the sparse mean-reverting oracle (util/oracle/SparseMeanRevertingOracle.py:88-229),
// the latency models (model/LatencyModel.py:113-145, Kernel.py:378-381) and numpy's legacy
// MT19937 transforms -- as a single persistent kernel in which every warp owns one instance:
//
//   * event queue  : three tiers ordered by two moving key limits.  "near" (shared memory, <= 32 events
//                    below near_limit kept SORTED, one per lane: pop = read the tail, insert = ballot
//                    rank + one warp-wide shift), list A (HBM, events below the horizon) and list B
//                    (HBM, the rest).  A and B are unsorted append-only lists with tombstones, scanned
//                    32 keys per coalesced load only when the tier above runs dry: B -> A pulls a few
//                    hundred events, A -> near a couple of dozen, so the thousands of long-dated
//                    wakeups in B are touched once per few thousand pops;
//   * order book   : per side the best <= BOOK_TOP orders price-time sorted in shared memory (matching,
//                    level sums, inserts and cancels are ballot/popc scans and warp-wide shifts) over
//                    an unsorted "pool" list in HBM in the same key/payload format as the event lists
//                    (key = price rank, arrival number | order id).  Pool slots never move (tombstones
//                    are reused), the slot is reported with ORDER_ACCEPTED and handed back with the
//                    cancel, so deep inserts and deep cancels are O(1); the pool -> top refill reuses
//                    the event queue's window selection;
//   * agent state  : one 128-byte record per agent in HBM, moved to/from shared memory with one
//                    coalesced transaction per dispatch (skipped when the same agent is dispatched
//                    again, e.g. busy re-queues); handler code is warp-uniform;
//   * RNG          : per-stream MT19937 state in HBM (reference-exact mode, warp-parallel twist)
//                    or counter-based Philox4x32-10 (throughput mode).
//
// The kernel is bound by INSTRUCTION FETCH (profiles/): one warp walks ~670 mostly cold
// instructions per event; the instruction cache holds 32 KB per TPC -- two SMs share it
// (scripts/microbench/icache_tpc.cu) -- the hot set of the engine is larger, and ncu shows the GPC-level
// instruction cache ~90 % busy in every version of this kernel (run time == its request count / 7.1e9 per
// second).  So the code is kept SMALL rather than "fast": events travel in registers, big bodies are
// __noinline__ and shared by all call sites, the four HBM lists (event lists A/B, two book pools) share one
// implementation, the batch descriptor lives in constant memory.  (Inlining hot helpers to save instructions
// has repeatedly made it slower.)
//
// No tensor cores: there is no dense contraction on this path (integer / latency bound).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <limits.h>
#include <math.h>
#include <mutex>
#include <string>
#include <vector>

#include "abides_b200.h"
#include "abides_dd_math.h"
#include "lane_host_common.h"
#include "lane_kernels.h"

#define FULL 0xffffffffu
#define UNLIKELY(x) __builtin_expect(!!(x), 0)
#define LIKELY(x) __builtin_expect(!!(x), 1)

// ----------------------------------------------------------------------------------------------
// the LANE engine (lane_kernels.cu / lane_device.inc): one instance per thread, used for large batches
// ----------------------------------------------------------------------------------------------
#define LN_INL __device__ __forceinline__
#define LN_BIG __device__ __noinline__
// load-time RNG work shared by both engines: seeding of streams described by their seed, constructor-time draws
namespace rngsetup {
#include "rng_setup_device.inc"
__global__ void rng_materialize_kernel(const __grid_constant__ RngSetup p) {
  rng_materialize_one(p, blockIdx.x, blockIdx.y * blockDim.x + threadIdx.x);
}
__global__ void rng_ctor_draws_kernel(const __grid_constant__ RngSetup p) {
  rng_ctor_draws_one(p, blockIdx.x, blockIdx.y * blockDim.x + threadIdx.x);
}
// Streams from which the config drew millions of words before the simulation starts (the sparse_zi configs draw an
// N x N latency matrix from the global stream and read one row of it): one WARP per stream, state in shared memory,
// every generation step of 624 words done by the 32 lanes together.
constexpr int SKIP_MIN_WORDS = 16 * ABIDES_MT_N;
constexpr int SKIP_WARPS = 4;
__device__ __forceinline__ void coop_twist(unsigned* mt, int lane) {
  const unsigned U = 0x80000000u, L = 0x7fffffffu, MAG = 0x9908b0dfu;
#pragma unroll 1
  for (int base = 0; base < 227; base += 32) {
    const int k = base + lane; const bool act = k < 227; unsigned v = 0;
    if (act) { const unsigned y = (mt[k] & U) | (mt[k + 1] & L); v = mt[k + 397] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u); }
    __syncwarp();
    if (act) mt[k] = v;
    __syncwarp();
  }
#pragma unroll 1
  for (int base = 227; base < 623; base += 32) {
    const int k = base + lane; const bool act = k < 623; unsigned v = 0;
    if (act) { const unsigned y = (mt[k] & U) | (mt[k + 1] & L); v = mt[k - 227] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u); }
    __syncwarp();
    if (act) mt[k] = v;
    __syncwarp();
  }
  if (lane == 0) { const unsigned y = (mt[623] & U) | (mt[0] & L); mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u); }
  __syncwarp();
}
__global__ void __launch_bounds__(32 * SKIP_WARPS) rng_skip_kernel(const __grid_constant__ RngSetup p, const int* list, int n) {
  __shared__ unsigned s_mt[SKIP_WARPS][ABIDES_MT_N];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * SKIP_WARPS + w;
  if (i >= n) return;
  const long long s = list[i];
  const long long row = p.dev_row ? (long long)p.dev_row[s] : s;
  if (row < 0) return;
  unsigned* mt = s_mt[w];
  if (lane == 0) {
    unsigned x = p.seed[s];
    mt[0] = x;
    for (int j = 1; j < ABIDES_MT_N; j++) { x = 1812433253u * (x ^ (x >> 30)) + (unsigned)j; mt[j] = x; }
  }
  __syncwarp();
  const long long k = p.pos[s];
  const long long twists = (k + ABIDES_MT_N - 1) / ABIDES_MT_N;
#pragma unroll 1
  for (long long t = 0; t < twists; t++) coop_twist(mt, lane);
  unsigned* out = p.mt + (size_t)row * ABIDES_MT_N;
  for (int j = lane; j < ABIDES_MT_N; j += 32) out[j] = mt[j];
  __syncwarp();
  if (lane == 0) p.pos[s] = (int)(k - (twists - 1) * ABIDES_MT_N);
}
// util/oracle/MeanRevertingOracle.py:49-81 -- the DENSE mean-reverting fundamental: one value per time step,
//   r[0] = r_bar;  r[t] = max(0, kappa r_bar + (1 - kappa) r[t-1] + shock[t]),  shock = normal(scale = sqrt(sigma_s), size = T)
// drawn from the stream whose state is handed in (the reference uses the process-global np.random).  The recurrence is
// sequential in t, so the kernel vectorises ACROSS series: one thread per (seed, symbol) series, states in HBM.
__global__ void dense_fundamental_kernel(int n_series, unsigned* mt, int* pos, int* has_gauss, double* gauss, double r_bar,
                                         double kappa, double sigma_s, long long n_steps, long long* out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_series) return;
  SeqRng r;
  r.mt = mt + (size_t)s * ABIDES_MT_N; r.pos = pos[s]; r.has_gauss = has_gauss[s]; r.gauss = gauss[s];
  const double scale = sqrt(sigma_s);
  long long* o = out + (size_t)s * n_steps;
  double prev = r_bar;
  for (long long t = 0; t < n_steps; t++) {
    const double shock = 0.0 + scale * seq_gauss(r);         // np.random.normal(loc=0.0, scale): loc + scale * gauss
    if (t > 0) {
      double v = (kappa * r_bar) + ((1 - kappa) * prev);
      v = v + shock;
      prev = v < 0 ? 0.0 : v;                                // max(0, v)
    }
    o[t] = (long long)rint(prev);                            // np.round(r).astype(int)
  }
  pos[s] = r.pos; has_gauss[s] = r.has_gauss; gauss[s] = r.gauss;
}
}  // namespace rngsetup

namespace {

extern __shared__ __align__(128) unsigned char smem_raw[];

constexpr int NEAR_CAP = 32;            // events in the shared-memory tier (one per lane, sorted)
constexpr int FILL_MAX = 24;            // A -> near pulls at most this many ...
constexpr int FILL_MIN = 8;             // ... and widens its window when it gets fewer
constexpr int A_CAP = 4096;             // list A capacity per instance
constexpr int A_FILL_MAX = 384;         // B -> A pulls
constexpr int A_FILL_MIN = 96;
constexpr int A_REBALANCE = 2048;       // live entries in A above which the horizon is pulled in
constexpr int BOOK_TOP = 64;            // orders per side kept (sorted) in shared memory
constexpr int TOP_FILL = 32;            // pool -> top pulls at most this many (one per lane) ...
constexpr int TOP_FILL_MIN = 8;         // ... and widens its price window when it gets fewer
enum { L_A = 0, L_B = 1, L_POOL = 2, N_LISTS = 4 };   // HBM lists: event lists A, B; book pools L_POOL + side
constexpr int MOM_RING = 64;            // >= 51 prefix sums
constexpr int TXN_CAP = 1 << 16;        // ring of transaction rows since the last volume query (per instance)
constexpr int UNR_CAP = 1 << 16;        // ring of de-duplicated rows the volume queries sum over
constexpr int TRACE_FUNDAMENTAL = 64;   // trace code of an oracle step (time = ts, p0 = value)
constexpr int SID_AGENT = 4;            // RNG stream id of the agent being dispatched (0..3: ABIDES_STREAM_*)
constexpr int N_PROF = 16;
constexpr int OLOG_CAP = 1 << 18;       // ring of order-log rows (HBL populations only; a bucket = all orders between two trades)
constexpr int HBL_RANGE_CAP = 16384;    // widest price range (ticks) an HBL belief table may span

struct __align__(16) EvKey { long long t; unsigned long long k2; };   // k2 = rcpt<<33 | wakeup<<32 | uniq
struct __align__(16) EvPay { int w[8]; };
struct __align__(16) BookOrd { int price, qty, id, agent; };
struct __align__(16) OpenOrd { int key, id, price, sqty; };            // sqty > 0 buy, < 0 sell
struct __align__(16) TxnRow { long long t; int q; int seq; };          // OrderBook.history transaction + bucket of its record
struct __align__(16) OLogRow { int bucket, price, flags, pad; };        // OrderBook.history[bucket][order]: side (bit 0), has transactions (bit 1)
struct __align__(16) RngState { int pos; unsigned flags; double gauss; };   // flags bit 0: cached gaussian valid

enum : unsigned {
  F_HAS_OPEN = 1u, F_HAS_CLOSE = 2u, F_MKT_CLOSED = 4u, F_HAS_LAST_TRADE = 8u, F_HAS_DAILY_CLOSE = 16u,
  F_KNOWN_BOOK = 32u, F_TRADING = 64u, F_HAS_PREV_WAKE = 128u, F_SUB_REQUESTED = 256u,
  F_MM_AWAIT_SPREAD = 512u, F_MM_AWAIT_TV = 1024u, F_HAS_LAST_MID = 2048u, F_HAS_AVG20 = 4096u,
  F_HAS_AVG50 = 8192u, F_HAS_TV = 16384u, F_STATE_SHIFT = 16u, F_STATE_MASK = 7u << 16
};
enum { ST_AWAITING_WAKEUP = 0, ST_INACTIVE = 1, ST_AWAITING_SPREAD = 2, ST_ACTIVE = 3, ST_AWAITING_STREAM = 4,
       ST_AWAITING_MARKET_DATA = 5 };
enum { BID = 0, ASK = 1 };              // book side index; a BUY order rests on BID and matches ASK

struct __align__(128) AgentRec {        // 128 bytes: one coalesced transaction
  long long busy;                       // Kernel.agentCurrentTimes[a]
  long long cash;                       // holdings['CASH']
  long long cur_time;                   // Agent.currentTime
  long long holdings;                   // holdings[symbol]
  int last_trade;
  int bid, bid_vol, ask, ask_vol;       // known_bids[0] / known_asks[0], price -1 = empty
  unsigned flags;
  int ord_off;                          // open orders: slice of the instance's arena
  unsigned short ord_cnt, ord_cap;
  RngState rng;                         // MT19937 position / Philox word counter, numpy's cached gaussian
  int size;                             // Noise / Value / Momentum order size
  int type_idx;                         // absolute row in the types table
  union {
    struct { double r_t, sigma_t; long long prev_wake; int hint_id, hint;      // ZI / Value / HBL (+ pool slot of the last
             int stream_n, stream_B; } est;                                    //  accepted order; HBL: order-stream window)
    struct { long long wakeup_time; } noise;
    struct { double csum; int n_mids; int ring; double avg20, avg50; } mom;
    struct { double last_spread, window_size; long long transacted_volume;
             int buy_size, sell_size, tick_size, last_mid; } mm;
    struct { long long executed, transacted_volume; } pov;                         // POVExecutionAgent
    struct { long long executed; double arrival; } sx;                             // TWAP / VWAP ExecutionAgent
  } u;
};
static_assert(sizeof(AgentRec) == 128, "AgentRec must be 128 bytes");

// market-data subscriptions (ExchangeAgent.subscription_dict :286-297, in insertion order) and the bodies of the
// MARKET_DATA messages that carry more than one level (a ring: the message holds the sequence number)
constexpr int MAX_SUBS = 256;
constexpr int MD_RING = 4096;
struct SubRec { int agent; int levels; double freq; long long last; };
struct MdSnap { unsigned seq; int nb, na, pad; int bp[ABIDES_MD_MAX_LEVELS]; int ap[ABIDES_MD_MAX_LEVELS]; };

struct Hot {
  long long now, ttl, delivered, n_fills, volume;
  long long exch_busy, exch_cd, exch_cur_time, additional_delay;
  long long or_pt; double or_pv; long long ms_time; double ms_value;
  long long next_order_id;
  long long lim_t; unsigned long long lim_k;                    // near_limit
  long long hz_t; unsigned long long hz_k;                      // horizon
  long long span[N_LISTS];                                      // adaptive windows: A -> near, B -> A, pool -> top (x2)
  long long sel_t; unsigned long long sel_k;                    // result of choose_limit
  unsigned uniq;
  int n_near;
  int l_total[N_LISTS], l_live[N_LISTS];   // per HBM list: slots used, live entries
  int n_top[2];                         // book: entries per side in shared memory
  unsigned near_free;                   // free payload slots of the near tier
  int n_pfree[2];                       // book pools: free (tombstoned) slots available for reuse
  int last_hint;                        // pool slot + 1 of the order book_enter just placed (0: shared part)
  long long olog_len;                   // order-log rows written (HBL populations)
  unsigned book_arr;                    // arrival counter of resting orders (time priority)
  long long pb_t[2]; unsigned long long pb_k[2];   // per side: lower bound of every pool key
  int last_trade, has_last_trade;
  int arena_top, status, n_trace, queue_peak, book_peak_orders, done;
  int find_idx;                         // result of side_find
  int hist_B, tv_prev_len;              // open history bucket; OrderBook._transacted_volume previous length
  long long log_len, log_mark, unr_len; // transaction rows logged / scanned by the last query / unrolled rows
  long long book_update_ts;             // OrderBook.last_update_ts (LLONG_MIN: None)
  int n_subs; unsigned md_seq;          // subscribers registered; MARKET_DATA bodies written to the ring
  RngState rs[ABIDES_EXTRA_STREAMS];
};

struct __align__(128) InstShared {      // persisted per instance between launches (abides_run(until))
  long long nt[NEAR_CAP];               // sorted DESCENDING by (nt, nk): entry n_near-1 is the next event
  unsigned long long nk[NEAR_CAP];
  int nslot[NEAR_CAP];                  // payload slot of each entry
  EvPay npay[NEAR_CAP];
  BookOrd top[2][BOOK_TOP];             // best first
  int tseq[2][BOOK_TOP];                // history bucket in which each resting order was entered
  int tarr[2][BOOK_TOP];                // arrival number (time priority across top and pool)
  int tlog[2][BOOK_TOP];                // order-log row of each resting order (-1: no log)
  AgentRec ag;
  Hot hot;
};

struct DevBatch {
  int n_instances, rng_mode, trace_cap, b_cap, arena_cap, n_types, deep_cap;
  int trace_flags;                      // enum abides_trace_flags
  const abides_instance_cfg* cfg;
  const abides_agent_type* types;
  abm_dd* type_log1mk;                  // log(1 - kappa) per type, double-double
  double* type_omp2;                    // 1 - (1 - kappa) ** 2 per type (ZeroIntelligenceAgent.py:214)
  const int* agent_type;
  const double* lat_to;
  const double* lat_from;
  const long long* size;
  const long long* wakeup_time;
  const int* agent_theta;
  const int* theta;
  unsigned* mt;
  const int* mt_pos;
  const int* mt_has_gauss;
  const double* mt_gauss;
  AgentRec* agents;
  InstShared* state;
  EvKey* a_keys; EvPay* a_pay;          // [n_instances][A_CAP]
  EvKey* b_keys; EvPay* b_pay;          // [n_instances][b_cap]
  EvKey* p_keys; EvPay* p_pay;          // [n_instances][2][deep_cap] book pools
  int* p_free;                          // [n_instances][2][deep_cap] free-slot stacks of the pools
  unsigned* task_ctr; int* slice_done;  // dynamic (slice, instance) task hand-out: next task, slices stored per instance
  SubRec* subs;                         // [n_instances][MAX_SUBS], or NULL (no subscribing agent in the batch)
  MdSnap* md;                           // [n_instances][MD_RING], or NULL
  OLogRow* olog;                        // [n_instances][OLOG_CAP] order log, or NULL (no HBL agent in the batch)
  int* hbl_cnt;                         // [n_instances][4][HBL_RANGE_CAP] belief-table scratch, or NULL
  OpenOrd* arena;
  TxnRow* txn_log;                      // [n_instances][TXN_CAP]
  TxnRow* unrolled;                     // [n_instances][UNR_CAP]
  double* mom_ring;
  const int* mom_index;                 // per agent row: ring index or -1
  abides_trace_rec* trace;
  abides_instance_result* res;
  abides_agent_result* ares;
  unsigned long long* prof;             // [N_PROF] cycle / count accumulators (profiling builds)
};

__constant__ DevBatch c_db;             // uploaded in-stream before every launch

struct Run {                            // per-launch context, shared memory, not persisted
  abides_instance_cfg cfg;
  EvKey* lk[N_LISTS]; EvPay* lp[N_LISTS]; int lcap[N_LISTS];
  OpenOrd* arena; TxnRow* txn_log; TxnRow* unrolled;
  int* pfree[2];
  OLogRow* olog; int* hbl_cnt;
  SubRec* subs; MdSnap* md;
  abides_trace_rec* trace;
  unsigned* mt0;                        // MT state of this instance's stream 0 (agent 0)
  const double* lat_to; const double* lat_from;   // this instance's rows
  AgentRec* agents;
  long long aoff;                       // agent_offset
  int inst, n_agents;
  int cur_agent;                        // agent whose record sits in st.ag (-1 none)
  int rp_op, rp_max, rp_base;           // order-book replay harness: current op, events per op, first op
  abides_book_event* rp_events; int* rp_counts;
  unsigned long long prof[N_PROF];
};
struct __align__(128) Smem { InstShared st; Run run; };
// ABM_CTA_WARPS instances share one CTA (warp y = instance blockIdx.x * ABM_CTA_WARPS + y) and meet at a barrier before
// every event, so that the warps of an SM walk the same code at about the same time and share instruction-cache lines
#ifndef ABM_CTA_WARPS
#define ABM_CTA_WARPS 1
#endif
// ABM_TIME_SLICES > 1: instances are run as that many time slices handed to the warps dynamically (engine_device.inc)
#ifndef ABM_TIME_SLICES
#define ABM_TIME_SLICES 1
#endif
static_assert(ABM_TIME_SLICES == 1 || ABM_CTA_WARPS == 1, "time slices need one instance per CTA");
#if ABM_CTA_WARPS > 1
#define SM (reinterpret_cast<Smem*>(smem_raw + threadIdx.y * (unsigned)sizeof(Smem)))
#else
#define SM (reinterpret_cast<Smem*>(smem_raw))
#endif
#define ST (SM->st)
#define HOT (SM->st.hot)
#define RUN (SM->run)
#define LANE ((int)threadIdx.x)

#ifdef ABIDES_PROFILE
#define PROF_T0() long long prof_t0_ = clock64()
#define PROF_ADD(slot) do { long long c_ = clock64(); RUN.prof[slot] += (unsigned long long)(c_ - prof_t0_); prof_t0_ = c_; } while (0)
#define PROF_CNT(slot) do { RUN.prof[slot] += 1ull; } while (0)
#else
#define PROF_T0() do {} while (0)
#define PROF_ADD(slot) do {} while (0)
#define PROF_CNT(slot) do {} while (0)
#endif
enum { P_POP = 0, P_REFILL, P_REC, P_EXCH_ORDER, P_EXCH_CANCEL, P_EXCH_QUERY, P_EXCH_OTHER, P_AG_WAKE, P_AG_SPREAD,
       P_AG_OTHER, P_REQUEUE, P_N_REFILL, P_N_BPULL, P_N_REQUEUE, P_N_RECLOAD, P_TOTAL };

// ----------------------------------------------------------------------------------------------
// warp helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ long long warp_min_ll(long long v) {
  unsigned hi = (unsigned)((unsigned long long)v >> 32) ^ 0x80000000u;     // order-preserving for signed v
  unsigned mh = __reduce_min_sync(FULL, hi);
  unsigned lo = hi == mh ? (unsigned)v : 0xffffffffu;
  unsigned ml = __reduce_min_sync(FULL, lo);
  return (long long)(((unsigned long long)(mh ^ 0x80000000u) << 32) | ml);
}
__device__ __forceinline__ unsigned long long warp_min_ull(unsigned long long v) {
  unsigned hi = (unsigned)(v >> 32);
  unsigned mh = __reduce_min_sync(FULL, hi);
  unsigned lo = hi == mh ? (unsigned)v : 0xffffffffu;
  unsigned ml = __reduce_min_sync(FULL, lo);
  return ((unsigned long long)mh << 32) | ml;
}
__device__ __forceinline__ unsigned long long warp_max_ull(unsigned long long v) {
  unsigned hi = (unsigned)(v >> 32);
  unsigned mh = __reduce_max_sync(FULL, hi);
  unsigned lo = hi == mh ? (unsigned)v : 0u;
  unsigned ml = __reduce_max_sync(FULL, lo);
  return ((unsigned long long)mh << 32) | ml;
}
__device__ __forceinline__ int warp_sum_i(int v) { return __reduce_add_sync(FULL, v); }
__device__ __forceinline__ bool key_less(long long at, unsigned long long ak, long long bt, unsigned long long bk) {
  return at < bt || (at == bt && ak < bk);
}
// IEEE double division, one shared copy (the inline expansion is ~25 instructions per site; instruction bytes are
// what this kernel pays for -- see the header comment)
__device__ __noinline__ double fdiv(double a, double b) { return a / b; }

// ----------------------------------------------------------------------------------------------
// kernels
// ----------------------------------------------------------------------------------------------

// one warp per instance: build the initial device state (Kernel.runner prologue, Kernel.py:97-172)
__global__ void __launch_bounds__(32) init_kernel() {
  const DevBatch& db = c_db;
  int inst = blockIdx.x, lane = threadIdx.x;
  if (inst >= db.n_instances) return;
  const abides_instance_cfg* cfg = &db.cfg[inst];
  InstShared* G = &db.state[inst];
  int n = cfg->n_agents;
  long long aoff = cfg->agent_offset;
  // zero the persistent block
  for (size_t i = lane; i < sizeof(InstShared) / 4; i += 32) reinterpret_cast<unsigned*>(G)[i] = 0;
  __syncwarp();
  EvKey* fk = db.b_keys + (size_t)inst * db.b_cap;
  EvPay* fp = db.b_pay + (size_t)inst * db.b_cap;
  for (int a = lane; a < n; a += 32) {
    long long row = aoff + a;
    AgentRec A;
    memset(&A, 0, sizeof A);
    const abides_agent_type& ty = db.types[cfg->type_base + db.agent_type[row]];
    A.type_idx = cfg->type_base + db.agent_type[row];
    A.busy = cfg->start_time;
    A.cash = ty.starting_cash;
    A.bid = -1; A.ask = -1;
    A.size = (int)db.size[row];
    long long stream = aoff + (long long)ABIDES_EXTRA_STREAMS * inst + a;
    if (db.rng_mode == ABIDES_RNG_MT19937) {
      A.rng.pos = db.mt_pos[stream];
      A.rng.gauss = db.mt_gauss[stream];
      A.rng.flags = db.mt_has_gauss[stream] ? 1u : 0u;
    }
    switch (ty.klass) {
      case ABIDES_ZI: case ABIDES_VALUE: case ABIDES_HBL: A.u.est.r_t = ty.r_bar; A.u.est.sigma_t = 0; break;
      case ABIDES_NOISE: A.u.noise.wakeup_time = db.wakeup_time[row]; break;
      case ABIDES_SCHEDULE_EXEC: A.u.sx.executed = 0; A.u.sx.arrival = nan(""); break;
      case ABIDES_MOMENTUM: A.u.mom.ring = db.mom_index[row]; break;
      case ABIDES_ADAPTIVE_MM:
        A.flags |= F_MM_AWAIT_SPREAD | F_MM_AWAIT_TV;
        A.u.mm.buy_size = (int)ty.min_order_size; A.u.mm.sell_size = (int)ty.min_order_size;
        A.u.mm.last_spread = 50;
        A.u.mm.window_size = ty.window_size >= 0 ? (double)ty.window_size : 0;
        A.u.mm.tick_size = ty.window_size >= 0 ? (int)ceil(50 * ty.level_spacing) : 0;
        break;
      default: break;
    }
    db.agents[row] = A;
    // Agent.kernelStarting: setWakeup(startTime) (agent/Agent.py:77) -- all start in list B
    EvKey k; k.t = cfg->start_time; k.k2 = ((unsigned long long)(unsigned)a << 33) | (1ull << 32);
    fk[a] = k;
    EvPay p;
    for (int i = 0; i < 8; i++) p.w[i] = 0;
    fp[a] = p;
  }
  __syncwarp();
  if (lane == 0) {
    Hot& h = G->hot;
    h.now = cfg->start_time;
    h.exch_busy = cfg->start_time;
    h.exch_cd = cfg->default_computation_delay;
    h.exch_cur_time = cfg->start_time;
    h.or_pt = cfg->mkt_open; h.or_pv = cfg->r_bar;
    h.ms_time = cfg->megashock_time0; h.ms_value = cfg->megashock_value0;
    h.lim_t = LLONG_MIN; h.lim_k = 0;
    h.hz_t = LLONG_MIN; h.hz_k = 0;
    h.span[L_A] = 1000000000ll; h.span[L_B] = 64000000000ll;
    h.span[L_POOL + BID] = 64; h.span[L_POOL + ASK] = 64;
    h.near_free = 0xffffffffu;
    h.l_total[L_B] = n; h.l_live[L_B] = n; h.queue_peak = n;
    h.last_trade = (int)cfg->r_bar; h.has_last_trade = 1;
    h.book_update_ts = LLONG_MIN;
    for (int s = 0; s < ABIDES_EXTRA_STREAMS; s++) {
      long long stream = aoff + (long long)ABIDES_EXTRA_STREAMS * inst + n + s;
      if (db.rng_mode == ABIDES_RNG_MT19937) {
        h.rs[s].pos = db.mt_pos[stream]; h.rs[s].gauss = db.mt_gauss[stream];
        h.rs[s].flags = db.mt_has_gauss[stream] ? 1u : 0u;
      }
    }
  }
}

__global__ void init_types_kernel() {
  const DevBatch& db = c_db;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= db.n_types) return;
  double base = 1 - db.types[i].kappa;
  abm_dd l = (base > 0 && base != 1.0) ? abm_log_dd(base) : abm_mk(0.0, 0.0);
  db.type_log1mk[i] = l;
  db.type_omp2[i] = 1 - abm_pow_with_log(l, base, 2.0);
}

// ----------------------------------------------------------------------------------------------
// kernel variants (engine_device.inc): same code, different feature sets; abides_load picks one per batch
// ----------------------------------------------------------------------------------------------
#define CLS(k) (1 << (k))
namespace v_full {                      // everything (tracing, Philox, every agent class, book replay)
#define ABM_V_TRACE 1
#define ABM_V_PHILOX 1
#define ABM_V_CLASSES 0x7ff
#define ABM_V_LAT (-1)
#define ABM_V_OLOG 1
#define ABM_V_TVLOG 1
#define ABM_V_BOOKEXT 1
#define ABM_V_FULL 1
#define ABM_V_MD 1
#include "engine_device.inc"
}
namespace v_zi_legacy {                 // sparse_zi_1000: ZI agents, agentLatency matrix + noise, MT19937, no trace
#define ABM_V_TRACE 0
#define ABM_V_PHILOX 0
#define ABM_V_CLASSES CLS(ABIDES_ZI)
#define ABM_V_LAT ABIDES_LAT_LEGACY_NOISE
#define ABM_V_OLOG 0
#define ABM_V_TVLOG 0
#define ABM_V_BOOKEXT 0
#define ABM_V_FULL 0
#define ABM_V_MD 0
#include "engine_device.inc"
}
namespace v_zi_cubic {                  // sparse_zi_100: ZI agents, cubic latency model
#define ABM_V_TRACE 0
#define ABM_V_PHILOX 0
#define ABM_V_CLASSES CLS(ABIDES_ZI)
#define ABM_V_LAT ABIDES_LAT_CUBIC
#define ABM_V_OLOG 0
#define ABM_V_TVLOG 0
#define ABM_V_BOOKEXT 0
#define ABM_V_FULL 0
#define ABM_V_MD 0
#include "engine_device.inc"
}
namespace v_rmsc {                      // rmsc03 family: noise / value / momentum / adaptive MM / POV, deterministic latency
#define ABM_V_TRACE 0
#define ABM_V_PHILOX 0
#define ABM_V_CLASSES (CLS(ABIDES_NOISE) | CLS(ABIDES_VALUE) | CLS(ABIDES_MOMENTUM) | CLS(ABIDES_ADAPTIVE_MM) | CLS(ABIDES_POV_EXEC))
#define ABM_V_LAT ABIDES_LAT_DETERMINISTIC
#define ABM_V_OLOG 0
#define ABM_V_TVLOG 1
#define ABM_V_BOOKEXT 1
#define ABM_V_FULL 0
#define ABM_V_MD 0
#include "engine_device.inc"
}
namespace v_rmsc01 {                    // literal rmsc01: market maker (polling) / ZI / HBL / momentum, deterministic latency
#define ABM_V_TRACE 0
#define ABM_V_PHILOX 0
#define ABM_V_CLASSES (CLS(ABIDES_ZI) | CLS(ABIDES_HBL) | CLS(ABIDES_MOMENTUM) | CLS(ABIDES_MARKET_MAKER))
#define ABM_V_LAT ABIDES_LAT_DETERMINISTIC
#define ABM_V_OLOG 1
#define ABM_V_TVLOG 0
#define ABM_V_BOOKEXT 0
#define ABM_V_FULL 0
#define ABM_V_MD 0
#include "engine_device.inc"
}
namespace v_rmsc02 {                    // rmsc02: the rmsc01 classes with market-data subscriptions
#define ABM_V_TRACE 0
#define ABM_V_PHILOX 0
#define ABM_V_CLASSES (CLS(ABIDES_ZI) | CLS(ABIDES_HBL) | CLS(ABIDES_MOMENTUM) | CLS(ABIDES_MARKET_MAKER) | CLS(ABIDES_SUBSCRIPTION))
#define ABM_V_LAT ABIDES_LAT_DETERMINISTIC
#define ABM_V_OLOG 1
#define ABM_V_TVLOG 0
#define ABM_V_BOOKEXT 0
#define ABM_V_FULL 0
#define ABM_V_MD 1
#include "engine_device.inc"
}
namespace v_exec {                      // config/execution.py: the rmsc03 classes plus a TWAP / VWAP execution agent
#define ABM_V_TRACE 0
#define ABM_V_PHILOX 0
#define ABM_V_CLASSES (CLS(ABIDES_NOISE) | CLS(ABIDES_VALUE) | CLS(ABIDES_MOMENTUM) | CLS(ABIDES_ADAPTIVE_MM) | CLS(ABIDES_POV_EXEC) | CLS(ABIDES_SCHEDULE_EXEC))
#define ABM_V_LAT ABIDES_LAT_DETERMINISTIC
#define ABM_V_OLOG 0
#define ABM_V_TVLOG 1
#define ABM_V_BOOKEXT 1
#define ABM_V_FULL 0
#define ABM_V_MD 0
#include "engine_device.inc"
}
enum { VAR_FULL = 0, VAR_ZI_LEGACY, VAR_ZI_CUBIC, VAR_RMSC, VAR_RMSC01, VAR_RMSC02, VAR_EXEC, N_VARIANTS };

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
thread_local std::string g_err;
// The warp engine's kernels read their batch descriptor from ONE __constant__ symbol per device (c_db), uploaded
// in-stream before each launch: two handles on the same device must not overlap a warp-engine load / reset / run.
// (The lane engine passes its descriptor by value as a __grid_constant__ parameter and needs no lock.)
std::mutex g_cdb_mutex[64];
struct CdbLock {
  std::mutex* m;
  CdbLock(const abides_handle* h, bool warp);
  ~CdbLock() { if (m) m->unlock(); }
};
int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail(ABIDES_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));              \
  } while (0)

}  // namespace

namespace {
struct ShapeSig {
  int ni = 0, n_types = 0, rng_mode = -1, trace_cap = -1, max_agents = 0, n_mom = 0;
  bool has_hbl = false, needs_tv = false, has_md = false;
  int lane_variant = lane::LV_NONE;     // LV_NONE: the warp-per-instance engine
  long long na = -1, n_theta = -1, n_dev_rows = -1, n_mt_rows = 0, n_skip = 0;
  bool seeded = false;
  bool operator==(const ShapeSig& o) const {
    return lane_variant == o.lane_variant && n_dev_rows == o.n_dev_rows && seeded == o.seeded && n_mt_rows == o.n_mt_rows && n_skip == o.n_skip && ni == o.ni && n_types == o.n_types && rng_mode == o.rng_mode && trace_cap == o.trace_cap &&
           max_agents == o.max_agents && n_mom == o.n_mom && na == o.na && n_theta == o.n_theta && has_hbl == o.has_hbl && needs_tv == o.needs_tv && has_md == o.has_md;
  }
};
}  // namespace

struct abides_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  DevBatch db{};
  std::vector<void*> allocs;
  bool loaded = false;
  int n_instances = 0;
  long long n_agents_total = 0;
  int variant = VAR_FULL;               // which kernel variant the loaded batch runs on
  int engine_pref = ABIDES_ENGINE_AUTO; // abides_set_engine
  int lane_variant = lane::LV_NONE;     // != LV_NONE: the loaded batch runs on the lane engine
  lane::LaneBatch lb{};
  float last_ms = 0.f, last_reset_ms = 0.f;
  int last_launches = 0;
  int launches_since_load = 0;
  long long launches_total = 0;          // kernel launches of earlier loads (abides_launch_count)
  unsigned* d_mt_seed = nullptr; int* d_key_row = nullptr; int* d_dev_row = nullptr;   // compact stream description (seeded batches)
  // seeded batches without uploaded rows keep no pristine copy of the states: every (re)initialisation seeds them again
  bool rematerialize = false;
  rngsetup::RngSetup rs{};
  int *pos0 = nullptr, *hg0 = nullptr; double* g0 = nullptr;    // the uploaded per-stream tables (rng setup rewrites the live ones)
  int* d_skip_list = nullptr; int n_skip = 0;
  int max_agents = 0;
  unsigned* mt_pristine = nullptr;
  size_t n_streams = 0, n_dev_rows = 0; // streams of the batch; MT19937 states resident on the device (seed-only streams have none)
  int n_mom = 0;
  ShapeSig sig;
};

namespace {
CdbLock::CdbLock(const abides_handle* h, bool warp) : m(nullptr) {
  if (warp && h->device >= 0 && h->device < 64) { m = &g_cdb_mutex[h->device]; m->lock(); }
}

template <typename T>
int dev_alloc(abides_handle* h, T** out, size_t count) {
  void* p = nullptr;
  size_t bytes = (count ? count : 1) * sizeof(T);
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return fail(ABIDES_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  h->allocs.push_back(p);
  *out = reinterpret_cast<T*>(p);
  return 0;
}
template <typename T>
int dev_upload(abides_handle* h, T** out, const T* src, size_t count) {
  int rc = dev_alloc(h, out, count);
  if (rc) return rc;
  if (count) CK(cudaMemcpyAsync(*out, src, count * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  return 0;
}
void free_all(abides_handle* h) {
  for (void* p : h->allocs) cudaFree(p);
  h->allocs.clear();
  h->loaded = false;
}

}  // namespace

namespace {
template <typename T>
int h2d(abides_handle* h, const T* dst, const T* src, size_t count) {
  if (count) CK(cudaMemcpyAsync(const_cast<T*>(dst), src, count * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  return 0;
}
// seed / fast-forward the streams the host described by their seeds, then the agents' constructor draws, into `mt`
int run_rng_setup(abides_handle* h, unsigned* mt, const unsigned* staging) {
  DevBatch& d = h->db;
  rngsetup::RngSetup p = h->rs;
  p.mt = mt; p.staging = staging;
  if (h->rematerialize) {
    const size_t ns = h->n_streams;
    CK(cudaMemcpyAsync(const_cast<int*>(d.mt_pos), h->pos0, ns * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(const_cast<int*>(d.mt_has_gauss), h->hg0, ns * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(const_cast<double*>(d.mt_gauss), h->g0, ns * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  }
  const dim3 grid(d.n_instances, (h->max_agents + ABIDES_EXTRA_STREAMS + 63) / 64);
  if (p.key_row) {
    rngsetup::rng_materialize_kernel<<<grid, 64, 0, h->stream>>>(p);
    if (h->n_skip > 0)
      rngsetup::rng_skip_kernel<<<(h->n_skip + rngsetup::SKIP_WARPS - 1) / rngsetup::SKIP_WARPS, 32 * rngsetup::SKIP_WARPS, 0, h->stream>>>(
          p, h->d_skip_list, h->n_skip);
    h->launches_since_load += 1 + (h->n_skip > 0);
  }
  if (p.draw_flags) { rngsetup::rng_ctor_draws_kernel<<<grid, 64, 0, h->stream>>>(p); h->launches_since_load += 1; }
  CK(cudaGetLastError());
  return 0;
}

int launch_init(abides_handle* h) {
  DevBatch& d = h->db;
  if (d.rng_mode == ABIDES_RNG_MT19937 && h->rematerialize) { int rc = run_rng_setup(h, d.mt, nullptr); if (rc) return rc; }
  else if (d.rng_mode == ABIDES_RNG_MT19937)
    CK(cudaMemcpyAsync(d.mt, h->mt_pristine, h->n_dev_rows * ABIDES_MT_N * sizeof(unsigned), cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemsetAsync(d.mom_ring, 0, (size_t)(h->n_mom > 0 ? h->n_mom : 1) * MOM_RING * sizeof(double), h->stream));
  CK(cudaMemsetAsync(d.res, 0, (size_t)d.n_instances * sizeof(abides_instance_result), h->stream));
  CK(cudaMemsetAsync(d.prof, 0, N_PROF * sizeof(unsigned long long), h->stream));
  if (h->lane_variant != lane::LV_NONE) {
    CK(lane::launch_init(h->lane_variant, h->lb, h->stream));
    CK(cudaGetLastError());
    h->launches_since_load += 2;
    return 0;
  }
  CK(cudaMemcpyToSymbolAsync(c_db, &d, sizeof(DevBatch), 0, cudaMemcpyHostToDevice, h->stream));
  init_types_kernel<<<(d.n_types + 63) / 64, 64, 0, h->stream>>>();
  init_kernel<<<d.n_instances, 32, 0, h->stream>>>();
  CK(cudaGetLastError());
  h->launches_since_load += 2;
  return 0;
}
}  // namespace

extern "C" {

const char* abides_last_error(void) { return g_err.c_str(); }
const char* abides_version(void) {
#ifdef ABIDES_PROFILE
  return "abides_b200 0.2 (sm_100a, profiling build)";
#else
  return "abides_b200 0.2 (sm_100a)";
#endif
}

int abides_create(int device, abides_handle** out) {
  if (!out) return fail(ABIDES_EINVAL, "abides_create: out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return fail(ABIDES_ECUDA, std::string("abides_create: no CUDA device (") + cudaGetErrorString(e) +
                                  "); this engine has no CPU fallback");
  if (device < 0 || device >= n) return fail(ABIDES_EINVAL, "abides_create: bad device index");
  CK(cudaSetDevice(device));
  abides_handle* h = new abides_handle();
  h->device = device;
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&h->ev0));
  CK(cudaEventCreate(&h->ev1));
  CK(cudaFuncSetAttribute(v_full::sim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ABM_CTA_WARPS * sizeof(Smem))));
  CK(cudaFuncSetAttribute(v_zi_legacy::sim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ABM_CTA_WARPS * sizeof(Smem))));
  CK(cudaFuncSetAttribute(v_zi_cubic::sim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ABM_CTA_WARPS * sizeof(Smem))));
  CK(cudaFuncSetAttribute(v_rmsc::sim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ABM_CTA_WARPS * sizeof(Smem))));
  CK(cudaFuncSetAttribute(v_rmsc01::sim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ABM_CTA_WARPS * sizeof(Smem))));
  CK(cudaFuncSetAttribute(v_rmsc02::sim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ABM_CTA_WARPS * sizeof(Smem))));
  CK(cudaFuncSetAttribute(v_exec::sim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ABM_CTA_WARPS * sizeof(Smem))));
  *out = h;
  return 0;
}

void abides_destroy(abides_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  free_all(h);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}


int abides_set_engine(abides_handle* h, int32_t engine) {
  if (!h) return fail(ABIDES_EINVAL, "NULL handle");
  if (engine != ABIDES_ENGINE_AUTO && engine != ABIDES_ENGINE_WARP && engine != ABIDES_ENGINE_LANE)
    return fail(ABIDES_EINVAL, "abides_set_engine: unknown engine");
  h->engine_pref = engine;
  return 0;
}

int abides_engine_info(abides_handle* h, int32_t* engine, int32_t* variant) {
  if (!h || !h->loaded) return fail(ABIDES_ESTATE, "abides_engine_info: no batch loaded");
  const bool ln = h->lane_variant != lane::LV_NONE;
  if (engine) *engine = ln ? ABIDES_ENGINE_LANE : ABIDES_ENGINE_WARP;
  if (variant) *variant = ln ? lane::LANE_VARIANT_BASE + h->lane_variant : h->variant;
  return 0;
}

int abides_load(abides_handle* h, const abides_batch* b) {
  if (!h || !b) return fail(ABIDES_EINVAL, "abides_load: NULL argument");
  if (b->n_instances <= 0 || !b->instances || !b->types || !b->agent_type || !b->lat_to_exch || !b->lat_from_exch ||
      !b->agent_size || !b->agent_wakeup_time || !b->agent_theta)
    return fail(ABIDES_EINVAL, "abides_load: missing table");
  const bool seeded = b->mt_key_row != nullptr && b->rng_mode == ABIDES_RNG_MT19937;
  if (b->rng_mode == ABIDES_RNG_MT19937 && (!b->mt_pos || !b->mt_has_gauss || !b->mt_gauss || (!seeded && !b->mt_key)))
    return fail(ABIDES_EINVAL, "abides_load: MT19937 mode needs mt_key/mt_pos/mt_has_gauss/mt_gauss");
  if (seeded && (b->rng_mode != ABIDES_RNG_MT19937 || !b->mt_seed || b->n_mt_rows < 0 || (b->n_mt_rows > 0 && !b->mt_key)))
    return fail(ABIDES_EINVAL, "abides_load: mt_key_row needs MT19937 mode, mt_seed and n_mt_rows rows of mt_key");
  if (b->draw_flags & ~(ABIDES_DRAW_ZI_THETA | ABIDES_DRAW_MOMENTUM_SIZE)) return fail(ABIDES_EINVAL, "abides_load: unknown draw_flags");
  if (b->trace_flags & ~ABIDES_TRACE_EXCHANGE_LOG) return fail(ABIDES_EINVAL, "abides_load: unknown trace_flags");
  if (b->draw_flags && (!b->mt_seed && b->mt_key_row)) return fail(ABIDES_EINVAL, "abides_load: draw_flags with mt_key_row need mt_seed");
  if (b->draw_flags && b->rng_mode == ABIDES_RNG_PHILOX && (!b->mt_key_row || !b->mt_seed || !b->mt_pos || !b->mt_has_gauss || !b->mt_gauss))
    return fail(ABIDES_EINVAL, "abides_load: in Philox mode the constructor draws need every agent stream described by its seed");
  const bool draw_theta = (b->draw_flags & ABIDES_DRAW_ZI_THETA) != 0;
  if (b->rng_mode != ABIDES_RNG_MT19937 && b->rng_mode != ABIDES_RNG_PHILOX) return fail(ABIDES_EINVAL, "abides_load: bad rng_mode");
  CK(cudaSetDevice(h->device));
  int ni = b->n_instances;
  long long na = b->n_agents_total;
  int max_agents = 0;
  long long off = 0;
  if (na <= 0 || b->n_types <= 0 || b->n_theta_total < 0) return fail(ABIDES_EINVAL, "abides_load: n_agents_total / n_types / n_theta_total out of range");
  std::vector<int> mom_index((size_t)na, -1);
  int n_mom = 0, n_hbl = 0;
  unsigned class_mask = 0, lat_mask = 0;
  long long n_sub = 0;
  for (int i = 0; i < ni; i++) {
    const abides_instance_cfg& ic = b->instances[i];
    if (ic.latency_kind < ABIDES_LAT_DETERMINISTIC || ic.latency_kind > ABIDES_LAT_LEGACY_NOISE)
      return fail(ABIDES_EINVAL, "abides_load: bad latency_kind");
    lat_mask |= 1u << ic.latency_kind;
    if (ic.latency_kind == ABIDES_LAT_LEGACY_NOISE && ic.n_latency_noise < 1) return fail(ABIDES_EINVAL, "abides_load: the legacy latency model needs n_latency_noise >= 1");
    if (ic.theta_offset < 0 || ic.theta_offset > b->n_theta_total) return fail(ABIDES_EINVAL, "abides_load: bad theta_offset");
    if (ic.agent_offset != off || ic.n_agents <= 0) return fail(ABIDES_EINVAL, "abides_load: agent_offset must be the prefix sum of n_agents");
    if (ic.type_base < 0 || ic.type_base >= b->n_types) return fail(ABIDES_EINVAL, "abides_load: bad type_base");
    if (ic.r_bar < 0 || ic.r_bar > 2.0e9) return fail(ABIDES_EINVAL, "abides_load: r_bar out of the int32 price range");
    if (ic.flags & ~(int64_t)ABIDES_INST_NO_ORACLE) return fail(ABIDES_EINVAL, "abides_load: unknown instance flags");
    if (off + ic.n_agents > na) return fail(ABIDES_EINVAL, "abides_load: n_agents_total mismatch");
    for (int a = 0; a < ic.n_agents; a++) {
      int t = ic.type_base + b->agent_type[off + a];
      if (t < 0 || t >= b->n_types) return fail(ABIDES_EINVAL, "abides_load: agent_type out of range");
      int k = b->types[t].klass;
      if ((a == 0) != (k == ABIDES_EXCHANGE)) return fail(ABIDES_EINVAL, "abides_load: agent 0 (and only agent 0) must be the exchange");
      if (k < 0 || k > ABIDES_MARKET_REPLAY) return fail(ABIDES_EINVAL, "abides_load: unknown agent class");
      if (k == ABIDES_MARKET_REPLAY) {                    // the tape: six ints per row in the agent's theta slice
        const long long rows = b->agent_size ? b->agent_size[off + a] : 0;
        if (rows < 1 || b->agent_theta[off + a] < 0 || !b->theta || draw_theta ||
            ic.theta_offset + b->agent_theta[off + a] + 6 * rows > b->n_theta_total)
          return fail(ABIDES_EINVAL, "abides_load: MarketReplayAgent without a tape (theta rows, agent_size = row count)");
      }
      if (k == ABIDES_HBL) { n_hbl++; if (b->types[t].hbl_L < 1) return fail(ABIDES_EINVAL, "abides_load: HBL agent with L < 1"); }
      if (a > 0) class_mask |= 1u << k;
      if (b->types[t].subscribe) {
        if (k != ABIDES_MARKET_MAKER && k != ABIDES_MOMENTUM && k != ABIDES_SUBSCRIPTION)
          return fail(ABIDES_EINVAL, "abides_load: subscribe set on an agent class without a subscription mode");
        if (b->types[t].sub_levels < 1 || b->types[t].sub_levels > ABIDES_MD_MAX_LEVELS || !(b->types[t].subscribe_freq >= 0))
          return fail(ABIDES_EINVAL, "abides_load: bad subscription levels / freq");
        n_sub++;
      } else if (k == ABIDES_SUBSCRIPTION) return fail(ABIDES_EINVAL, "abides_load: SubscriptionAgent without subscribe");
      if (k == ABIDES_MOMENTUM) mom_index[off + a] = n_mom++;
      if ((k == ABIDES_ZI || k == ABIDES_HBL) && (b->agent_theta[off + a] < 0 || (!b->theta && !draw_theta))) return fail(ABIDES_EINVAL, "abides_load: ZI / HBL agent without theta");
      if ((k == ABIDES_ZI || k == ABIDES_HBL) && !draw_theta && b->agent_theta[off + a] >= 0 &&
          ic.theta_offset + b->agent_theta[off + a] + 2ll * b->types[t].q_max > b->n_theta_total)
        return fail(ABIDES_EINVAL, "abides_load: a ZI / HBL agent's theta rows run past n_theta_total");
      if ((k == ABIDES_ZI || k == ABIDES_HBL) && draw_theta && (!(b->types[t].sigma_pv >= 0) || b->types[t].q_max < 1 ||
          ic.theta_offset + b->agent_theta[off + a] + 2ll * b->types[t].q_max > b->n_theta_total))
        return fail(ABIDES_EINVAL, "abides_load: ABIDES_DRAW_ZI_THETA needs sigma_pv, q_max and room for 2 q_max values per agent");
      if (k == ABIDES_SCHEDULE_EXEC && b->types[t].exec_trade) {
        const abides_agent_type& ty = b->types[t];
        if (ty.wake_up_freq <= 0 || ty.exec_end_time < ty.exec_start_time + 2 * ty.wake_up_freq ||
            (ty.exec_end_time - ty.exec_start_time) % ty.wake_up_freq != 0)
          return fail(ABIDES_EINVAL, "abides_load: execution agent horizon is not a regular grid of >= 3 points");
        if (b->agent_theta[off + a] < 0 || !b->theta) return fail(ABIDES_EINVAL, "abides_load: execution agent without schedule");
      }
    }
    if (ic.n_agents > max_agents) max_agents = ic.n_agents;
    off += ic.n_agents;
  }
  if (off != na) return fail(ABIDES_EINVAL, "abides_load: n_agents_total mismatch");
  ShapeSig sig;
  sig.ni = ni; sig.n_types = b->n_types; sig.rng_mode = b->rng_mode; sig.trace_cap = b->trace_capacity;
  sig.max_agents = max_agents; sig.n_mom = n_mom; sig.na = na; sig.n_theta = b->n_theta_total; sig.has_hbl = n_hbl > 0;
  const bool needs_tv = (class_mask & ((1u << ABIDES_ADAPTIVE_MM) | (1u << ABIDES_POV_EXEC))) != 0;
  sig.needs_tv = needs_tv;
  sig.has_md = n_sub > 0;
  // which engine: the lane engine (one instance per thread) needs thousands of instances to fill the chip; below
  // that the warp-per-instance engine is faster.  abides_set_engine / ABIDES_ENGINE override the choice.
  int lane_variant = lane::lane_pick_variant(class_mask, lat_mask, b->rng_mode, n_sub);
  {
    int pref = h->engine_pref;
    if (pref == ABIDES_ENGINE_AUTO) {
      const char* e = getenv("ABIDES_ENGINE");
      if (e && !strcmp(e, "lane")) pref = ABIDES_ENGINE_LANE;
      else if (e && !strcmp(e, "warp")) pref = ABIDES_ENGINE_WARP;
    }
    if (pref == ABIDES_ENGINE_LANE && lane_variant == lane::LV_NONE)
      return fail(ABIDES_EINVAL, "abides_load: the lane engine was requested but does not carry this batch (HBL agents, "
                                 "market-data subscriptions and Philox streams run on the warp engine)");
    const bool zi_only = lane_variant == lane::LV_ZI_LEGACY || lane_variant == lane::LV_ZI_CUBIC;
    const int auto_min = zi_only ? ABIDES_LANE_AUTO_MIN_INSTANCES_ZI : ABIDES_LANE_AUTO_MIN_INSTANCES;
    bool to_warp = pref == ABIDES_ENGINE_WARP || (pref == ABIDES_ENGINE_AUTO && ni < auto_min);
    if (class_mask & (1u << ABIDES_MARKET_REPLAY)) {      // order-tape replay agents run on the lane engine only
      if (pref == ABIDES_ENGINE_WARP || lane_variant == lane::LV_NONE)
        return fail(ABIDES_EINVAL, "abides_load: MarketReplayAgent runs on the lane engine (not with HBL agents / subscriptions)");
      to_warp = false;
    }
    if (to_warp && pref == ABIDES_ENGINE_AUTO && lane_variant != lane::LV_NONE) {
      // AUTO never picks an engine whose tables cannot fit: the warp engine keeps a 2.5 KB MT19937 state per stream
      // (twice for batches uploaded with their states) and the larger per-instance capacities below -- ~20 MB for an
      // rmsc03 instance, i.e. ~8 000 of them in 180 GB -- where the lane engine holds 33 000
      const double streams = (double)na + (double)ABIDES_EXTRA_STREAMS * ni;
      double need = b->rng_mode == ABIDES_RNG_MT19937 ? streams * ABIDES_MT_N * 4.0 * (seeded && b->n_mt_rows == 0 ? 1.0 : 2.0) : 0.0;
      need += (double)na * (sizeof(AgentRec) + 64.0);
      need += (double)ni * ((double)sizeof(InstShared) + (double)A_CAP * (sizeof(EvKey) + sizeof(EvPay)) +
                            (4.0 * max_agents + 4096) * (sizeof(EvKey) + sizeof(EvPay)) +
                            (8.0 * max_agents + 4096) * sizeof(OpenOrd) +
                            2.0 * (4.0 * max_agents + 1024) * (sizeof(EvKey) + sizeof(EvPay) + 4.0) +
                            (needs_tv ? (double)(TXN_CAP + UNR_CAP) * sizeof(TxnRow) : 0.0) +
                            (double)(b->trace_capacity > 0 ? b->trace_capacity : 0) * sizeof(abides_trace_rec));
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && need > 0.92 * (double)total_b) to_warp = false;
    }
    if (to_warp) lane_variant = lane::LV_NONE;
  }
  sig.lane_variant = lane_variant;
  const bool use_lane = lane_variant != lane::LV_NONE;
  // MT19937 states on the device: one row per stream, except that the lane engine keeps NOISE agents' seeded streams
  // as their seed only (they draw one or two words a day, NoiseAgent.py:162)
  const size_t ns_all = (size_t)na + (size_t)ABIDES_EXTRA_STREAMS * ni;
  std::vector<int> dev_row;
  size_t n_dev_rows = ns_all;
  if (b->rng_mode == ABIDES_RNG_MT19937 && !seeded)
    for (size_t s2 = 0; s2 < ns_all; s2++)
      if (b->mt_pos[s2] < 0 || b->mt_pos[s2] > ABIDES_MT_N) return fail(ABIDES_EINVAL, "abides_load: mt_pos outside 0..624");
  if (seeded) {
    for (size_t s2 = 0; s2 < ns_all; s2++)
      if (b->mt_key_row[s2] >= 0 && (b->mt_pos[s2] < 0 || b->mt_pos[s2] > ABIDES_MT_N)) return fail(ABIDES_EINVAL, "abides_load: mt_pos outside 0..624");
    for (size_t s2 = 0; s2 < ns_all; s2++)
      if (b->mt_key_row[s2] >= b->n_mt_rows || b->mt_key_row[s2] < -1 || (b->mt_key_row[s2] < 0 && b->mt_pos[s2] < 0))
        return fail(ABIDES_EINVAL, "abides_load: bad mt_key_row / mt_pos");
    if (use_lane) {
      n_dev_rows = lane::lane_assign_rows(b, dev_row);
      if (n_dev_rows > 0x7fffffffull) return fail(ABIDES_EINVAL, "abides_load: too many streams");
    }
  }
  sig.seeded = seeded;
  sig.n_mt_rows = seeded ? b->n_mt_rows : 0;
  const bool rematerialize = seeded && b->n_mt_rows == 0;
  std::vector<int> skip_list;                   // streams fast-forwarded by a warp each
  if (seeded)
    for (size_t s2 = 0; s2 < ns_all; s2++)
      if (b->mt_key_row[s2] < 0 && b->mt_pos[s2] >= rngsetup::SKIP_MIN_WORDS && (!use_lane || dev_row[s2] >= 0)) skip_list.push_back((int)s2);
  sig.n_skip = (long long)skip_list.size();
  sig.n_dev_rows = (long long)n_dev_rows;
  const lane::LaneCaps lcap = lane::lane_caps(max_agents, lane_variant, needs_tv);
  DevBatch& d = h->db;
  int rc;
  size_t ns = (size_t)na + (size_t)ABIDES_EXTRA_STREAMS * ni;
  if (!(h->loaded && sig == h->sig)) {          // (re)allocate; same-shaped batches reuse the device buffers
    free_all(h);
    memset(&d, 0, sizeof d);
    d.n_instances = ni; d.rng_mode = b->rng_mode; d.trace_cap = b->trace_capacity; d.n_types = b->n_types;
    d.b_cap = 4 * max_agents + 4096;
    d.arena_cap = 8 * max_agents + 4096;
    d.deep_cap = 4 * max_agents + 1024;
    abides_instance_cfg* dcfg; abides_agent_type* dty; int* dat; double *dlt, *dlf; long long *dsz, *dwk; int *dath, *dth, *dmi;
    if ((rc = dev_alloc(h, &dcfg, (size_t)ni))) return rc;
    if ((rc = dev_alloc(h, &dty, (size_t)b->n_types))) return rc;
    if ((rc = dev_alloc(h, &dat, (size_t)na))) return rc;
    if ((rc = dev_alloc(h, &dlt, (size_t)na))) return rc;
    if ((rc = dev_alloc(h, &dlf, (size_t)na))) return rc;
    if ((rc = dev_alloc(h, &dsz, (size_t)na))) return rc;
    if ((rc = dev_alloc(h, &dwk, (size_t)na))) return rc;
    if ((rc = dev_alloc(h, &dath, (size_t)na))) return rc;
    if ((rc = dev_alloc(h, &dth, (size_t)(b->n_theta_total > 0 ? b->n_theta_total : 0)))) return rc;
    if ((rc = dev_alloc(h, &dmi, (size_t)na))) return rc;
    d.cfg = dcfg; d.types = dty; d.agent_type = dat; d.lat_to = dlt; d.lat_from = dlf; d.size = dsz; d.wakeup_time = dwk;
    d.agent_theta = dath; d.theta = dth; d.mom_index = dmi;
    if (b->rng_mode == ABIDES_RNG_MT19937) {
      int *dpos, *dhg; double* dg;
      if ((rc = dev_alloc(h, &d.mt, n_dev_rows * ABIDES_MT_N))) return rc;
      h->mt_pristine = nullptr; h->pos0 = nullptr; h->hg0 = nullptr; h->g0 = nullptr; h->d_skip_list = nullptr;
      if (!rematerialize) { if ((rc = dev_alloc(h, &h->mt_pristine, n_dev_rows * ABIDES_MT_N))) return rc; }
      else {
        if ((rc = dev_alloc(h, &h->pos0, ns))) return rc;
        if ((rc = dev_alloc(h, &h->hg0, ns))) return rc;
        if ((rc = dev_alloc(h, &h->g0, ns))) return rc;
      }
      if (!skip_list.empty()) { if ((rc = dev_alloc(h, &h->d_skip_list, skip_list.size()))) return rc; }
      if (seeded) {
        if ((rc = dev_alloc(h, &h->d_mt_seed, ns))) return rc;
        if ((rc = dev_alloc(h, &h->d_key_row, ns))) return rc;
        if (use_lane) { if ((rc = dev_alloc(h, &h->d_dev_row, ns))) return rc; }
      }
      if ((rc = dev_alloc(h, &dpos, ns))) return rc;
      if ((rc = dev_alloc(h, &dhg, ns))) return rc;
      if ((rc = dev_alloc(h, &dg, ns))) return rc;
      d.mt_pos = dpos; d.mt_has_gauss = dhg; d.mt_gauss = dg;
    }
    if ((rc = dev_alloc(h, &d.type_log1mk, (size_t)b->n_types))) return rc;
    if ((rc = dev_alloc(h, &d.type_omp2, (size_t)b->n_types))) return rc;
    if ((rc = dev_alloc(h, &d.prof, (size_t)N_PROF))) return rc;
    if ((rc = dev_alloc(h, &d.mom_ring, (size_t)(n_mom > 0 ? n_mom : 1) * MOM_RING))) return rc;
    if ((rc = dev_alloc(h, &d.trace, (size_t)ni * (d.trace_cap > 0 ? d.trace_cap : 0)))) return rc;
    if ((rc = dev_alloc(h, &d.res, (size_t)ni))) return rc;
    if ((rc = dev_alloc(h, &d.ares, (size_t)na))) return rc;
    memset(&h->lb, 0, sizeof h->lb);
    if (use_lane) {
      lane::LaneBatch& lb = h->lb;
      lb.hcap = lcap.hcap; lb.scap = lcap.scap; lb.arena_cap = lcap.arena_cap; lb.deep_cap = lcap.deep_cap;
      lb.txn_cap = lcap.txn_cap; lb.unr_cap = lcap.unr_cap; lb.top_k = lcap.top_k;
      if ((rc = dev_alloc(h, &lb.agents, (size_t)na))) return rc;
      if ((rc = dev_alloc(h, &lb.state, (size_t)ni))) return rc;
      if ((rc = dev_alloc(h, &lb.hkeys, (size_t)ni * lane::heap_stride(lcap.hcap)))) return rc;
      if ((rc = dev_alloc(h, &lb.hpay, (size_t)ni * lcap.scap))) return rc;
      if ((rc = dev_alloc(h, &lb.hfree, (size_t)ni * lcap.scap))) return rc;
      if ((rc = dev_alloc(h, &lb.top, (size_t)ni * 2 * lcap.top_k))) return rc;
      if ((rc = dev_alloc(h, &lb.pl_price, (size_t)ni * 2 * lcap.deep_cap))) return rc;
      if ((rc = dev_alloc(h, &lb.pl_ord, (size_t)ni * 2 * lcap.deep_cap))) return rc;
      if ((rc = dev_alloc(h, &lb.pl_free, (size_t)ni * 2 * lcap.deep_cap))) return rc;
      if ((rc = dev_alloc(h, &lb.arena, (size_t)ni * lcap.arena_cap))) return rc;
      if (needs_tv) {
        if ((rc = dev_alloc(h, &lb.txn_log, (size_t)ni * lcap.txn_cap))) return rc;
        if ((rc = dev_alloc(h, &lb.unrolled, (size_t)ni * lcap.unr_cap))) return rc;
      }
    } else {
    if ((rc = dev_alloc(h, &d.agents, (size_t)na))) return rc;
    if ((rc = dev_alloc(h, &d.state, (size_t)ni))) return rc;
    if ((rc = dev_alloc(h, &d.task_ctr, (size_t)1))) return rc;
    if ((rc = dev_alloc(h, &d.slice_done, (size_t)ni))) return rc;
    if ((rc = dev_alloc(h, &d.a_keys, (size_t)ni * A_CAP))) return rc;
    if ((rc = dev_alloc(h, &d.a_pay, (size_t)ni * A_CAP))) return rc;
    if ((rc = dev_alloc(h, &d.b_keys, (size_t)ni * d.b_cap))) return rc;
    if ((rc = dev_alloc(h, &d.b_pay, (size_t)ni * d.b_cap))) return rc;
    if ((rc = dev_alloc(h, &d.arena, (size_t)ni * d.arena_cap))) return rc;
    if ((rc = dev_alloc(h, &d.p_keys, (size_t)ni * 2 * d.deep_cap))) return rc;
    if ((rc = dev_alloc(h, &d.p_pay, (size_t)ni * 2 * d.deep_cap))) return rc;
    if ((rc = dev_alloc(h, &d.p_free, (size_t)ni * 2 * d.deep_cap))) return rc;
    if (n_hbl > 0) {
      if ((rc = dev_alloc(h, &d.olog, (size_t)ni * OLOG_CAP))) return rc;
      if ((rc = dev_alloc(h, &d.hbl_cnt, (size_t)ni * 4 * HBL_RANGE_CAP))) return rc;
    }
    if (n_sub > 0) {
      if ((rc = dev_alloc(h, &d.subs, (size_t)ni * MAX_SUBS))) return rc;
      if ((rc = dev_alloc(h, &d.md, (size_t)ni * MD_RING))) return rc;
    }
    if (needs_tv) {                           // 2 MB per instance: only where some agent queries transacted volume
      if ((rc = dev_alloc(h, &d.txn_log, (size_t)ni * TXN_CAP))) return rc;
      if ((rc = dev_alloc(h, &d.unrolled, (size_t)ni * UNR_CAP))) return rc;
    }
    }
    if (use_lane) {                             // the lane engine reads the shared input tables through its own descriptor
      lane::LaneBatch& lb = h->lb;
      lb.n_instances = ni; lb.trace_cap = d.trace_cap; lb.n_types = d.n_types; lb.rng_mode = b->rng_mode;
      { const char* ra = getenv("ABIDES_LANE_RUN_AHEAD"); lb.run_ahead = ra ? atoi(ra) : 0; if (lb.run_ahead < 0) lb.run_ahead = 0; }
      lb.cfg = d.cfg; lb.types = d.types; lb.type_log1mk = d.type_log1mk; lb.type_omp2 = d.type_omp2;
      lb.agent_type = d.agent_type; lb.lat_to = d.lat_to; lb.lat_from = d.lat_from; lb.size = d.size;
      lb.wakeup_time = d.wakeup_time; lb.agent_theta = d.agent_theta; lb.theta = d.theta; lb.mom_index = d.mom_index;
      lb.mt = d.mt; lb.mt_row = seeded ? h->d_dev_row : nullptr; lb.mt_seed = seeded ? h->d_mt_seed : nullptr;
      lb.mt_pos = d.mt_pos; lb.mt_has_gauss = d.mt_has_gauss; lb.mt_gauss = d.mt_gauss;
      lb.mom_ring = d.mom_ring; lb.trace = d.trace; lb.res = d.res; lb.ares = d.ares; lb.prof = d.prof;
    }
    h->sig = sig;
  }
  d.trace_flags = b->trace_flags;
  if (use_lane) h->lb.trace_flags = b->trace_flags;
  h->lane_variant = lane_variant;
  h->n_streams = ns; h->n_mom = n_mom; h->n_dev_rows = n_dev_rows;
  h->rematerialize = rematerialize; h->max_agents = max_agents; h->n_skip = (int)skip_list.size();
  // kernel variant: the leanest one whose feature set covers the batch (engine_device.inc)
  h->variant = VAR_FULL;
  if (b->rng_mode == ABIDES_RNG_MT19937 && b->trace_capacity == 0) {
    const unsigned zi = 1u << ABIDES_ZI;
    const unsigned rmsc01 = (1u << ABIDES_ZI) | (1u << ABIDES_HBL) | (1u << ABIDES_MOMENTUM) | (1u << ABIDES_MARKET_MAKER);
    const unsigned rmsc = (1u << ABIDES_NOISE) | (1u << ABIDES_VALUE) | (1u << ABIDES_MOMENTUM) | (1u << ABIDES_ADAPTIVE_MM) |
                          (1u << ABIDES_POV_EXEC);
    if ((class_mask & ~zi) == 0 && lat_mask == (1u << ABIDES_LAT_LEGACY_NOISE)) h->variant = VAR_ZI_LEGACY;
    else if ((class_mask & ~zi) == 0 && lat_mask == (1u << ABIDES_LAT_CUBIC)) h->variant = VAR_ZI_CUBIC;
    else if ((class_mask & ~rmsc) == 0 && lat_mask == (1u << ABIDES_LAT_DETERMINISTIC) && n_sub == 0) h->variant = VAR_RMSC;
    else if ((class_mask & ~rmsc01) == 0 && lat_mask == (1u << ABIDES_LAT_DETERMINISTIC) && n_sub == 0) h->variant = VAR_RMSC01;
    else if ((class_mask & ~(rmsc01 | (1u << ABIDES_SUBSCRIPTION))) == 0 && lat_mask == (1u << ABIDES_LAT_DETERMINISTIC)) h->variant = VAR_RMSC02;
    else if ((class_mask & ~(rmsc | (1u << ABIDES_SCHEDULE_EXEC))) == 0 && lat_mask == (1u << ABIDES_LAT_DETERMINISTIC) && n_sub == 0) h->variant = VAR_EXEC;
  }
  if ((rc = h2d(h, d.cfg, b->instances, (size_t)ni))) return rc;
  if ((rc = h2d(h, d.types, b->types, (size_t)b->n_types))) return rc;
  if ((rc = h2d(h, d.agent_type, b->agent_type, (size_t)na))) return rc;
  if ((rc = h2d(h, d.lat_to, b->lat_to_exch, (size_t)na))) return rc;
  if ((rc = h2d(h, d.lat_from, b->lat_from_exch, (size_t)na))) return rc;
  if ((rc = h2d(h, d.size, reinterpret_cast<const long long*>(b->agent_size), (size_t)na))) return rc;
  if ((rc = h2d(h, d.wakeup_time, reinterpret_cast<const long long*>(b->agent_wakeup_time), (size_t)na))) return rc;
  if ((rc = h2d(h, d.agent_theta, b->agent_theta, (size_t)na))) return rc;
  if (!draw_theta && (rc = h2d(h, d.theta, b->theta, (size_t)(b->n_theta_total > 0 ? b->n_theta_total : 0)))) return rc;
  if ((rc = h2d(h, d.mom_index, mom_index.data(), (size_t)na))) return rc;
  if (b->rng_mode == ABIDES_RNG_MT19937) {
    unsigned* staging = nullptr;
    if (!seeded) {
      if ((rc = h2d(h, (const unsigned*)h->mt_pristine, b->mt_key, ns * ABIDES_MT_N))) return rc;
    } else if (b->n_mt_rows > 0) {
      if (use_lane) {                           // uploaded rows keep their index: straight into place
        if ((rc = h2d(h, (const unsigned*)h->mt_pristine, b->mt_key, (size_t)b->n_mt_rows * ABIDES_MT_N))) return rc;
      } else {                                  // the warp engine indexes rows by stream: scatter from a staging copy
        cudaError_t e = cudaMalloc(&staging, (size_t)b->n_mt_rows * ABIDES_MT_N * sizeof(unsigned));
        if (e != cudaSuccess) return fail(ABIDES_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        if ((rc = h2d(h, (const unsigned*)staging, b->mt_key, (size_t)b->n_mt_rows * ABIDES_MT_N))) { cudaFree(staging); return rc; }
      }
    }
    if (rematerialize) {
      if ((rc = h2d(h, (const int*)h->pos0, b->mt_pos, ns))) return rc;
      if ((rc = h2d(h, (const int*)h->hg0, b->mt_has_gauss, ns))) return rc;
      if ((rc = h2d(h, (const double*)h->g0, b->mt_gauss, ns))) return rc;
    } else {
      if ((rc = h2d(h, d.mt_pos, b->mt_pos, ns))) return rc;
      if ((rc = h2d(h, d.mt_has_gauss, b->mt_has_gauss, ns))) return rc;
      if ((rc = h2d(h, d.mt_gauss, b->mt_gauss, ns))) return rc;
    }
    if (seeded) {
      if ((rc = h2d(h, (const unsigned*)h->d_mt_seed, b->mt_seed, ns))) return rc;
      if ((rc = h2d(h, (const int*)h->d_key_row, b->mt_key_row, ns))) return rc;
      if (use_lane && (rc = h2d(h, (const int*)h->d_dev_row, dev_row.data(), ns))) return rc;
      if (!skip_list.empty() && (rc = h2d(h, (const int*)h->d_skip_list, skip_list.data(), skip_list.size()))) return rc;
    }
    rngsetup::RngSetup& p = h->rs;
    memset(&p, 0, sizeof p);
    p.n_instances = ni; p.draw_flags = b->draw_flags; p.skip_min = skip_list.empty() ? 0 : rngsetup::SKIP_MIN_WORDS;
    p.cfg = d.cfg; p.types = d.types; p.agent_type = d.agent_type; p.agent_theta = d.agent_theta;
    p.theta = const_cast<int*>(d.theta); p.size = const_cast<long long*>(d.size);
    p.key_row = seeded ? h->d_key_row : nullptr; p.dev_row = (seeded && use_lane) ? h->d_dev_row : nullptr;
    p.seed = h->d_mt_seed;
    p.pos = const_cast<int*>(d.mt_pos); p.has_gauss = const_cast<int*>(d.mt_has_gauss); p.gauss = const_cast<double*>(d.mt_gauss);
    if (!rematerialize && (seeded || b->draw_flags)) {         // once, into the pristine copy every reset restores
      rc = run_rng_setup(h, h->mt_pristine, staging);
      cudaError_t e = rc ? cudaSuccess : cudaStreamSynchronize(h->stream);
      if (staging) cudaFree(staging);
      if (rc) return rc;
      if (e != cudaSuccess) return fail(ABIDES_ECUDA, std::string("rng setup: ") + cudaGetErrorString(e));
    }
  }
  if (b->rng_mode == ABIDES_RNG_PHILOX && b->draw_flags) {
    // Philox streams carry the run, but the agents' constructor draws still come from RandomState(seed): once, from
    // throw-away generators (rng_setup_device.inc)
    const size_t ns2 = (size_t)na + (size_t)ABIDES_EXTRA_STREAMS * ni;
    unsigned* dseed = nullptr; int *dpos = nullptr, *dhg = nullptr; double* dg = nullptr;
    auto cleanup = [&]() { cudaFree(dseed); cudaFree(dpos); cudaFree(dhg); cudaFree(dg); };
    cudaError_t e = cudaMalloc(&dseed, ns2 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMalloc(&dpos, ns2 * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&dhg, ns2 * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&dg, ns2 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(dseed, b->mt_seed, ns2 * sizeof(unsigned), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dpos, b->mt_pos, ns2 * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dhg, b->mt_has_gauss, ns2 * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dg, b->mt_gauss, ns2 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
      rngsetup::RngSetup p;
      memset(&p, 0, sizeof p);
      p.n_instances = ni; p.draw_flags = b->draw_flags;
      p.cfg = d.cfg; p.types = d.types; p.agent_type = d.agent_type; p.agent_theta = d.agent_theta;
      p.theta = const_cast<int*>(d.theta); p.size = const_cast<long long*>(d.size);
      p.seed = dseed; p.pos = dpos; p.has_gauss = dhg; p.gauss = dg;
      const dim3 grid(ni, (max_agents + ABIDES_EXTRA_STREAMS + 63) / 64);
      rngsetup::rng_ctor_draws_kernel<<<grid, 64, 0, h->stream>>>(p);
      e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    }
    cleanup();
    if (e != cudaSuccess) return fail(ABIDES_ECUDA, std::string("constructor draws: ") + cudaGetErrorString(e));
  }
  h->launches_total += h->launches_since_load;
  h->launches_since_load = 0;
  {
    CdbLock lock(h, h->lane_variant == lane::LV_NONE);
    if ((rc = launch_init(h))) return rc;
    CK(cudaStreamSynchronize(h->stream));      // mom_index (a local vector) must outlive the copy
  }
  h->n_instances = ni; h->n_agents_total = na; h->loaded = true;
  return 0;
}

/* Re-arm the loaded batch (inputs stay resident in HBM): restores the MT19937 states and rebuilds the
 * initial device state, so the same batch can be run again. */
int abides_reset(abides_handle* h) {
  if (!h || !h->loaded) return fail(ABIDES_ESTATE, "abides_reset: no batch loaded");
  CK(cudaSetDevice(h->device));
  CdbLock lock(h, h->lane_variant == lane::LV_NONE);
  CK(cudaEventRecord(h->ev0, h->stream));
  int rc = launch_init(h);
  if (rc) return rc;
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventElapsedTime(&h->last_reset_ms, h->ev0, h->ev1));
  return 0;
}

int abides_last_reset_ms(abides_handle* h, float* ms) {
  if (!h) return fail(ABIDES_EINVAL, "NULL handle");
  if (ms) *ms = h->last_reset_ms;
  return 0;
}

int abides_run(abides_handle* h, int64_t until_time) {
  if (!h || !h->loaded) return fail(ABIDES_ESTATE, "abides_run: no batch loaded");
  CK(cudaSetDevice(h->device));
  CdbLock lock(h, h->lane_variant == lane::LV_NONE);
  CK(cudaEventRecord(h->ev0, h->stream));
  if (h->lane_variant != lane::LV_NONE) {
    CK(lane::launch_sim(h->lane_variant, h->lb, (long long)until_time, h->stream));
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    h->last_launches = 1;
    h->launches_since_load += 1;
    return 0;
  }
  CK(cudaMemcpyToSymbolAsync(c_db, &h->db, sizeof(DevBatch), 0, cudaMemcpyHostToDevice, h->stream));
  if (ABM_TIME_SLICES > 1) {
    CK(cudaMemsetAsync(h->db.task_ctr, 0, sizeof(unsigned), h->stream));
    CK(cudaMemsetAsync(h->db.slice_done, 0, sizeof(int) * (size_t)h->n_instances, h->stream));
  }
#define LAUNCH_SIM(ns)                                                                                         \
  ns::sim_kernel<<<(h->n_instances + ABM_CTA_WARPS - 1) / ABM_CTA_WARPS, dim3(32, ABM_CTA_WARPS),              \
                   ABM_CTA_WARPS * sizeof(Smem), h->stream>>>((long long)until_time)
  switch (h->variant) {
    case VAR_ZI_LEGACY: LAUNCH_SIM(v_zi_legacy); break;
    case VAR_ZI_CUBIC: LAUNCH_SIM(v_zi_cubic); break;
    case VAR_RMSC01: LAUNCH_SIM(v_rmsc01); break;
    case VAR_RMSC02: LAUNCH_SIM(v_rmsc02); break;
    case VAR_EXEC: LAUNCH_SIM(v_exec); break;
    case VAR_RMSC: LAUNCH_SIM(v_rmsc); break;
    default: LAUNCH_SIM(v_full); break;
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  h->last_launches = 1;
  h->launches_since_load += 1;
  return 0;
}

int abides_read_profile(abides_handle* h, uint64_t* out, int32_t n) {
  if (!h || !h->loaded || !out) return fail(ABIDES_ESTATE, "abides_read_profile: no batch loaded");
  CK(cudaSetDevice(h->device));
  unsigned long long tmp[N_PROF];
  CK(cudaMemcpy(tmp, h->db.prof, sizeof tmp, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++) out[i] = i < N_PROF ? tmp[i] : 0;
  return 0;
}

int abides_dense_fundamental(abides_handle* h, int32_t n_series, uint32_t* mt_key, int32_t* mt_pos, int32_t* mt_has_gauss,
                             double* mt_gauss, double r_bar, double kappa, double sigma_s, int64_t n_steps, int64_t* out) {
  if (!h || n_series <= 0 || !mt_key || !mt_pos || !mt_has_gauss || !mt_gauss || !out || n_steps <= 0)
    return fail(ABIDES_EINVAL, "abides_dense_fundamental: bad argument");
  if (!(sigma_s >= 0)) return fail(ABIDES_EINVAL, "abides_dense_fundamental: sigma_s must be a variance >= 0");
  for (int i = 0; i < n_series; i++)
    if (mt_pos[i] < 0 || mt_pos[i] > ABIDES_MT_N) return fail(ABIDES_EINVAL, "abides_dense_fundamental: bad mt_pos");
  CK(cudaSetDevice(h->device));
  unsigned* d_mt = nullptr; int *d_pos = nullptr, *d_hg = nullptr; double* d_g = nullptr; long long* d_out = nullptr;
  auto cleanup = [&]() { cudaFree(d_mt); cudaFree(d_pos); cudaFree(d_hg); cudaFree(d_g); cudaFree(d_out); };
#define RCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(ABIDES_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
  const size_t n = (size_t)n_series;
  RCK(cudaMalloc(&d_mt, n * ABIDES_MT_N * sizeof(unsigned)));
  RCK(cudaMalloc(&d_pos, n * sizeof(int)));
  RCK(cudaMalloc(&d_hg, n * sizeof(int)));
  RCK(cudaMalloc(&d_g, n * sizeof(double)));
  RCK(cudaMalloc(&d_out, n * (size_t)n_steps * sizeof(long long)));
  RCK(cudaMemcpyAsync(d_mt, mt_key, n * ABIDES_MT_N * sizeof(unsigned), cudaMemcpyHostToDevice, h->stream));
  RCK(cudaMemcpyAsync(d_pos, mt_pos, n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  RCK(cudaMemcpyAsync(d_hg, mt_has_gauss, n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  RCK(cudaMemcpyAsync(d_g, mt_gauss, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  RCK(cudaEventRecord(h->ev0, h->stream));
  rngsetup::dense_fundamental_kernel<<<(n_series + 63) / 64, 64, 0, h->stream>>>(n_series, d_mt, d_pos, d_hg, d_g, r_bar, kappa,
                                                                                  sigma_s, (long long)n_steps, d_out);
  RCK(cudaGetLastError());
  RCK(cudaEventRecord(h->ev1, h->stream));
  RCK(cudaMemcpyAsync(out, d_out, n * (size_t)n_steps * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  RCK(cudaMemcpyAsync(mt_key, d_mt, n * ABIDES_MT_N * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  RCK(cudaMemcpyAsync(mt_pos, d_pos, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  RCK(cudaMemcpyAsync(mt_has_gauss, d_hg, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  RCK(cudaMemcpyAsync(mt_gauss, d_g, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  RCK(cudaStreamSynchronize(h->stream));
  RCK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
#undef RCK
  h->last_launches = 1;
  h->launches_since_load += 1;
  cleanup();
  return 0;
}

int abides_launch_count(abides_handle* h, int64_t* n) {
  if (!h || !n) return fail(ABIDES_EINVAL, "NULL argument");
  *n = h->launches_total + h->launches_since_load;
  return 0;
}

int abides_last_run_ms(abides_handle* h, float* ms, int32_t* n_launches) {
  if (!h) return fail(ABIDES_EINVAL, "NULL handle");
  if (ms) *ms = h->last_ms;
  if (n_launches) *n_launches = h->last_launches;
  return 0;
}

int abides_read_results(abides_handle* h, abides_instance_result* inst_out, abides_agent_result* agents_out) {
  if (!h || !h->loaded) return fail(ABIDES_ESTATE, "abides_read_results: no batch loaded");
  CK(cudaSetDevice(h->device));
  if (inst_out)
    CK(cudaMemcpyAsync(inst_out, h->db.res, (size_t)h->n_instances * sizeof(abides_instance_result),
                       cudaMemcpyDeviceToHost, h->stream));
  if (agents_out)
    CK(cudaMemcpyAsync(agents_out, h->db.ares, (size_t)h->n_agents_total * sizeof(abides_agent_result),
                       cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int abides_read_trace(abides_handle* h, int32_t instance, abides_trace_rec* out, int64_t capacity, int64_t* n_out) {
  if (!h || !h->loaded) return fail(ABIDES_ESTATE, "abides_read_trace: no batch loaded");
  if (instance < 0 || instance >= h->n_instances) return fail(ABIDES_EINVAL, "abides_read_trace: bad instance");
  CK(cudaSetDevice(h->device));
  int n = 0;
  if (h->lane_variant != lane::LV_NONE) {
    const char* base = reinterpret_cast<const char*>(&h->lb.state[instance]);
    CK(cudaMemcpy(&n, base + offsetof(lane::LHot, n_trace), sizeof(int), cudaMemcpyDeviceToHost));
  } else {
  const char* base = reinterpret_cast<const char*>(&h->db.state[instance]);
  CK(cudaMemcpy(&n, base + offsetof(InstShared, hot) + offsetof(Hot, n_trace), sizeof(int), cudaMemcpyDeviceToHost));
  }
  if (n_out) *n_out = n;
  int64_t k = n;
  if (k > h->db.trace_cap) k = h->db.trace_cap;
  if (k > capacity) k = capacity;
  if (out && k > 0)
    CK(cudaMemcpy(out, h->db.trace + (size_t)instance * h->db.trace_cap, (size_t)k * sizeof(abides_trace_rec),
                  cudaMemcpyDeviceToHost));
  return 0;
}

int abides_replay_book(abides_handle* h, int32_t n_books, const int64_t* op_offset, const abides_book_op* ops,
                       int32_t stream_history, int32_t max_events_per_op, abides_book_event* events_out,
                       int32_t* event_count_out, abides_book_state* states_out) {
  if (!h || n_books <= 0 || !op_offset || !ops || !events_out || !event_count_out || !states_out || max_events_per_op <= 0)
    return fail(ABIDES_EINVAL, "abides_replay_book: bad argument");
  CK(cudaSetDevice(h->device));
  long long n_ops = op_offset[n_books];
  if (n_ops <= 0) return 0;
  long long longest = 0;
  for (int b = 0; b < n_books; b++) {
    long long len = op_offset[b + 1] - op_offset[b];
    if (len < 0) return fail(ABIDES_EINVAL, "abides_replay_book: op_offset must be non-decreasing");
    if (len > longest) longest = len;
  }
  for (long long i = 0; i < n_ops; i++)
    if (ops[i].price < 0 || ops[i].price > 2000000000ll || ops[i].qty > 2000000000ll || ops[i].order_id >= (1ll << 30))
      return fail(ABIDES_EINVAL, "abides_replay_book: price / quantity / order id outside the device's int32 range");
  int deep_cap = (int)(longest + 64);
  if (h->engine_pref == ABIDES_ENGINE_LANE) {          // the lane engine's book: one thread per book
    long long* d_off = nullptr; abides_book_op* d_ops = nullptr; abides_book_event* d_ev = nullptr; int* d_cnt = nullptr;
    abides_book_state* d_st = nullptr;
    lane::LaneBatch lb;
    memset(&lb, 0, sizeof lb);
    lb.top_k = lane::lane_top_k(lane::LV_FULL); lb.deep_cap = deep_cap;
    auto cleanup = [&]() { cudaFree(d_off); cudaFree(d_ops); cudaFree(d_ev); cudaFree(d_cnt); cudaFree(d_st); cudaFree(lb.top);
                           cudaFree(lb.pl_price); cudaFree(lb.pl_ord); cudaFree(lb.pl_free); };
#define RCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(ABIDES_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
    RCK(cudaMalloc(&d_off, sizeof(long long) * (n_books + 1)));
    RCK(cudaMalloc(&d_ops, sizeof(abides_book_op) * n_ops));
    RCK(cudaMalloc(&d_ev, sizeof(abides_book_event) * n_ops * max_events_per_op));
    RCK(cudaMalloc(&d_cnt, sizeof(int) * n_ops));
    RCK(cudaMalloc(&d_st, sizeof(abides_book_state) * n_ops));
    RCK(cudaMalloc(&lb.top, sizeof(lane::LOrd) * (size_t)n_books * 2 * lb.top_k));
    RCK(cudaMalloc(&lb.pl_price, sizeof(int) * (size_t)n_books * 2 * deep_cap));
    RCK(cudaMalloc(&lb.pl_ord, sizeof(lane::LOrd) * (size_t)n_books * 2 * deep_cap));
    RCK(cudaMalloc(&lb.pl_free, sizeof(int) * (size_t)n_books * 2 * deep_cap));
    RCK(cudaMemcpyAsync(d_off, op_offset, sizeof(long long) * (n_books + 1), cudaMemcpyHostToDevice, h->stream));
    RCK(cudaMemcpyAsync(d_ops, ops, sizeof(abides_book_op) * n_ops, cudaMemcpyHostToDevice, h->stream));
    RCK(cudaMemsetAsync(d_ev, 0, sizeof(abides_book_event) * n_ops * max_events_per_op, h->stream));
    RCK(cudaEventRecord(h->ev0, h->stream));
    RCK(lane::launch_replay(lb, n_books, d_off, d_ops, stream_history, max_events_per_op, d_ev, d_cnt, d_st, h->stream));
    RCK(cudaGetLastError());
    RCK(cudaEventRecord(h->ev1, h->stream));
    RCK(cudaMemcpyAsync(events_out, d_ev, sizeof(abides_book_event) * n_ops * max_events_per_op, cudaMemcpyDeviceToHost, h->stream));
    RCK(cudaMemcpyAsync(event_count_out, d_cnt, sizeof(int) * n_ops, cudaMemcpyDeviceToHost, h->stream));
    RCK(cudaMemcpyAsync(states_out, d_st, sizeof(abides_book_state) * n_ops, cudaMemcpyDeviceToHost, h->stream));
    RCK(cudaStreamSynchronize(h->stream));
    RCK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
#undef RCK
    h->last_launches = 1;
    cleanup();
    return 0;
  }
  long long* d_off = nullptr; abides_book_op* d_ops = nullptr; abides_book_event* d_ev = nullptr; int* d_cnt = nullptr;
  abides_book_state* d_st = nullptr; EvKey* d_deep = nullptr; EvPay* d_seq = nullptr; int* d_free = nullptr;
  int rc = 0;
  auto cleanup = [&]() { cudaFree(d_off); cudaFree(d_ops); cudaFree(d_ev); cudaFree(d_cnt); cudaFree(d_st); cudaFree(d_deep); cudaFree(d_seq); cudaFree(d_free); };
#define RCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(ABIDES_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
  RCK(cudaMalloc(&d_off, sizeof(long long) * (n_books + 1)));
  RCK(cudaMalloc(&d_ops, sizeof(abides_book_op) * n_ops));
  RCK(cudaMalloc(&d_ev, sizeof(abides_book_event) * n_ops * max_events_per_op));
  RCK(cudaMalloc(&d_cnt, sizeof(int) * n_ops));
  RCK(cudaMalloc(&d_st, sizeof(abides_book_state) * n_ops));
  RCK(cudaMalloc(&d_deep, sizeof(EvKey) * (size_t)n_books * 2 * deep_cap));
  RCK(cudaMalloc(&d_seq, sizeof(EvPay) * (size_t)n_books * 2 * deep_cap));
  RCK(cudaMalloc(&d_free, sizeof(int) * (size_t)n_books * 2 * deep_cap));
  RCK(cudaMemcpyAsync(d_off, op_offset, sizeof(long long) * (n_books + 1), cudaMemcpyHostToDevice, h->stream));
  RCK(cudaMemcpyAsync(d_ops, ops, sizeof(abides_book_op) * n_ops, cudaMemcpyHostToDevice, h->stream));
  RCK(cudaMemsetAsync(d_ev, 0, sizeof(abides_book_event) * n_ops * max_events_per_op, h->stream));
  RCK(cudaFuncSetAttribute(v_full::replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  RCK(cudaEventRecord(h->ev0, h->stream));
  v_full::replay_kernel<<<n_books, 32, sizeof(Smem), h->stream>>>(n_books, d_off, d_ops, stream_history, max_events_per_op, d_ev, d_cnt,
                                                         d_st, d_deep, d_seq, d_free, deep_cap);
  RCK(cudaGetLastError());
  RCK(cudaEventRecord(h->ev1, h->stream));
  RCK(cudaMemcpyAsync(events_out, d_ev, sizeof(abides_book_event) * n_ops * max_events_per_op, cudaMemcpyDeviceToHost, h->stream));
  RCK(cudaMemcpyAsync(event_count_out, d_cnt, sizeof(int) * n_ops, cudaMemcpyDeviceToHost, h->stream));
  RCK(cudaMemcpyAsync(states_out, d_st, sizeof(abides_book_state) * n_ops, cudaMemcpyDeviceToHost, h->stream));
  RCK(cudaStreamSynchronize(h->stream));
  RCK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
#undef RCK
  h->last_launches = 1;
  cleanup();
  (void)rc;
  return 0;
}

}  // extern "C"
