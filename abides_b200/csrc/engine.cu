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
#include <string>
#include <vector>

#include "abides_b200.h"
#include "abides_dd_math.h"

#define FULL 0xffffffffu
#define UNLIKELY(x) __builtin_expect(!!(x), 0)
#define LIKELY(x) __builtin_expect(!!(x), 1)

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
  F_KNOWN_BOOK = 32u, F_TRADING = 64u, F_HAS_PREV_WAKE = 128u,
  F_MM_AWAIT_SPREAD = 512u, F_MM_AWAIT_TV = 1024u, F_HAS_LAST_MID = 2048u, F_HAS_AVG20 = 4096u,
  F_HAS_AVG50 = 8192u, F_HAS_TV = 16384u, F_STATE_SHIFT = 16u, F_STATE_MASK = 7u << 16
};
enum { ST_AWAITING_WAKEUP = 0, ST_INACTIVE = 1, ST_AWAITING_SPREAD = 2, ST_ACTIVE = 3, ST_AWAITING_STREAM = 4 };
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
  } u;
};
static_assert(sizeof(AgentRec) == 128, "AgentRec must be 128 bytes");

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
#define SM (reinterpret_cast<Smem*>(smem_raw))
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
// RNG: numpy legacy RandomState transforms over MT19937 (HBM-resident state) or Philox4x32-10.
// Streams are named by a small id: 0..3 = the instance's extra streams (state in HOT.rs), SID_AGENT
// = the stream of the agent record currently in shared memory (state in ST.ag.rng).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ RngState& rng_state(int sid) {
  unsigned off = sid == SID_AGENT ? (unsigned)(offsetof(Smem, st) + offsetof(InstShared, ag) + offsetof(AgentRec, rng))
                                  : (unsigned)(offsetof(Smem, st) + offsetof(InstShared, hot) + offsetof(Hot, rs)) +
                                        (unsigned)sid * (unsigned)sizeof(RngState);
  return *reinterpret_cast<RngState*>(smem_raw + off);
}
__device__ __forceinline__ int rng_stream(int sid) { return sid == SID_AGENT ? RUN.cur_agent : RUN.n_agents + sid; }

__device__ __noinline__ void mt_twist(unsigned* mt) {
  const unsigned U = 0x80000000u, L = 0x7fffffffu, MAG = 0x9908b0dfu;
  const int lane = LANE;
  __syncwarp();
#pragma unroll 1
  for (int base = 0; base < 227; base += 32) {
    int k = base + lane; bool act = k < 227; unsigned v = 0;
    if (act) { unsigned y = (mt[k] & U) | (mt[k + 1] & L); v = mt[k + 397] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u); }
    __syncwarp();
    if (act) mt[k] = v;
    __syncwarp();
  }
#pragma unroll 1
  for (int base = 227; base < 623; base += 32) {
    int k = base + lane; bool act = k < 623; unsigned v = 0;
    if (act) { unsigned y = (mt[k] & U) | (mt[k + 1] & L); v = mt[k - 227] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u); }
    __syncwarp();
    if (act) mt[k] = v;
    __syncwarp();
  }
  if (lane == 0) { unsigned y = (mt[623] & U) | (mt[0] & L); mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u); }
  __syncwarp();
}

__device__ __forceinline__ void philox4x32_10(unsigned long long ctr, unsigned long long key, unsigned out[4]) {
  unsigned c0 = (unsigned)ctr, c1 = (unsigned)(ctr >> 32), c2 = 0, c3 = 0;
  unsigned k0 = (unsigned)key, k1 = (unsigned)(key >> 32);
#pragma unroll 1
  for (int i = 0; i < 10; i++) {
    unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ unsigned long long philox_key(unsigned long long seed, int stream) {
  return seed * 0x9E3779B97F4A7C15ull + (unsigned long long)stream * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull;
}

__device__ __noinline__ unsigned philox_u32(int sid) {
  RngState& r = rng_state(sid);
  unsigned n = (unsigned)r.pos;
  unsigned o[4];
  philox4x32_10(n >> 2, philox_key(RUN.cfg.seed, rng_stream(sid)), o);
  r.pos = (int)(n + 1);
  unsigned sel = n & 3u;
  return sel == 0 ? o[0] : sel == 1 ? o[1] : sel == 2 ? o[2] : o[3];
}
__device__ __noinline__ unsigned rng_u32(int sid) {
  if (UNLIKELY(c_db.rng_mode == ABIDES_RNG_PHILOX)) return philox_u32(sid);
  RngState& r = rng_state(sid);
  unsigned* mt = RUN.mt0 + (size_t)rng_stream(sid) * ABIDES_MT_N;
  int p = r.pos;
  if (UNLIKELY(p >= ABIDES_MT_N)) { mt_twist(mt); p = 0; }
  unsigned y = mt[p];
  r.pos = p + 1;
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
__device__ __noinline__ double rng_double(int sid) {
  int a = (int)(rng_u32(sid) >> 5), b = (int)(rng_u32(sid) >> 6);
  return (a * 67108864.0 + b) / 9007199254740992.0;
}
__device__ __noinline__ double rng_gauss(int sid) {
  RngState& r = rng_state(sid);
  if (r.flags & 1u) {
    double t = r.gauss;
    r.flags = 0u;
    r.gauss = 0.0;
    return t;
  }
  double f, x1, x2, r2;
  do {
    x1 = 2.0 * rng_double(sid) - 1.0;
    x2 = 2.0 * rng_double(sid) - 1.0;
    r2 = x1 * x1 + x2 * x2;
  } while (r2 >= 1.0 || r2 == 0.0);
  f = sqrt(fdiv(-2.0 * abm_log(r2), r2));
  r.gauss = f * x1;
  r.flags = 1u;
  return f * x2;
}
__device__ __forceinline__ double rng_normal(int sid, double loc, double scale) { return loc + scale * rng_gauss(sid); }
__device__ __forceinline__ double rng_exponential(int sid, double scale) { return scale * (-abm_log(1.0 - rng_double(sid))); }
__device__ __forceinline__ double rng_uniform(int sid, double lo, double hi) { return lo + (hi - lo) * rng_double(sid); }
__device__ __noinline__ long long rng_randint(int sid, long long lo, long long hi) {
  unsigned long long rng = (unsigned long long)(hi - 1 - lo);
  if (rng == 0) return lo;
  if (rng == 0xFFFFFFFFull) return lo + (long long)rng_u32(sid);
  unsigned mask = (unsigned)rng;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  unsigned v;
  while ((v = (rng_u32(sid) & mask)) > (unsigned)rng) {}
  return lo + (long long)v;
}

// ----------------------------------------------------------------------------------------------
// trace
// ----------------------------------------------------------------------------------------------
__device__ __noinline__ void trace_write(long long time, int agent, int code, long long p0, long long p1, long long p2, long long p3) {
  int i = HOT.n_trace;
  HOT.n_trace = i + 1;
  if (!RUN.trace || i >= c_db.trace_cap) return;
  if (LANE == 0) {
    abides_trace_rec r;
    r.time = time; r.agent = agent; r.code = code; r.p0 = p0; r.p1 = p1; r.p2 = p2; r.p3 = p3;
    RUN.trace[i] = r;
  }
  __syncwarp();
}

// ----------------------------------------------------------------------------------------------
// event queue.  Invariants: every near event < near_limit (HOT.lim_*) <= every event of list A and B;
// every A event < horizon (HOT.hz_*) <= every B event.  Lists are unsorted, append-only, with
// tombstones (t == LLONG_MAX); which tier an event sits in never changes the pop order.
// ----------------------------------------------------------------------------------------------
__device__ __noinline__ void list_compact(int L) {
  EvKey* keys = RUN.lk[L]; EvPay* pay = RUN.lp[L];
  int n = HOT.l_total[L], w = 0;
  __syncwarp();
#pragma unroll 1
  for (int base = 0; base < n; base += 32) {
    int i = base + LANE;
    bool live = false; EvKey k; EvPay p;
    if (i < n) { k = keys[i]; live = k.t != LLONG_MAX; if (live) p = pay[i]; }
    unsigned m = __ballot_sync(FULL, live);
    __syncwarp();
    if (live) {
      int j = w + __popc(m & ((1u << LANE) - 1u));
      if (j != i) { keys[j] = k; pay[j] = p; }
    }
    w += __popc(m);
    __syncwarp();
  }
  HOT.l_total[L] = w;
}

__device__ __noinline__ void list_append(int L, long long t, unsigned long long k2, int w0, int w1, int w2, int w3,
                                         int w4, int w5) {
  int i = HOT.l_total[L];
  if (UNLIKELY(i >= RUN.lcap[L])) {
    list_compact(L);
    i = HOT.l_total[L];
    if (i >= RUN.lcap[L]) { HOT.status = ABIDES_EOVERFLOW; return; }
  }
  if (LANE == 0) { EvKey k; k.t = t; k.k2 = k2; RUN.lk[L][i] = k; }
  else if (LANE == 1) *reinterpret_cast<int4*>(&RUN.lp[L][i].w[0]) = make_int4(w0, w1, w2, w3);
  else if (LANE == 2) *reinterpret_cast<int4*>(&RUN.lp[L][i].w[4]) = make_int4(w4, w5, 0, 0);
  __syncwarp();
  HOT.l_total[L] = i + 1;
  HOT.l_live[L] += 1;
}

// ---- near tier: <= 32 events sorted descending in shared memory, payloads in free-listed slots ----
__device__ __forceinline__ int near_alloc(int w0, int w1, int w2, int w3, int w4, int w5) {
  const unsigned fm = HOT.near_free;
  const int slot = __ffs(fm) - 1;
  if (LANE == 2) *reinterpret_cast<int4*>(&ST.npay[slot].w[0]) = make_int4(w0, w1, w2, w3);
  else if (LANE == 3) *reinterpret_cast<int4*>(&ST.npay[slot].w[4]) = make_int4(w4, w5, 0, 0);
  HOT.near_free = fm & ~(1u << slot);
  return slot;
}
// insert (t, k2) -> payload slot; needs n_near < NEAR_CAP
__device__ __forceinline__ void near_insert_body(long long t, unsigned long long k2, int slot) {
  const int n = HOT.n_near, lane = LANE;
  long long et = 0; unsigned long long ek = 0; int es = 0;
  bool gt = false;
  if (lane < n) { et = ST.nt[lane]; ek = ST.nk[lane]; es = ST.nslot[lane]; gt = key_less(t, k2, et, ek); }
  const int pos = __popc(__ballot_sync(FULL, gt));            // entries above the new key keep their place
  if (lane >= pos && lane < n) { ST.nt[lane + 1] = et; ST.nk[lane + 1] = ek; ST.nslot[lane + 1] = es; }
  if (lane == 0) { ST.nt[pos] = t; ST.nk[pos] = k2; ST.nslot[pos] = slot; }
  HOT.n_near = n + 1;
  __syncwarp();
}
__device__ __noinline__ void near_insert(long long t, unsigned long long k2, int slot) { near_insert_body(t, k2, slot); }
// near tier full: hand its largest key (entry 0) to list A and pull near_limit down to it
__device__ __noinline__ void near_spill() {
  const int lane = LANE;
  const long long mt = ST.nt[0]; const unsigned long long mk = ST.nk[0]; const int ms = ST.nslot[0];
  const int4 p0 = *reinterpret_cast<const int4*>(&ST.npay[ms].w[0]);
  const int4 p1 = *reinterpret_cast<const int4*>(&ST.npay[ms].w[4]);
  long long et = 0; unsigned long long ek = 0; int es = 0;
  if (lane >= 1) { et = ST.nt[lane]; ek = ST.nk[lane]; es = ST.nslot[lane]; }
  __syncwarp();
  if (lane >= 1) { ST.nt[lane - 1] = et; ST.nk[lane - 1] = ek; ST.nslot[lane - 1] = es; }
  HOT.n_near = NEAR_CAP - 1;
  HOT.near_free |= 1u << ms;
  __syncwarp();
  list_append(L_A, mt, mk, p0.x, p0.y, p0.z, p0.w, p1.x, p1.y);
  HOT.lim_t = mt; HOT.lim_k = mk;
}

__device__ __noinline__ int list_count_lt(const EvKey* __restrict__ keys, int n, long long pt, unsigned long long pk) {
  int cnt = 0;
#pragma unroll 4
  for (int i = LANE; i < n; i += 32) {
    EvKey k = keys[i];
    cnt += key_less(k.t, k.k2, pt, pk) ? 1 : 0;
  }
  return warp_sum_i(cnt);
}

// Pick a key limit for list L such that 1 <= #(entries < limit) <= kmax (aiming at >= kmin), limit <= (ubt, ubk).
// The window width HOT.span[S] adapts.  More than kmax events on one timestamp are split on the secondary key
// (recipient, kind, uniq).  Result in HOT.sel_*; returns the count.
__device__ __noinline__ int choose_limit(int L, int S, int kmax, int kmin, int hard_cap, long long ubt,
                                         unsigned long long ubk) {
  const EvKey* keys = RUN.lk[L];
  const int n = HOT.l_total[L];
  __syncwarp();
  long long tmin = LLONG_MAX;
#pragma unroll 4
  for (int i = LANE; i < n; i += 32) { long long t = keys[i].t; tmin = t < tmin ? t : tmin; }
  tmin = warp_min_ll(tmin);
  long long dT = HOT.span[S], pt; unsigned long long pk = 0; int cnt;
  bool burst = false, clamped = false;
  for (;;) {
    pt = (tmin > LLONG_MAX - 2 - dT) ? LLONG_MAX - 1 : tmin + dT; pk = 0;
    clamped = !key_less(pt, pk, ubt, ubk);
    if (clamped) { pt = ubt; pk = ubk; }
    cnt = list_count_lt(keys, n, pt, pk);
    if (cnt <= kmax) break;
    if (dT <= 1) { burst = true; break; }
    dT >>= 2; if (dT < 1) dT = 1;
  }
  if (burst) {
    // more than kmax events share the timestamp tmin: interpolate (recipients are spread evenly), then bisect
    unsigned long long kmin_k = ~0ull, kmax_k = 0;
    for (int i = LANE; i < n; i += 32) {
      EvKey k = keys[i];
      if (k.t == tmin) { kmin_k = k.k2 < kmin_k ? k.k2 : kmin_k; kmax_k = k.k2 > kmax_k ? k.k2 : kmax_k; }
    }
    kmin_k = warp_min_ull(kmin_k); kmax_k = warp_max_ull(kmax_k);
    unsigned long long lo = kmin_k, hi = kmax_k + 1;         // count(<lo) == 0, count(<hi) == cnt > kmax
    unsigned long long best = 0; int best_cnt = 0;
    double frac = (double)(kmax * 3 / 4) / (double)cnt;
    unsigned long long guess = kmin_k + (unsigned long long)((double)(kmax_k - kmin_k) * frac) + 1;
    for (int it = 0; it < 80; it++) {
      unsigned long long mid = it == 0 ? guess : lo + (hi - lo) / 2;
      if (mid <= lo) mid = lo + 1;
      if (mid >= hi) break;
      int cc = list_count_lt(keys, n, tmin, mid);
      if (cc > kmax) hi = mid;
      else { lo = mid; if (cc > best_cnt) { best = mid; best_cnt = cc; } if (cc >= kmin) break; }
    }
    if (best_cnt == 0) { best = hi; best_cnt = list_count_lt(keys, n, tmin, hi); }   // identical keys
    pt = tmin; pk = best; cnt = best_cnt;
    if (cnt > hard_cap) { HOT.status = ABIDES_EOVERFLOW; cnt = 0; pk = 0; }
  } else {
    if (!clamped && cnt < kmin && dT < (1ll << 50)) dT <<= 1;
    HOT.span[S] = dT;
  }
  HOT.sel_t = pt; HOT.sel_k = pk;
  return cnt;
}

// list A ran dry: move the next few hundred events of B into it and advance the horizon
__device__ __noinline__ void b_pull() {
  PROF_CNT(P_N_BPULL);
  if (HOT.l_total[1] > 1024 && HOT.l_live[1] * 2 < HOT.l_total[1]) list_compact(1);
  int cnt = choose_limit(L_B, L_B, A_FILL_MAX, A_FILL_MIN, A_CAP / 2, LLONG_MAX - 1, 0ull);
  const long long pt = HOT.sel_t; const unsigned long long pk = HOT.sel_k;
  EvKey* bk = RUN.lk[1]; EvPay* bp = RUN.lp[1]; EvKey* ak = RUN.lk[0]; EvPay* ap = RUN.lp[0];
  const int n = HOT.l_total[1];
  int na = 0;
#pragma unroll 1
  for (int base = 0; base < n; base += 32) {
    int i = base + LANE;
    bool take = false; EvKey k;
    if (i < n) { k = bk[i]; take = key_less(k.t, k.k2, pt, pk); }
    unsigned m = __ballot_sync(FULL, take);
    if (m) {
      if (take) {
        int pos = na + __popc(m & ((1u << LANE) - 1u));
        ak[pos] = k; ap[pos] = bp[i];
        bk[i].t = LLONG_MAX;
      }
      na += __popc(m);
    }
  }
  __syncwarp();
  HOT.l_total[0] = na; HOT.l_live[0] = na;
  HOT.l_live[1] -= cnt;
  HOT.hz_t = pt; HOT.hz_k = pk;
  __syncwarp();
}

// list A grew past A_REBALANCE (a burst inside the horizon): pull the horizon in and hand the tail back to B
__device__ __noinline__ void a_rebalance() {
  list_compact(0);
  int cnt = choose_limit(L_A, L_B, A_FILL_MAX, A_FILL_MIN, A_CAP / 2, HOT.hz_t, HOT.hz_k);
  if (HOT.status) return;
  const long long pt = HOT.sel_t; const unsigned long long pk = HOT.sel_k;
  const int n = HOT.l_total[0];
  int moved = n - cnt;
  if (HOT.l_total[1] + moved > RUN.lcap[1]) {
    list_compact(1);
    if (HOT.l_total[1] + moved > RUN.lcap[1]) { HOT.status = ABIDES_EOVERFLOW; return; }
  }
  EvKey* bk = RUN.lk[1]; EvPay* bp = RUN.lp[1]; EvKey* ak = RUN.lk[0]; EvPay* ap = RUN.lp[0];
  int nb = HOT.l_total[1];
#pragma unroll 1
  for (int base = 0; base < n; base += 32) {
    int i = base + LANE;
    bool take = false; EvKey k;
    if (i < n) { k = ak[i]; take = !key_less(k.t, k.k2, pt, pk); }
    unsigned m = __ballot_sync(FULL, take);
    if (m) {
      if (take) {
        int pos = nb + __popc(m & ((1u << LANE) - 1u));
        bk[pos] = k; bp[pos] = ap[i];
        ak[i].t = LLONG_MAX;
      }
      nb += __popc(m);
    }
  }
  __syncwarp();
  HOT.l_total[1] = nb; HOT.l_live[1] += moved;
  HOT.l_live[0] = cnt;
  HOT.hz_t = pt; HOT.hz_k = pk;
  __syncwarp();
  list_compact(0);
}

__device__ __noinline__ void q_push(long long t, unsigned long long k2, int w0, int w1, int w2, int w3, int w4, int w5) {
  if (key_less(t, k2, HOT.lim_t, HOT.lim_k)) {
    if (UNLIKELY(HOT.n_near == NEAR_CAP)) near_spill();
    if (key_less(t, k2, HOT.lim_t, HOT.lim_k)) near_insert(t, k2, near_alloc(w0, w1, w2, w3, w4, w5));
    else list_append(L_A, t, k2, w0, w1, w2, w3, w4, w5);
  } else if (key_less(t, k2, HOT.hz_t, HOT.hz_k)) {
    list_append(L_A, t, k2, w0, w1, w2, w3, w4, w5);
    if (UNLIKELY(HOT.l_live[L_A] > A_REBALANCE)) a_rebalance();
  } else {
    list_append(L_B, t, k2, w0, w1, w2, w3, w4, w5);
  }
}
__device__ __forceinline__ void track_queue_peak() {      // sampled at refills (diagnostic only)
  int tot = HOT.n_near + HOT.l_live[L_A] + HOT.l_live[L_B];
  if (tot > HOT.queue_peak) HOT.queue_peak = tot;
}

// near tier is empty: pull the next batch (<= FILL_MAX smallest keys) out of list A (refilled from B when dry)
__device__ __noinline__ void q_refill() {
  PROF_CNT(P_N_REFILL);
  track_queue_peak();
  if (HOT.l_live[L_A] == 0) {
    HOT.l_total[L_A] = 0;
    if (HOT.l_live[L_B] == 0) return;
    b_pull();
    if (HOT.status) return;
  }
  if (HOT.l_total[L_A] > 256 && HOT.l_live[L_A] * 2 < HOT.l_total[L_A]) list_compact(L_A);
  const int cnt = choose_limit(L_A, L_A, FILL_MAX, FILL_MIN, NEAR_CAP, HOT.hz_t, HOT.hz_k);
  const long long pt = HOT.sel_t; const unsigned long long pk = HOT.sel_k;
  EvKey* keys = RUN.lk[L_A]; const EvPay* pay = RUN.lp[L_A];
  const int n = HOT.l_total[L_A];
  int nn = 0;                                   // the near tier is empty: payload slot i <-> i-th taken event
#pragma unroll 1
  for (int base = 0; base < n; base += 32) {
    int i = base + LANE;
    bool take = false; EvKey k;
    if (i < n) { k = keys[i]; take = key_less(k.t, k.k2, pt, pk); }
    unsigned m = __ballot_sync(FULL, take);
    if (m) {
      if (take) {
        int pos = nn + __popc(m & ((1u << LANE) - 1u));
        ST.nt[pos] = k.t; ST.nk[pos] = k.k2; ST.npay[pos] = pay[i];
        keys[i].t = LLONG_MAX;
      }
      nn += __popc(m);
    }
  }
  __syncwarp();
  // rank sort, descending: lane i owns taken event i
  long long et = 0; unsigned long long ek = 0; int rank = 0;
  if (LANE < nn) { et = ST.nt[LANE]; ek = ST.nk[LANE]; }
#pragma unroll 1
  for (int j = 0; j < nn; j++) {
    long long ot = ST.nt[j]; unsigned long long ok = ST.nk[j];
    rank += (key_less(et, ek, ot, ok) || (et == ot && ek == ok && j < LANE)) ? 1 : 0;
  }
  __syncwarp();
  if (LANE < nn) { ST.nt[rank] = et; ST.nk[rank] = ek; ST.nslot[rank] = LANE; }
  HOT.n_near = nn;
  HOT.near_free = nn >= 32 ? 0u : ~((1u << nn) - 1u);
  HOT.l_live[L_A] -= cnt;
  HOT.lim_t = pt; HOT.lim_k = pk;
  __syncwarp();
}

// ----------------------------------------------------------------------------------------------
// Kernel.sendMessage / setWakeup  (Kernel.py:329-418).  Payload words: w0 = code | mkt_closed<<8 | is_buy<<9;
// order-carrying: w1 sender (to exchange), w2 id, w3 price, w4 qty, w5 fill price / new qty;
// QUERY_SPREAD reply: w1 bid, w2 bid volume, w3 ask, w4 ask volume, w5 last trade.
// ----------------------------------------------------------------------------------------------
__device__ __noinline__ void k_send(int sender, int rcpt, long long delay, int w0, int w1, int w2, int w3, int w4, int w5) {
  const abides_instance_cfg& cfg = RUN.cfg;
  const double minlat = sender == 0 ? RUN.lat_from[rcpt] : RUN.lat_to[sender];
  const long long cd = sender == 0 ? HOT.exch_cd : cfg.default_computation_delay;
  const long long sent = HOT.now + cd + HOT.additional_delay + delay;
  const int kind = cfg.latency_kind;
  long long lat;
  if (kind == ABIDES_LAT_LEGACY_NOISE) {
    // random_state.choice(len(noise), 1, noise) == randint(0, len(noise)) (Kernel.py:380): masked rejection
    const unsigned rng = (unsigned)(cfg.n_latency_noise - 1);
    unsigned v = 0;
    if (rng != 0) {
      unsigned mask = rng;
      mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
      while ((v = (rng_u32(ABIDES_STREAM_KERNEL) & mask)) > rng) {}
    }
    lat = (long long)(minlat + (double)(long long)v);
  } else if (kind == ABIDES_LAT_DETERMINISTIC) {
    lat = (long long)minlat;
  } else {
    double x = rng_uniform(ABIDES_STREAM_LATENCY, cfg.jitter_clip, 1.0);
    double l = minlat + (fdiv(cfg.jitter, abm_cube(x)) * fdiv(minlat, cfg.jitter_unit));
    lat = (long long)l;
  }
  const unsigned u = HOT.uniq;
  HOT.uniq = u + 1;
  const long long t = sent + lat;
  const unsigned long long k2 = ((unsigned long long)(unsigned)rcpt << 33) | (unsigned long long)u;
  // fast path: straight into the sorted shared-memory tier
  if (key_less(t, k2, HOT.lim_t, HOT.lim_k) && HOT.n_near < NEAR_CAP) near_insert_body(t, k2, near_alloc(w0, w1, w2, w3, w4, w5));
  else q_push(t, k2, w0, w1, w2, w3, w4, w5);
}
__device__ __noinline__ void k_wakeup(int agent, long long t) {
  if (t < HOT.now) { HOT.status = ABIDES_EINVAL; return; }
  q_push(t, ((unsigned long long)(unsigned)agent << 33) | (1ull << 32), 0, 0, 0, 0, 0, 0);
}
__device__ __forceinline__ int pay_code(int w0) { return w0 & 0xff; }
__device__ __forceinline__ bool pay_closed(int w0) { return (w0 >> 8) & 1; }
__device__ __forceinline__ bool pay_is_buy(int w0) { return (w0 >> 9) & 1; }
__device__ __forceinline__ int order_w0(int code, bool is_buy) { return code | (is_buy ? 1 << 9 : 0); }

// ----------------------------------------------------------------------------------------------
// oracle (util/oracle/SparseMeanRevertingOracle.py)
// ----------------------------------------------------------------------------------------------
__device__ __noinline__ double oracle_step(long long ts, double v_adj, long long pt, double pv) {
  const abides_instance_cfg& cfg = RUN.cfg;
  double d = (double)(ts - pt);
  double mu = cfg.r_bar, gamma = cfg.kappa, theta = cfg.fund_vol;
  double e1 = abm_exp(-gamma * d);
  double loc = mu + (pv - mu) * e1;
  double e2 = abm_exp(-2 * gamma * d);
  double scale = fdiv(theta, 2 * gamma) * (1 - e2);
  double v = rng_normal(ABIDES_STREAM_ORACLE, loc, scale);
  v += v_adj;
  if (!(v > 0)) v = 0;
  v = rint(v);
  HOT.or_pt = ts; HOT.or_pv = v;
  if (UNLIKELY(RUN.trace != nullptr)) trace_write(ts, -1, TRACE_FUNDAMENTAL, (long long)v, 0, 0, 0);
  return v;
}
__device__ __noinline__ double oracle_advance(long long t) {
  const abides_instance_cfg& cfg = RUN.cfg;
  long long pt = HOT.or_pt; double pv = HOT.or_pv;
  if (t <= pt) return pv;
  while (UNLIKELY(HOT.ms_time < t)) {
    double v = oracle_step(HOT.ms_time, HOT.ms_value, pt, pv);
    pt = HOT.ms_time; pv = v;
    double dt = rng_exponential(ABIDES_STREAM_GLOBAL, fdiv(1.0, cfg.megashock_lambda_a));
    HOT.ms_time = pt + (long long)dt;
    double msv = rng_normal(ABIDES_STREAM_ORACLE, cfg.megashock_mean, sqrt(cfg.megashock_var));
    HOT.ms_value = rng_randint(ABIDES_STREAM_ORACLE, 0, 2) == 0 ? msv : -msv;
  }
  return oracle_step(t, 0.0, pt, pv);
}
__device__ __noinline__ long long oracle_observe(long long t, double sigma_n, int sid) {
  double r_t = (t >= RUN.cfg.mkt_close) ? oracle_advance(RUN.cfg.mkt_close - 1) : oracle_advance(t);
  if (sigma_n == 0) return (long long)r_t;
  return (long long)rint(rng_normal(sid, r_t, sqrt(sigma_n)));
}

// ----------------------------------------------------------------------------------------------
// order book (util/OrderBook.py).  Per side s = BID / ASK the best <= BOOK_TOP resting orders are kept
// price-time sorted in shared memory (index 0 = best price, oldest first).  Every other resting order
// sits in the side's unsorted POOL, an HBM list in the event-list format: key.t = price rank (smaller =
// better), key.k2 = arrival number << 32 | order id, payload = price, qty, id, agent, history bucket,
// arrival number.  Invariant: every top entry precedes every pool entry in price-time order, and
// (pb_t, pb_k)[s] is a lower bound of the pool's keys.  Deep inserts are O(1) appends, deep cancels one
// coalesced scan over 16-byte keys, and the pool -> top refill reuses choose_limit (a price window).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ bool px_better_eq(int s, int p, int price) { return s == BID ? p >= price : p <= price; }
__device__ __forceinline__ long long pool_rank(int s, int price) { return s == ASK ? (long long)price : (long long)(0x7fffffff - price); }
__device__ __forceinline__ unsigned long long pool_k2(unsigned arr, int id) { return ((unsigned long long)arr << 32) | (unsigned)id; }

__device__ __noinline__ void top_remove(int s, int n, int idx) {                 // delete slot idx from top[s][0..n)
  __syncwarp();
#pragma unroll 1
  for (int base = idx; base < n - 1; base += 32) {
    int j = base + LANE; BookOrd t; int tq = 0, ta = 0, tl = 0;
    bool act = j < n - 1;
    if (act) { t = ST.top[s][j + 1]; tq = ST.tseq[s][j + 1]; ta = ST.tarr[s][j + 1]; tl = ST.tlog[s][j + 1]; }
    __syncwarp();
    if (act) { ST.top[s][j] = t; ST.tseq[s][j] = tq; ST.tarr[s][j] = ta; ST.tlog[s][j] = tl; }
    __syncwarp();
  }
}
__device__ __noinline__ void top_insert(int s, int n, int idx, int price, int qty, int id, int agent, int oseq, int arr, int olog) {
  __syncwarp();
  int cnt = n - idx;
#pragma unroll 1
  for (int done = 0; done < cnt; done += 32) {
    int j = n - 1 - done - LANE; BookOrd t; int tq = 0, ta = 0, tl = 0;
    bool act = j >= idx;
    if (act) { t = ST.top[s][j]; tq = ST.tseq[s][j]; ta = ST.tarr[s][j]; tl = ST.tlog[s][j]; }
    __syncwarp();
    if (act) { ST.top[s][j + 1] = t; ST.tseq[s][j + 1] = tq; ST.tarr[s][j + 1] = ta; ST.tlog[s][j + 1] = tl; }
    __syncwarp();
  }
  if (LANE == 0) {
    BookOrd o; o.price = price; o.qty = qty; o.id = id; o.agent = agent;
    ST.top[s][idx] = o; ST.tseq[s][idx] = oseq; ST.tarr[s][idx] = arr; ST.tlog[s][idx] = olog;
  }
  __syncwarp();
}
// count top entries whose price is better than or equal to `price` for this side
__device__ __forceinline__ int top_count_better_eq(int s, int n, int price) {
  __syncwarp();
  int cnt = 0;
  for (int j = LANE; j < n; j += 32) cnt += px_better_eq(s, ST.top[s][j].price, price) ? 1 : 0;
  return warp_sum_i(cnt);
}
__device__ __noinline__ void pool_append(int s, int price, int qty, int id, int agent, int hseq, int arr, int olog) {
  const int L = L_POOL + s;
  const long long kt = pool_rank(s, price); const unsigned long long kk = pool_k2((unsigned)arr, id);
  if (HOT.l_live[L] == 0 || key_less(kt, kk, HOT.pb_t[s], HOT.pb_k[s])) { HOT.pb_t[s] = kt; HOT.pb_k[s] = kk; }
  // slots never move (no compaction): a tombstoned slot is reused, so the slot handed back as a hint stays valid
  int i;
  const int nf = HOT.n_pfree[s];
  if (nf > 0) { i = RUN.pfree[s][nf - 1]; HOT.n_pfree[s] = nf - 1; }
  else {
    i = HOT.l_total[L];
    if (UNLIKELY(i >= RUN.lcap[L])) { HOT.status = ABIDES_EOVERFLOW; HOT.last_hint = 0; return; }
    HOT.l_total[L] = i + 1;
  }
  if (LANE == 0) { EvKey k; k.t = kt; k.k2 = kk; RUN.lk[L][i] = k; }
  else if (LANE == 1) *reinterpret_cast<int4*>(&RUN.lp[L][i].w[0]) = make_int4(price, qty, id, agent);
  else if (LANE == 2) *reinterpret_cast<int4*>(&RUN.lp[L][i].w[4]) = make_int4(hseq, arr, olog, 0);
  __syncwarp();
  HOT.l_live[L] += 1;
  HOT.last_hint = i + 1;
}
__device__ __forceinline__ void pool_free_slot(int s, int i) {                  // tombstone slot i (all lanes call)
  const int nf = HOT.n_pfree[s];
  __syncwarp();
  if (LANE == 0) { RUN.lk[L_POOL + s][i].t = LLONG_MAX; RUN.pfree[s][nf] = i; }
  __syncwarp();
  HOT.n_pfree[s] = nf + 1;
  HOT.l_live[L_POOL + s] -= 1;
}
// the shared-memory part ran empty: pull the best price levels of the pool back in
__device__ __noinline__ void side_refill(int s) {
  const int L = L_POOL + s;
  if (HOT.n_top[s] > 0 || HOT.l_live[L] == 0) return;
  const int cnt = choose_limit(L, L, TOP_FILL, TOP_FILL_MIN, TOP_FILL, LLONG_MAX - 1, 0ull);
  if (HOT.status) return;
  const long long pt = HOT.sel_t; const unsigned long long pk = HOT.sel_k;
  EvKey* keys = RUN.lk[L]; const EvPay* pay = RUN.lp[L];
  int* fr = RUN.pfree[s];
  const int n = HOT.l_total[L], nf = HOT.n_pfree[s];
  int nn = 0;
#pragma unroll 1
  for (int base = 0; base < n; base += 32) {
    int i = base + LANE;
    bool take = false; EvKey k;
    if (i < n) { k = keys[i]; take = key_less(k.t, k.k2, pt, pk); }
    unsigned m = __ballot_sync(FULL, take);
    if (m) {
      if (take) {
        int pos = nn + __popc(m & ((1u << LANE) - 1u));
        EvPay p = pay[i];
        BookOrd o; o.price = p.w[0]; o.qty = p.w[1]; o.id = p.w[2]; o.agent = p.w[3];
        ST.top[s][pos] = o; ST.tseq[s][pos] = p.w[4]; ST.tarr[s][pos] = p.w[5]; ST.tlog[s][pos] = p.w[6];
        keys[i].t = LLONG_MAX;
        fr[nf + pos] = i;
      }
      nn += __popc(m);
    }
  }
  HOT.n_pfree[s] = nf + nn;
  __syncwarp();
  // rank sort (<= TOP_FILL = 32 entries, lane i owns entry i): better price first, then arrival
  BookOrd e; int eq = 0, ea = 0, el = 0, rank = 0;
  if (LANE < nn) { e = ST.top[s][LANE]; eq = ST.tseq[s][LANE]; ea = ST.tarr[s][LANE]; el = ST.tlog[s][LANE]; }
  else { e.price = 0; e.qty = 0; e.id = 0; e.agent = 0; }
  const long long er = pool_rank(s, e.price);
#pragma unroll 1
  for (int j = 0; j < nn; j++) {
    long long orr = pool_rank(s, ST.top[s][j].price); unsigned oa = (unsigned)ST.tarr[s][j];
    rank += (orr < er || (orr == er && oa < (unsigned)ea)) ? 1 : 0;
  }
  __syncwarp();
  if (LANE < nn) { ST.top[s][rank] = e; ST.tseq[s][rank] = eq; ST.tarr[s][rank] = ea; ST.tlog[s][rank] = el; }
  HOT.n_top[s] = nn;
  HOT.l_live[L] -= cnt;
  HOT.pb_t[s] = pt; HOT.pb_k[s] = pk;
  __syncwarp();
}
__device__ __forceinline__ void side_pop_front(int s) {                          // remove the best entry
  top_remove(s, HOT.n_top[s], 0);
  HOT.n_top[s] -= 1;
  if (HOT.n_top[s] == 0) side_refill(s);
}
// sum of quantities resting in the pool at `price` (a level that runs past the shared part)
__device__ __noinline__ int pool_level_qty(int s, int price) {
  const int L = L_POOL + s;
  const EvKey* keys = RUN.lk[L]; const EvPay* pay = RUN.lp[L];
  const int n = HOT.l_total[L];
  const long long kt = pool_rank(s, price);
  int v = 0;
#pragma unroll 2
  for (int i = LANE; i < n; i += 32) if (keys[i].t == kt) v += pay[i].w[1];
  return warp_sum_i(v);
}
// getInsideBids(1)/getInsideAsks(1) (:378-399): price and total quantity of the best level
__device__ __noinline__ int2 side_inside(int s) {
  __syncwarp();
  const int ns = HOT.n_top[s];
  if (ns == 0) return make_int2(-1, 0);
  const int p = ST.top[s][0].price;
  int v = 0;
  for (int j = LANE; j < ns; j += 32) { BookOrd e = ST.top[s][j]; v += e.price == p ? e.qty : 0; }
  v = warp_sum_i(v);
  // the level may continue in the pool only if it runs to the end of the shared part
  if (ST.top[s][ns - 1].price == p && HOT.l_live[L_POOL + s] > 0 && HOT.pb_t[s] <= pool_rank(s, p)) v += pool_level_qty(s, p);
  return make_int2(p, v);
}
// OrderBook.enterOrder (:272-298): position = number of resting orders with better-or-equal price
__device__ __noinline__ void book_enter(int s, int price, int qty, int id, int agent, int olog) {
  int ns = HOT.n_top[s];
  const int oseq = HOT.hist_B;                           // history[0][order_id] = {...} (:60-64)
  const int arr = (int)HOT.book_arr;
  HOT.book_arr = (unsigned)arr + 1u;
  const int cs = top_count_better_eq(s, ns, price);
  bool to_top = cs < ns;
  if (!to_top && ns < BOOK_TOP)                          // behind all of top: only if it also precedes all of the pool
    to_top = HOT.l_live[L_POOL + s] == 0 ||
             key_less(pool_rank(s, price), pool_k2((unsigned)arr, id), HOT.pb_t[s], HOT.pb_k[s]);
  if (to_top) {
    if (UNLIKELY(ns == BOOK_TOP)) {                      // spill the worst shared entry to the pool
      BookOrd w = ST.top[s][ns - 1];
      pool_append(s, w.price, w.qty, w.id, w.agent, ST.tseq[s][ns - 1], ST.tarr[s][ns - 1], ST.tlog[s][ns - 1]);
      if (HOT.status) return;
      ns--;
    }
    top_insert(s, ns, cs, price, qty, id, agent, oseq, arr, olog);
    HOT.n_top[s] = ns + 1;
    HOT.last_hint = 0;
  } else {
    pool_append(s, price, qty, id, agent, oseq, arr, olog);
  }
  int tot = HOT.n_top[0] + HOT.n_top[1] + HOT.l_live[L_POOL] + HOT.l_live[L_POOL + 1];
  if (tot > HOT.book_peak_orders) HOT.book_peak_orders = tot;
}
// find (price, id) on a side -- the FIRST match in book order; returns 0 = absent, 1 = in the shared part,
// 2 = in the pool; index in HOT.find_idx
__device__ __noinline__ int side_find(int s, int price, int id) {
  __syncwarp();
  const int ns = HOT.n_top[s];
  for (int base = 0; base < ns; base += 32) {
    int j = base + LANE;
    bool hit = false;
    if (j < ns) { BookOrd e = ST.top[s][j]; hit = e.price == price && e.id == id; }
    unsigned m = __ballot_sync(FULL, hit);
    if (m) { HOT.find_idx = base + __ffs(m) - 1; return 1; }
  }
  const int L = L_POOL + s;
  if (HOT.l_live[L] == 0) return 0;
  const EvKey* keys = RUN.lk[L];
  const int n = HOT.l_total[L];
  const long long kt = pool_rank(s, price);
  unsigned long long best = ~0ull; int bi = -1;
#pragma unroll 4
  for (int i = LANE; i < n; i += 32) {
    EvKey k = keys[i];
    if (k.t == kt && (unsigned)k.k2 == (unsigned)id && k.k2 < best) { best = k.k2; bi = i; }
  }
  unsigned long long mb = warp_min_ull(best);
  if (mb == ~0ull) return 0;
  unsigned m = __ballot_sync(FULL, best == mb);
  HOT.find_idx = __shfl_sync(FULL, bi, __ffs(m) - 1);
  return 2;
}

// ---- OrderBook.history / get_transacted_volume (:401-471) ------------------------------------
// The reference keeps, per history bucket (one per trade-producing order), the (time, qty) transactions
// of the orders ENTERED in that bucket, and a volume query unrolls only a window of "recent" buckets.
// Equivalent device form: every transaction is one TxnRow tagged with the bucket of its record; a query
// scans the rows logged since the previous query, keeps those whose bucket lies in the window, drops
// duplicate (time, qty) rows of that batch and appends the rest to the unrolled ring it then sums over.
__device__ __forceinline__ void txn_log_row(long long t, int q, int seq) {
  long long i = HOT.log_len;
  __syncwarp();
  if (LANE == 0 && RUN.txn_log) { TxnRow r; r.t = t; r.q = q; r.seq = seq; RUN.txn_log[i & (TXN_CAP - 1)] = r; }
  __syncwarp();
  HOT.log_len = i + 1;
}
__device__ __noinline__ long long book_transacted_volume(long long now, long long lookback) {
  Hot& h = HOT;
  const int SH = RUN.cfg.stream_history;
  int B = h.hist_B;
  int n = (B + 1 < SH + 1) ? B + 1 : SH + 1, p = h.tv_prev_len;
  int lo = 1, hi = 0;
  if (p == 0) { lo = B - (n - 1); hi = B; h.tv_prev_len = n; }
  else if (p != n) { int cnt = n - p - 1; if (cnt > 0) { lo = B - cnt + 1; hi = B; } h.tv_prev_len = n; }
  long long start = h.log_mark, end = h.log_len;
  if (end - start > TXN_CAP) { h.status = ABIDES_EOVERFLOW; return 0; }
  const TxnRow* lg = RUN.txn_log;
  TxnRow* un = RUN.unrolled;
  long long ulen = h.unr_len;
  __syncwarp();
  if (hi >= lo) {
    for (long long base = start; base < end; base += 32) {
      long long i = base + LANE;
      bool keep = false; TxnRow r;
      if (i < end) {
        r = lg[i & (TXN_CAP - 1)];
        keep = r.seq >= lo && r.seq <= hi;
        if (keep) {                                     // drop_duplicates over this batch (:448-451)
          for (long long j = i - 1; j >= start; j--) {
            TxnRow e = lg[j & (TXN_CAP - 1)];
            if (e.t != r.t) break;
            if (e.q == r.q && e.seq >= lo && e.seq <= hi) { keep = false; break; }
          }
        }
      }
      unsigned m = __ballot_sync(FULL, keep);
      if (keep) un[(ulen + __popc(m & ((1u << LANE) - 1u))) & (UNR_CAP - 1)] = r;
      ulen += __popc(m);
    }
  }
  __syncwarp();
  h.unr_len = ulen;
  h.log_mark = end;
  __syncwarp();
  // sum quantities of unrolled rows with time >= now - lookback (rows are in time order)
  long long thr = now - lookback, sum = 0;
  long long ul = h.unr_len, floor_i = ul > UNR_CAP ? ul - UNR_CAP : 0;
  for (long long base = ul; base > floor_i; base -= 32) {
    long long i = base - 1 - LANE;
    bool in = false; TxnRow r;
    if (i >= floor_i) { r = un[i & (UNR_CAP - 1)]; in = r.t >= thr; }
    sum += in ? r.q : 0;
    if (__any_sync(FULL, i >= floor_i && !in)) { floor_i = -1; break; }
  }
  if (floor_i > 0 && ul > UNR_CAP) {                     // the whole ring lies inside the window: rows were lost
    h.status = ABIDES_EOVERFLOW;
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
  return sum;
}

// outbound notification of the book: to the event queue (engine) or to the replay tape (K2 harness)
__device__ __noinline__ void replay_emit(int rcpt, int code, bool is_buy, int id, int price, int qty, int fill) {
  Run& r = RUN;
  int k = r.rp_counts[r.rp_op];
  __syncwarp();
  if (LANE == 0) {
    if (k < r.rp_max) {
      abides_book_event e;
      e.op_index = r.rp_op - r.rp_base; e.kind = code; e.agent = rcpt; e.is_buy = is_buy ? 1 : 0;
      e.order_id = id; e.price = price; e.qty = qty; e.fill_price = fill;
      r.rp_events[(size_t)r.rp_op * r.rp_max + k] = e;
    }
    r.rp_counts[r.rp_op] = k + 1;
  }
  __syncwarp();
}
template <bool RP>
__device__ __forceinline__ void book_send(int rcpt, int code, bool is_buy, int id, int price, int qty, int fill, int w1 = 0) {
  if (RP) replay_emit(rcpt, code, is_buy, id, price, qty, fill);
  else k_send(0, rcpt, code == ABIDES_MSG_ORDER_MODIFIED ? 0 : RUN.cfg.exch_pipeline_delay, order_w0(code, is_buy), w1,
              id, price, qty, fill);        // ORDER_ACCEPTED/CANCELLED/EXECUTED get +pipeline_delay (ExchangeAgent.py:404-407)
}

// ---- order log: OrderBook.history as the HBL agents see it (QUERY_ORDER_STREAM, ExchangeAgent.py:210-225) --------
// One row per limit order the book accepted for processing (history[0][order_id] = {...}, :60-64), tagged with its
// bucket; bit 1 is set when a transaction is appended to that order's record (:245-253).  Kept only when the batch
// holds HBL agents (RUN.olog != NULL).
__device__ __forceinline__ int olog_append(int bucket, int price, bool is_buy) {
  if (RUN.olog == nullptr) return -1;
  const long long i = HOT.olog_len;
  __syncwarp();
  if (LANE == 0) { OLogRow r; r.bucket = bucket; r.price = price; r.flags = is_buy ? 1 : 0; r.pad = 0; RUN.olog[i & (OLOG_CAP - 1)] = r; }
  __syncwarp();
  HOT.olog_len = i + 1;
  return (int)(i & 0x7fffffff);
}
__device__ __forceinline__ void olog_mark(int row) {
  if (RUN.olog == nullptr || row < 0) return;
  if ((HOT.olog_len & 0x7fffffff) - row > OLOG_CAP) return;          // the row has left the ring (older than any window)
  __syncwarp();
  if (LANE == 0) RUN.olog[row & (OLOG_CAP - 1)].flags |= 2;
  __syncwarp();
}

// OrderBook.handleLimitOrder (:46-158)
template <bool RP>
__device__ __noinline__ void book_handle_limit(int agent, int id, int price, int qty, bool is_buy) {
  Hot& h = HOT;
  if (qty <= 0) return;
  long long ex_qty = 0, ex_pq = 0; int executed = 0;
  const int os = is_buy ? ASK : BID;
  const int my_log = RP ? -1 : olog_append(h.hist_B, price, is_buy);
  for (;;) {
    __syncwarp();
    bool match = false; BookOrd top;
    if (h.n_top[os] > 0) { top = ST.top[os][0]; match = is_buy ? price >= top.price : price <= top.price; }
    if (match) {                                               // executeOrder (:190-256)
      int fill, tseq = ST.tseq[os][0];
      const int rest_log = ST.tlog[os][0];
      if (qty >= top.qty) { fill = top.qty; side_pop_front(os); }
      else { fill = qty; __syncwarp(); if (LANE == 0) ST.top[os][0].qty = top.qty - qty; __syncwarp(); }
      // history (:245-253): the aggressor logs its REMAINING quantity, the resting order the fill
      txn_log_row(h.now, qty, h.hist_B);
      if (h.hist_B - tseq <= RUN.cfg.stream_history) { txn_log_row(h.now, fill, tseq); if (!RP) olog_mark(rest_log); }
      if (!RP) olog_mark(my_log);
      qty -= fill;
      book_send<RP>(agent, ABIDES_MSG_ORDER_EXECUTED, is_buy, id, price, fill, top.price);
      book_send<RP>(top.agent, ABIDES_MSG_ORDER_EXECUTED, !is_buy, top.id, top.price, fill, top.price);
      ex_qty += fill; ex_pq += (long long)top.price * fill; executed++;
      h.n_fills++; h.volume += fill;
      if (qty <= 0) break;
    } else {
      book_enter(is_buy ? BID : ASK, price, qty, id, agent, my_log);
      book_send<RP>(agent, ABIDES_MSG_ORDER_ACCEPTED, is_buy, id, price, qty, -1, HOT.last_hint);   // w1: pool slot hint
      break;
    }
  }
  if (executed) {
    h.last_trade = (int)(long long)rint((double)ex_pq / (double)ex_qty);     // int(round(.)) (:130)
    h.has_last_trade = 1;
    h.hist_B++;                                                                // history.insert(0, {}) (:137)
  }
}
// OrderBook.cancelOrder (:300-348)
template <bool RP>
__device__ __noinline__ void book_cancel(int agent, int id, int price, bool is_buy, int hint) {
  const int s = is_buy ? BID : ASK;
  int where = 0, idx = 0;
  if (hint > 0 && hint <= HOT.l_total[L_POOL + s]) {       // the slot the exchange reported when it accepted the order
    const EvKey k = RUN.lk[L_POOL + s][hint - 1];
    if (k.t == pool_rank(s, price) && (unsigned)k.k2 == (unsigned)id) { where = 2; idx = hint - 1; }
  }
  if (where == 0) {
    where = side_find(s, price, id);
    if (where == 0) return;
    idx = HOT.find_idx;
  }
  int oid, oprice, oqty;
  if (where == 1) {
    BookOrd o = ST.top[s][idx];
    oid = o.id; oprice = o.price; oqty = o.qty;
    top_remove(s, HOT.n_top[s], idx);
    HOT.n_top[s] -= 1;
    if (HOT.n_top[s] == 0) side_refill(s);
  } else {
    const int L = L_POOL + s;
    const int4 p = *reinterpret_cast<const int4*>(&RUN.lp[L][idx].w[0]);
    oprice = p.x; oqty = p.y; oid = p.z;
    pool_free_slot(s, idx);
  }
  book_send<RP>(agent, ABIDES_MSG_ORDER_CANCELLED, is_buy, oid, oprice, oqty, -1);
}

// OrderBook.handleMarketOrder (:160-188): one fresh limit order per opposite level until the quantity is
// covered.  Each child consumes exactly the level it was sized from, so walking the live book level by
// level is equivalent to the reference's snapshot walk; an uncovered remainder is dropped.
template <bool RP>
__device__ __noinline__ void book_handle_market(int agent, int qty, bool is_buy) {
  if (qty <= 0) return;
  int rem = qty;
  for (;;) {
    int2 in = side_inside(is_buy ? ASK : BID);
    int p = in.x, v = in.y;
    if (p < 0) break;
    int q = rem <= v ? rem : v;
    int id = (int)HOT.next_order_id++;
    book_handle_limit<RP>(agent, id, p, q, is_buy);
    if (rem <= v) break;
    rem -= v;
  }
}
// OrderBook.modifyOrder (:350-373), including the reference's index-0 overwrite (:359): the modified order
// replaces the FIRST order of its price level (which keeps its queue position), not the matched one.
template <bool RP>
__device__ __noinline__ void book_modify(int agent, int id, int price, int new_qty, bool is_buy) {
  const int s = is_buy ? BID : ASK;
  const int L = L_POOL + s;
  int where = side_find(s, price, id);
  if (where == 0) return;
  int idx = HOT.find_idx;
  const int oseq = where == 1 ? ST.tseq[s][idx] : RUN.lp[L][idx].w[4];
  // first entry of that price level
  const int ns = HOT.n_top[s];
  int first = -1;
  __syncwarp();
  for (int base = 0; base < ns && first < 0; base += 32) {
    int j = base + LANE;
    unsigned m = __ballot_sync(FULL, j < ns && ST.top[s][j].price == price);
    if (m) first = base + __ffs(m) - 1;
  }
  if (first >= 0) {
    __syncwarp();
    if (LANE == 0) {
      BookOrd no; no.price = price; no.qty = new_qty; no.id = id; no.agent = agent;
      ST.top[s][first] = no; ST.tseq[s][first] = oseq;
    }
    __syncwarp();
  } else {
    // the whole level sits in the pool: its first order is the one with the smallest arrival number
    EvKey* keys = RUN.lk[L];
    const int n = HOT.l_total[L];
    const long long kt = pool_rank(s, price);
    unsigned long long best = ~0ull; int bi = -1;
    for (int i = LANE; i < n; i += 32) {
      EvKey k = keys[i];
      if (k.t == kt && k.k2 < best) { best = k.k2; bi = i; }
    }
    unsigned long long mb = warp_min_ull(best);
    unsigned m = __ballot_sync(FULL, best == mb);
    bi = __shfl_sync(FULL, bi, __ffs(m) - 1);
    __syncwarp();
    if (LANE == 0) {
      keys[bi].k2 = (mb & 0xffffffff00000000ull) | (unsigned)id;
      EvPay p = RUN.lp[L][bi];
      p.w[0] = price; p.w[1] = new_qty; p.w[2] = id; p.w[3] = agent; p.w[4] = oseq;
      RUN.lp[L][bi] = p;
    }
    __syncwarp();
  }
  if (HOT.hist_B - oseq <= RUN.cfg.stream_history)         // one notification per history bucket holding the id
    book_send<RP>(agent, ABIDES_MSG_ORDER_MODIFIED, is_buy, id, price, new_qty, -1);
}
// number of distinct prices on a side (replay harness only)
__device__ __noinline__ int side_levels(int s) {
  const int L = L_POOL + s;
  const int ns = HOT.n_top[s], n = HOT.l_total[L];
  const EvKey* keys = RUN.lk[L];
  int c = 0;
  __syncwarp();
  for (int j = LANE; j < ns; j += 32) c += (j == 0 || ST.top[s][j].price != ST.top[s][j - 1].price) ? 1 : 0;
  const long long worst = ns > 0 ? pool_rank(s, ST.top[s][ns - 1].price) : -1;
  for (int i = LANE; i < n; i += 32) {
    EvKey k = keys[i];
    if (k.t == LLONG_MAX || k.t == worst) continue;
    bool firstOfLevel = true;                              // no other pool entry of this price arrived earlier
    for (int j = 0; j < n; j++) { EvKey o = keys[j]; if (o.t == k.t && o.k2 < k.k2) { firstOfLevel = false; break; } }
    c += firstOfLevel ? 1 : 0;
  }
  return warp_sum_i(c);
}

// ----------------------------------------------------------------------------------------------
// TradingAgent (agent/TradingAgent.py)
// ----------------------------------------------------------------------------------------------
struct Msg { int w0, w1, w2, w3, w4, w5; };

__device__ __forceinline__ int ag_state(const AgentRec& A) { return (A.flags & F_STATE_MASK) >> F_STATE_SHIFT; }
__device__ __forceinline__ void ag_set_state(AgentRec& A, int s) { A.flags = (A.flags & ~F_STATE_MASK) | ((unsigned)s << F_STATE_SHIFT); }

__device__ __forceinline__ void agent_send(int a, int w0, int w2, int w3, int w4, int w5) { k_send(a, 0, 0, w0, a, w2, w3, w4, w5); }
__device__ __forceinline__ void ta_query_spread(int a) { agent_send(a, ABIDES_MSG_QUERY_SPREAD, 1, 0, 0, 0); }

__device__ __noinline__ void ord_reserve(int initial_cap) {
  AgentRec& A = ST.ag; Hot& h = HOT;
  if (A.ord_cnt < A.ord_cap) return;
  int ncap = A.ord_cap ? A.ord_cap * 2 : initial_cap;
  if (ncap > 65535 || h.arena_top + ncap > c_db.arena_cap) { h.status = ABIDES_EOVERFLOW; return; }
  int noff = h.arena_top;
  h.arena_top += ncap;
  __syncwarp();
  for (int j = LANE; j < A.ord_cnt; j += 32) RUN.arena[noff + j] = RUN.arena[A.ord_off + j];
  __syncwarp();
  A.ord_off = noff; A.ord_cap = (unsigned short)ncap;
}
// TradingAgent.placeLimitOrder (:295-330), ignore_risk=True
__device__ __noinline__ void ta_place_limit(int a, long long qty, bool is_buy, long long price) {
  AgentRec& A = ST.ag; Hot& h = HOT;
  const abides_agent_type& ty = c_db.types[A.type_idx];
  int id = (int)h.next_order_id++;
  if (qty <= 0) return;
  OpenOrd oo; oo.key = id; oo.id = id; oo.price = (int)price; oo.sqty = is_buy ? (int)qty : -(int)qty;
  if (UNLIKELY(id == 0)) oo.id = (int)h.next_order_id++;            // deepcopy of order id 0 takes a fresh id (Order.py:29)
  int init_cap = ty.klass == ABIDES_ADAPTIVE_MM ? (int)(2 * ty.num_ticks + 8)
               : ty.klass == ABIDES_MARKET_MAKER ? 8 * ty.mm_num_levels + 8 : (ty.klass == ABIDES_MOMENTUM ? 16 : 4);
  ord_reserve(init_cap);
  if (UNLIKELY(h.status != 0)) return;
  __syncwarp();
  if (LANE == 0) RUN.arena[A.ord_off + A.ord_cnt] = oo;
  __syncwarp();
  A.ord_cnt++;
  agent_send(a, order_w0(ABIDES_MSG_LIMIT_ORDER, is_buy), id, (int)price, (int)qty, -1);
  if (UNLIKELY(ty.log_orders && id == 0)) h.next_order_id++;
}
__device__ __noinline__ void ta_cancel_all(int a) {
  AgentRec& A = ST.ag;
  const int klass = c_db.types[A.type_idx].klass;
  const bool hinted = klass == ABIDES_ZI || klass == ABIDES_VALUE || klass == ABIDES_HBL;
  __syncwarp();
  for (int i = 0; i < A.ord_cnt; i++) {
    OpenOrd o = RUN.arena[A.ord_off + i];
    int hint = -1;                                       // w5 of a CANCEL_ORDER: pool slot hint (not part of the reference message)
    if (hinted && o.id == A.u.est.hint_id && A.u.est.hint > 0) hint = A.u.est.hint;
    agent_send(a, order_w0(ABIDES_MSG_CANCEL_ORDER, o.sqty > 0), o.id, o.price, o.sqty < 0 ? -o.sqty : o.sqty, hint);
  }
}
__device__ __noinline__ int ord_find(int key) {
  const AgentRec& A = ST.ag;
  const int n = A.ord_cnt;
  const OpenOrd* s = RUN.arena + A.ord_off;
  __syncwarp();
  for (int base = 0; base < n; base += 32) {
    int j = base + LANE;
    bool hit = j < n && s[j].key == key;
    unsigned m = __ballot_sync(FULL, hit);
    if (m) return base + __ffs(m) - 1;
  }
  return -1;
}
__device__ __noinline__ void ord_remove(int idx) {
  AgentRec& A = ST.ag;
  OpenOrd* s = RUN.arena + A.ord_off;
  int n = A.ord_cnt;
  __syncwarp();
  for (int base = idx; base < n - 1; base += 32) {
    int j = base + LANE; OpenOrd t;
    bool act = j < n - 1;
    if (act) t = s[j + 1];
    __syncwarp();
    if (act) s[j] = t;
    __syncwarp();
  }
  A.ord_cnt = (unsigned short)(n - 1);
}
// TradingAgent.wakeup (:152-168); returns can_trade
__device__ __forceinline__ bool ta_wakeup(int a) {
  AgentRec& A = ST.ag;
  A.cur_time = HOT.now;
  if (UNLIKELY(!(A.flags & F_HAS_OPEN))) {
    agent_send(a, ABIDES_MSG_WHEN_MKT_OPEN, 0, 0, 0, 0);
    agent_send(a, ABIDES_MSG_WHEN_MKT_CLOSE, 0, 0, 0, 0);
  }
  return (A.flags & F_HAS_OPEN) && (A.flags & F_HAS_CLOSE) && !(A.flags & F_MKT_CLOSED);
}
// TradingAgent.receiveMessage (:181-264)
__device__ __forceinline__ void ta_receive(int a, const Msg& m, int klass, const abides_agent_type& ty) {
  AgentRec& A = ST.ag;
  A.cur_time = HOT.now;
  bool had = (A.flags & F_HAS_OPEN) && (A.flags & F_HAS_CLOSE);
  int code = pay_code(m.w0) & ~ABIDES_MSG_REPLY;
  switch (code) {
    case ABIDES_MSG_WHEN_MKT_OPEN: A.flags |= F_HAS_OPEN; break;
    case ABIDES_MSG_WHEN_MKT_CLOSE: A.flags |= F_HAS_CLOSE; break;
    case ABIDES_MSG_ORDER_EXECUTED: {                         // orderExecuted (:390-426)
      long long q = pay_is_buy(m.w0) ? m.w4 : -(long long)m.w4;
      A.holdings += q;
      A.cash -= q * (long long)m.w5;
      int idx = ord_find(m.w2);
      if (idx >= 0) {
        OpenOrd o = RUN.arena[A.ord_off + idx];
        int oq = o.sqty < 0 ? -o.sqty : o.sqty;
        if (m.w4 >= oq) ord_remove(idx);
        else { __syncwarp(); if (LANE == 0) RUN.arena[A.ord_off + idx].sqty = o.sqty > 0 ? oq - m.w4 : -(oq - m.w4); __syncwarp(); }
      }
      break;
    }
    case ABIDES_MSG_ORDER_ACCEPTED:                           // (not in the reference) remember where the order rests
      if (klass == ABIDES_ZI || klass == ABIDES_VALUE || klass == ABIDES_HBL) { A.u.est.hint_id = m.w2; A.u.est.hint = m.w1; }
      break;
    case ABIDES_MSG_ORDER_CANCELLED: {
      int idx = ord_find(m.w2);
      if (idx >= 0) ord_remove(idx);
      break;
    }
    case ABIDES_MSG_MKT_CLOSED: A.flags |= F_MKT_CLOSED; break;
    case ABIDES_MSG_QUERY_LAST_TRADE:
      if (pay_closed(m.w0)) A.flags |= F_MKT_CLOSED;
      A.last_trade = m.w5; A.flags |= F_HAS_LAST_TRADE;
      if (A.flags & F_MKT_CLOSED) A.flags |= F_HAS_DAILY_CLOSE;
      break;
    case ABIDES_MSG_QUERY_SPREAD:                             // querySpread (:482-501)
      if (pay_closed(m.w0)) A.flags |= F_MKT_CLOSED;
      A.last_trade = m.w5; A.flags |= F_HAS_LAST_TRADE;
      if (A.flags & F_MKT_CLOSED) A.flags |= F_HAS_DAILY_CLOSE;
      A.bid = m.w1; A.bid_vol = m.w2; A.ask = m.w3; A.ask_vol = m.w4; A.flags |= F_KNOWN_BOOK;
      break;
    case ABIDES_MSG_QUERY_ORDER_STREAM:                       // queryOrderStream (:514-520): which buckets we were handed
      if (pay_closed(m.w0)) A.flags |= F_MKT_CLOSED;
      if (klass == ABIDES_HBL) { A.u.est.stream_n = m.w1; A.u.est.stream_B = m.w2; }
      break;
    case ABIDES_MSG_QUERY_TRANSACTED_VOLUME:
      if (pay_closed(m.w0)) A.flags |= F_MKT_CLOSED;
      if (klass == ABIDES_ADAPTIVE_MM)
        A.u.mm.transacted_volume = ((long long)(unsigned)m.w2) | ((long long)m.w3 << 32);
      else if (klass == ABIDES_POV_EXEC) {
        A.u.pov.transacted_volume = ((long long)(unsigned)m.w2) | ((long long)m.w3 << 32);
        A.flags |= F_HAS_TV;
      }
      break;
    default: break;
  }
  bool have = (A.flags & F_HAS_OPEN) && (A.flags & F_HAS_CLOSE);
  if (UNLIKELY(have && !had)) {
    long long off;
    if (klass == ABIDES_ZI || klass == ABIDES_VALUE || klass == ABIDES_NOISE || klass == ABIDES_HBL) off = rng_randint(SID_AGENT, 0, 100);
    else off = ty.wake_up_freq;
    k_wakeup(a, RUN.cfg.mkt_open + off);
  }
}

// ZI / Value estimator (ZeroIntelligenceAgent.py:163-247, ValueAgent.py:123-186)
__device__ __noinline__ long long estimator_update(long long obs_t, const abides_agent_type& ty, abm_dd log1mk, double omp2) {
  AgentRec& A = ST.ag;
  if (!(A.flags & F_HAS_PREV_WAKE)) { A.u.est.prev_wake = RUN.cfg.mkt_open; A.flags |= F_HAS_PREV_WAKE; }
  double base = 1 - ty.kappa;
  double delta = (double)(A.cur_time - A.u.est.prev_wake);
  double pw = abm_pow_with_log(log1mk, base, delta);
  double r_tprime = (1 - pw) * ty.r_bar;
  r_tprime += pw * A.u.est.r_t;
  double pw2 = abm_pow_with_log(log1mk, base, 2 * delta);
  double sigma_tprime = pw2 * A.u.est.sigma_t;
  sigma_tprime += fdiv(1 - pw2, omp2) * ty.sigma_s;
  double r_t = fdiv(ty.sigma_n, ty.sigma_n + sigma_tprime) * r_tprime;
  r_t += fdiv(sigma_tprime, ty.sigma_n + sigma_tprime) * (double)obs_t;
  A.u.est.r_t = r_t;
  A.u.est.sigma_t = fdiv(ty.sigma_n * A.u.est.sigma_t, ty.sigma_n + A.u.est.sigma_t);
  double d2 = (double)(RUN.cfg.mkt_close - A.cur_time);
  if (!(d2 > 0)) d2 = 0;
  double pw3 = abm_pow_with_log(log1mk, base, d2);
  double r_T = (1 - pw3) * ty.r_bar;
  r_T += pw3 * A.u.est.r_t;
  A.u.est.prev_wake = A.cur_time;
  return (long long)rint(r_T);
}
__device__ __forceinline__ long long round_hundreds(long long hld) {     // int(round(h, -2) / 100)
  long long q = hld / 100, r = hld % 100;
  if (r < 0) { q -= 1; r += 100; }
  if (r > 50 || (r == 50 && (q & 1))) q += 1;
  return q;
}
// ZeroIntelligenceAgent.placeOrder (:249-281)
// ZeroIntelligenceAgent.updateEstimates (:163-247): observation, side, total unit valuation (side in HOT.find_idx)
__device__ __noinline__ long long zi_update_estimates(int a, const abides_agent_type& ty) {
  AgentRec& A = ST.ag;
  long long obs = oracle_observe(A.cur_time, ty.sigma_n, SID_AGENT);
  long long q = A.holdings / 100;
  bool buy;
  if (UNLIKELY(q >= ty.q_max)) buy = false;
  else if (UNLIKELY(q <= -ty.q_max)) buy = true;
  else buy = rng_randint(SID_AGENT, 0, 2) != 0;
  long long r_T = estimator_update(obs, ty, c_db.type_log1mk[A.type_idx], c_db.type_omp2[A.type_idx]);
  q += ty.q_max - 1;
  long long row = RUN.aoff + a;
  long long theta = c_db.theta[RUN.cfg.theta_offset + c_db.agent_theta[row] + (buy ? q + 1 : q)];
  HOT.find_idx = buy ? 1 : 0;
  return r_T + theta;
}
__device__ __noinline__ void zi_place_order(int a, const abides_agent_type& ty) {
  AgentRec& A = ST.ag;
  const long long v = zi_update_estimates(a, ty);
  const bool buy = HOT.find_idx != 0;
  long long R = rng_randint(SID_AGENT, ty.R_min, ty.R_max + 1);
  long long p = buy ? v - R : v + R;
  if (buy && A.ask_vol > 0) {
    long long R_ask = v - A.ask;
    if ((double)R_ask >= ty.eta * (double)R) p = A.ask;
  } else if (!buy && A.bid_vol > 0) {
    long long R_bid = A.bid - v;
    if ((double)R_bid >= ty.eta * (double)R) p = A.bid;
  }
  ta_place_limit(a, 100, buy, p);
}
__device__ __forceinline__ int warp_incl_scan(int v) {         // inclusive prefix sum over lanes 0..31
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int x = __shfl_up_sync(FULL, v, o); if (LANE >= o) v += x; }
  return v;
}
// HeuristicBeliefLearningAgent.placeOrder (:42-153).  The buckets it was handed are [stream_B - stream_n, stream_B):
// it holds the exchange's history dicts themselves, so "has transactions" is read as of NOW from the order log.
// nd[:, 0..3] = sa, sb, ua, ub per price; cumulative sums; Pr = num / denom; Es = Pr * surplus; first arg-max.
__device__ __noinline__ void hbl_place_order(int a, const abides_agent_type& ty) {
  AgentRec& A = ST.ag;
  if (A.u.est.stream_n < ty.hbl_L) { zi_place_order(a, ty); return; }      // not enough history: exactly ZI
  const long long v = zi_update_estimates(a, ty);
  const bool buy = HOT.find_idx != 0;
  const int lo_b = A.u.est.stream_B - A.u.est.stream_n, hi_b = A.u.est.stream_B - 1;
  const OLogRow* lg = RUN.olog;
  const long long end = HOT.olog_len;
  // rows are in bucket order: walk back (newest first, lane 0 = newest) to the first row of bucket lo_b
  long long start = -1;
  int low_p = INT_MAX, high_p = 0;
#pragma unroll 1
  for (long long base = end; start < 0; base -= 32) {
    const long long i = base - 1 - LANE;
    const bool lost = i >= 0 && end - i > OLOG_CAP;          // the window reaches past the ring
    bool stop = i < 0 || lost;
    if (!stop) {
      OLogRow r = lg[i & (OLOG_CAP - 1)];
      stop = r.bucket < lo_b;
      if (!stop && r.bucket <= hi_b) { low_p = min(low_p, r.price); high_p = max(high_p, r.price); }
    }
    const unsigned m = __ballot_sync(FULL, stop), ml = __ballot_sync(FULL, lost);
    if (m) {
      const int k = __ffs(m) - 1;
      if ((ml >> k) & 1u) { HOT.status = ABIDES_EOVERFLOW; return; }
      start = base - k;                                        // first row that is not older than the window
    }
  }
  low_p = __reduce_min_sync(FULL, low_p);
  high_p = __reduce_max_sync(FULL, high_p);
  const int R = high_p - low_p + 1;
  if (R <= 0) { HOT.status = ABIDES_ESTATE; return; }                       // np.zeros((negative, 8)) raises in the reference
  if (R > HBL_RANGE_CAP) { HOT.status = ABIDES_EOVERFLOW; return; }
  int* c0 = RUN.hbl_cnt;                       // sa: successful asks   (rows of HBL_RANGE_CAP ints)
  int* c1 = c0 + HBL_RANGE_CAP;                // sb: successful bids
  int* c2 = c1 + HBL_RANGE_CAP;                // ua: unsuccessful asks
  int* c3 = c2 + HBL_RANGE_CAP;                // ub: unsuccessful bids
  for (int i = LANE; i < R; i += 32) { c0[i] = 0; c1[i] = 0; c2[i] = 0; c3[i] = 0; }
  __syncwarp();
  for (long long i = start + LANE; i < end; i += 32) {
    OLogRow r = lg[i & (OLOG_CAP - 1)];
    if (r.bucket < lo_b || r.bucket > hi_b) continue;
    int* col = (r.flags & 1) ? ((r.flags & 2) ? c1 : c3) : ((r.flags & 2) ? c0 : c2);
    atomicAdd(&col[r.price - low_p], 1);
  }
  __syncwarp();
  double best_es = -INFINITY; int best_i = INT_MAX;
  const int nchunk = (R + 31) / 32;
  if (buy) {
    // ub becomes a suffix sum (in place), then one ascending pass with the three prefix sums
    int carry = 0;
    for (int ch = nchunk - 1; ch >= 0; ch--) {
      int i = ch * 32 + (31 - LANE);
      int x = i < R ? c3[i] : 0;
      int sc = warp_incl_scan(x) + carry;
      if (i < R) c3[i] = sc;
      carry = __shfl_sync(FULL, sc, 31);
    }
    __syncwarp();
    int k0 = 0, k1 = 0, k2 = 0;
    for (int ch = 0; ch < nchunk; ch++) {
      int i = ch * 32 + LANE;
      bool in = i < R;
      int s0 = warp_incl_scan(in ? c0[i] : 0) + k0, s1 = warp_incl_scan(in ? c1[i] : 0) + k1, s2 = warp_incl_scan(in ? c2[i] : 0) + k2;
      k0 = __shfl_sync(FULL, s0, 31); k1 = __shfl_sync(FULL, s1, 31); k2 = __shfl_sync(FULL, s2, 31);
      if (in) {
        double num = ((double)s0 + (double)s1) + (double)s2, den = num + (double)c3[i];
        double pr = den == 0 ? 0.0 : fdiv(num, den);
        double es = pr * (double)(v - ((long long)low_p + i));
        if (es > best_es) { best_es = es; best_i = i; }
      }
    }
  } else {
    // ua becomes a prefix sum (in place), then one descending pass with the three suffix sums
    int carry = 0;
    for (int ch = 0; ch < nchunk; ch++) {
      int i = ch * 32 + LANE;
      int x = i < R ? c2[i] : 0;
      int sc = warp_incl_scan(x) + carry;
      if (i < R) c2[i] = sc;
      carry = __shfl_sync(FULL, sc, 31);
    }
    __syncwarp();
    int k0 = 0, k1 = 0, k3 = 0;
    for (int ch = nchunk - 1; ch >= 0; ch--) {
      int i = ch * 32 + (31 - LANE);
      bool in = i < R;
      int s0 = warp_incl_scan(in ? c0[i] : 0) + k0, s1 = warp_incl_scan(in ? c1[i] : 0) + k1, s3 = warp_incl_scan(in ? c3[i] : 0) + k3;
      k0 = __shfl_sync(FULL, s0, 31); k1 = __shfl_sync(FULL, s1, 31); k3 = __shfl_sync(FULL, s3, 31);
      if (in) {
        double num = ((double)s0 + (double)s1) + (double)s3, den = (((double)s0 + (double)s1) + (double)c2[i]) + (double)s3;
        double pr = den == 0 ? 0.0 : fdiv(num, den);
        double es = pr * (double)(((long long)low_p + i) - v);
        if (es >= best_es) { best_es = es; best_i = i; }          // descending walk: a later equal value has the smaller index
      }
    }
  }
  // np.argmax: the first (lowest-index) maximum over all lanes
#pragma unroll 1
  for (int o = 16; o; o >>= 1) {
    double oe = __shfl_xor_sync(FULL, best_es, o); int oi = __shfl_xor_sync(FULL, best_i, o);
    if (oe > best_es || (oe == best_es && oi < best_i)) { best_es = oe; best_i = oi; }
  }
  if (best_es > 0) ta_place_limit(a, 100, buy, (long long)low_p + best_i);
}
// ValueAgent.placeOrder (:188-220)
__device__ __noinline__ void value_place_order(int a, const abides_agent_type& ty) {
  AgentRec& A = ST.ag;
  long long obs = oracle_observe(A.cur_time, ty.sigma_n, SID_AGENT);
  long long r_T = estimator_update(obs, ty, c_db.type_log1mk[A.type_idx], c_db.type_omp2[A.type_idx]);
  bool buy; long long p;
  const int g = ABIDES_STREAM_GLOBAL;
  if (A.bid > 0 && A.ask > 0) {
    long long mid = (long long)((double)((long long)A.ask + A.bid) / 2);
    long long spread = A.ask > A.bid ? A.ask - A.bid : A.bid - A.ask;
    long long adj;
    if (rng_double(g) < ty.percent_aggr) adj = 0;
    else adj = rng_randint(g, 0, ty.depth_spread * spread);
    if (r_T < mid) { buy = false; p = A.bid + adj; }
    else { buy = true; p = A.ask - adj; }
  } else {
    buy = rng_randint(g, 0, 2) != 0;
    p = r_T;
  }
  ta_place_limit(a, A.size, buy, p);
}
// NoiseAgent.placeOrder (:113-123)
__device__ __noinline__ void noise_place_order(int a) {
  AgentRec& A = ST.ag;
  bool buy = rng_randint(ABIDES_STREAM_GLOBAL, 0, 2) != 0;
  if (buy && A.ask > 0) ta_place_limit(a, A.size, true, A.ask);
  else if (!buy && A.bid > 0) ta_place_limit(a, A.size, false, A.bid);
}
// MomentumAgent.placeOrders (:56-66): running prefix sum + ring of the last MOM_RING prefix sums
__device__ __noinline__ void momentum_place_orders(int a) {
  AgentRec& A = ST.ag;
  if (!(A.bid > 0 && A.ask > 0)) return;
  double* ring = c_db.mom_ring + (size_t)A.u.mom.ring * MOM_RING;
  double mid = (double)((long long)A.bid + A.ask) / 2;
  double cs = A.u.mom.csum + mid;                         // np.cumsum: sequential float adds
  int n = A.u.mom.n_mids;                                 // index of this mid
  __syncwarp();
  if (LANE == 0) ring[n % MOM_RING] = cs;
  __syncwarp();
  A.u.mom.csum = cs; A.u.mom.n_mids = n + 1;
  int len = n + 1;
  if (len > 20) { double prev = ring[(n - 20) % MOM_RING]; A.u.mom.avg20 = rint(((cs - prev) / 20.0) * 100.0) / 100.0; A.flags |= F_HAS_AVG20; }
  if (len > 50) { double prev = ring[(n - 50) % MOM_RING]; A.u.mom.avg50 = rint(((cs - prev) / 50.0) * 100.0) / 100.0; A.flags |= F_HAS_AVG50; }
  if ((A.flags & F_HAS_AVG20) && (A.flags & F_HAS_AVG50)) {
    if (A.u.mom.avg20 >= A.u.mom.avg50) ta_place_limit(a, A.size, true, A.ask);
    else ta_place_limit(a, A.size, false, A.bid);
  }
}
// AdaptiveMarketMakerAgent (:112-281)
__device__ __forceinline__ void mm_adaptive_update(AgentRec& A, const abides_agent_type& ty) {
  A.u.mm.window_size = A.u.mm.last_spread;
  int t = (int)(long long)rint(ty.level_spacing * A.u.mm.window_size);
  A.u.mm.tick_size = t == 0 ? 1 : t;
}
__device__ __noinline__ double sigmoid_stable(double x, double beta) {
  if (x >= 0) { double z = abm_exp(-beta * x); return 1 / (1 + z); }
  double z = abm_exp(beta * x);
  return z / (1 + z);
}
__device__ __noinline__ void mm_update_order_size(const abides_agent_type& ty) {
  AgentRec& A = ST.ag;
  long long qty = (long long)rint(ty.pov * (double)A.u.mm.transacted_volume);
  if (ty.skew_beta == 0) {
    long long s = qty >= ty.min_order_size ? qty : ty.min_order_size;
    A.u.mm.buy_size = (int)s; A.u.mm.sell_size = (int)s;
  } else {
    double ps = sigmoid_stable((double)A.holdings, ty.skew_beta);
    long long sell = (long long)ceil(ps * (double)qty);
    long long buy = (long long)floor((1 - ps) * (double)qty);
    A.u.mm.buy_size = (int)(buy >= ty.min_order_size ? buy : ty.min_order_size);
    A.u.mm.sell_size = (int)(sell >= ty.min_order_size ? sell : ty.min_order_size);
  }
}
__device__ __noinline__ void mm_place_orders(int a, long long mid, const abides_agent_type& ty) {
  AgentRec& A = ST.ag;
  long long highest_bid, lowest_ask;
  double w = A.u.mm.window_size;
  if (ty.anchor == 0) { highest_bid = mid - (long long)floor(0.5 * w); lowest_ask = mid + (long long)ceil(0.5 * w); }
  else if (ty.anchor == 1) { highest_bid = mid - 1; lowest_ask = (long long)((double)mid + w); }
  else { highest_bid = (long long)((double)mid - w); lowest_ask = mid + 1; }
  long long tick = A.u.mm.tick_size, nt = ty.num_ticks;
  long long lowest_bid = highest_bid - (nt - 1) * tick;
  long long first_bid = 0, last_ask = nt - 1;
  if (ty.backstop_quantity >= 0) {
    ta_place_limit(a, ty.backstop_quantity, true, lowest_bid);
    first_bid = 1;
    ta_place_limit(a, ty.backstop_quantity, false, lowest_ask + (nt - 1) * tick);
    last_ask = nt - 2;
  }
  for (long long i = first_bid; i < nt; i++) ta_place_limit(a, A.u.mm.buy_size, true, lowest_bid + i * tick);
  for (long long i = 0; i <= last_ask; i++) ta_place_limit(a, A.u.mm.sell_size, false, lowest_ask + i * tick);
}

// wakeup() of ZeroIntelligence / Value / Noise agents (ZeroIntelligenceAgent.py:96-160, ValueAgent.py:82-121,
// NoiseAgent.py:77-111)
__device__ __noinline__ void zvn_wakeup(int a, const abides_agent_type& ty) {
  AgentRec& A = ST.ag; Hot& h = HOT;
  ta_wakeup(a);
  ag_set_state(A, ST_INACTIVE);
  if (UNLIKELY(!(A.flags & F_HAS_OPEN) || !(A.flags & F_HAS_CLOSE))) return;
  A.flags |= F_TRADING;
  if (UNLIKELY((A.flags & F_MKT_CLOSED) && (A.flags & F_HAS_DAILY_CLOSE))) return;
  if (ty.klass == ABIDES_NOISE) {
    if (A.u.noise.wakeup_time > h.now) k_wakeup(a, A.u.noise.wakeup_time);
  } else {
    double dt = rng_exponential(SID_AGENT, fdiv(1.0, ty.lambda_a));
    k_wakeup(a, h.now + (long long)rint(dt));
  }
  if (UNLIKELY((A.flags & F_MKT_CLOSED) && !(A.flags & F_HAS_DAILY_CLOSE))) { ta_query_spread(a); ag_set_state(A, ST_AWAITING_SPREAD); return; }
  if (ty.klass != ABIDES_NOISE) ta_cancel_all(a);
  if (ty.klass == ABIDES_HBL) {                               // HBL.wakeup (:29-40): state 'ACTIVE' -> ask for the order stream
    agent_send(a, ABIDES_MSG_QUERY_ORDER_STREAM, ty.hbl_L, 0, 0, 0);
    ag_set_state(A, ST_AWAITING_STREAM);
    return;
  }
  ta_query_spread(a);
  ag_set_state(A, ST_AWAITING_SPREAD);
}
// wakeup() of Momentum / AdaptiveMarketMaker / POV-heartbeat agents
__device__ __noinline__ void other_wakeup(int a, const abides_agent_type& ty) {
  AgentRec& A = ST.ag; Hot& h = HOT;
  switch (ty.klass) {
    case ABIDES_MOMENTUM:
      if (ta_wakeup(a)) { ta_query_spread(a); ag_set_state(A, ST_AWAITING_SPREAD); }
      return;
    case ABIDES_ADAPTIVE_MM:
      if (ta_wakeup(a)) {
        ta_cancel_all(a);
        h.additional_delay += ty.cancel_limit_delay;
        ta_query_spread(a);
        agent_send(a, ABIDES_MSG_QUERY_TRANSACTED_VOLUME, (int)(unsigned)(ty.wake_up_freq & 0xffffffffll),
                   (int)(ty.wake_up_freq >> 32), 0, 0);
      }
      return;
    case ABIDES_MARKET_MAKER:                                   // MarketMakerAgent.wakeup (:56-66), polling mode
      if (ta_wakeup(a)) {
        ta_cancel_all(a);
        agent_send(a, ABIDES_MSG_QUERY_SPREAD, ty.mm_num_levels, 0, 0, 0);   // depth = subscribe_num_levels
        ag_set_state(A, ST_AWAITING_SPREAD);
      }
      return;
    case ABIDES_POV_EXEC: {                                     // POVExecutionAgent.wakeup (:40-48)
      const bool can_trade = ta_wakeup(a);
      k_wakeup(a, h.now + ty.wake_up_freq);
      if (!can_trade) return;
      if (ty.exec_trade && ty.exec_quantity - (double)A.u.pov.executed > 0 && ty.exec_start_time < h.now &&
          h.now < ty.exec_end_time) {
        ta_cancel_all(a);                                       // its market orders cannot be cancelled: no-op
        agent_send(a, ABIDES_MSG_QUERY_SPREAD, -1, 0, 0, 0);    // getCurrentSpread(depth=sys.maxsize)
        agent_send(a, ABIDES_MSG_QUERY_TRANSACTED_VOLUME, (int)(unsigned)(ty.exec_lookback & 0xffffffffll),
                   (int)(ty.exec_lookback >> 32), 0, 0);
        A.flags |= F_MM_AWAIT_TV;                               // state = 'AWAITING_TRANSACTED_VOLUME' (never reset)
      }
      return;
    }
    default: return;
  }
}
__device__ __forceinline__ void agent_wakeup(int a) {
  const abides_agent_type& ty = c_db.types[ST.ag.type_idx];
  if (ty.klass <= ABIDES_NOISE || ty.klass == ABIDES_HBL) zvn_wakeup(a, ty);
  else other_wakeup(a, ty);
}

// receiveMessage() tails of Momentum / AdaptiveMarketMaker agents
__device__ __noinline__ void other_receive(int a, int code, int w4_of_last_msg, const abides_agent_type& ty) {
  AgentRec& A = ST.ag; Hot& h = HOT;
  switch (ty.klass) {
    case ABIDES_MOMENTUM:
      if (ag_state(A) == ST_AWAITING_SPREAD && code == ABIDES_MSG_QUERY_SPREAD) {
        momentum_place_orders(a);
        k_wakeup(a, h.now + ty.wake_up_freq);
        ag_set_state(A, ST_AWAITING_WAKEUP);
      }
      return;
    case ABIDES_ADAPTIVE_MM: {
      long long mid = A.u.mm.last_mid;
      if (ty.window_size < 0) mm_adaptive_update(A, ty);
      if (code == ABIDES_MSG_QUERY_TRANSACTED_VOLUME && (A.flags & F_MM_AWAIT_TV)) {
        mm_update_order_size(ty);
        A.flags &= ~F_MM_AWAIT_TV;
      }
      if (code == ABIDES_MSG_QUERY_SPREAD && (A.flags & F_MM_AWAIT_SPREAD)) {
        if (A.bid > 0 && A.ask > 0) {
          mid = (long long)((double)((long long)A.ask + A.bid) / 2);
          A.u.mm.last_mid = (int)mid; A.flags |= F_HAS_LAST_MID;
          if (ty.window_size < 0) {
            double spread = (double)(A.ask - A.bid);
            double ewma = ty.spread_alpha * spread + (1 - ty.spread_alpha) * A.u.mm.last_spread;
            A.u.mm.window_size = ewma; A.u.mm.last_spread = ewma;
          }
        }
        A.flags &= ~F_MM_AWAIT_SPREAD;
      }
      if (!(A.flags & F_MM_AWAIT_SPREAD) && !(A.flags & F_MM_AWAIT_TV)) {
        if (!(A.flags & F_HAS_LAST_MID)) { h.status = ABIDES_ESTATE; return; }
        mm_place_orders(a, mid, ty);
        A.flags |= F_MM_AWAIT_SPREAD | F_MM_AWAIT_TV;
        k_wakeup(a, h.now + ty.wake_up_freq);
      }
      return;
    }
    case ABIDES_MARKET_MAKER:                                   // MarketMakerAgent.receiveMessage (:68-91)
      if (ag_state(A) == ST_AWAITING_SPREAD && code == ABIDES_MSG_QUERY_SPREAD) {
        ta_cancel_all(a);
        long long mid = A.last_trade, spread = 10;              // self.last_spread is never updated
        if (A.bid > 0 && A.ask > 0) {
          mid = (long long)((double)((long long)A.ask + A.bid) / 2);
          long long d = (long long)A.ask - A.bid; if (d < 0) d = -d;
          spread = (long long)((double)d / 2);
        }
#pragma unroll 1
        for (long long i = 0; i < 2 * ty.mm_num_levels; i++) {
          // round(randint(min, max) / 2): Python rounds x.5 to even
          const long long size = (long long)rint((double)rng_randint(SID_AGENT, ty.mm_min_size, ty.mm_max_size) / 2);
          ta_place_limit(a, size, true, mid - spread - i);
          ta_place_limit(a, size, false, mid + spread + i);
        }
        k_wakeup(a, h.now + ty.wake_up_freq);
        ag_set_state(A, ST_AWAITING_WAKEUP);
      }
      return;
    case ABIDES_POV_EXEC: {                                     // POVExecutionAgent.receiveMessage (:53-72)
      if (code == ABIDES_MSG_ORDER_EXECUTED) A.u.pov.executed += w4_of_last_msg;
      if (h.now > ty.exec_end_time) return;
      if (ty.exec_quantity - (double)A.u.pov.executed > 0 && (A.flags & F_MM_AWAIT_TV) &&
          code == ABIDES_MSG_QUERY_TRANSACTED_VOLUME && (A.flags & F_HAS_TV) && h.now > ty.exec_start_time) {
        const long long qty = (long long)rint(ty.pov * (double)A.u.pov.transacted_volume);
        ta_cancel_all(a);
        // TradingAgent.placeMarketOrder (:333-365): the MarketOrder takes an id even when it is then ignored
        const int id = (int)h.next_order_id++;
        if (qty > 0) {
          if (UNLIKELY(id == 0)) h.next_order_id++;
          agent_send(a, order_w0(ABIDES_MSG_MARKET_ORDER, ty.exec_is_buy != 0), id, 0, (int)qty, -1);
          if (UNLIKELY(ty.log_orders && id == 0)) h.next_order_id++;
        }
      }
      return;
    }
    default: return;
  }
}

__device__ __noinline__ void agent_receive(int a, int w0, int w1, int w2, int w3, int w4, int w5) {
  AgentRec& A = ST.ag;
  const abides_agent_type& ty = c_db.types[A.type_idx];
  Msg m; m.w0 = w0; m.w1 = w1; m.w2 = w2; m.w3 = w3; m.w4 = w4; m.w5 = w5;
  const int code = pay_code(w0) & ~ABIDES_MSG_REPLY;
  const int klass = ty.klass;
  ta_receive(a, m, klass, ty);
  if (klass <= ABIDES_NOISE || klass == ABIDES_HBL) {
    if (ag_state(A) == ST_AWAITING_SPREAD && code == ABIDES_MSG_QUERY_SPREAD) {
      if (A.flags & F_MKT_CLOSED) return;
      if (klass == ABIDES_ZI) zi_place_order(a, ty);
      else if (klass == ABIDES_HBL) hbl_place_order(a, ty);
      else if (klass == ABIDES_VALUE) value_place_order(a, ty);
      else noise_place_order(a);
      ag_set_state(A, ST_AWAITING_WAKEUP);
    }
    if (klass == ABIDES_HBL && ag_state(A) == ST_AWAITING_STREAM && code == ABIDES_MSG_QUERY_ORDER_STREAM) {
      if (A.flags & F_MKT_CLOSED) return;                      // HBL.receiveMessage (:155-173)
      ta_query_spread(a);
      ag_set_state(A, ST_AWAITING_SPREAD);
    }
    return;
  }
  other_receive(a, code, w4, ty);
}

// ----------------------------------------------------------------------------------------------
// ExchangeAgent (agent/ExchangeAgent.py:113-283, :398-411)
// ----------------------------------------------------------------------------------------------
__device__ __noinline__ void exchange_receive(int w0, int w1, int w2, int w3, int w4, int w5) {
  Hot& h = HOT;
  const abides_instance_cfg& cfg = RUN.cfg;
  h.exch_cur_time = h.now;
  h.exch_cd = cfg.exch_computation_delay;
  const int code = pay_code(w0), sender = w1;
  const bool is_order = code == ABIDES_MSG_LIMIT_ORDER || code == ABIDES_MSG_MARKET_ORDER || code == ABIDES_MSG_CANCEL_ORDER ||
                        code == ABIDES_MSG_MODIFY_ORDER;
  const bool is_query = code == ABIDES_MSG_QUERY_LAST_TRADE || code == ABIDES_MSG_QUERY_SPREAD ||
                        code == ABIDES_MSG_QUERY_ORDER_STREAM || code == ABIDES_MSG_QUERY_TRANSACTED_VOLUME;
  const bool closed = h.now > cfg.mkt_close;
  const int cflag = closed ? 1 << 8 : 0;
  if (UNLIKELY(closed && !is_query)) { k_send(0, sender, 0, ABIDES_MSG_MKT_CLOSED, 0, 0, 0, 0, 0); return; }
  int oid = w2;
  if (is_order) {
    if (UNLIKELY(oid == 0)) { if (cfg.exch_log_orders) h.next_order_id++; oid = (int)h.next_order_id++; }
  }
  switch (code) {
    case ABIDES_MSG_WHEN_MKT_OPEN:
      h.exch_cd = 0;
      k_send(0, sender, 0, ABIDES_MSG_WHEN_MKT_OPEN | ABIDES_MSG_REPLY, 0, 0, 0, 0, 0);
      break;
    case ABIDES_MSG_WHEN_MKT_CLOSE:
      h.exch_cd = 0;
      k_send(0, sender, 0, ABIDES_MSG_WHEN_MKT_CLOSE | ABIDES_MSG_REPLY, 0, 0, 0, 0, 0);
      break;
    case ABIDES_MSG_QUERY_LAST_TRADE:
      k_send(0, sender, 0, (ABIDES_MSG_QUERY_LAST_TRADE | ABIDES_MSG_REPLY) | cflag, 0, 0, 0, 0, h.last_trade);
      break;
    case ABIDES_MSG_QUERY_ORDER_STREAM: {                      // history[1:length+1] (ExchangeAgent.py:210-225)
      int avail = h.hist_B < cfg.stream_history ? h.hist_B : cfg.stream_history;
      if (w2 < avail) avail = w2;
      k_send(0, sender, 0, (ABIDES_MSG_QUERY_ORDER_STREAM | ABIDES_MSG_REPLY) | cflag, avail, h.hist_B, 0, 0, 0);
      break;
    }
    case ABIDES_MSG_QUERY_SPREAD: {
      int2 b = side_inside(BID), a = side_inside(ASK);
      k_send(0, sender, 0, (ABIDES_MSG_QUERY_SPREAD | ABIDES_MSG_REPLY) | cflag, b.x, b.y, a.x, a.y, h.last_trade);
      break;
    }
    case ABIDES_MSG_QUERY_TRANSACTED_VOLUME: {
      long long lookback = ((long long)(unsigned)w2) | ((long long)w3 << 32);
      long long vol = book_transacted_volume(h.now, lookback);
      k_send(0, sender, 0, (ABIDES_MSG_QUERY_TRANSACTED_VOLUME | ABIDES_MSG_REPLY) | cflag, 0,
             (int)(unsigned)(vol & 0xffffffffll), (int)(vol >> 32), 0, 0);
      break;
    }
    case ABIDES_MSG_LIMIT_ORDER:
      book_handle_limit<false>(sender, oid, w3, w4, pay_is_buy(w0));
      break;
    case ABIDES_MSG_MARKET_ORDER:
      book_handle_market<false>(sender, w4, pay_is_buy(w0));
      break;
    case ABIDES_MSG_CANCEL_ORDER:
      book_cancel<false>(sender, oid, w3, pay_is_buy(w0), w5 > 0 ? w5 : 0);
      break;
    case ABIDES_MSG_MODIFY_ORDER:
      book_modify<false>(sender, oid, w3, w5, pay_is_buy(w0));
      break;
    default: break;
  }
}

// ----------------------------------------------------------------------------------------------
// trace of a dispatched event (layout of include/abides_b200.h: abides_trace_rec)
// ----------------------------------------------------------------------------------------------
__device__ __noinline__ void trace_event(long long t, unsigned long long k2, int w0, int w1, int w2, int w3, int w4, int w5) {
  int rcpt = (int)(k2 >> 33);
  if ((k2 >> 32) & 1) { trace_write(t, rcpt, 0, 0, 0, 0, 0); return; }
  int full = pay_code(w0), code = full & ~ABIDES_MSG_REPLY;
  long long p0 = 0, p1 = 0, p2 = 0, p3 = 0;
  if (rcpt == 0) {
    p0 = w1;
    if (code == ABIDES_MSG_LIMIT_ORDER || code == ABIDES_MSG_MARKET_ORDER || code == ABIDES_MSG_CANCEL_ORDER ||
        code == ABIDES_MSG_MODIFY_ORDER) {
      p1 = w2; p2 = w3; p3 = pay_is_buy(w0) ? w4 : -(long long)w4;
      if (code == ABIDES_MSG_CANCEL_ORDER) p3 = pay_is_buy(w0) ? 1 : -1;
    } else if (code == ABIDES_MSG_QUERY_SPREAD) p1 = w2 < 0 ? LLONG_MAX : w2;   // depth (-1: sys.maxsize)
  } else {
    switch (code) {
      case ABIDES_MSG_ORDER_ACCEPTED: case ABIDES_MSG_ORDER_EXECUTED: case ABIDES_MSG_ORDER_CANCELLED:
      case ABIDES_MSG_ORDER_MODIFIED:
        p0 = w2; p1 = w3; p2 = pay_is_buy(w0) ? w4 : -(long long)w4; p3 = w5; break;
      case ABIDES_MSG_QUERY_SPREAD: p0 = w1; p1 = w2; p2 = w3; p3 = w4; break;
      case ABIDES_MSG_WHEN_MKT_OPEN: p0 = RUN.cfg.mkt_open; break;
      case ABIDES_MSG_WHEN_MKT_CLOSE: p0 = RUN.cfg.mkt_close; break;
      case ABIDES_MSG_QUERY_TRANSACTED_VOLUME: p0 = ((long long)(unsigned)w2) | ((long long)w3 << 32); break;
      case ABIDES_MSG_QUERY_LAST_TRADE: p0 = w5; break;
      default: break;
    }
  }
  trace_write(t, rcpt, full, p0, p1, p2, p3);
}

// ----------------------------------------------------------------------------------------------
// agent record movement: one coalesced 128-byte transaction each way
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void rec_load(int a) {
  __syncwarp();
  const unsigned* src = reinterpret_cast<const unsigned*>(&RUN.agents[a]);
  reinterpret_cast<unsigned*>(&ST.ag)[LANE] = src[LANE];
  RUN.cur_agent = a;
  __syncwarp();
}
__device__ __forceinline__ void rec_store(int a) {
  __syncwarp();
  unsigned* dst = reinterpret_cast<unsigned*>(&RUN.agents[a]);
  dst[LANE] = reinterpret_cast<const unsigned*>(&ST.ag)[LANE];
  __syncwarp();
}

// kernelStopping in id order (Kernel.py:289-291; TradingAgent.py:122-147 and subclasses)
__device__ __noinline__ void finalize() {
  Hot& h = HOT;
  long long slowest = h.exch_busy > RUN.cfg.start_time ? h.exch_busy : RUN.cfg.start_time;
  for (int a = 1; a < RUN.n_agents; a++) {
    rec_load(a);
    AgentRec& A = ST.ag;
    const abides_agent_type& ty = c_db.types[A.type_idx];
    if (A.busy > slowest) slowest = A.busy;
    long long cash = A.cash;
    if (A.holdings != 0) cash += (long long)A.last_trade * A.holdings;
    double fv = nan("");
    if (ty.klass == ABIDES_ZI || ty.klass == ABIDES_HBL) {
      long long H = round_hundreds(A.holdings);
      long long rT = oracle_observe(A.cur_time, 0, SID_AGENT);
      long long surplus = 0;
      const int* th = c_db.theta + RUN.cfg.theta_offset + c_db.agent_theta[RUN.aoff + a];
      // past q_max lots the reference indexes theta out of range: IndexError when long (H > q_max), Python's
      // negative-index wrap-around when short (ZeroIntelligenceAgent.py:75-81)
      if (H > ty.q_max) { if (h.status == 0) h.status = ABIDES_ESTATE; H = ty.q_max; }
      if (H > 0) for (long long x = 1; x <= H; x++) surplus += th[x + ty.q_max - 1];
      else if (H < 0) {
        for (long long x = H + 1; x <= 0; x++) { long long j = x + ty.q_max - 1; surplus += th[j < 0 ? j + 2 * ty.q_max : j]; }
        surplus = -surplus;
      }
      surplus += rT * H;
      surplus += A.cash - ty.starting_cash;
      fv = (double)surplus;
    } else if (ty.klass == ABIDES_VALUE) {
      long long H = round_hundreds(A.holdings);
      long long rT = oracle_observe(A.cur_time, 0, SID_AGENT);
      long long surplus = rT * H + A.cash - ty.starting_cash;
      fv = (double)surplus / (double)ty.starting_cash;
    } else if (ty.klass == ABIDES_NOISE) {
      long long H = round_hundreds(A.holdings);
      double rT;
      if ((A.flags & F_KNOWN_BOOK) && A.bid > 0 && A.ask > 0) rT = (double)((long long)A.bid + A.ask) / 2;
      else rT = (double)A.last_trade;
      double surplus = rT * (double)H;
      surplus += (double)(A.cash - ty.starting_cash);
      fv = surplus / (double)ty.starting_cash;
    }
    if (LANE == 0) {
      abides_agent_result r;
      r.cash = A.cash; r.holdings = A.holdings; r.marked_to_market = cash; r.final_valuation = fv;
      c_db.ares[RUN.aoff + a] = r;
    }
  }
  if (LANE == 0) {
    abides_agent_result r; r.cash = 0; r.holdings = 0; r.marked_to_market = 0; r.final_valuation = nan("");
    c_db.ares[RUN.aoff] = r;
    abides_instance_result R;
    memset(&R, 0, sizeof R);
    R.ttl_messages = h.ttl; R.delivered = h.delivered; R.final_time = h.now;
    R.slowest_agent_finish_time = slowest;
    R.last_trade = h.has_last_trade ? h.last_trade : -1;
    R.final_fundamental = (long long)h.or_pv;
    R.n_orders = h.next_order_id; R.n_fills = h.n_fills; R.traded_volume = h.volume;
    R.status = h.status; R.queue_peak = h.queue_peak; R.book_peak_orders = h.book_peak_orders;
    c_db.res[RUN.inst] = R;
  }
  RUN.cur_agent = -1;
}

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

__device__ __noinline__ void run_setup(int inst) {
  const DevBatch& db = c_db;
  Run& r = RUN;
  r.cfg = db.cfg[inst];
  r.lk[L_A] = db.a_keys + (size_t)inst * A_CAP; r.lp[L_A] = db.a_pay + (size_t)inst * A_CAP; r.lcap[L_A] = A_CAP;
  r.lk[L_B] = db.b_keys + (size_t)inst * db.b_cap; r.lp[L_B] = db.b_pay + (size_t)inst * db.b_cap; r.lcap[L_B] = db.b_cap;
  for (int s = 0; s < 2; s++) {
    r.lk[L_POOL + s] = db.p_keys + ((size_t)inst * 2 + s) * db.deep_cap;
    r.lp[L_POOL + s] = db.p_pay + ((size_t)inst * 2 + s) * db.deep_cap;
    r.lcap[L_POOL + s] = db.deep_cap;
    r.pfree[s] = db.p_free + ((size_t)inst * 2 + s) * db.deep_cap;
  }
  r.olog = db.olog ? db.olog + (size_t)inst * OLOG_CAP : nullptr;
  r.hbl_cnt = db.hbl_cnt ? db.hbl_cnt + (size_t)inst * 4 * HBL_RANGE_CAP : nullptr;
  r.arena = db.arena + (size_t)inst * db.arena_cap;
  r.txn_log = db.txn_log + (size_t)inst * TXN_CAP;
  r.unrolled = db.unrolled + (size_t)inst * UNR_CAP;
  r.trace = db.trace_cap > 0 ? db.trace + (size_t)inst * db.trace_cap : nullptr;
  r.aoff = r.cfg.agent_offset; r.inst = inst; r.n_agents = r.cfg.n_agents; r.cur_agent = -1;
  r.mt0 = db.mt ? db.mt + (size_t)(r.aoff + (long long)ABIDES_EXTRA_STREAMS * inst) * ABIDES_MT_N : nullptr;
  r.lat_to = db.lat_to + r.aoff; r.lat_from = db.lat_from + r.aoff;
  r.agents = db.agents + r.aoff;
  for (int i = 0; i < N_PROF; i++) r.prof[i] = 0;
}

// the persistent event loop: Kernel.runner (Kernel.py:190-271), one warp per instance
__global__ void __launch_bounds__(32) sim_kernel(long long until_time) {
  const int inst = blockIdx.x, lane = threadIdx.x;
  if (inst >= c_db.n_instances) return;
  InstShared* G = &c_db.state[inst];
  {
    const int4* src = reinterpret_cast<const int4*>(G);
    int4* dst = reinterpret_cast<int4*>(&ST);
#pragma unroll 1
    for (int i = lane; i < (int)(sizeof(InstShared) / 16); i += 32) dst[i] = src[i];
  }
  if (lane == 0) run_setup(inst);
  __syncwarp();
  Hot& h = HOT;
  const long long stop_time = RUN.cfg.stop_time;
  const long long default_cd = RUN.cfg.default_computation_delay;
  const bool tracing = RUN.trace != nullptr;
#ifdef ABIDES_PROFILE
  const long long prof_start = clock64();
#endif

  if (!h.done) {
    bool finished = false;
    for (;;) {
      PROF_T0();
      if (UNLIKELY(h.status != 0)) { finished = true; break; }
      if (UNLIKELY(!(h.now <= stop_time))) { finished = true; break; }
      if (UNLIKELY(h.n_near == 0)) {
        if (h.l_live[0] + h.l_live[1] == 0) { finished = true; break; }
        q_refill();
        PROF_ADD(P_REFILL);
        if (h.n_near == 0) { finished = true; break; }
      }
      const int top_i = h.n_near - 1;
      const long long t = ST.nt[top_i];
      const unsigned long long k2 = ST.nk[top_i];
      const int slot = ST.nslot[top_i];
      if (UNLIKELY(t > until_time)) break;
      h.n_near = top_i;
      h.now = t;
      h.ttl++;
      h.additional_delay = 0;
      const int rcpt = (int)(k2 >> 33);
      const bool is_wakeup = (k2 >> 32) & 1;
      PROF_ADD(P_POP);
      if (rcpt != 0 && rcpt != RUN.cur_agent) { rec_load(rcpt); PROF_CNT(P_N_RECLOAD); PROF_ADD(P_REC); }
      const long long busy = rcpt == 0 ? h.exch_busy : ST.ag.busy;
      if (busy > t) {
        // agent in the future: re-queue at its busy time with the same uniq (Kernel.py:217-223)
        PROF_CNT(P_N_REQUEUE);
        if (key_less(busy, k2, h.lim_t, h.lim_k)) {
          near_insert(busy, k2, slot);
        } else {
          const int4 p0 = *reinterpret_cast<const int4*>(&ST.npay[slot].w[0]);
          const int4 p1 = *reinterpret_cast<const int4*>(&ST.npay[slot].w[4]);
          h.near_free |= 1u << slot;
          __syncwarp();
          q_push(busy, k2, p0.x, p0.y, p0.z, p0.w, p1.x, p1.y);
        }
        PROF_ADD(P_REQUEUE);
        continue;
      }
      const int4 p0 = *reinterpret_cast<const int4*>(&ST.npay[slot].w[0]);
      const int4 p1 = *reinterpret_cast<const int4*>(&ST.npay[slot].w[4]);
      h.near_free |= 1u << slot;
      __syncwarp();
      h.delivered++;
      if (UNLIKELY(tracing)) trace_event(t, k2, p0.x, p0.y, p0.z, p0.w, p1.x, p1.y);
      if (rcpt == 0) {
        h.exch_busy = t;
        if (is_wakeup) h.exch_cur_time = t;
        else exchange_receive(p0.x, p0.y, p0.z, p0.w, p1.x, p1.y);
        h.exch_busy += h.exch_cd + h.additional_delay;
#ifdef ABIDES_PROFILE
        { int c = pay_code(p0.x);
          PROF_ADD(c == ABIDES_MSG_LIMIT_ORDER ? P_EXCH_ORDER : c == ABIDES_MSG_CANCEL_ORDER ? P_EXCH_CANCEL :
                   c == ABIDES_MSG_QUERY_SPREAD ? P_EXCH_QUERY : P_EXCH_OTHER); }
#endif
      } else {
        ST.ag.busy = t;
        if (is_wakeup) agent_wakeup(rcpt);
        else agent_receive(rcpt, p0.x, p0.y, p0.z, p0.w, p1.x, p1.y);
        ST.ag.busy += default_cd + h.additional_delay;
        rec_store(rcpt);
#ifdef ABIDES_PROFILE
        PROF_ADD(is_wakeup ? P_AG_WAKE : (pay_code(p0.x) & ~ABIDES_MSG_REPLY) == ABIDES_MSG_QUERY_SPREAD ? P_AG_SPREAD : P_AG_OTHER);
#endif
      }
    }
    if (finished) {
      finalize();
      h.done = 1;
    }
  }
  __syncwarp();
#ifdef ABIDES_PROFILE
  if (lane == 0) {
    RUN.prof[P_TOTAL] += (unsigned long long)(clock64() - prof_start);
    for (int i = 0; i < N_PROF; i++) atomicAdd(&c_db.prof[i], RUN.prof[i]);
  }
  __syncwarp();
#endif
  {
    int4* dst = reinterpret_cast<int4*>(G);
    const int4* src = reinterpret_cast<const int4*>(&ST);
#pragma unroll 1
    for (int i = lane; i < (int)(sizeof(InstShared) / 16); i += 32) dst[i] = src[i];
  }
}

// ----------------------------------------------------------------------------------------------
// K2 parity harness: replay recorded order tapes into the book, one warp per book
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) replay_kernel(int n_books, const long long* op_offset, const abides_book_op* ops,
                                                    int stream_history, int max_events, abides_book_event* events,
                                                    int* counts, abides_book_state* states, EvKey* pool_keys, EvPay* pool_pay,
                                                    int* pool_free, int deep_cap) {
  int b = blockIdx.x, lane = threadIdx.x;
  if (b >= n_books) return;
  for (int i = lane; i < (int)(sizeof(Smem) / 4); i += 32) reinterpret_cast<unsigned*>(SM)[i] = 0;
  __syncwarp();
  if (lane == 0) {
    Run& r = RUN;
    r.cfg.stream_history = stream_history;
    for (int s = 0; s < 2; s++) {
      r.lk[L_POOL + s] = pool_keys + ((size_t)b * 2 + s) * deep_cap;
      r.lp[L_POOL + s] = pool_pay + ((size_t)b * 2 + s) * deep_cap;
      r.lcap[L_POOL + s] = deep_cap;
      r.pfree[s] = pool_free + ((size_t)b * 2 + s) * deep_cap;
      HOT.span[L_POOL + s] = 64;
    }
    r.txn_log = nullptr;
    r.rp_events = events; r.rp_counts = counts; r.rp_max = max_events; r.rp_base = (int)op_offset[b];
    HOT.next_order_id = 1 << 30;                 // ids of market-order children: outside the tape's id space
  }
  __syncwarp();
  Hot& h = HOT;
  for (long long i = op_offset[b]; i < op_offset[b + 1]; i++) {
    abides_book_op op = ops[i];
    __syncwarp();
    if (lane == 0) { RUN.rp_op = (int)i; counts[i] = 0; }
    __syncwarp();
    h.now = op.time;
    bool buy = op.is_buy != 0;
    switch (op.kind) {
      case ABIDES_MSG_LIMIT_ORDER: book_handle_limit<true>(op.agent, (int)op.order_id, (int)op.price, (int)op.qty, buy); break;
      case ABIDES_MSG_MARKET_ORDER: book_handle_market<true>(op.agent, (int)op.qty, buy); break;
      case ABIDES_MSG_CANCEL_ORDER: book_cancel<true>(op.agent, (int)op.order_id, (int)op.price, buy, 0); break;
      case ABIDES_MSG_MODIFY_ORDER: book_modify<true>(op.agent, (int)op.order_id, (int)op.price, (int)op.new_qty, buy); break;
      default: break;
    }
    int2 bi = side_inside(BID), ai = side_inside(ASK);
    int nbl = side_levels(BID), nal = side_levels(ASK);
    if (lane == 0) {
      abides_book_state st;
      st.bid = bi.x; st.bid_vol = bi.y; st.ask = ai.x; st.ask_vol = ai.y;
      st.last_trade = h.has_last_trade ? h.last_trade : -1;
      st.n_bid_levels = nbl; st.n_ask_levels = nal;
      states[i] = st;
    }
    __syncwarp();
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
thread_local std::string g_err;
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
  bool has_hbl = false;
  long long na = -1, n_theta = -1;
  bool operator==(const ShapeSig& o) const {
    return ni == o.ni && n_types == o.n_types && rng_mode == o.rng_mode && trace_cap == o.trace_cap &&
           max_agents == o.max_agents && n_mom == o.n_mom && na == o.na && n_theta == o.n_theta && has_hbl == o.has_hbl;
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
  float last_ms = 0.f, last_reset_ms = 0.f;
  int last_launches = 0;
  int launches_since_load = 0;
  unsigned* mt_pristine = nullptr;
  size_t n_streams = 0;
  int n_mom = 0;
  ShapeSig sig;
};

namespace {

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
int launch_init(abides_handle* h) {
  DevBatch& d = h->db;
  if (d.rng_mode == ABIDES_RNG_MT19937)
    CK(cudaMemcpyAsync(d.mt, h->mt_pristine, h->n_streams * ABIDES_MT_N * sizeof(unsigned), cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemsetAsync(d.mom_ring, 0, (size_t)(h->n_mom > 0 ? h->n_mom : 1) * MOM_RING * sizeof(double), h->stream));
  CK(cudaMemsetAsync(d.res, 0, (size_t)d.n_instances * sizeof(abides_instance_result), h->stream));
  CK(cudaMemsetAsync(d.prof, 0, N_PROF * sizeof(unsigned long long), h->stream));
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
  CK(cudaFuncSetAttribute(sim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
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


int abides_load(abides_handle* h, const abides_batch* b) {
  if (!h || !b) return fail(ABIDES_EINVAL, "abides_load: NULL argument");
  if (b->n_instances <= 0 || !b->instances || !b->types || !b->agent_type || !b->lat_to_exch || !b->lat_from_exch ||
      !b->agent_size || !b->agent_wakeup_time || !b->agent_theta)
    return fail(ABIDES_EINVAL, "abides_load: missing table");
  if (b->rng_mode == ABIDES_RNG_MT19937 && (!b->mt_key || !b->mt_pos || !b->mt_has_gauss || !b->mt_gauss))
    return fail(ABIDES_EINVAL, "abides_load: MT19937 mode needs mt_key/mt_pos/mt_has_gauss/mt_gauss");
  if (b->rng_mode != ABIDES_RNG_MT19937 && b->rng_mode != ABIDES_RNG_PHILOX) return fail(ABIDES_EINVAL, "abides_load: bad rng_mode");
  CK(cudaSetDevice(h->device));
  int ni = b->n_instances;
  long long na = b->n_agents_total;
  int max_agents = 0;
  long long off = 0;
  std::vector<int> mom_index((size_t)na, -1);
  int n_mom = 0, n_hbl = 0;
  for (int i = 0; i < ni; i++) {
    const abides_instance_cfg& ic = b->instances[i];
    if (ic.agent_offset != off || ic.n_agents <= 0) return fail(ABIDES_EINVAL, "abides_load: agent_offset must be the prefix sum of n_agents");
    if (ic.type_base < 0 || ic.type_base >= b->n_types) return fail(ABIDES_EINVAL, "abides_load: bad type_base");
    if (ic.r_bar < 0 || ic.r_bar > 2.0e9) return fail(ABIDES_EINVAL, "abides_load: r_bar out of the int32 price range");
    if (off + ic.n_agents > na) return fail(ABIDES_EINVAL, "abides_load: n_agents_total mismatch");
    for (int a = 0; a < ic.n_agents; a++) {
      int t = ic.type_base + b->agent_type[off + a];
      if (t < 0 || t >= b->n_types) return fail(ABIDES_EINVAL, "abides_load: agent_type out of range");
      int k = b->types[t].klass;
      if ((a == 0) != (k == ABIDES_EXCHANGE)) return fail(ABIDES_EINVAL, "abides_load: agent 0 (and only agent 0) must be the exchange");
      if (k < 0 || k > ABIDES_HBL) return fail(ABIDES_EINVAL, "abides_load: unknown agent class");
      if (k == ABIDES_HBL) { n_hbl++; if (b->types[t].hbl_L < 1) return fail(ABIDES_EINVAL, "abides_load: HBL agent with L < 1"); }
      if (k == ABIDES_MOMENTUM) mom_index[off + a] = n_mom++;
      if ((k == ABIDES_ZI || k == ABIDES_HBL) && (b->agent_theta[off + a] < 0 || !b->theta)) return fail(ABIDES_EINVAL, "abides_load: ZI / HBL agent without theta");
    }
    if (ic.n_agents > max_agents) max_agents = ic.n_agents;
    off += ic.n_agents;
  }
  if (off != na) return fail(ABIDES_EINVAL, "abides_load: n_agents_total mismatch");
  ShapeSig sig;
  sig.ni = ni; sig.n_types = b->n_types; sig.rng_mode = b->rng_mode; sig.trace_cap = b->trace_capacity;
  sig.max_agents = max_agents; sig.n_mom = n_mom; sig.na = na; sig.n_theta = b->n_theta_total; sig.has_hbl = n_hbl > 0;
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
      if ((rc = dev_alloc(h, &d.mt, ns * ABIDES_MT_N))) return rc;
      if ((rc = dev_alloc(h, &h->mt_pristine, ns * ABIDES_MT_N))) return rc;
      if ((rc = dev_alloc(h, &dpos, ns))) return rc;
      if ((rc = dev_alloc(h, &dhg, ns))) return rc;
      if ((rc = dev_alloc(h, &dg, ns))) return rc;
      d.mt_pos = dpos; d.mt_has_gauss = dhg; d.mt_gauss = dg;
    }
    if ((rc = dev_alloc(h, &d.type_log1mk, (size_t)b->n_types))) return rc;
    if ((rc = dev_alloc(h, &d.type_omp2, (size_t)b->n_types))) return rc;
    if ((rc = dev_alloc(h, &d.agents, (size_t)na))) return rc;
    if ((rc = dev_alloc(h, &d.state, (size_t)ni))) return rc;
    if ((rc = dev_alloc(h, &d.a_keys, (size_t)ni * A_CAP))) return rc;
    if ((rc = dev_alloc(h, &d.a_pay, (size_t)ni * A_CAP))) return rc;
    if ((rc = dev_alloc(h, &d.b_keys, (size_t)ni * d.b_cap))) return rc;
    if ((rc = dev_alloc(h, &d.b_pay, (size_t)ni * d.b_cap))) return rc;
    if ((rc = dev_alloc(h, &d.prof, (size_t)N_PROF))) return rc;
    if ((rc = dev_alloc(h, &d.arena, (size_t)ni * d.arena_cap))) return rc;
    if ((rc = dev_alloc(h, &d.p_keys, (size_t)ni * 2 * d.deep_cap))) return rc;
    if ((rc = dev_alloc(h, &d.p_pay, (size_t)ni * 2 * d.deep_cap))) return rc;
    if ((rc = dev_alloc(h, &d.p_free, (size_t)ni * 2 * d.deep_cap))) return rc;
    if (n_hbl > 0) {
      if ((rc = dev_alloc(h, &d.olog, (size_t)ni * OLOG_CAP))) return rc;
      if ((rc = dev_alloc(h, &d.hbl_cnt, (size_t)ni * 4 * HBL_RANGE_CAP))) return rc;
    }
    if ((rc = dev_alloc(h, &d.txn_log, (size_t)ni * TXN_CAP))) return rc;
    if ((rc = dev_alloc(h, &d.unrolled, (size_t)ni * UNR_CAP))) return rc;
    if ((rc = dev_alloc(h, &d.mom_ring, (size_t)(n_mom > 0 ? n_mom : 1) * MOM_RING))) return rc;
    if ((rc = dev_alloc(h, &d.trace, (size_t)ni * (d.trace_cap > 0 ? d.trace_cap : 0)))) return rc;
    if ((rc = dev_alloc(h, &d.res, (size_t)ni))) return rc;
    if ((rc = dev_alloc(h, &d.ares, (size_t)na))) return rc;
    h->sig = sig;
  }
  h->n_streams = ns; h->n_mom = n_mom;
  if ((rc = h2d(h, d.cfg, b->instances, (size_t)ni))) return rc;
  if ((rc = h2d(h, d.types, b->types, (size_t)b->n_types))) return rc;
  if ((rc = h2d(h, d.agent_type, b->agent_type, (size_t)na))) return rc;
  if ((rc = h2d(h, d.lat_to, b->lat_to_exch, (size_t)na))) return rc;
  if ((rc = h2d(h, d.lat_from, b->lat_from_exch, (size_t)na))) return rc;
  if ((rc = h2d(h, d.size, reinterpret_cast<const long long*>(b->agent_size), (size_t)na))) return rc;
  if ((rc = h2d(h, d.wakeup_time, reinterpret_cast<const long long*>(b->agent_wakeup_time), (size_t)na))) return rc;
  if ((rc = h2d(h, d.agent_theta, b->agent_theta, (size_t)na))) return rc;
  if ((rc = h2d(h, d.theta, b->theta, (size_t)(b->n_theta_total > 0 ? b->n_theta_total : 0)))) return rc;
  if ((rc = h2d(h, d.mom_index, mom_index.data(), (size_t)na))) return rc;
  if (b->rng_mode == ABIDES_RNG_MT19937) {
    if ((rc = h2d(h, (const unsigned*)h->mt_pristine, b->mt_key, ns * ABIDES_MT_N))) return rc;
    if ((rc = h2d(h, d.mt_pos, b->mt_pos, ns))) return rc;
    if ((rc = h2d(h, d.mt_has_gauss, b->mt_has_gauss, ns))) return rc;
    if ((rc = h2d(h, d.mt_gauss, b->mt_gauss, ns))) return rc;
  }
  h->launches_since_load = 0;
  if ((rc = launch_init(h))) return rc;
  CK(cudaStreamSynchronize(h->stream));      // mom_index (a local vector) must outlive the copy
  h->n_instances = ni; h->n_agents_total = na; h->loaded = true;
  return 0;
}

/* Re-arm the loaded batch (inputs stay resident in HBM): restores the MT19937 states and rebuilds the
 * initial device state, so the same batch can be run again. */
int abides_reset(abides_handle* h) {
  if (!h || !h->loaded) return fail(ABIDES_ESTATE, "abides_reset: no batch loaded");
  CK(cudaSetDevice(h->device));
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
  CK(cudaEventRecord(h->ev0, h->stream));
  CK(cudaMemcpyToSymbolAsync(c_db, &h->db, sizeof(DevBatch), 0, cudaMemcpyHostToDevice, h->stream));
  sim_kernel<<<h->n_instances, 32, sizeof(Smem), h->stream>>>((long long)until_time);
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
  const char* base = reinterpret_cast<const char*>(&h->db.state[instance]);
  CK(cudaMemcpy(&n, base + offsetof(InstShared, hot) + offsetof(Hot, n_trace), sizeof(int), cudaMemcpyDeviceToHost));
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
  RCK(cudaFuncSetAttribute(replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  RCK(cudaEventRecord(h->ev0, h->stream));
  replay_kernel<<<n_books, 32, sizeof(Smem), h->stream>>>(n_books, d_off, d_ops, stream_history, max_events_per_op, d_ev, d_cnt,
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
