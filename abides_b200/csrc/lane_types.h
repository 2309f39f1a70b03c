// lane_types.h -- data layout of the LANE engine: one simulation instance per THREAD (a warp = 32 instances).
//
// Every structure here is per instance and touched by exactly one thread, so nothing is warp-cooperative:
// the same code compiles for the device (engine.cu) and, for the CPU-side logic tests only, for the host
// (tests/hostsim/, never loaded by the package).
//
//   * event queue : 4-ary min-heap of 16-byte keys in HBM, root at index 3 so that every sibling group is one
//                   aligned 64-byte block.  A key packs (time - start, recipient, wakeup flag, uniq, payload slot):
//                   a = t_rel << 16 | rcpt >> 4,  b = (rcpt & 15) << 60 | wakeup << 59 | uniq << 27 | slot, compared
//                   as an unsigned 128-bit number -- the reference's (deliverAt, (recipient, MESSAGE < WAKEUP, uniq))
//                   order (Kernel.py:387,:417, message/Message.py:32-45).  Payloads (32 B) live in free-listed slots.
//   * order book  : per side a sorted circular deque of the best <= K orders (O(1) at both ends: matching pops the
//                   front, market-maker ladders grow from either end) over an unsorted pool with a free list.
//   * agent state : one 128-byte record per agent, fields read and written in place (only the touched 32-byte sectors
//                   move; a busy re-queue reads 8 bytes).
//   * RNG         : MT19937 per stream, regenerated ONE WORD PER DRAW (no 624-word twist spikes that would stall the
//                   other 31 lanes), or seed-only "lazy" streams for agents that draw a handful of words per day.
#ifndef ABIDES_LANE_TYPES_H
#define ABIDES_LANE_TYPES_H

#include <stdint.h>
#include <stddef.h>
#include "abides_b200.h"
#include "abides_dd_math.h"

#if defined(__CUDACC__)
#define LN_ALIGN(n) __align__(n)
#define LN_HD __host__ __device__
#else
#define LN_ALIGN(n) alignas(n)
#define LN_HD
#endif

// agent classes as preprocessor constants (the variants' class masks are tested with #if)
#define LN_CLS_ZI 1
#define LN_CLS_VALUE 2
#define LN_CLS_NOISE 3
#define LN_CLS_MOMENTUM 4
#define LN_CLS_ADAPTIVE_MM 5
#define LN_CLS_POV_EXEC 6
#define LN_CLS_MARKET_MAKER 7
#define LN_CLS_HBL 8
#define LN_CLS_SCHEDULE_EXEC 9
#define LN_CLS_MARKET_REPLAY 11
static_assert(LN_CLS_ZI == ABIDES_ZI && LN_CLS_VALUE == ABIDES_VALUE && LN_CLS_NOISE == ABIDES_NOISE &&
              LN_CLS_MOMENTUM == ABIDES_MOMENTUM && LN_CLS_ADAPTIVE_MM == ABIDES_ADAPTIVE_MM &&
              LN_CLS_POV_EXEC == ABIDES_POV_EXEC && LN_CLS_MARKET_MAKER == ABIDES_MARKET_MAKER &&
              LN_CLS_HBL == ABIDES_HBL && LN_CLS_SCHEDULE_EXEC == ABIDES_SCHEDULE_EXEC &&
              LN_CLS_MARKET_REPLAY == ABIDES_MARKET_REPLAY, "class constants");
#define LN_CLS(k) (1 << (k))

namespace lane {

#ifndef LANE_CTA_THREADS
#define LANE_CTA_THREADS 128
#endif
// warps of a CTA that only run the heavy handler classes (-DLANE_HEAVY_WARPS=1: measured, no gain -- the order-placement
// handler alone is the critical path of a round, whichever warp runs it)
#ifndef LANE_HEAVY_WARPS
#define LANE_HEAVY_WARPS 0
#endif
// LANE_STAGE: copy the instance's hot scalars and the recipient's record to thread-local memory around each half of a
// round (1) or work on them in place (0).  LANE_ONE_CTX: one LCtx per thread, re-pointed per half, instead of two.
// Measured at the bench size (sparse_zi_1000 x 45 056, profiles/r2_lane_staging_ab.json): staging 579 M msg/s and 3.0 KB of
// DRAM traffic per message, in place 659 M msg/s and 1.65 KB, in place with one context 620 M msg/s -- the staged copies
// live in local memory, which no longer fits the L2 at that many resident threads.
#ifndef LANE_STAGE
#define LANE_STAGE 0
#endif
#ifndef LANE_ONE_CTX
#define LANE_ONE_CTX 0
#endif
constexpr int LANE_CTA = LANE_CTA_THREADS;   // instances per CTA of the event loop (one owner thread each)
constexpr int LANE_HEAVY = 32 * LANE_HEAVY_WARPS;   // extra threads of a CTA that only run the heavy handler classes
constexpr int LANE_KINDS = 128;              // handler classes the CTA regroups popped events by
constexpr int LANE_PROF_SLOTS = 8;            // rounds, pop cycles, sort cycles, handle cycles, events, ...
constexpr int HEAP_ROOT = 3;                 // children of node i: 4 * (i - 2) .. + 3; parent of j: j / 4 + 2
constexpr int OUTBOX_CAP = 4;                // messages / wakeups an event may emit before the outbox is flushed early
constexpr int MOM_RING = 64;                 // >= 51 prefix sums (MomentumAgent)
constexpr int DEAD_PRICE = INT32_MIN;        // tombstone in a pool's price array
constexpr int LAZY_MAX_WORDS = 227;          // a seed-only stream can hand out this many 32-bit words
constexpr int TRACE_FUNDAMENTAL = 64;
constexpr int SID_AGENT = 4;                 // RNG stream id of the agent being dispatched (0..3: ABIDES_STREAM_*)

struct LN_ALIGN(16) LKey { unsigned long long a, b; };
struct LN_ALIGN(16) LPay { int w[8]; };
struct LN_ALIGN(16) LOrd { int price, qty, id, agent; int seq, arr, log, pad; };   // a resting order (top deque / pool)
struct LN_ALIGN(16) OpenOrd { int key, id, price, sqty; };      // TradingAgent.orders entry; sqty > 0 buy, < 0 sell; key -1 = removed
struct LN_ALIGN(16) TxnRow { long long t; int q; int seq; };    // OrderBook.history transaction + bucket of its record
struct LN_ALIGN(16) RngState { int pos; unsigned flags; double gauss; };
enum : unsigned { RNG_HAS_GAUSS = 1u, RNG_INC = 2u, RNG_LAZY = 4u };

enum : unsigned {
  F_HAS_OPEN = 1u, F_HAS_CLOSE = 2u, F_MKT_CLOSED = 4u, F_HAS_LAST_TRADE = 8u, F_HAS_DAILY_CLOSE = 16u,
  F_KNOWN_BOOK = 32u, F_TRADING = 64u, F_HAS_PREV_WAKE = 128u, F_SUB_REQUESTED = 256u,
  F_MM_AWAIT_SPREAD = 512u, F_MM_AWAIT_TV = 1024u, F_HAS_LAST_MID = 2048u, F_HAS_AVG20 = 4096u,
  F_HAS_AVG50 = 8192u, F_HAS_TV = 16384u, F_STATE_SHIFT = 16u, F_STATE_MASK = 7u << 16
};
enum { ST_AWAITING_WAKEUP = 0, ST_INACTIVE = 1, ST_AWAITING_SPREAD = 2, ST_ACTIVE = 3, ST_AWAITING_STREAM = 4,
       ST_AWAITING_MARKET_DATA = 5 };
enum { BID = 0, ASK = 1 };

struct LN_ALIGN(128) LAgent {           // 128 bytes = four 32-byte sectors; sector 0 is all a busy re-queue reads
  long long busy;                       // Kernel.agentCurrentTimes[a]
  long long cash;                       // holdings['CASH']
  long long cur_time;                   // Agent.currentTime
  long long holdings;                   // holdings[symbol]
  int last_trade;
  int bid, bid_vol, ask, ask_vol;       // known_bids[0] / known_asks[0], price -1 = empty
  unsigned flags;
  int ord_off;                          // open orders: slice of the instance's arena
  unsigned short ord_cnt, ord_cap;      // entries used (incl. removed ones) / reserved
  RngState rng;
  unsigned short ord_head, ord_live;    // first entry that may be live / live entries
  int type_idx;                         // absolute row in the types table
  union {
    struct { double r_t, sigma_t; long long prev_wake; int hint_id, hint; int stream_n, stream_B; } est;
    struct { long long wakeup_time; } noise;
    struct { double csum; int n_mids; int ring; double avg20, avg50; } mom;
    struct { double last_spread, window_size; long long transacted_volume;
             int buy_size, sell_size, tick_size, last_mid; } mm;
    struct { long long executed, transacted_volume; } pov;
    struct { long long executed; double arrival; } sx;
    struct { int row0, started; } rp;   // MarketReplayAgent: first tape row of the group the next wake places
  } u;
};
static_assert(sizeof(LAgent) == 128, "LAgent must be 128 bytes");

struct LN_ALIGN(16) LHot {              // per-instance scalars, in HBM; worked on in place (LANE_STAGE=1: staged in thread-local memory while a half runs)
  // first 64 bytes: everything the pop half reads or writes (loaded / stored as four 16-byte vectors)
  long long now, ttl, delivered, exch_busy;
  int hn;                               // events in the heap
  int n_hfree, slot_bump, spare_slot;   // payload slots: free stack, never-used frontier, one cached free slot
  int status, done, n_trace, queue_peak;
  // the rest: only the handle half (and kernelStopping)
  long long n_fills, volume;
  long long exch_cd, exch_cur_time, additional_delay;
  long long or_pt; double or_pv; long long ms_time; double ms_value;
  long long next_order_id;
  long long log_len, log_mark, unr_len;
  long long book_update_ts;
  unsigned long long pb[2];             // per side: lower bound of every pool key (price rank << 32 | arrival)
  unsigned uniq;
  int hd[2], n_top[2];                  // book deques: physical index of the best entry, entries
  int pl_total[2], pl_live[2], n_pfree[2];
  int last_hint;                        // pool slot + 1 of the order book_enter just placed (0: in the deque)
  unsigned book_arr;                    // arrival counter of resting orders (time priority)
  int last_trade, has_last_trade;
  int arena_top, book_peak_orders;
  int hist_B, tv_prev_len;
  int pad_[3];
  RngState rs[ABIDES_EXTRA_STREAMS];
};
constexpr int LHOT_POP_BYTES = 64;
static_assert(sizeof(LHot) % 16 == 0 && offsetof(LHot, n_fills) == LHOT_POP_BYTES, "LHot layout");

// everything a launch needs; passed BY VALUE as a __grid_constant__ kernel parameter (no __constant__ symbol shared
// between handles)
struct LaneBatch {
  int n_instances, trace_cap, n_types, top_k;
  int hcap, scap, arena_cap, deep_cap, txn_cap, unr_cap;
  int rng_mode;                         // enum abides_rng_mode
  int run_ahead;                        // light events an instance may run past its sorted one per round
  int trace_flags;                      // enum abides_trace_flags
  const abides_instance_cfg* cfg;
  const abides_agent_type* types;
  abm_dd* type_log1mk;
  double* type_omp2;
  const int* agent_type;
  const double* lat_to;
  const double* lat_from;
  const long long* size;
  const long long* wakeup_time;
  const int* agent_theta;
  const int* theta;
  const int* mom_index;
  unsigned* mt;                         // [rows][624]
  const int* mt_row;                    // [n_streams] row of a stream's state in mt, -1 = seed-only stream; NULL = identity
  const unsigned* mt_seed;              // [n_streams] init_genrand seed of seed-only / device-seeded streams, or NULL
  const int* mt_pos;
  const int* mt_has_gauss;
  const double* mt_gauss;
  LAgent* agents;
  LHot* state;
  LKey* hkeys;                          // [ni][hcap + HEAP_ROOT + 1]
  LPay* hpay;                           // [ni][scap]
  int* hfree;                           // [ni][scap]
  LOrd* top;                            // [ni][2][top_k]
  int* pl_price;                        // [ni][2][deep_cap]
  LOrd* pl_ord;                         // [ni][2][deep_cap]
  int* pl_free;                         // [ni][2][deep_cap]
  OpenOrd* arena;                       // [ni][arena_cap]
  TxnRow* txn_log;                      // [ni][txn_cap] or NULL
  TxnRow* unrolled;                     // [ni][unr_cap] or NULL
  double* mom_ring;
  abides_trace_rec* trace;
  abides_instance_result* res;
  abides_agent_result* ares;
  unsigned long long* prof;             // [LANE_PROF_SLOTS] phase cycle counters (profiling builds) or NULL
};

LN_HD inline int heap_stride(int hcap) { return (hcap + HEAP_ROOT + 1 + 7) & ~7; }      // whole 128-byte lines per instance

}  // namespace lane
#endif
