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
//   * event queue  : a small "near" tier in shared memory (all events below a moving key limit;
//                    pop = warp arg-min over <=128 keys) over an unsorted "far" list in HBM that is
//                    scanned, 32 keys per coalesced load, only when the near tier runs dry;
//   * order book   : both sides as price-time sorted arrays in shared memory; matching, level
//                    sums, inserts and cancels are ballot/popc scans and warp-wide shifts;
//   * agent state  : one 128-byte record per agent in HBM, moved to/from shared memory with one
//                    coalesced transaction per dispatch; handler code is warp-uniform;
//   * RNG          : per-stream MT19937 state in HBM (reference-exact mode, warp-parallel twist)
//                    or counter-based Philox4x32-10 (throughput mode).
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

namespace {

extern __shared__ __align__(128) unsigned char smem_raw[];

constexpr int NEAR_CAP = 128;
constexpr int FILL_MAX = 64;
constexpr int FILL_MIN = 16;
constexpr int BOOK_TOP = 128;          // orders per side kept in shared memory
constexpr int MOM_RING = 64;            // >= 51 prefix sums
constexpr int TXN_CAP = 1 << 16;        // ring of transaction rows since the last volume query (per instance)
constexpr int UNR_CAP = 1 << 16;        // ring of de-duplicated rows the volume queries sum over
constexpr int TRACE_FUNDAMENTAL = 64;   // trace code of an oracle step (time = ts, p0 = value)

struct __align__(16) EvKey { long long t; unsigned long long k2; };   // k2 = rcpt<<33 | wakeup<<32 | uniq
struct __align__(16) EvPay { int w[8]; };
struct __align__(16) BookOrd { int price, qty, id, agent; };
struct __align__(16) OpenOrd { int key, id, price, sqty; };            // sqty > 0 buy, < 0 sell
struct __align__(16) TxnRow { long long t; int q; int seq; };          // OrderBook.history transaction + bucket of its record

enum : unsigned {
  F_HAS_OPEN = 1u, F_HAS_CLOSE = 2u, F_MKT_CLOSED = 4u, F_HAS_LAST_TRADE = 8u, F_HAS_DAILY_CLOSE = 16u,
  F_KNOWN_BOOK = 32u, F_TRADING = 64u, F_HAS_PREV_WAKE = 128u, F_HAS_GAUSS = 256u,
  F_MM_AWAIT_SPREAD = 512u, F_MM_AWAIT_TV = 1024u, F_HAS_LAST_MID = 2048u, F_HAS_AVG20 = 4096u,
  F_HAS_AVG50 = 8192u, F_STATE_SHIFT = 16u, F_STATE_MASK = 3u << 16
};
enum { ST_AWAITING_WAKEUP = 0, ST_INACTIVE = 1, ST_AWAITING_SPREAD = 2, ST_ACTIVE = 3 };

struct __align__(128) AgentRec {        // 128 bytes: one coalesced transaction
  long long busy;                       // Kernel.agentCurrentTimes[a]
  long long cash;                       // holdings['CASH']
  long long cur_time;                   // Agent.currentTime
  long long holdings;                   // holdings[symbol]
  int last_trade;
  int bid, bid_vol, ask, ask_vol;       // known_bids[0] / known_asks[0], price -1 = empty
  unsigned flags;
  int ord_off, ord_cnt, ord_cap;        // open orders: slice of the instance's arena
  int rng_pos;                          // MT19937 position / Philox word counter
  double gauss;                         // numpy legacy cached gaussian
  int size;                             // Noise / Value / Momentum order size
  int type_idx;                         // absolute row in the types table
  union {
    struct { double r_t, sigma_t; long long prev_wake; } est;                     // ZI / Value
    struct { long long wakeup_time; } noise;
    struct { double csum; int n_mids; int ring; double avg20, avg50; } mom;
    struct { double last_spread, window_size; long long transacted_volume;
             int buy_size, sell_size, tick_size, last_mid; } mm;
  } u;
};
static_assert(sizeof(AgentRec) == 128, "AgentRec must be 128 bytes");

struct RngState { int pos; unsigned flags; double gauss; };

struct Hot {
  long long now, ttl, delivered, n_fills, volume;
  long long exch_busy, exch_cd, exch_cur_time, additional_delay;
  long long or_pt; double or_pv; long long ms_time; double ms_value;
  long long next_order_id;
  long long near_limit_t; unsigned long long near_limit_k; long long dT;
  unsigned uniq;
  int n_near, n_far_total, n_far_live;
  int n_bids, n_asks, nd_bids, nd_asks, last_trade, has_last_trade;
  int arena_top, status, n_trace, queue_peak, book_peak_orders, done;
  int hist_B, tv_prev_len;              // open history bucket; OrderBook._transacted_volume previous length
  long long log_len, log_mark, unr_len; // transaction rows logged / scanned by the last query / unrolled rows
  RngState rs[ABIDES_EXTRA_STREAMS];
};

struct __align__(128) InstShared {
  EvKey nkey[NEAR_CAP];
  EvPay npay[NEAR_CAP];
  BookOrd bids[BOOK_TOP];
  BookOrd asks[BOOK_TOP];
  int bid_seq[BOOK_TOP];                // history bucket in which each resting order was entered
  int ask_seq[BOOK_TOP];
  AgentRec ag;
  Hot hot;
};

struct DevBatch {
  int n_instances, rng_mode, trace_cap, far_cap, arena_cap, n_types, deep_cap;
  const abides_instance_cfg* cfg;
  const abides_agent_type* types;
  abm_dd* type_log1mk;                  // log(1 - kappa) per type, double-double
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
  EvKey* far_keys;
  EvPay* far_pay;
  OpenOrd* arena;
  BookOrd* deep_book;                   // [n_instances][2][deep_cap]
  int* deep_seq;                        // [n_instances][2][deep_cap]
  TxnRow* txn_log;                      // [n_instances][TXN_CAP]
  TxnRow* unrolled;                     // [n_instances][UNR_CAP]
  double* mom_ring;
  const int* mom_index;                 // per agent row: ring index or -1
  abides_trace_rec* trace;
  abides_instance_result* res;
  abides_agent_result* ares;
};

struct Run {                            // per-launch context, shared memory, not persisted
  DevBatch db;
  abides_instance_cfg cfg;
  EvKey* far_keys; EvPay* far_pay; OpenOrd* arena; BookOrd* deep; int* deep_seq; TxnRow* txn_log; TxnRow* unrolled;
  abides_trace_rec* trace;
  long long aoff;                       // agent_offset
  int inst, n_agents;
  int cur_agent;                        // agent whose record sits in st.ag (-1 none)
  int rp_op, rp_max, rp_base;           // order-book replay harness: current op, events per op, first op
  abides_book_event* rp_events; int* rp_counts;
};
struct __align__(128) Smem { InstShared st; Run run; };
#define SM (reinterpret_cast<Smem*>(smem_raw))
#define LANE ((int)threadIdx.x)

// ----------------------------------------------------------------------------------------------
// warp helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ long long warp_min_ll(long long v) {
  for (int o = 16; o; o >>= 1) { long long x = __shfl_xor_sync(FULL, v, o); v = x < v ? x : v; }
  return v;
}
__device__ __forceinline__ unsigned long long warp_min_ull(unsigned long long v) {
  for (int o = 16; o; o >>= 1) { unsigned long long x = __shfl_xor_sync(FULL, v, o); v = x < v ? x : v; }
  return v;
}
__device__ __forceinline__ unsigned long long warp_max_ull(unsigned long long v) {
  for (int o = 16; o; o >>= 1) { unsigned long long x = __shfl_xor_sync(FULL, v, o); v = x > v ? x : v; }
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ bool key_less(long long at, unsigned long long ak, long long bt, unsigned long long bk) {
  return at < bt || (at == bt && ak < bk);
}

// ----------------------------------------------------------------------------------------------
// RNG: numpy legacy RandomState transforms over MT19937 (HBM-resident state) or Philox4x32-10
// ----------------------------------------------------------------------------------------------
struct RngH { unsigned* mt; int* pos; double* gauss; unsigned* flagword; unsigned flagbit; unsigned long long key; };

__device__ __noinline__ void mt_twist(unsigned* mt, int lane) {
  const unsigned U = 0x80000000u, L = 0x7fffffffu, MAG = 0x9908b0dfu;
  __syncwarp();
  for (int base = 0; base < 227; base += 32) {
    int k = base + lane; bool act = k < 227; unsigned v = 0;
    if (act) { unsigned y = (mt[k] & U) | (mt[k + 1] & L); v = mt[k + 397] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u); }
    __syncwarp();
    if (act) mt[k] = v;
    __syncwarp();
  }
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
#pragma unroll
  for (int i = 0; i < 10; i++) {
    unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __noinline__ unsigned rng_u32(const RngH& r) {
  if (SM->run.db.rng_mode == ABIDES_RNG_PHILOX) {
    unsigned n = (unsigned)*r.pos;
    unsigned o[4];
    philox4x32_10(n >> 2, r.key, o);
    *r.pos = (int)(n + 1);
    unsigned sel = n & 3u;
    return sel == 0 ? o[0] : sel == 1 ? o[1] : sel == 2 ? o[2] : o[3];
  }
  int p = *r.pos;
  if (p >= ABIDES_MT_N) { mt_twist(r.mt, LANE); p = 0; }
  unsigned y = r.mt[p];
  *r.pos = p + 1;
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
__device__ double rng_double(const RngH& r) {
  int a = (int)(rng_u32(r) >> 5), b = (int)(rng_u32(r) >> 6);
  return (a * 67108864.0 + b) / 9007199254740992.0;
}
__device__ __noinline__ double rng_gauss(const RngH& r) {
  if (*r.flagword & r.flagbit) {
    double t = *r.gauss;
    *r.flagword &= ~r.flagbit;
    *r.gauss = 0.0;
    return t;
  }
  double f, x1, x2, r2;
  do {
    x1 = 2.0 * rng_double(r) - 1.0;
    x2 = 2.0 * rng_double(r) - 1.0;
    r2 = x1 * x1 + x2 * x2;
  } while (r2 >= 1.0 || r2 == 0.0);
  f = sqrt(-2.0 * abm_log(r2) / r2);
  *r.gauss = f * x1;
  *r.flagword |= r.flagbit;
  return f * x2;
}
__device__ double rng_normal(const RngH& r, double loc, double scale) { return loc + scale * rng_gauss(r); }
__device__ double rng_exponential(const RngH& r, double scale) {
  return scale * (-abm_log(1.0 - rng_double(r)));
}
__device__ double rng_uniform(const RngH& r, double lo, double hi) { return lo + (hi - lo) * rng_double(r); }
__device__ long long rng_randint(const RngH& r, long long lo, long long hi) {
  unsigned long long rng = (unsigned long long)(hi - 1 - lo);
  if (rng == 0) return lo;
  if (rng == 0xFFFFFFFFull) return lo + (long long)rng_u32(r);
  unsigned mask = (unsigned)rng;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  unsigned v;
  while ((v = (rng_u32(r) & mask)) > (unsigned)rng) {}
  return lo + (long long)v;
}

__device__ __forceinline__ unsigned long long philox_key(unsigned long long seed, int stream) {
  return seed * 0x9E3779B97F4A7C15ull + (unsigned long long)stream * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull;
}
__device__ __forceinline__ RngH agent_rng() {      // stream of the agent currently in S->ag
  RngH r;
  AgentRec& A = SM->st.ag;
  r.mt = SM->run.db.mt ? SM->run.db.mt + (size_t)(SM->run.aoff + (long long)ABIDES_EXTRA_STREAMS * SM->run.inst + SM->run.cur_agent) * ABIDES_MT_N : nullptr;
  r.pos = &A.rng_pos; r.gauss = &A.gauss; r.flagword = &A.flags; r.flagbit = F_HAS_GAUSS;
  r.key = philox_key(SM->run.cfg.seed, SM->run.cur_agent);
  return r;
}
__device__ __forceinline__ RngH extra_rng(int which) {
  RngH r;
  RngState& s = SM->st.hot.rs[which];
  int stream = SM->run.n_agents + which;
  r.mt = SM->run.db.mt ? SM->run.db.mt + (size_t)(SM->run.aoff + (long long)ABIDES_EXTRA_STREAMS * SM->run.inst + stream) * ABIDES_MT_N : nullptr;
  r.pos = &s.pos; r.gauss = &s.gauss; r.flagword = &s.flags; r.flagbit = 1u;
  r.key = philox_key(SM->run.cfg.seed, stream);
  return r;
}

// ----------------------------------------------------------------------------------------------
// trace
// ----------------------------------------------------------------------------------------------
__device__ __noinline__ void trace_write(long long time, int agent, int code, long long p0, long long p1, long long p2, long long p3) {
  Hot& h = SM->st.hot;
  int i = h.n_trace;
  h.n_trace = i + 1;
  if (!SM->run.trace || i >= SM->run.db.trace_cap) return;
  if (LANE == 0) {
    abides_trace_rec r;
    r.time = time; r.agent = agent; r.code = code; r.p0 = p0; r.p1 = p1; r.p2 = p2; r.p3 = p3;
    SM->run.trace[i] = r;
  }
  __syncwarp();
}

// ----------------------------------------------------------------------------------------------
// event queue: near tier (shared memory) over an unsorted far list (HBM)
// ----------------------------------------------------------------------------------------------
__device__ __noinline__ void far_compact() {
  Hot& h = SM->st.hot;
  int n = h.n_far_total, w = 0;
  __syncwarp();
  for (int base = 0; base < n; base += 32) {
    int i = base + LANE;
    bool live = false; EvKey k; EvPay p;
    if (i < n) { k = SM->run.far_keys[i]; live = k.t != LLONG_MAX; if (live) p = SM->run.far_pay[i]; }
    unsigned m = __ballot_sync(FULL, live);
    __syncwarp();
    if (live) {
      int j = w + __popc(m & ((1u << LANE) - 1u));
      if (j != i) { SM->run.far_keys[j] = k; SM->run.far_pay[j] = p; }
    }
    w += __popc(m);
    __syncwarp();
  }
  h.n_far_total = w;
}

__device__ __noinline__ void far_append(const EvKey& k, const EvPay& p) {
  Hot& h = SM->st.hot;
  if (h.n_far_total >= SM->run.db.far_cap) {
    far_compact();
    if (h.n_far_total >= SM->run.db.far_cap) { h.status = ABIDES_EOVERFLOW; return; }
  }
  int i = h.n_far_total;
  if (LANE == 0) SM->run.far_keys[i] = k;
  else if (LANE == 1) *reinterpret_cast<int4*>(&SM->run.far_pay[i].w[0]) = make_int4(p.w[0], p.w[1], p.w[2], p.w[3]);
  else if (LANE == 2) *reinterpret_cast<int4*>(&SM->run.far_pay[i].w[4]) = make_int4(p.w[4], p.w[5], p.w[6], p.w[7]);
  __syncwarp();
  h.n_far_total = i + 1;
  h.n_far_live++;
}

// arg-min (or arg-max) of the near keys; ties broken by slot index so the result is warp-uniform
__device__ __noinline__ int near_argext(bool want_max) {
  const InstShared* S = &SM->st;
  int n = S->hot.n_near;
  long long bt = 0; unsigned long long bk = 0; int bi = -1;
  __syncwarp();
  for (int i = LANE; i < n; i += 32) {
    EvKey k = S->nkey[i];
    bool better = bi < 0 || (want_max ? key_less(bt, bk, k.t, k.k2) : key_less(k.t, k.k2, bt, bk));
    if (better) { bt = k.t; bk = k.k2; bi = i; }
  }
  for (int o = 16; o; o >>= 1) {
    long long ot = __shfl_xor_sync(FULL, bt, o);
    unsigned long long ok = __shfl_xor_sync(FULL, bk, o);
    int oi = __shfl_xor_sync(FULL, bi, o);
    bool better;
    if (oi < 0) better = false;
    else if (bi < 0) better = true;
    else if (ot == bt && ok == bk) better = oi < bi;
    else better = want_max ? key_less(bt, bk, ot, ok) : key_less(ot, ok, bt, bk);
    if (better) { bt = ot; bk = ok; bi = oi; }
  }
  return bi;
}

__device__ __noinline__ void near_remove(int idx) {
  InstShared* S = &SM->st;
  int last = S->hot.n_near - 1;
  if (idx != last) { S->nkey[idx] = S->nkey[last]; S->npay[idx] = S->npay[last]; }
  S->hot.n_near = last;
  __syncwarp();
}

__device__ __noinline__ void q_push(const EvKey& k, const EvPay& p) {
  InstShared* S = &SM->st;
  Hot& h = S->hot;
  if (key_less(k.t, k.k2, h.near_limit_t, h.near_limit_k)) {
    if (h.n_near == NEAR_CAP) {
      // near tier full: move its largest key out to the far list and pull the limit down to it
      int mi = near_argext(true);
      EvKey mk = S->nkey[mi]; EvPay mp = S->npay[mi];
      near_remove(mi);
      far_append(mk, mp);
      h.near_limit_t = mk.t; h.near_limit_k = mk.k2;
      if (!key_less(k.t, k.k2, mk.t, mk.k2)) { far_append(k, p); goto done; }
    }
    int n = h.n_near;
    S->nkey[n] = k; S->npay[n] = p;
    h.n_near = n + 1;
  } else {
    far_append(k, p);
  }
done:
  int tot = h.n_near + h.n_far_live;
  if (tot > h.queue_peak) h.queue_peak = tot;
}

__device__ __noinline__ int far_count_lt(int n, long long pt, unsigned long long pk) {
  int cnt = 0;
  for (int i = LANE; i < n; i += 32) {
    EvKey k = SM->run.far_keys[i];
    cnt += key_less(k.t, k.k2, pt, pk) ? 1 : 0;
  }
  return warp_sum_i(cnt);
}

// near tier is empty: pull the next batch (<= FILL_MAX smallest keys) out of the far list
__device__ __noinline__ void q_refill() {
  InstShared* S = &SM->st;
  Hot& h = S->hot;
  if (h.n_far_total > 1024 && h.n_far_live * 2 < h.n_far_total) far_compact();
  int n = h.n_far_total;
  __syncwarp();
  long long tmin = LLONG_MAX;
  for (int i = LANE; i < n; i += 32) { long long t = SM->run.far_keys[i].t; tmin = t < tmin ? t : tmin; }
  tmin = warp_min_ll(tmin);
  long long dT = h.dT, pt; unsigned long long pk = 0; int cnt;
  bool burst = false;
  for (;;) {
    pt = (tmin > LLONG_MAX - 2 - dT) ? LLONG_MAX - 1 : tmin + dT;
    cnt = far_count_lt(n, pt, 0);
    if (cnt <= FILL_MAX) break;
    if (dT <= 1) { burst = true; break; }
    dT >>= 2; if (dT < 1) dT = 1;
  }
  if (burst) {
    // more than FILL_MAX events share the timestamp tmin: split them on the secondary key
    // (recipient, kind, uniq).  Interpolate first (recipients are spread evenly), then bisect.
    unsigned long long kmin = ~0ull, kmax = 0;
    for (int i = LANE; i < n; i += 32) {
      EvKey k = SM->run.far_keys[i];
      if (k.t == tmin) { kmin = k.k2 < kmin ? k.k2 : kmin; kmax = k.k2 > kmax ? k.k2 : kmax; }
    }
    kmin = warp_min_ull(kmin); kmax = warp_max_ull(kmax);
    unsigned long long lo = kmin, hi = kmax + 1;          // count(<lo) == 0, count(<hi) == cnt > FILL_MAX
    unsigned long long best = 0; int best_cnt = 0;
    double frac = (double)(FILL_MAX * 3 / 4) / (double)cnt;
    unsigned long long guess = kmin + (unsigned long long)((double)(kmax - kmin) * frac) + 1;
    for (int it = 0; it < 80; it++) {
      unsigned long long mid = it == 0 ? guess : lo + (hi - lo) / 2;
      if (mid <= lo) mid = lo + 1;
      if (mid >= hi) break;
      int cc = far_count_lt(n, tmin, mid);
      if (cc > FILL_MAX) hi = mid;
      else { lo = mid; if (cc > best_cnt) { best = mid; best_cnt = cc; } if (cc >= FILL_MIN) break; }
    }
    if (best_cnt == 0) { best = hi; best_cnt = far_count_lt(n, tmin, hi); }   // identical keys
    pt = tmin; pk = best; cnt = best_cnt;
    if (cnt > NEAR_CAP) { h.status = ABIDES_EOVERFLOW; cnt = 0; pk = 0; }
  }
  int nn = h.n_near;
  for (int base = 0; base < n; base += 32) {
    int i = base + LANE;
    bool take = false; EvKey k;
    if (i < n) { k = SM->run.far_keys[i]; take = key_less(k.t, k.k2, pt, pk); }
    unsigned m = __ballot_sync(FULL, take);
    if (m) {
      if (take) {
        int pos = nn + __popc(m & ((1u << LANE) - 1u));
        S->nkey[pos] = k; S->npay[pos] = SM->run.far_pay[i];
        SM->run.far_keys[i].t = LLONG_MAX;
      }
      nn += __popc(m);
    }
  }
  __syncwarp();
  h.n_near = nn;
  h.n_far_live -= cnt;
  h.near_limit_t = pt; h.near_limit_k = pk;
  if (!burst) {
    if (cnt < FILL_MIN && dT < (1ll << 50)) dT <<= 1;
    h.dT = dT;
  }
  __syncwarp();
}

// ----------------------------------------------------------------------------------------------
// Kernel.sendMessage / setWakeup  (Kernel.py:329-418)
// ----------------------------------------------------------------------------------------------
__device__ __noinline__ void k_send(int sender, int rcpt, const EvPay& pay, long long delay) {
  Hot& h = SM->st.hot;
  const abides_instance_cfg* cfg = &SM->run.cfg;
  long long cd = sender == 0 ? h.exch_cd : cfg->default_computation_delay;
  long long sent = h.now + cd + h.additional_delay + delay;
  int other = sender == 0 ? rcpt : sender;
  double minlat = sender == 0 ? SM->run.db.lat_from[SM->run.aoff + other] : SM->run.db.lat_to[SM->run.aoff + other];
  long long lat;
  if (cfg->latency_kind == ABIDES_LAT_DETERMINISTIC) {
    lat = (long long)minlat;
  } else if (cfg->latency_kind == ABIDES_LAT_CUBIC) {
    double x = rng_uniform(extra_rng(ABIDES_STREAM_LATENCY), cfg->jitter_clip, 1.0);
    double l = minlat + ((cfg->jitter / abm_cube(x)) * (minlat / cfg->jitter_unit));
    lat = (long long)l;
  } else {
    long long noise = rng_randint(extra_rng(ABIDES_STREAM_KERNEL), 0, cfg->n_latency_noise);
    lat = (long long)(minlat + (double)noise);
  }
  EvKey k;
  k.t = sent + lat;
  k.k2 = ((unsigned long long)(unsigned)rcpt << 33) | (unsigned long long)h.uniq;
  h.uniq++;
  q_push(k, pay);
}
__device__ __noinline__ void k_wakeup(int agent, long long t) {
  Hot& h = SM->st.hot;
  if (t < h.now) { h.status = ABIDES_EINVAL; return; }
  EvKey k; k.t = t; k.k2 = ((unsigned long long)(unsigned)agent << 33) | (1ull << 32);
  EvPay p;
#pragma unroll
  for (int i = 0; i < 8; i++) p.w[i] = 0;
  q_push(k, p);
}

// ----------------------------------------------------------------------------------------------
// oracle (util/oracle/SparseMeanRevertingOracle.py)
// ----------------------------------------------------------------------------------------------
__device__ __noinline__ double oracle_step(long long ts, double v_adj, long long pt, double pv) {
  const abides_instance_cfg* cfg = &SM->run.cfg;
  Hot& h = SM->st.hot;
  double d = (double)(ts - pt);
  double mu = cfg->r_bar, gamma = cfg->kappa, theta = cfg->fund_vol;
  double e1 = abm_exp(-gamma * d);
  double loc = mu + (pv - mu) * e1;
  double e2 = abm_exp(-2 * gamma * d);
  double scale = (theta / (2 * gamma)) * (1 - e2);
  double v = rng_normal(extra_rng(ABIDES_STREAM_ORACLE), loc, scale);
  v += v_adj;
  if (!(v > 0)) v = 0;
  v = rint(v);
  h.or_pt = ts; h.or_pv = v;
  if (SM->run.trace) trace_write(ts, -1, TRACE_FUNDAMENTAL, (long long)v, 0, 0, 0);
  return v;
}
__device__ __noinline__ double oracle_advance(long long t) {
  const abides_instance_cfg* cfg = &SM->run.cfg;
  Hot& h = SM->st.hot;
  long long pt = h.or_pt; double pv = h.or_pv;
  if (t <= pt) return pv;
  while (h.ms_time < t) {
    double v = oracle_step(h.ms_time, h.ms_value, pt, pv);
    pt = h.ms_time; pv = v;
    double dt = rng_exponential(extra_rng(ABIDES_STREAM_GLOBAL), 1.0 / cfg->megashock_lambda_a);
    h.ms_time = pt + (long long)dt;
    double msv = rng_normal(extra_rng(ABIDES_STREAM_ORACLE), cfg->megashock_mean, sqrt(cfg->megashock_var));
    h.ms_value = rng_randint(extra_rng(ABIDES_STREAM_ORACLE), 0, 2) == 0 ? msv : -msv;
  }
  return oracle_step(t, 0.0, pt, pv);
}
__device__ __noinline__ long long oracle_observe(long long t, double sigma_n, const RngH& rs) {
  double r_t = (t >= SM->run.cfg.mkt_close) ? oracle_advance(SM->run.cfg.mkt_close - 1) : oracle_advance(t);
  if (sigma_n == 0) return (long long)r_t;
  return (long long)rint(rng_normal(rs, r_t, sqrt(sigma_n)));
}

// ----------------------------------------------------------------------------------------------
// payload helpers.  w[0] = code | mkt_closed<<8 | is_buy<<9
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ EvPay pay_zero(int code) {
  EvPay p;
#pragma unroll
  for (int i = 0; i < 8; i++) p.w[i] = 0;
  p.w[0] = code;
  return p;
}
__device__ __forceinline__ int pay_code(const EvPay& p) { return p.w[0] & 0xff; }
__device__ __forceinline__ bool pay_closed(const EvPay& p) { return (p.w[0] >> 8) & 1; }
__device__ __forceinline__ bool pay_is_buy(const EvPay& p) { return (p.w[0] >> 9) & 1; }
// order-carrying payloads: w[1] sender (to exchange), w[2] id, w[3] price, w[4] qty, w[5] fill price / new qty
__device__ __forceinline__ EvPay pay_order(int code, int sender, int id, int price, int qty, bool is_buy, int fill) {
  EvPay p = pay_zero(code | (is_buy ? 1 << 9 : 0));
  p.w[1] = sender; p.w[2] = id; p.w[3] = price; p.w[4] = qty; p.w[5] = fill;
  return p;
}

// ----------------------------------------------------------------------------------------------
// order book (util/OrderBook.py).  Each side is ONE price-time sorted sequence (index 0 = best
// price, oldest first).  Its best BOOK_TOP entries live in shared memory; the rest -- the "deep
// book" -- is spilled to HBM in REVERSE order (deep[nd-1] is the best deep entry), so that moving
// entries across the boundary is an O(1) push/pop at the array's end and every scan or shift over
// the deep part is a run of coalesced 512-byte warp transactions.
// ----------------------------------------------------------------------------------------------
struct Side { BookOrd* top; int* tseq; int* ns; BookOrd* deep; int* dseq; int* nd; bool is_buy; };
__device__ __forceinline__ Side book_side(bool is_buy) {
  Side sd; InstShared* S = &SM->st;
  sd.top = is_buy ? S->bids : S->asks;
  sd.tseq = is_buy ? S->bid_seq : S->ask_seq;
  sd.ns = is_buy ? &S->hot.n_bids : &S->hot.n_asks;
  sd.deep = SM->run.deep + (is_buy ? 0 : SM->run.db.deep_cap);
  sd.dseq = SM->run.deep_seq + (is_buy ? 0 : SM->run.db.deep_cap);
  sd.nd = is_buy ? &S->hot.nd_bids : &S->hot.nd_asks;
  sd.is_buy = is_buy;
  return sd;
}
// generic warp-parallel single-slot shifts on an (order, bucket) array pair in shared or global memory
__device__ __noinline__ void arr_remove(BookOrd* s, int* q, int n, int idx) {   // delete slot idx from [0..n)
  __syncwarp();
  for (int base = idx; base < n - 1; base += 32) {
    int j = base + LANE; BookOrd t; int tq = 0;
    bool act = j < n - 1;
    if (act) { t = s[j + 1]; tq = q[j + 1]; }
    __syncwarp();
    if (act) { s[j] = t; q[j] = tq; }
    __syncwarp();
  }
}
__device__ __noinline__ void arr_insert(BookOrd* s, int* q, int n, int idx, const BookOrd& o, int oseq) {
  __syncwarp();
  int cnt = n - idx;
  for (int done = 0; done < cnt; done += 32) {
    int j = n - 1 - done - LANE; BookOrd t; int tq = 0;
    bool act = j >= idx;
    if (act) { t = s[j]; tq = q[j]; }
    __syncwarp();
    if (act) { s[j + 1] = t; q[j + 1] = tq; }
    __syncwarp();
  }
  if (LANE == 0) { s[idx] = o; q[idx] = oseq; }
  __syncwarp();
}
// count entries of s[0..n) whose price is better than or equal to `price` for this side
__device__ __noinline__ int arr_count_better_eq(const BookOrd* s, int n, int price, bool is_buy) {
  __syncwarp();
  int cnt = 0;
  for (int j = LANE; j < n; j += 32) {
    int p = s[j].price;
    cnt += (is_buy ? p >= price : p <= price) ? 1 : 0;
  }
  return warp_sum_i(cnt);
}
// the shared-memory part ran empty: pull the best entries of the deep book back in
__device__ __noinline__ void side_refill(const Side& sd) {
  int nd = *sd.nd, ns = *sd.ns;
  if (ns > 0 || nd == 0) return;
  int k = nd < BOOK_TOP / 2 ? nd : BOOK_TOP / 2;
  __syncwarp();
  for (int i = LANE; i < k; i += 32) { sd.top[i] = sd.deep[nd - 1 - i]; sd.tseq[i] = sd.dseq[nd - 1 - i]; }
  __syncwarp();
  *sd.ns = k; *sd.nd = nd - k;
}
__device__ __noinline__ void side_pop_front(const Side& sd) {                   // remove the best entry
  arr_remove(sd.top, sd.tseq, *sd.ns, 0);
  *sd.ns -= 1;
  side_refill(sd);
}
// getInsideBids(1)/getInsideAsks(1) (:378-399): price and total quantity of the best level
__device__ __noinline__ void side_inside(const Side& sd, int& price, int& vol) {
  __syncwarp();
  int ns = *sd.ns, nd = *sd.nd;
  if (ns == 0) { price = -1; vol = 0; return; }
  int p = sd.top[0].price, v = 0;
  bool whole = true;                       // does the level run to the end of the shared part?
  for (int base = 0; base < ns; base += 32) {
    int j = base + LANE;
    bool in = j < ns && sd.top[j].price == p;
    v += in ? sd.top[j].qty : 0;
    if (__any_sync(FULL, j < ns && !in)) { whole = false; break; }
  }
  if (whole) {
    for (int base = 0; base < nd; base += 32) {
      int j = nd - 1 - base - LANE;
      bool in = j >= 0 && sd.deep[j].price == p;
      v += in ? sd.deep[j].qty : 0;
      if (__any_sync(FULL, j >= 0 && !in)) break;
    }
  }
  price = p; vol = warp_sum_i(v);
}
// OrderBook.enterOrder (:272-298): position = number of resting orders with better-or-equal price
__device__ __noinline__ void book_enter(const BookOrd& o, bool is_buy) {
  Hot& h = SM->st.hot;
  Side sd = book_side(is_buy);
  int ns = *sd.ns, nd = *sd.nd;
  int oseq = h.hist_B;                                   // history[0][order_id] = {...} (:60-64)
  int cs = arr_count_better_eq(sd.top, ns, o.price, is_buy);
  bool to_top = cs < ns || nd == 0;
  int cd = 0;
  if (!to_top) {
    cd = arr_count_better_eq(sd.deep, nd, o.price, is_buy);
    if (cd == 0 && ns < BOOK_TOP) to_top = true;        // worse than all of top, better than all of deep
  }
  if (to_top && ns == BOOK_TOP && cs == ns) to_top = false;   // full shared part and o is its new worst: goes deep
  if (to_top) {
    if (ns == BOOK_TOP) {                                // spill the worst shared entry to the deep book's best end
      if (nd >= SM->run.db.deep_cap) { h.status = ABIDES_EOVERFLOW; return; }
      __syncwarp();
      if (LANE == 0) { sd.deep[nd] = sd.top[ns - 1]; sd.dseq[nd] = sd.tseq[ns - 1]; }
      __syncwarp();
      *sd.nd = nd + 1; ns--;
    }
    arr_insert(sd.top, sd.tseq, ns, cs, o, oseq);
    *sd.ns = ns + 1;
  } else {
    if (nd >= SM->run.db.deep_cap) { h.status = ABIDES_EOVERFLOW; return; }
    arr_insert(sd.deep, sd.dseq, nd, nd - cd, o, oseq);  // logical index cd <-> physical nd - cd (reversed)
    *sd.nd = nd + 1;
  }
  int tot = h.n_bids + h.n_asks + h.nd_bids + h.nd_asks;
  if (tot > h.book_peak_orders) h.book_peak_orders = tot;
}
// find (price, id) on a side; returns 0 = absent, 1 = in the shared part, 2 = in the deep part; idx physical
__device__ __noinline__ int side_find(const Side& sd, int price, int id, int& idx) {
  __syncwarp();
  int ns = *sd.ns, nd = *sd.nd;
  for (int base = 0; base < ns; base += 32) {
    int j = base + LANE;
    bool hit = j < ns && sd.top[j].price == price && sd.top[j].id == id;
    unsigned m = __ballot_sync(FULL, hit);
    if (m) { idx = base + __ffs(m) - 1; return 1; }
  }
  // deep part, scanned from its best end (highest physical index) so the FIRST match in book order wins
  for (int base = 0; base < nd; base += 32) {
    int j = nd - 1 - base - LANE;
    bool hit = false;
    if (j >= 0) { BookOrd e = sd.deep[j]; hit = e.price == price && e.id == id; }
    unsigned m = __ballot_sync(FULL, hit);
    if (m) { idx = nd - 1 - base - (__ffs(m) - 1); return 2; }
  }
  return 0;
}

// ---- OrderBook.history / get_transacted_volume (:401-471) ------------------------------------
// The reference keeps, per history bucket (one per trade-producing order), the (time, qty) transactions
// of the orders ENTERED in that bucket, and a volume query unrolls only a window of "recent" buckets.
// Equivalent device form: every transaction is one TxnRow tagged with the bucket of its record; a query
// scans the rows logged since the previous query, keeps those whose bucket lies in the window, drops
// duplicate (time, qty) rows of that batch and appends the rest to the unrolled ring it then sums over.
__device__ __forceinline__ void txn_log_row(long long t, int q, int seq) {
  Hot& h = SM->st.hot;
  long long i = h.log_len;
  __syncwarp();
  if (LANE == 0 && SM->run.txn_log) { TxnRow r; r.t = t; r.q = q; r.seq = seq; SM->run.txn_log[i & (TXN_CAP - 1)] = r; }
  __syncwarp();
  h.log_len = i + 1;
}
__device__ __noinline__ long long book_transacted_volume(long long now, long long lookback) {
  Hot& h = SM->st.hot;
  const int SH = SM->run.cfg.stream_history;
  int B = h.hist_B;
  int n = (B + 1 < SH + 1) ? B + 1 : SH + 1, p = h.tv_prev_len;
  int lo = 1, hi = 0;
  if (p == 0) { lo = B - (n - 1); hi = B; h.tv_prev_len = n; }
  else if (p != n) { int cnt = n - p - 1; if (cnt > 0) { lo = B - cnt + 1; hi = B; } h.tv_prev_len = n; }
  long long start = h.log_mark, end = h.log_len;
  if (end - start > TXN_CAP) { h.status = ABIDES_EOVERFLOW; return 0; }
  const TxnRow* lg = SM->run.txn_log;
  TxnRow* un = SM->run.unrolled;
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

// OrderBook.handleLimitOrder (:46-158)
template <typename Send>
__device__ __noinline__ void book_handle_limit(Send send, int agent, int id, int price, int qty, bool is_buy) {
  Hot& h = SM->st.hot;
  if (qty <= 0) return;
  long long ex_qty = 0, ex_pq = 0; int executed = 0;
  Side opp = book_side(!is_buy);
  for (;;) {
    __syncwarp();
    bool match = false; BookOrd top;
    if (*opp.ns > 0) { top = opp.top[0]; match = is_buy ? price >= top.price : price <= top.price; }
    if (match) {                                               // executeOrder (:190-256)
      int fill, tseq = opp.tseq[0];
      if (qty >= top.qty) { fill = top.qty; side_pop_front(opp); }
      else { fill = qty; __syncwarp(); if (LANE == 0) opp.top[0].qty = top.qty - qty; __syncwarp(); }
      // history (:245-253): the aggressor logs its REMAINING quantity, the resting order the fill
      txn_log_row(h.now, qty, h.hist_B);
      if (h.hist_B - tseq <= SM->run.cfg.stream_history) txn_log_row(h.now, fill, tseq);
      qty -= fill;
      send(agent, pay_order(ABIDES_MSG_ORDER_EXECUTED, 0, id, price, fill, is_buy, top.price));
      send(top.agent, pay_order(ABIDES_MSG_ORDER_EXECUTED, 0, top.id, top.price, fill, !is_buy, top.price));
      ex_qty += fill; ex_pq += (long long)top.price * fill; executed++;
      h.n_fills++; h.volume += fill;
      if (qty <= 0) break;
    } else {
      BookOrd o; o.price = price; o.qty = qty; o.id = id; o.agent = agent;
      book_enter(o, is_buy);
      send(agent, pay_order(ABIDES_MSG_ORDER_ACCEPTED, 0, id, price, qty, is_buy, -1));
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
template <typename Send>
__device__ __noinline__ void book_cancel(Send send, int agent, int id, int price, bool is_buy) {
  Side sd = book_side(is_buy);
  int idx = 0;
  int where = side_find(sd, price, id, idx);
  if (where == 0) return;
  BookOrd o;
  if (where == 1) {
    o = sd.top[idx];
    arr_remove(sd.top, sd.tseq, *sd.ns, idx);
    *sd.ns -= 1;
    side_refill(sd);
  } else {
    o = sd.deep[idx];
    arr_remove(sd.deep, sd.dseq, *sd.nd, idx);
    *sd.nd -= 1;
  }
  send(agent, pay_order(ABIDES_MSG_ORDER_CANCELLED, 0, o.id, o.price, o.qty, is_buy, -1));
}

// OrderBook.handleMarketOrder (:160-188): one fresh limit order per opposite level until the quantity is
// covered.  Each child consumes exactly the level it was sized from, so walking the live book level by
// level is equivalent to the reference's snapshot walk; an uncovered remainder is dropped.
template <typename Send>
__device__ __noinline__ void book_handle_market(Send send, int agent, int qty, bool is_buy) {
  Hot& h = SM->st.hot;
  if (qty <= 0) return;
  int rem = qty;
  for (;;) {
    int p, v;
    side_inside(book_side(!is_buy), p, v);
    if (p < 0) break;
    int q = rem <= v ? rem : v;
    int id = (int)h.next_order_id++;
    book_handle_limit(send, agent, id, p, q, is_buy);
    if (rem <= v) break;
    rem -= v;
  }
}
// OrderBook.modifyOrder (:350-373), including the reference's index-0 overwrite (:359)
template <typename Send>
__device__ __noinline__ void book_modify(Send send, int agent, int id, int price, int new_qty, bool is_buy) {
  Hot& h = SM->st.hot;
  Side sd = book_side(is_buy);
  int idx = 0;
  int where = side_find(sd, price, id, idx);
  if (where == 0) return;
  int oseq = where == 1 ? sd.tseq[idx] : sd.dseq[idx];
  // first entry of that price level = number of entries with a strictly better price
  int ns = *sd.ns, nd = *sd.nd;
  int better = 0;
  __syncwarp();
  for (int j = LANE; j < ns; j += 32) { int pp = sd.top[j].price; better += (is_buy ? pp > price : pp < price) ? 1 : 0; }
  for (int j = LANE; j < nd; j += 32) { int pp = sd.deep[j].price; better += (is_buy ? pp > price : pp < price) ? 1 : 0; }
  better = warp_sum_i(better);
  BookOrd no; no.price = price; no.qty = new_qty; no.id = id; no.agent = agent;
  __syncwarp();
  if (LANE == 0) {
    if (better < ns) { sd.top[better] = no; sd.tseq[better] = oseq; }
    else { sd.deep[nd - 1 - (better - ns)] = no; sd.dseq[nd - 1 - (better - ns)] = oseq; }
  }
  __syncwarp();
  if (h.hist_B - oseq <= SM->run.cfg.stream_history)       // one notification per history bucket holding the id
    send(agent, pay_order(ABIDES_MSG_ORDER_MODIFIED, 0, id, price, new_qty, is_buy, -1));
}
__device__ __noinline__ int side_levels(const Side& sd) {                      // number of distinct prices
  int ns = *sd.ns, nd = *sd.nd, c = 0;
  __syncwarp();
  for (int j = LANE; j < ns + nd; j += 32) {
    int p = j < ns ? sd.top[j].price : sd.deep[nd - 1 - (j - ns)].price;
    int q = j == 0 ? p - 1 : (j - 1 < ns ? sd.top[j - 1].price : sd.deep[nd - 1 - (j - 1 - ns)].price);
    c += (j == 0 || p != q) ? 1 : 0;
  }
  return warp_sum_i(c);
}

// ----------------------------------------------------------------------------------------------
// TradingAgent (agent/TradingAgent.py)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ int ag_state(const AgentRec& A) { return (A.flags & F_STATE_MASK) >> F_STATE_SHIFT; }
__device__ __forceinline__ void ag_set_state(AgentRec& A, int s) { A.flags = (A.flags & ~F_STATE_MASK) | ((unsigned)s << F_STATE_SHIFT); }

__device__ void agent_send(int a, EvPay p) { p.w[1] = a; k_send(a, 0, p, 0); }
__device__ void ta_query_spread(int a) { EvPay p = pay_zero(ABIDES_MSG_QUERY_SPREAD); p.w[2] = 1; agent_send(a, p); }

__device__ __noinline__ void ord_reserve(AgentRec& A, int initial_cap) {
  Hot& h = SM->st.hot;
  if (A.ord_cnt < A.ord_cap) return;
  int ncap = A.ord_cap ? A.ord_cap * 2 : initial_cap;
  if (h.arena_top + ncap > SM->run.db.arena_cap) { h.status = ABIDES_EOVERFLOW; return; }
  int noff = h.arena_top;
  h.arena_top += ncap;
  __syncwarp();
  for (int j = LANE; j < A.ord_cnt; j += 32) SM->run.arena[noff + j] = SM->run.arena[A.ord_off + j];
  __syncwarp();
  A.ord_off = noff; A.ord_cap = ncap;
}
// TradingAgent.placeLimitOrder (:295-330), ignore_risk=True
__device__ __noinline__ void ta_place_limit(int a, long long qty, bool is_buy, long long price) {
  AgentRec& A = SM->st.ag; Hot& h = SM->st.hot;
  const abides_agent_type& ty = SM->run.db.types[A.type_idx];
  int id = (int)h.next_order_id++;
  if (qty <= 0) return;
  OpenOrd oo; oo.key = id; oo.id = id; oo.price = (int)price; oo.sqty = is_buy ? (int)qty : -(int)qty;
  if (id == 0) oo.id = (int)h.next_order_id++;            // deepcopy of order id 0 takes a fresh id (Order.py:29)
  int init_cap = ty.klass == ABIDES_ADAPTIVE_MM ? (int)(2 * ty.num_ticks + 8) : (ty.klass == ABIDES_MOMENTUM ? 16 : 4);
  ord_reserve(A, init_cap);
  if (h.status) return;
  __syncwarp();
  if (LANE == 0) SM->run.arena[A.ord_off + A.ord_cnt] = oo;
  __syncwarp();
  A.ord_cnt++;
  agent_send(a, pay_order(ABIDES_MSG_LIMIT_ORDER, a, id, (int)price, (int)qty, is_buy, -1));
  if (ty.log_orders && id == 0) h.next_order_id++;
}
__device__ __noinline__ void ta_cancel_all(int a) {
  AgentRec& A = SM->st.ag;
  __syncwarp();
  for (int i = 0; i < A.ord_cnt; i++) {
    OpenOrd o = SM->run.arena[A.ord_off + i];
    agent_send(a, pay_order(ABIDES_MSG_CANCEL_ORDER, a, o.id, o.price, o.sqty < 0 ? -o.sqty : o.sqty, o.sqty > 0, -1));
  }
}
__device__ __noinline__ int ord_find(const AgentRec& A, int key) {
  __syncwarp();
  for (int base = 0; base < A.ord_cnt; base += 32) {
    int j = base + LANE;
    bool hit = j < A.ord_cnt && SM->run.arena[A.ord_off + j].key == key;
    unsigned m = __ballot_sync(FULL, hit);
    if (m) return base + __ffs(m) - 1;
  }
  return -1;
}
__device__ __noinline__ void ord_remove(AgentRec& A, int idx) {
  OpenOrd* s = SM->run.arena + A.ord_off;
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
  A.ord_cnt = n - 1;
}
// TradingAgent.wakeup (:152-168); returns can_trade
__device__ bool ta_wakeup(int a) {
  AgentRec& A = SM->st.ag;
  A.cur_time = SM->st.hot.now;
  if (!(A.flags & F_HAS_OPEN)) {
    agent_send(a, pay_zero(ABIDES_MSG_WHEN_MKT_OPEN));
    agent_send(a, pay_zero(ABIDES_MSG_WHEN_MKT_CLOSE));
  }
  return (A.flags & F_HAS_OPEN) && (A.flags & F_HAS_CLOSE) && !(A.flags & F_MKT_CLOSED);
}
// TradingAgent.receiveMessage (:181-264)
__device__ __noinline__ void ta_receive(int a, const EvPay& m, int klass, const abides_agent_type& ty) {
  AgentRec& A = SM->st.ag;
  A.cur_time = SM->st.hot.now;
  bool had = (A.flags & F_HAS_OPEN) && (A.flags & F_HAS_CLOSE);
  int code = pay_code(m) & ~ABIDES_MSG_REPLY;
  switch (code) {
    case ABIDES_MSG_WHEN_MKT_OPEN: A.flags |= F_HAS_OPEN; break;
    case ABIDES_MSG_WHEN_MKT_CLOSE: A.flags |= F_HAS_CLOSE; break;
    case ABIDES_MSG_ORDER_EXECUTED: {                         // orderExecuted (:390-426)
      long long q = pay_is_buy(m) ? m.w[4] : -(long long)m.w[4];
      A.holdings += q;
      A.cash -= q * (long long)m.w[5];
      int idx = ord_find(A, m.w[2]);
      if (idx >= 0) {
        OpenOrd o = SM->run.arena[A.ord_off + idx];
        int oq = o.sqty < 0 ? -o.sqty : o.sqty;
        if (m.w[4] >= oq) ord_remove(A, idx);
        else { __syncwarp(); if (LANE == 0) SM->run.arena[A.ord_off + idx].sqty = o.sqty > 0 ? oq - m.w[4] : -(oq - m.w[4]); __syncwarp(); }
      }
      break;
    }
    case ABIDES_MSG_ORDER_CANCELLED: {
      int idx = ord_find(A, m.w[2]);
      if (idx >= 0) ord_remove(A, idx);
      break;
    }
    case ABIDES_MSG_MKT_CLOSED: A.flags |= F_MKT_CLOSED; break;
    case ABIDES_MSG_QUERY_LAST_TRADE:
      if (pay_closed(m)) A.flags |= F_MKT_CLOSED;
      A.last_trade = m.w[5]; A.flags |= F_HAS_LAST_TRADE;
      if (A.flags & F_MKT_CLOSED) A.flags |= F_HAS_DAILY_CLOSE;
      break;
    case ABIDES_MSG_QUERY_SPREAD:                             // querySpread (:482-501)
      if (pay_closed(m)) A.flags |= F_MKT_CLOSED;
      A.last_trade = m.w[5]; A.flags |= F_HAS_LAST_TRADE;
      if (A.flags & F_MKT_CLOSED) A.flags |= F_HAS_DAILY_CLOSE;
      A.bid = m.w[1]; A.bid_vol = m.w[2]; A.ask = m.w[3]; A.ask_vol = m.w[4]; A.flags |= F_KNOWN_BOOK;
      break;
    case ABIDES_MSG_QUERY_TRANSACTED_VOLUME:
      if (pay_closed(m)) A.flags |= F_MKT_CLOSED;
      if (klass == ABIDES_ADAPTIVE_MM)
        A.u.mm.transacted_volume = ((long long)(unsigned)m.w[2]) | ((long long)m.w[3] << 32);
      break;
    default: break;
  }
  bool have = (A.flags & F_HAS_OPEN) && (A.flags & F_HAS_CLOSE);
  if (have && !had) {
    long long off;
    if (klass == ABIDES_ZI || klass == ABIDES_VALUE || klass == ABIDES_NOISE) off = rng_randint(agent_rng(), 0, 100);
    else off = ty.wake_up_freq;
    k_wakeup(a, SM->run.cfg.mkt_open + off);
  }
}

// ZI / Value estimator (ZeroIntelligenceAgent.py:163-247, ValueAgent.py:123-186)
__device__ __noinline__ long long estimator_update(long long obs_t, const abides_agent_type& ty, abm_dd log1mk) {
  AgentRec& A = SM->st.ag;
  if (!(A.flags & F_HAS_PREV_WAKE)) { A.u.est.prev_wake = SM->run.cfg.mkt_open; A.flags |= F_HAS_PREV_WAKE; }
  double base = 1 - ty.kappa;
  double delta = (double)(A.cur_time - A.u.est.prev_wake);
  double pw = abm_pow_with_log(log1mk, base, delta);
  double r_tprime = (1 - pw) * ty.r_bar;
  r_tprime += pw * A.u.est.r_t;
  double pw2 = abm_pow_with_log(log1mk, base, 2 * delta);
  double sigma_tprime = pw2 * A.u.est.sigma_t;
  sigma_tprime += ((1 - pw2) / (1 - abm_pow_with_log(log1mk, base, 2.0))) * ty.sigma_s;
  double r_t = (ty.sigma_n / (ty.sigma_n + sigma_tprime)) * r_tprime;
  r_t += (sigma_tprime / (ty.sigma_n + sigma_tprime)) * (double)obs_t;
  A.u.est.r_t = r_t;
  A.u.est.sigma_t = (ty.sigma_n * A.u.est.sigma_t) / (ty.sigma_n + A.u.est.sigma_t);
  double d2 = (double)(SM->run.cfg.mkt_close - A.cur_time);
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
__device__ __noinline__ void zi_place_order(int a, const abides_agent_type& ty) {
  AgentRec& A = SM->st.ag;
  RngH rs = agent_rng();
  long long obs = oracle_observe(A.cur_time, ty.sigma_n, rs);
  long long q = A.holdings / 100;
  bool buy;
  if (q >= ty.q_max) buy = false;
  else if (q <= -ty.q_max) buy = true;
  else buy = rng_randint(rs, 0, 2) != 0;
  long long r_T = estimator_update(obs, ty, SM->run.db.type_log1mk[A.type_idx]);
  q += ty.q_max - 1;
  long long row = SM->run.aoff + a;
  long long theta = SM->run.db.theta[SM->run.cfg.theta_offset + SM->run.db.agent_theta[row] + (buy ? q + 1 : q)];
  long long v = r_T + theta;
  long long R = rng_randint(rs, ty.R_min, ty.R_max + 1);
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
// ValueAgent.placeOrder (:188-220)
__device__ __noinline__ void value_place_order(int a, const abides_agent_type& ty) {
  AgentRec& A = SM->st.ag;
  long long obs = oracle_observe(A.cur_time, ty.sigma_n, agent_rng());
  long long r_T = estimator_update(obs, ty, SM->run.db.type_log1mk[A.type_idx]);
  bool buy; long long p;
  RngH g = extra_rng(ABIDES_STREAM_GLOBAL);
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
  AgentRec& A = SM->st.ag;
  bool buy = rng_randint(extra_rng(ABIDES_STREAM_GLOBAL), 0, 2) != 0;
  if (buy && A.ask > 0) ta_place_limit(a, A.size, true, A.ask);
  else if (!buy && A.bid > 0) ta_place_limit(a, A.size, false, A.bid);
}
// MomentumAgent.placeOrders (:56-66): running prefix sum + ring of the last MOM_RING prefix sums
__device__ __noinline__ void momentum_place_orders(int a) {
  AgentRec& A = SM->st.ag;
  if (!(A.bid > 0 && A.ask > 0)) return;
  double* ring = SM->run.db.mom_ring + (size_t)A.u.mom.ring * MOM_RING;
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
__device__ void mm_adaptive_update(AgentRec& A, const abides_agent_type& ty) {
  A.u.mm.window_size = A.u.mm.last_spread;
  int t = (int)(long long)rint(ty.level_spacing * A.u.mm.window_size);
  A.u.mm.tick_size = t == 0 ? 1 : t;
}
__device__ double sigmoid_stable(double x, double beta) {
  if (x >= 0) { double z = abm_exp(-beta * x); return 1 / (1 + z); }
  double z = abm_exp(beta * x);
  return z / (1 + z);
}
__device__ __noinline__ void mm_update_order_size(AgentRec& A, const abides_agent_type& ty) {
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
  AgentRec& A = SM->st.ag;
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

__device__ __noinline__ void agent_wakeup(int a) {
  AgentRec& A = SM->st.ag; Hot& h = SM->st.hot;
  const abides_agent_type& ty = SM->run.db.types[A.type_idx];
  switch (ty.klass) {
    case ABIDES_ZI: case ABIDES_VALUE: case ABIDES_NOISE: {
      ta_wakeup(a);
      ag_set_state(A, ST_INACTIVE);
      if (!(A.flags & F_HAS_OPEN) || !(A.flags & F_HAS_CLOSE)) return;
      A.flags |= F_TRADING;
      if ((A.flags & F_MKT_CLOSED) && (A.flags & F_HAS_DAILY_CLOSE)) return;
      if (ty.klass == ABIDES_NOISE) {
        if (A.u.noise.wakeup_time > h.now) k_wakeup(a, A.u.noise.wakeup_time);
      } else {
        double dt = rng_exponential(agent_rng(), 1.0 / ty.lambda_a);
        k_wakeup(a, h.now + (long long)rint(dt));
      }
      if ((A.flags & F_MKT_CLOSED) && !(A.flags & F_HAS_DAILY_CLOSE)) { ta_query_spread(a); ag_set_state(A, ST_AWAITING_SPREAD); return; }
      if (ty.klass != ABIDES_NOISE) ta_cancel_all(a);
      ta_query_spread(a);
      ag_set_state(A, ST_AWAITING_SPREAD);
      return;
    }
    case ABIDES_MOMENTUM:
      if (ta_wakeup(a)) { ta_query_spread(a); ag_set_state(A, ST_AWAITING_SPREAD); }
      return;
    case ABIDES_ADAPTIVE_MM:
      if (ta_wakeup(a)) {
        ta_cancel_all(a);
        h.additional_delay += ty.cancel_limit_delay;
        ta_query_spread(a);
        EvPay p = pay_zero(ABIDES_MSG_QUERY_TRANSACTED_VOLUME);
        p.w[2] = (int)(unsigned)(ty.wake_up_freq & 0xffffffffll); p.w[3] = (int)(ty.wake_up_freq >> 32);
        agent_send(a, p);
      }
      return;
    case ABIDES_POV_EXEC:
      ta_wakeup(a);
      k_wakeup(a, h.now + ty.wake_up_freq);
      return;
    default: return;
  }
}

__device__ __noinline__ void agent_receive(int a, const EvPay& m) {
  AgentRec& A = SM->st.ag; Hot& h = SM->st.hot;
  const abides_agent_type& ty = SM->run.db.types[A.type_idx];
  int code = pay_code(m) & ~ABIDES_MSG_REPLY;
  ta_receive(a, m, ty.klass, ty);
  switch (ty.klass) {
    case ABIDES_ZI: case ABIDES_VALUE: case ABIDES_NOISE:
      if (ag_state(A) == ST_AWAITING_SPREAD && code == ABIDES_MSG_QUERY_SPREAD) {
        if (A.flags & F_MKT_CLOSED) return;
        if (ty.klass == ABIDES_ZI) zi_place_order(a, ty);
        else if (ty.klass == ABIDES_VALUE) value_place_order(a, ty);
        else noise_place_order(a);
        ag_set_state(A, ST_AWAITING_WAKEUP);
      }
      return;
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
        mm_update_order_size(A, ty);
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
    default: return;
  }
}

// ----------------------------------------------------------------------------------------------
// ExchangeAgent (agent/ExchangeAgent.py:113-283, :398-411)
// ----------------------------------------------------------------------------------------------
struct ExchSend {
  __device__ void operator()(int rcpt, const EvPay& p) const {
    int code = pay_code(p);
    long long delay = (code == ABIDES_MSG_ORDER_ACCEPTED || code == ABIDES_MSG_ORDER_CANCELLED ||
                       code == ABIDES_MSG_ORDER_EXECUTED) ? SM->run.cfg.exch_pipeline_delay : 0;
    k_send(0, rcpt, p, delay);
  }
};

__device__ __noinline__ void exchange_receive(const EvPay& m) {
  Hot& h = SM->st.hot;
  const abides_instance_cfg* cfg = &SM->run.cfg;
  ExchSend send;
  h.exch_cur_time = h.now;
  h.exch_cd = cfg->exch_computation_delay;
  int code = pay_code(m), sender = m.w[1];
  bool is_order = code == ABIDES_MSG_LIMIT_ORDER || code == ABIDES_MSG_MARKET_ORDER || code == ABIDES_MSG_CANCEL_ORDER ||
                  code == ABIDES_MSG_MODIFY_ORDER;
  bool is_query = code == ABIDES_MSG_QUERY_LAST_TRADE || code == ABIDES_MSG_QUERY_SPREAD ||
                  code == ABIDES_MSG_QUERY_ORDER_STREAM || code == ABIDES_MSG_QUERY_TRANSACTED_VOLUME;
  bool closed = h.now > cfg->mkt_close;
  int cflag = closed ? 1 << 8 : 0;
  if (closed && !is_query) { send(sender, pay_zero(ABIDES_MSG_MKT_CLOSED)); return; }
  int oid = m.w[2];
  if (is_order) {
    if (cfg->exch_log_orders && oid == 0) h.next_order_id++;
    if (oid == 0) oid = (int)h.next_order_id++;
  }
  switch (code) {
    case ABIDES_MSG_WHEN_MKT_OPEN:
      h.exch_cd = 0;
      send(sender, pay_zero(ABIDES_MSG_WHEN_MKT_OPEN | ABIDES_MSG_REPLY));
      break;
    case ABIDES_MSG_WHEN_MKT_CLOSE:
      h.exch_cd = 0;
      send(sender, pay_zero(ABIDES_MSG_WHEN_MKT_CLOSE | ABIDES_MSG_REPLY));
      break;
    case ABIDES_MSG_QUERY_LAST_TRADE: {
      EvPay r = pay_zero((ABIDES_MSG_QUERY_LAST_TRADE | ABIDES_MSG_REPLY) | cflag);
      r.w[5] = h.last_trade;
      send(sender, r);
      break;
    }
    case ABIDES_MSG_QUERY_SPREAD: {
      EvPay r = pay_zero((ABIDES_MSG_QUERY_SPREAD | ABIDES_MSG_REPLY) | cflag);
      side_inside(book_side(true), r.w[1], r.w[2]);
      side_inside(book_side(false), r.w[3], r.w[4]);
      r.w[5] = h.last_trade;
      send(sender, r);
      break;
    }
    case ABIDES_MSG_QUERY_TRANSACTED_VOLUME: {
      long long lookback = ((long long)(unsigned)m.w[2]) | ((long long)m.w[3] << 32);
      long long vol = book_transacted_volume(h.now, lookback);
      EvPay r = pay_zero((ABIDES_MSG_QUERY_TRANSACTED_VOLUME | ABIDES_MSG_REPLY) | cflag);
      r.w[2] = (int)(unsigned)(vol & 0xffffffffll); r.w[3] = (int)(vol >> 32);
      send(sender, r);
      break;
    }
    case ABIDES_MSG_LIMIT_ORDER:
      book_handle_limit(send, sender, oid, m.w[3], m.w[4], pay_is_buy(m));
      break;
    case ABIDES_MSG_MARKET_ORDER:
      book_handle_market(send, sender, m.w[4], pay_is_buy(m));
      break;
    case ABIDES_MSG_CANCEL_ORDER:
      book_cancel(send, sender, oid, m.w[3], pay_is_buy(m));
      break;
    case ABIDES_MSG_MODIFY_ORDER:
      book_modify(send, sender, oid, m.w[3], m.w[5], pay_is_buy(m));
      break;
    default: break;
  }
}

// ----------------------------------------------------------------------------------------------
// trace of a dispatched event (layout of include/abides_b200.h: abides_trace_rec)
// ----------------------------------------------------------------------------------------------
__device__ __noinline__ void trace_event(const EvKey& k, const EvPay& m) {
  int rcpt = (int)(k.k2 >> 33);
  if ((k.k2 >> 32) & 1) { trace_write(k.t, rcpt, 0, 0, 0, 0, 0); return; }
  int full = pay_code(m), code = full & ~ABIDES_MSG_REPLY;
  long long p0 = 0, p1 = 0, p2 = 0, p3 = 0;
  if (rcpt == 0) {
    p0 = m.w[1];
    if (code == ABIDES_MSG_LIMIT_ORDER || code == ABIDES_MSG_MARKET_ORDER || code == ABIDES_MSG_CANCEL_ORDER ||
        code == ABIDES_MSG_MODIFY_ORDER) {
      p1 = m.w[2]; p2 = m.w[3]; p3 = pay_is_buy(m) ? m.w[4] : -(long long)m.w[4];
      if (code == ABIDES_MSG_CANCEL_ORDER) p3 = pay_is_buy(m) ? 1 : -1;
    } else if (code == ABIDES_MSG_QUERY_SPREAD) p1 = m.w[2];
  } else {
    switch (code) {
      case ABIDES_MSG_ORDER_ACCEPTED: case ABIDES_MSG_ORDER_EXECUTED: case ABIDES_MSG_ORDER_CANCELLED:
      case ABIDES_MSG_ORDER_MODIFIED:
        p0 = m.w[2]; p1 = m.w[3]; p2 = pay_is_buy(m) ? m.w[4] : -(long long)m.w[4]; p3 = m.w[5]; break;
      case ABIDES_MSG_QUERY_SPREAD: p0 = m.w[1]; p1 = m.w[2]; p2 = m.w[3]; p3 = m.w[4]; break;
      case ABIDES_MSG_WHEN_MKT_OPEN: p0 = SM->run.cfg.mkt_open; break;
      case ABIDES_MSG_WHEN_MKT_CLOSE: p0 = SM->run.cfg.mkt_close; break;
      case ABIDES_MSG_QUERY_TRANSACTED_VOLUME: p0 = ((long long)(unsigned)m.w[2]) | ((long long)m.w[3] << 32); break;
      case ABIDES_MSG_QUERY_LAST_TRADE: p0 = m.w[5]; break;
      default: break;
    }
  }
  trace_write(k.t, rcpt, full, p0, p1, p2, p3);
}

// ----------------------------------------------------------------------------------------------
// agent record movement: one coalesced 128-byte transaction each way
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void rec_load(int a) {
  __syncwarp();
  const unsigned* src = reinterpret_cast<const unsigned*>(&SM->run.db.agents[SM->run.aoff + a]);
  reinterpret_cast<unsigned*>(&SM->st.ag)[LANE] = src[LANE];
  SM->run.cur_agent = a;
  __syncwarp();
}
__device__ __forceinline__ void rec_store(int a) {
  __syncwarp();
  unsigned* dst = reinterpret_cast<unsigned*>(&SM->run.db.agents[SM->run.aoff + a]);
  dst[LANE] = reinterpret_cast<const unsigned*>(&SM->st.ag)[LANE];
  __syncwarp();
}

// kernelStopping in id order (Kernel.py:289-291; TradingAgent.py:122-147 and subclasses)
__device__ __noinline__ void finalize() {
  Hot& h = SM->st.hot;
  long long slowest = h.exch_busy > SM->run.cfg.start_time ? h.exch_busy : SM->run.cfg.start_time;
  for (int a = 1; a < SM->run.n_agents; a++) {
    rec_load(a);
    AgentRec& A = SM->st.ag;
    const abides_agent_type& ty = SM->run.db.types[A.type_idx];
    if (A.busy > slowest) slowest = A.busy;
    long long cash = A.cash;
    if (A.holdings != 0) cash += (long long)A.last_trade * A.holdings;
    double fv = nan("");
    if (ty.klass == ABIDES_ZI) {
      long long H = round_hundreds(A.holdings);
      long long rT = oracle_observe(A.cur_time, 0, agent_rng());
      long long surplus = 0;
      const int* th = SM->run.db.theta + SM->run.cfg.theta_offset + SM->run.db.agent_theta[SM->run.aoff + a];
      if (H > 0) for (long long x = 1; x <= H; x++) surplus += th[x + ty.q_max - 1];
      else if (H < 0) { for (long long x = H + 1; x <= 0; x++) surplus += th[x + ty.q_max - 1]; surplus = -surplus; }
      surplus += rT * H;
      surplus += A.cash - ty.starting_cash;
      fv = (double)surplus;
    } else if (ty.klass == ABIDES_VALUE) {
      long long H = round_hundreds(A.holdings);
      long long rT = oracle_observe(A.cur_time, 0, agent_rng());
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
      SM->run.db.ares[SM->run.aoff + a] = r;
    }
  }
  if (LANE == 0) {
    abides_agent_result r; r.cash = 0; r.holdings = 0; r.marked_to_market = 0; r.final_valuation = nan("");
    SM->run.db.ares[SM->run.aoff] = r;
    abides_instance_result R;
    memset(&R, 0, sizeof R);
    R.ttl_messages = h.ttl; R.delivered = h.delivered; R.final_time = h.now;
    R.slowest_agent_finish_time = slowest;
    R.last_trade = h.has_last_trade ? h.last_trade : -1;
    R.final_fundamental = (long long)h.or_pv;
    R.n_orders = h.next_order_id; R.n_fills = h.n_fills; R.traded_volume = h.volume;
    R.status = h.status; R.queue_peak = h.queue_peak; R.book_peak_orders = h.book_peak_orders;
    SM->run.db.res[SM->run.inst] = R;
  }
  SM->run.cur_agent = -1;
}

// ----------------------------------------------------------------------------------------------
// kernels
// ----------------------------------------------------------------------------------------------

// one warp per instance: build the initial device state (Kernel.runner prologue, Kernel.py:97-172)
__global__ void __launch_bounds__(32) init_kernel(DevBatch db) {
  int inst = blockIdx.x, lane = threadIdx.x;
  if (inst >= db.n_instances) return;
  const abides_instance_cfg* cfg = &db.cfg[inst];
  InstShared* G = &db.state[inst];
  int n = cfg->n_agents;
  long long aoff = cfg->agent_offset;
  // zero the persistent block
  for (size_t i = lane; i < sizeof(InstShared) / 4; i += 32) reinterpret_cast<unsigned*>(G)[i] = 0;
  __syncwarp();
  EvKey* fk = db.far_keys + (size_t)inst * db.far_cap;
  EvPay* fp = db.far_pay + (size_t)inst * db.far_cap;
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
      A.rng_pos = db.mt_pos[stream];
      A.gauss = db.mt_gauss[stream];
      if (db.mt_has_gauss[stream]) A.flags |= F_HAS_GAUSS;
    }
    switch (ty.klass) {
      case ABIDES_ZI: case ABIDES_VALUE: A.u.est.r_t = ty.r_bar; A.u.est.sigma_t = 0; break;
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
    // Agent.kernelStarting: setWakeup(startTime) (agent/Agent.py:77)
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
    h.near_limit_t = LLONG_MIN; h.near_limit_k = 0;
    h.dT = 1000000000ll;
    h.n_far_total = n; h.n_far_live = n; h.queue_peak = n;
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

__global__ void init_types_kernel(DevBatch db) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= db.n_types) return;
  double base = 1 - db.types[i].kappa;
  db.type_log1mk[i] = (base > 0 && base != 1.0) ? abm_log_dd(base) : abm_mk(0.0, 0.0);
}

// the persistent event loop: Kernel.runner (Kernel.py:190-271), one warp per instance
__global__ void __launch_bounds__(32) sim_kernel(DevBatch db, long long until_time) {
  int inst = blockIdx.x, lane = threadIdx.x;
  if (inst >= db.n_instances) return;
  InstShared* S = &SM->st;
  InstShared* G = &db.state[inst];
  {
    const int4* src = reinterpret_cast<const int4*>(G);
    int4* dst = reinterpret_cast<int4*>(S);
    for (int i = lane; i < (int)(sizeof(InstShared) / 16); i += 32) dst[i] = src[i];
  }
  if (lane == 0) {
    Run& r = SM->run;
    r.db = db; r.cfg = db.cfg[inst];
    r.far_keys = db.far_keys + (size_t)inst * db.far_cap;
    r.far_pay = db.far_pay + (size_t)inst * db.far_cap;
    r.arena = db.arena + (size_t)inst * db.arena_cap;
    r.deep = db.deep_book + (size_t)inst * 2 * db.deep_cap;
    r.deep_seq = db.deep_seq + (size_t)inst * 2 * db.deep_cap;
    r.txn_log = db.txn_log + (size_t)inst * TXN_CAP;
    r.unrolled = db.unrolled + (size_t)inst * UNR_CAP;
    r.trace = db.trace_cap > 0 ? db.trace + (size_t)inst * db.trace_cap : nullptr;
    r.aoff = db.cfg[inst].agent_offset; r.inst = inst; r.n_agents = db.cfg[inst].n_agents; r.cur_agent = -1;
  }
  __syncwarp();
  Hot& h = S->hot;
  const long long stop_time = SM->run.cfg.stop_time;
  const long long default_cd = SM->run.cfg.default_computation_delay;

  if (!h.done) {
    bool finished = false;
    for (;;) {
      if (h.status != 0) { finished = true; break; }
      if (!(h.now <= stop_time)) { finished = true; break; }
      if (h.n_near == 0) {
        if (h.n_far_live == 0) { finished = true; break; }
        q_refill();
        if (h.n_near == 0) { finished = true; break; }
      }
      int idx = near_argext(false);
      EvKey k = S->nkey[idx];
      if (k.t > until_time) break;
      EvPay m = S->npay[idx];
      near_remove(idx);
      h.now = k.t;
      h.ttl++;
      h.additional_delay = 0;
      int rcpt = (int)(k.k2 >> 33);
      bool is_wakeup = (k.k2 >> 32) & 1;
      if (rcpt == 0) {
        if (h.exch_busy > h.now) { k.t = h.exch_busy; q_push(k, m); continue; }
        h.exch_busy = h.now;
        h.delivered++;
        if (SM->run.trace) trace_event(k, m);
        if (is_wakeup) h.exch_cur_time = h.now;
        else exchange_receive(m);
        h.exch_busy += h.exch_cd + h.additional_delay;
      } else {
        rec_load(rcpt);
        AgentRec& A = S->ag;
        if (A.busy > h.now) { k.t = A.busy; q_push(k, m); continue; }
        A.busy = h.now;
        h.delivered++;
        if (SM->run.trace) trace_event(k, m);
        if (is_wakeup) agent_wakeup(rcpt);
        else agent_receive(rcpt, m);
        A.busy += default_cd + h.additional_delay;
        rec_store(rcpt);
      }
    }
    if (finished) {
      finalize();
      h.done = 1;
    }
  }
  __syncwarp();
  {
    int4* dst = reinterpret_cast<int4*>(G);
    const int4* src = reinterpret_cast<const int4*>(S);
    for (int i = lane; i < (int)(sizeof(InstShared) / 16); i += 32) dst[i] = src[i];
  }
}

// ----------------------------------------------------------------------------------------------
// K2 parity harness: replay recorded order tapes into the book, one warp per book
// ----------------------------------------------------------------------------------------------
struct ReplaySend {
  __device__ void operator()(int rcpt, const EvPay& p) const {
    Run& r = SM->run;
    int k = r.rp_counts[r.rp_op];
    __syncwarp();
    if (LANE == 0) {
      if (k < r.rp_max) {
        abides_book_event e;
        e.op_index = r.rp_op - r.rp_base; e.kind = pay_code(p); e.agent = rcpt; e.is_buy = pay_is_buy(p) ? 1 : 0;
        e.order_id = p.w[2]; e.price = p.w[3]; e.qty = p.w[4]; e.fill_price = p.w[5];
        r.rp_events[(size_t)r.rp_op * r.rp_max + k] = e;
      }
      r.rp_counts[r.rp_op] = k + 1;
    }
    __syncwarp();
  }
};

__global__ void __launch_bounds__(32) replay_kernel(int n_books, const long long* op_offset, const abides_book_op* ops,
                                                    int stream_history, int max_events, abides_book_event* events,
                                                    int* counts, abides_book_state* states, BookOrd* deep, int* deep_seq,
                                                    int deep_cap) {
  int b = blockIdx.x, lane = threadIdx.x;
  if (b >= n_books) return;
  for (int i = lane; i < (int)(sizeof(Smem) / 4); i += 32) reinterpret_cast<unsigned*>(SM)[i] = 0;
  __syncwarp();
  if (lane == 0) {
    Run& r = SM->run;
    r.db.deep_cap = deep_cap;
    r.cfg.stream_history = stream_history;
    r.deep = deep + (size_t)b * 2 * deep_cap;
    r.deep_seq = deep_seq + (size_t)b * 2 * deep_cap;
    r.txn_log = nullptr;
    r.rp_events = events; r.rp_counts = counts; r.rp_max = max_events; r.rp_base = (int)op_offset[b];
    SM->st.hot.next_order_id = 1 << 30;          // ids of market-order children: outside the tape's id space
  }
  __syncwarp();
  Hot& h = SM->st.hot;
  ReplaySend send;
  for (long long i = op_offset[b]; i < op_offset[b + 1]; i++) {
    abides_book_op op = ops[i];
    __syncwarp();
    if (lane == 0) { SM->run.rp_op = (int)i; counts[i] = 0; }
    __syncwarp();
    h.now = op.time;
    bool buy = op.is_buy != 0;
    switch (op.kind) {
      case ABIDES_MSG_LIMIT_ORDER: book_handle_limit(send, op.agent, (int)op.order_id, (int)op.price, (int)op.qty, buy); break;
      case ABIDES_MSG_MARKET_ORDER: book_handle_market(send, op.agent, (int)op.qty, buy); break;
      case ABIDES_MSG_CANCEL_ORDER: book_cancel(send, op.agent, (int)op.order_id, (int)op.price, buy); break;
      case ABIDES_MSG_MODIFY_ORDER: book_modify(send, op.agent, (int)op.order_id, (int)op.price, (int)op.new_qty, buy); break;
      default: break;
    }
    int bp, bv, ap, av;
    side_inside(book_side(true), bp, bv);
    side_inside(book_side(false), ap, av);
    int nbl = side_levels(book_side(true)), nal = side_levels(book_side(false));
    if (lane == 0) {
      abides_book_state st;
      st.bid = bp; st.bid_vol = bv; st.ask = ap; st.ask_vol = av;
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
  long long na = -1, n_theta = -1;
  bool operator==(const ShapeSig& o) const {
    return ni == o.ni && n_types == o.n_types && rng_mode == o.rng_mode && trace_cap == o.trace_cap &&
           max_agents == o.max_agents && n_mom == o.n_mom && na == o.na && n_theta == o.n_theta;
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
  init_types_kernel<<<(d.n_types + 63) / 64, 64, 0, h->stream>>>(d);
  init_kernel<<<d.n_instances, 32, 0, h->stream>>>(d);
  CK(cudaGetLastError());
  h->launches_since_load += 2;
  return 0;
}
}  // namespace

extern "C" {

const char* abides_last_error(void) { return g_err.c_str(); }
const char* abides_version(void) { return "abides_b200 0.1 (sm_100a)"; }

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
  int n_mom = 0;
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
      if (k < 0 || k > ABIDES_POV_EXEC) return fail(ABIDES_EINVAL, "abides_load: unknown agent class");
      if (k == ABIDES_MOMENTUM) mom_index[off + a] = n_mom++;
      if (k == ABIDES_ZI && (b->agent_theta[off + a] < 0 || !b->theta)) return fail(ABIDES_EINVAL, "abides_load: ZI agent without theta");
    }
    if (ic.n_agents > max_agents) max_agents = ic.n_agents;
    off += ic.n_agents;
  }
  if (off != na) return fail(ABIDES_EINVAL, "abides_load: n_agents_total mismatch");
  ShapeSig sig;
  sig.ni = ni; sig.n_types = b->n_types; sig.rng_mode = b->rng_mode; sig.trace_cap = b->trace_capacity;
  sig.max_agents = max_agents; sig.n_mom = n_mom; sig.na = na; sig.n_theta = b->n_theta_total;
  DevBatch& d = h->db;
  int rc;
  size_t ns = (size_t)na + (size_t)ABIDES_EXTRA_STREAMS * ni;
  if (!(h->loaded && sig == h->sig)) {          // (re)allocate; same-shaped batches reuse the device buffers
    free_all(h);
    memset(&d, 0, sizeof d);
    d.n_instances = ni; d.rng_mode = b->rng_mode; d.trace_cap = b->trace_capacity; d.n_types = b->n_types;
    d.far_cap = 4 * max_agents + 4096;
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
    if ((rc = dev_alloc(h, &d.agents, (size_t)na))) return rc;
    if ((rc = dev_alloc(h, &d.state, (size_t)ni))) return rc;
    if ((rc = dev_alloc(h, &d.far_keys, (size_t)ni * d.far_cap))) return rc;
    if ((rc = dev_alloc(h, &d.far_pay, (size_t)ni * d.far_cap))) return rc;
    if ((rc = dev_alloc(h, &d.arena, (size_t)ni * d.arena_cap))) return rc;
    if ((rc = dev_alloc(h, &d.deep_book, (size_t)ni * 2 * d.deep_cap))) return rc;
    if ((rc = dev_alloc(h, &d.deep_seq, (size_t)ni * 2 * d.deep_cap))) return rc;
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
  sim_kernel<<<h->n_instances, 32, sizeof(Smem), h->stream>>>(h->db, (long long)until_time);
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  h->last_launches = 1;
  h->launches_since_load += 1;
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
  abides_book_state* d_st = nullptr; BookOrd* d_deep = nullptr; int* d_seq = nullptr;
  int rc = 0;
  auto cleanup = [&]() { cudaFree(d_off); cudaFree(d_ops); cudaFree(d_ev); cudaFree(d_cnt); cudaFree(d_st); cudaFree(d_deep); cudaFree(d_seq); };
#define RCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(ABIDES_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
  RCK(cudaMalloc(&d_off, sizeof(long long) * (n_books + 1)));
  RCK(cudaMalloc(&d_ops, sizeof(abides_book_op) * n_ops));
  RCK(cudaMalloc(&d_ev, sizeof(abides_book_event) * n_ops * max_events_per_op));
  RCK(cudaMalloc(&d_cnt, sizeof(int) * n_ops));
  RCK(cudaMalloc(&d_st, sizeof(abides_book_state) * n_ops));
  RCK(cudaMalloc(&d_deep, sizeof(BookOrd) * (size_t)n_books * 2 * deep_cap));
  RCK(cudaMalloc(&d_seq, sizeof(int) * (size_t)n_books * 2 * deep_cap));
  RCK(cudaMemcpyAsync(d_off, op_offset, sizeof(long long) * (n_books + 1), cudaMemcpyHostToDevice, h->stream));
  RCK(cudaMemcpyAsync(d_ops, ops, sizeof(abides_book_op) * n_ops, cudaMemcpyHostToDevice, h->stream));
  RCK(cudaMemsetAsync(d_ev, 0, sizeof(abides_book_event) * n_ops * max_events_per_op, h->stream));
  RCK(cudaFuncSetAttribute(replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  RCK(cudaEventRecord(h->ev0, h->stream));
  replay_kernel<<<n_books, 32, sizeof(Smem), h->stream>>>(n_books, d_off, d_ops, stream_history, max_events_per_op, d_ev, d_cnt,
                                                         d_st, d_deep, d_seq, deep_cap);
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
