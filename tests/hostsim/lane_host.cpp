// lane_host.cpp -- TEST INFRASTRUCTURE ONLY: the lane engine's device code (abides_b200/csrc/lane_device.inc) compiled
// for the host, one instance after another, so that the -m "not gpu" suite can check the engine's LOGIC (event order,
// book, agents, RNG regeneration) against the reference goldens without a GPU.  The product package never loads this:
// abides_b200/_abi.py only opens the CUDA library and there is no CPU fallback.  Built by tests/hostsim/__init__.py
// into tests/hostsim/_build/.
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define LN_INL static inline
#define LN_BIG static
#define LN_HOT static inline
#define LN_KERNEL static
#define LN_PREFETCH(p) ((void)(p))
#define LN_GRID_CONSTANT
#define LN_THREAD_INDEX() g_thread_index
#define UNLIKELY(x) __builtin_expect(!!(x), 0)
#define LIKELY(x) __builtin_expect(!!(x), 1)

#include "lane_host_common.h"

namespace lane {
static thread_local int g_thread_index = 0;
#include "lane_variants.inc"
}  // namespace lane
namespace rngsetup {
#include "rng_setup_device.inc"
}

using namespace lane;

namespace {
template <typename T> T* zalloc(size_t n) { return (T*)calloc(n ? n : 1, sizeof(T)); }
template <typename T> T* dup(const T* src, size_t n) {
  T* p = zalloc<T>(n);
  if (n && src) memcpy(p, src, n * sizeof(T));
  return p;
}
struct Owned {
  std::vector<void*> ptrs;
  template <typename T> T* keep(T* p) { ptrs.push_back((void*)p); return p; }
  ~Owned() { for (void* p : ptrs) free(p); }
};
}  // namespace

extern "C" {

// Runs a whole batch to completion on the host.  force_variant: -1 = the variant the loader would pick, else LV_*.
// Returns 0, or a negative abides_error; *variant_out receives the variant that ran.
int hostsim_run(const abides_batch* b, int force_variant, abides_instance_result* res_out, abides_agent_result* ares_out,
                int32_t trace_instance, abides_trace_rec* trace_out, int64_t trace_capacity, int64_t* n_trace_out,
                int32_t* variant_out) {
  const int ni = b->n_instances;
  const long long na = b->n_agents_total;
  unsigned class_mask = 0, lat_mask = 0; long long n_sub = 0; int max_agents = 0, n_mom = 0;
  std::vector<int> mom_index((size_t)na, -1);
  for (int i = 0; i < ni; i++) {
    const abides_instance_cfg& ic = b->instances[i];
    lat_mask |= 1u << ic.latency_kind;
    for (int a = 0; a < ic.n_agents; a++) {
      const abides_agent_type& ty = b->types[ic.type_base + b->agent_type[ic.agent_offset + a]];
      if (a > 0) class_mask |= 1u << ty.klass;
      if (ty.subscribe) n_sub++;
      if (ty.klass == ABIDES_MOMENTUM) mom_index[ic.agent_offset + a] = n_mom++;
    }
    if (ic.n_agents > max_agents) max_agents = ic.n_agents;
  }
  int variant = lane_pick_variant(class_mask, lat_mask, b->rng_mode, n_sub);
  if (variant == LV_NONE) return ABIDES_EINVAL;
  if (force_variant >= 0) {
    if (force_variant != LV_FULL && force_variant != variant) return ABIDES_EINVAL;
    variant = force_variant;
  }
  if (variant_out) *variant_out = variant;
  const bool needs_tv = (class_mask & ((1u << ABIDES_ADAPTIVE_MM) | (1u << ABIDES_POV_EXEC))) != 0;
  LaneCaps cap = lane_caps(max_agents, variant, needs_tv);
  if (const char* ov = getenv("HOSTSIM_CAPS")) {       // tests of the overflow paths: "hcap,scap,arena,deep,txn,unr" (0 = keep)
    int v[6] = {0, 0, 0, 0, 0, 0};
    sscanf(ov, "%d,%d,%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]);
    if (v[0]) cap.hcap = v[0];
    if (v[1]) cap.scap = v[1];
    if (v[2]) cap.arena_cap = v[2];
    if (v[3]) cap.deep_cap = v[3];
    if (v[4] && needs_tv) cap.txn_cap = v[4];
    if (v[5] && needs_tv) cap.unr_cap = v[5];
  }
  const size_t ns = (size_t)na + (size_t)ABIDES_EXTRA_STREAMS * ni;
  Owned own;
  LaneBatch lb;
  memset(&lb, 0, sizeof lb);
  lb.n_instances = ni; lb.trace_cap = (int)trace_capacity; lb.n_types = b->n_types; lb.top_k = cap.top_k;
  lb.rng_mode = b->rng_mode;
  lb.run_ahead = getenv("ABIDES_LANE_RUN_AHEAD") ? atoi(getenv("ABIDES_LANE_RUN_AHEAD")) : 0;
  lb.trace_flags = b->trace_flags;
  lb.hcap = cap.hcap; lb.scap = cap.scap; lb.arena_cap = cap.arena_cap; lb.deep_cap = cap.deep_cap;
  lb.txn_cap = cap.txn_cap; lb.unr_cap = cap.unr_cap;
  lb.cfg = b->instances; lb.types = b->types;
  lb.type_log1mk = own.keep(zalloc<abm_dd>(b->n_types));
  lb.type_omp2 = own.keep(zalloc<double>(b->n_types));
  lb.agent_type = b->agent_type; lb.lat_to = b->lat_to_exch; lb.lat_from = b->lat_from_exch;
  lb.wakeup_time = (const long long*)b->agent_wakeup_time;
  lb.agent_theta = b->agent_theta;
  lb.mom_index = mom_index.data();
  // the loader's RNG work (abides_load): rows, seeding, constructor draws -- on writable copies of the tables
  const bool philox = b->rng_mode == ABIDES_RNG_PHILOX;
  const bool seeded = b->mt_key_row != nullptr && !philox;
  std::vector<int> dev_row;
  size_t n_rows = philox ? 0 : ns;
  if (seeded) n_rows = lane_assign_rows(b, dev_row);
  unsigned* mt = philox ? nullptr : own.keep(zalloc<unsigned>(n_rows * ABIDES_MT_N));
  if (b->mt_key && !philox) memcpy(mt, b->mt_key, (seeded ? (size_t)b->n_mt_rows : ns) * ABIDES_MT_N * sizeof(unsigned));
  int* pos = own.keep(b->mt_pos ? dup<int>(b->mt_pos, ns) : zalloc<int>(ns));
  int* hg = own.keep(b->mt_has_gauss ? dup<int>(b->mt_has_gauss, ns) : zalloc<int>(ns));
  double* gs = own.keep(b->mt_gauss ? dup<double>(b->mt_gauss, ns) : zalloc<double>(ns));
  int* theta = own.keep(zalloc<int>((size_t)b->n_theta_total + 1));
  if (!(b->draw_flags & ABIDES_DRAW_ZI_THETA) && b->theta) memcpy(theta, b->theta, (size_t)b->n_theta_total * sizeof(int));
  long long* size = own.keep(dup<long long>((const long long*)b->agent_size, (size_t)na));
  if (seeded || b->draw_flags) {
    rngsetup::RngSetup p;
    memset(&p, 0, sizeof p);
    p.n_instances = ni; p.draw_flags = b->draw_flags; p.cfg = b->instances; p.types = b->types;
    p.agent_type = b->agent_type; p.agent_theta = b->agent_theta; p.theta = theta; p.size = size;
    p.mt = mt; p.staging = nullptr; p.key_row = seeded ? b->mt_key_row : nullptr; p.dev_row = seeded ? dev_row.data() : nullptr;
    p.seed = b->mt_seed; p.pos = pos; p.has_gauss = hg; p.gauss = gs;
    for (int i = 0; i < ni; i++) {
      const int n = b->instances[i].n_agents + ABIDES_EXTRA_STREAMS;
      if (seeded) for (int a = 0; a < n; a++) rngsetup::rng_materialize_one(p, i, a);
      if (b->draw_flags) for (int a = 0; a < n; a++) rngsetup::rng_ctor_draws_one(p, i, a);
    }
  }
  lb.theta = theta; lb.size = size;
  lb.mt = mt;
  lb.mt_row = seeded ? dev_row.data() : nullptr; lb.mt_seed = b->mt_seed;
  lb.mt_pos = pos; lb.mt_has_gauss = hg; lb.mt_gauss = gs;
  lb.agents = own.keep((LAgent*)aligned_alloc(128, sizeof(LAgent) * (size_t)(na ? na : 1)));
  lb.state = own.keep(zalloc<LHot>(ni));
  lb.hkeys = own.keep(zalloc<LKey>((size_t)ni * heap_stride(cap.hcap)));
  lb.hpay = own.keep(zalloc<LPay>((size_t)ni * cap.scap));
  lb.hfree = own.keep(zalloc<int>((size_t)ni * cap.scap));
  lb.top = own.keep(zalloc<LOrd>((size_t)ni * 2 * cap.top_k));
  lb.pl_price = own.keep(zalloc<int>((size_t)ni * 2 * cap.deep_cap));
  lb.pl_ord = own.keep(zalloc<LOrd>((size_t)ni * 2 * cap.deep_cap));
  lb.pl_free = own.keep(zalloc<int>((size_t)ni * 2 * cap.deep_cap));
  lb.arena = own.keep(zalloc<OpenOrd>((size_t)ni * cap.arena_cap));
  if (needs_tv) {
    lb.txn_log = own.keep(zalloc<TxnRow>((size_t)ni * cap.txn_cap));
    lb.unrolled = own.keep(zalloc<TxnRow>((size_t)ni * cap.unr_cap));
  }
  lb.mom_ring = own.keep(zalloc<double>((size_t)(n_mom > 0 ? n_mom : 1) * MOM_RING));
  // only the traced instance gets a trace buffer on the host: give every instance the same capacity slot-wise
  lb.trace = trace_capacity > 0 ? own.keep(zalloc<abides_trace_rec>((size_t)ni * trace_capacity)) : nullptr;
  lb.res = res_out; lb.ares = ares_out;

#define RUN_VARIANT(ns_)                                                                              \
  do {                                                                                                \
    for (int i = 0; i < b->n_types; i++) { g_thread_index = i; ns_::lane_init_types_kernel(lb); }     \
    for (int i = 0; i < ni; i++) { g_thread_index = i; ns_::lane_init_kernel(lb); }                   \
    for (int i = 0; i < ni; i++) { g_thread_index = i; ns_::lane_sim_kernel(lb, LLONG_MAX); }         \
  } while (0)
  switch (variant) {
    case LV_ZI_LEGACY: RUN_VARIANT(lv_zi_legacy); break;
    case LV_ZI_CUBIC: RUN_VARIANT(lv_zi_cubic); break;
    case LV_RMSC: RUN_VARIANT(lv_rmsc); break;
    default: RUN_VARIANT(lv_full); break;
  }
#undef RUN_VARIANT
  if (trace_capacity > 0 && trace_instance >= 0 && trace_instance < ni) {
    const int n = lb.state[trace_instance].n_trace;
    if (n_trace_out) *n_trace_out = n;
    const int64_t k = n < trace_capacity ? n : trace_capacity;
    if (trace_out) memcpy(trace_out, lb.trace + (size_t)trace_instance * trace_capacity, (size_t)k * sizeof(abides_trace_rec));
  }
  return 0;
}

// util/OrderBook.py replay through the lane engine's book (same contract as abides_replay_book)
int hostsim_replay_book(int32_t n_books, const int64_t* op_offset, const abides_book_op* ops, int32_t stream_history,
                        int32_t max_events_per_op, abides_book_event* events_out, int32_t* event_count_out,
                        abides_book_state* states_out) {
  long long n_ops = op_offset[n_books], longest = 0;
  for (int b = 0; b < n_books; b++) { long long len = op_offset[b + 1] - op_offset[b]; if (len > longest) longest = len; }
  Owned own;
  LaneBatch lb;
  memset(&lb, 0, sizeof lb);
  lb.top_k = lane_top_k(LV_FULL);
  lb.deep_cap = (int)(longest + 64);
  lb.top = own.keep(zalloc<LOrd>((size_t)n_books * 2 * lb.top_k));
  lb.pl_price = own.keep(zalloc<int>((size_t)n_books * 2 * lb.deep_cap));
  lb.pl_ord = own.keep(zalloc<LOrd>((size_t)n_books * 2 * lb.deep_cap));
  lb.pl_free = own.keep(zalloc<int>((size_t)n_books * 2 * lb.deep_cap));
  memset(events_out, 0, sizeof(abides_book_event) * (size_t)n_ops * max_events_per_op);
  for (int b = 0; b < n_books; b++) {
    g_thread_index = b;
    lv_full::lane_replay_kernel(lb, n_books, (const long long*)op_offset, ops, stream_history, max_events_per_op, events_out,
                                event_count_out, states_out);
  }
  return 0;
}

}  // extern "C"
