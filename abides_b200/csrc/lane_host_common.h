// lane_host_common.h -- host-side decisions of the lane engine shared by engine.cu and the CPU-side test harness:
// which kernel variant covers a batch, and the per-instance capacities of its device structures.
#ifndef ABIDES_LANE_HOST_COMMON_H
#define ABIDES_LANE_HOST_COMMON_H
#include <vector>
#include "lane_types.h"

namespace lane {

enum { LV_NONE = -1, LV_ZI_LEGACY = 0, LV_ZI_CUBIC = 1, LV_RMSC = 2, LV_FULL = 3, LV_COUNT = 4 };
// the variant ids reported by abides_engine_info(): the warp engine's variants are 0..6, the lane engine's follow
constexpr int LANE_VARIANT_BASE = 16;

inline int lane_top_k(int variant) { return variant == LV_RMSC ? 128 : variant == LV_FULL ? 64 : 32; }

// the leanest lane variant whose feature set covers the batch, or LV_NONE (HBL agents and market-data subscriptions
// run on the warp-per-instance engine)
inline int lane_pick_variant(unsigned class_mask, unsigned lat_mask, int rng_mode, long long n_subscribers) {
  (void)rng_mode;                            // MT19937 and Philox streams both run on every variant
  if (n_subscribers > 0) return LV_NONE;
  const unsigned zi = 1u << ABIDES_ZI;
  const unsigned rmsc = (1u << ABIDES_NOISE) | (1u << ABIDES_VALUE) | (1u << ABIDES_MOMENTUM) | (1u << ABIDES_ADAPTIVE_MM) |
                        (1u << ABIDES_POV_EXEC);
  const unsigned full = zi | rmsc | (1u << ABIDES_MARKET_MAKER) | (1u << ABIDES_SCHEDULE_EXEC) | (1u << ABIDES_MARKET_REPLAY);
  if ((class_mask & ~zi) == 0 && lat_mask == (1u << ABIDES_LAT_LEGACY_NOISE)) return LV_ZI_LEGACY;
  if ((class_mask & ~zi) == 0 && lat_mask == (1u << ABIDES_LAT_CUBIC)) return LV_ZI_CUBIC;
  if ((class_mask & ~rmsc) == 0 && lat_mask == (1u << ABIDES_LAT_DETERMINISTIC)) return LV_RMSC;
  if ((class_mask & ~full) == 0) return LV_FULL;
  return LV_NONE;
}

// Device rows of the MT19937 states of a batch whose streams are described by their seeds (abides_batch.mt_key_row):
// uploaded rows keep their index, seeded streams follow, except that NOISE agents' seeded streams get no row at all
// (-1): they draw one or two words a day (NoiseAgent.py:162), which the lane engine recomputes from the seed.
inline size_t lane_assign_rows(const abides_batch* b, std::vector<int>& dev_row) {
  const size_t ns = (size_t)b->n_agents_total + (size_t)ABIDES_EXTRA_STREAMS * b->n_instances;
  dev_row.assign(ns, -1);
  size_t next = (size_t)b->n_mt_rows;
  for (int i = 0; i < b->n_instances; i++) {
    const abides_instance_cfg& ic = b->instances[i];
    const size_t s0 = (size_t)ic.agent_offset + (size_t)ABIDES_EXTRA_STREAMS * i;
    for (int a = 0; a < ic.n_agents + ABIDES_EXTRA_STREAMS; a++) {
      const size_t s = s0 + a;
      if (b->mt_key_row[s] >= 0) { dev_row[s] = b->mt_key_row[s]; continue; }
      const bool lazy = a > 0 && a < ic.n_agents &&
                        b->types[ic.type_base + b->agent_type[ic.agent_offset + a]].klass == ABIDES_NOISE &&
                        b->mt_pos[s] < LAZY_MAX_WORDS / 2;
      dev_row[s] = lazy ? -1 : (int)next++;
    }
  }
  return next;
}

struct LaneCaps { int hcap, scap, arena_cap, deep_cap, txn_cap, unr_cap, top_k; };
inline LaneCaps lane_caps(int max_agents, int variant, bool needs_tv) {
  // Capacities are per instance and decide how many instances fit in HBM, i.e. the engine's throughput; every one of
  // them is checked on the device (ABIDES_EOVERFLOW stops the instance, never silently).
  LaneCaps c;
  const int n = max_agents;
  c.hcap = 2 * n + 4096;                     // events in flight: peak 2 n - 1 at the opening handshake
  c.scap = c.hcap;                           // message payload slots
  c.arena_cap = 4 * n + 4096;                // open-order entries of all agents (lists grow by doubling)
  // pool slots per book side: the rmsc populations rest at most a few hundred orders (the deque of 128 holds nearly
  // all of them); a ZI population rests up to one order per agent, spread over thousands of ticks
  c.deep_cap = variant == LV_RMSC ? n / 4 + 1024 : n + 1024;
  c.txn_cap = needs_tv ? 1 << 14 : 0;        // transaction rows between two volume queries
  c.unr_cap = needs_tv ? 1 << 14 : 0;        // de-duplicated rows inside the longest look-back window
  c.top_k = lane_top_k(variant);
  return c;
}

}  // namespace lane
#endif
