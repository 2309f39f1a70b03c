// lane_kernels.cu -- the lane engine's kernels (lane_device.inc, one instance per thread), compiled as their own
// translation unit: here a fetched instruction serves a whole warp of instances, so -- unlike in engine.cu, where the
// warp-per-instance kernel is bound by instruction fetch and every big body is kept out of line -- the hot helpers
// (RNG words, the fp64 exp / log / pow, the estimator) are INLINED into their handlers: no call frames spilled to
// local memory, and independent chains (the estimator's three pow, the oracle's two exp) overlap.
#define ABM_INLINE_BIG 1
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "lane_kernels.h"

#define UNLIKELY(x) __builtin_expect(!!(x), 0)
#define LIKELY(x) __builtin_expect(!!(x), 1)
#define LN_INL __device__ __forceinline__
#define LN_BIG __device__ __noinline__
#ifdef LANE_NO_INLINE_HOT
#define LN_HOT __device__ __noinline__
#else
#define LN_HOT __device__ __forceinline__
#endif
#define LN_PREFETCH(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#define LN_KERNEL __global__
#define LN_GRID_CONSTANT __grid_constant__
#define LN_THREAD_INDEX() ((int)(blockIdx.x * blockDim.x + threadIdx.x))

namespace lane {
#include "lane_variants.inc"

constexpr int INIT_BLOCK = 32;

#define LANE_DISPATCH(variant, CALL)                       \
  switch (variant) {                                       \
    case LV_ZI_LEGACY: { using namespace lv_zi_legacy; CALL; } break; \
    case LV_ZI_CUBIC: { using namespace lv_zi_cubic; CALL; } break;   \
    case LV_RMSC: { using namespace lv_rmsc; CALL; } break;           \
    default: { using namespace lv_full; CALL; } break;                \
  }

cudaError_t launch_init(int variant, const LaneBatch& lb, cudaStream_t stream) {
  const int tg = (lb.n_types + 63) / 64, ig = (lb.n_instances + INIT_BLOCK - 1) / INIT_BLOCK;
  LANE_DISPATCH(variant, (lane_init_types_kernel<<<tg, 64, 0, stream>>>(lb), lane_init_kernel<<<ig, INIT_BLOCK, 0, stream>>>(lb)));
  return cudaGetLastError();
}

cudaError_t launch_sim(int variant, const LaneBatch& lb, long long until_time, cudaStream_t stream) {
  const int grid = (lb.n_instances + LANE_CTA - 1) / LANE_CTA;
  // shared memory / L1 split of the SM: the kernel needs 6.5 KB of shared memory per CTA and lives on its L1 hit rate
  // (thread-local memory); ABIDES_LANE_SMEM_CARVEOUT = preferred shared-memory share in percent (0 = as much L1 as
  // possible), unset = the driver's default choice
  static const int carve = [] { const char* e = getenv("ABIDES_LANE_SMEM_CARVEOUT"); return e && *e ? atoi(e) : -1; }();
  if (carve >= 0 && carve <= 100)
    LANE_DISPATCH(variant, cudaFuncSetAttribute(lane_sim_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
  LANE_DISPATCH(variant, (lane_sim_kernel<<<grid, LANE_CTA + LANE_HEAVY, 0, stream>>>(lb, until_time)));
  return cudaGetLastError();
}

cudaError_t launch_replay(const LaneBatch& lb, int n_books, const long long* op_offset, const abides_book_op* ops,
                          int stream_history, int max_events, abides_book_event* events, int* counts,
                          abides_book_state* states, cudaStream_t stream) {
  lv_full::lane_replay_kernel<<<(n_books + INIT_BLOCK - 1) / INIT_BLOCK, INIT_BLOCK, 0, stream>>>(
      lb, n_books, op_offset, ops, stream_history, max_events, events, counts, states);
  return cudaGetLastError();
}
}  // namespace lane
