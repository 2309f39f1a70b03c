// lane_kernels.h -- launch entry points of the lane engine's kernels (lane_kernels.cu), called by engine.cu.
#ifndef ABIDES_LANE_KERNELS_H
#define ABIDES_LANE_KERNELS_H
#include <cuda_runtime.h>
#include "lane_host_common.h"

namespace lane {
// every function enqueues on `stream` and returns cudaGetLastError()
cudaError_t launch_init(int variant, const LaneBatch& lb, cudaStream_t stream);        // type constants + initial state
cudaError_t launch_sim(int variant, const LaneBatch& lb, long long until_time, cudaStream_t stream);
cudaError_t launch_replay(const LaneBatch& lb, int n_books, const long long* op_offset, const abides_book_op* ops,
                          int stream_history, int max_events, abides_book_event* events, int* counts,
                          abides_book_state* states, cudaStream_t stream);
}  // namespace lane
#endif
