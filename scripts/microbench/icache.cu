// Instruction-cache capacity probe: one warp per SM walks a straight-line loop body of S KB.
// Prints cycles per instruction vs body size (L0 -> per-SM I$ -> GPC-level cache transitions).
#include <cstdio>
#include <cuda_runtime.h>
#define I1 "mad.lo.u32 %0, %0, %1, %2;\n\t"
#define I8 I1 I1 I1 I1 I1 I1 I1 I1
#define I64 I8 I8 I8 I8 I8 I8 I8 I8
#define I512 I64 I64 I64 I64 I64 I64 I64 I64      /* 512 instrs = 8 KB */
template <int NBLK>
__global__ void walk(unsigned* out, int iters, long long* cyc) {
  unsigned x = threadIdx.x, a = 3, b = blockIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < NBLK; k++) asm volatile(I512 : "+r"(x) : "r"(a), "r"(b));
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x * gridDim.y + blockIdx.y] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
template <int NBLK>
void run(int warps_per_sm, int nsm) {
  unsigned* out; long long* cyc;
  cudaMalloc(&out, sizeof(unsigned) * nsm * 32 * warps_per_sm * 4);
  cudaMalloc(&cyc, sizeof(long long) * nsm * warps_per_sm);
  int iters = 4096 / NBLK;
  dim3 grid(nsm * warps_per_sm);
  walk<NBLK><<<grid, 32>>>(out, 2, cyc);
  walk<NBLK><<<grid, 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[4096];
  cudaMemcpy(h, cyc, sizeof(long long) * nsm * warps_per_sm, cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < nsm * warps_per_sm; i++) s += h[i];
  s /= nsm * warps_per_sm;
  printf("body %4d KB  warps/SM %2d : %.2f cycles/instr\n", NBLK * 8, warps_per_sm, s / ((double)iters * NBLK * 512));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  int nsm = 148;
  for (int w : {1, 7}) {
    run<1>(w, nsm); run<2>(w, nsm); run<3>(w, nsm); run<4>(w, nsm); run<6>(w, nsm); run<8>(w, nsm);
    run<12>(w, nsm); run<16>(w, nsm); run<24>(w, nsm); run<32>(w, nsm);
  }
  return 0;
}
