// Instruction-cache probe 2: each warp jumps pseudo-randomly between 1 KB code blocks of an N KB region
// (switch dispatch), i.e. the access pattern of an event loop with many small handlers.
#include <cstdio>
#include <cuda_runtime.h>
#define I1 "mad.lo.u32 %0, %0, %1, %2;\n\t"
#define I8 I1 I1 I1 I1 I1 I1 I1 I1
#define I56 I8 I8 I8 I8 I8 I8 I8
#define BLK(k) case k: asm volatile(I56 : "+r"(x) : "r"(a), "r"(b)); break;   /* ~1 KB with the dispatch code */
#define B4(k) BLK(k) BLK(k+1) BLK(k+2) BLK(k+3)
#define B16(k) B4(k) B4(k+4) B4(k+8) B4(k+12)
#define B64(k) B16(k) B16(k+16) B16(k+32) B16(k+48)
__global__ void walk(unsigned* out, int iters, int nblk, int n2, long long* cyc) {
  unsigned x = threadIdx.x, a = 3, b = blockIdx.x;
  unsigned s = blockIdx.x * 2654435761u + (threadIdx.x >> 5) * 40503u + 12345u;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    s = s * 1664525u + 1013904223u;
    unsigned r = s >> 10; int k = (r % 100u) < 98u ? (int)((r / 100u) % (unsigned)nblk) : nblk + (int)((r / 100u) % (unsigned)(n2 > 0 ? n2 : 1)); if (n2 == 0) k = (int)((r / 100u) % (unsigned)nblk);
    switch (k) { B64(0) B64(64) B64(128) B64(192) default: break; }
  }
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
int main() {
  int nsm = 148;
  unsigned* out; long long* cyc;
  cudaMalloc(&out, sizeof(unsigned) * nsm * 32 * 16);
  cudaMalloc(&cyc, sizeof(long long) * nsm * 16);
  int cases[][2] = {{28, 0}, {28, 12}, {28, 36}, {24, 0}, {24, 16}, {24, 40}, {20, 0}, {20, 20}, {20, 44}, {16, 48}, {12, 52}};
  for (int warps : {7}) {
    for (auto& c : cases) {
      int iters = 20000;
      walk<<<nsm, 32 * warps>>>(out, 200, c[0], c[1], cyc);
      walk<<<nsm, 32 * warps>>>(out, iters, c[0], c[1], cyc);
      cudaDeviceSynchronize();
      long long h[148 * 16];
      cudaMemcpy(h, cyc, sizeof(long long) * nsm * warps, cudaMemcpyDeviceToHost);
      double sum = 0; for (int i = 0; i < nsm * warps; i++) sum += h[i];
      sum /= nsm * warps;
      printf("hot %2d KB (98%% of jumps) + tail %2d KB (2%%)  warps/SM %d : %.2f cyc/instr\n", c[0], c[1], warps, sum / iters / 60.0);
    }
  }
  return 0;
}
