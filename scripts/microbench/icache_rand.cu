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
__global__ void walk(unsigned* out, int iters, int nblk, long long* cyc) {
  unsigned x = threadIdx.x, a = 3, b = blockIdx.x;
  unsigned s = blockIdx.x * 2654435761u + (threadIdx.x >> 5) * 40503u + 12345u;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    s = s * 1664525u + 1013904223u;
    int k = (int)((s >> 10) % (unsigned)nblk);
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
  for (int warps : {1, 7, 14}) {
    for (int nblk : {4, 8, 16, 24, 28, 32, 40, 48, 64, 96, 128, 192, 256}) {
      int iters = 20000;
      walk<<<nsm, 32 * warps>>>(out, 200, nblk, cyc);
      walk<<<nsm, 32 * warps>>>(out, iters, nblk, cyc);
      cudaDeviceSynchronize();
      long long h[148 * 16];
      cudaMemcpy(h, cyc, sizeof(long long) * nsm * warps, cudaMemcpyDeviceToHost);
      double sum = 0; for (int i = 0; i < nsm * warps; i++) sum += h[i];
      sum /= nsm * warps;
      printf("region %3d KB  warps/SM %2d : %.1f cycles per 1 KB block (%.2f cyc/instr)\n", nblk, warps, sum / iters, sum / iters / 60.0);
    }
  }
  return 0;
}
