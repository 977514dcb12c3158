// Micro-benchmark: issue pipes of packed 16-bit min/max flavours on sm_100a.
// asm volatile keeps every instruction; 8 independent chains per thread.
//   mode 0: VIMNMX.U16x2 only          mode 1: HMNMX2 (f16x2) only
//   mode 2: half VIMNMX half HMNMX2    mode 3: VIMNMX + IMAD (known different pipes)
// If HMNMX2 issues on another pipe than VIMNMX, mode 2 runs ~2x faster than 0 / 1.
// Also checks that f16x2 min/max on bit patterns in [0x0400, 0x7C00] orders like uint16.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
template <int MODE>
__global__ void k(unsigned *out, unsigned seed) {
  unsigned a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = (seed * (i + 3) + threadIdx.x * 7) & 0x3fff3fffu; b[i] = (seed + i * 977 + threadIdx.x) & 0x3fff3fffu; }
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const bool h = (MODE == 1) || (MODE == 2 && (i & 1));
      if (MODE == 3 && (i & 1)) { asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(a[i]) : "r"(b[i])); }
      else if (h) { asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); asm volatile("min.f16x2 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i])); }
      else { asm volatile("max.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); asm volatile("min.u16x2 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i])); }
    }
  }
  unsigned r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= a[i] ^ b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void check(unsigned *bad) {
  // all pairs (x, y) with x = blockIdx-based, y = thread-based in [0x0400, 0x7C00]
  unsigned x = 0x0400u + blockIdx.x * 7u; if (x > 0x7C00u) x = 0x7C00u;
  for (unsigned y = 0x0400u + threadIdx.x; y <= 0x7C00u; y += blockDim.x) {
    unsigned p = x | (y << 16), q = y | (x << 16), mn, mx;
    asm volatile("min.f16x2 %0, %1, %2;" : "=r"(mn) : "r"(p), "r"(q));
    asm volatile("max.f16x2 %0, %1, %2;" : "=r"(mx) : "r"(p), "r"(q));
    unsigned lo = x < y ? x : y, hi = x < y ? y : x;
    if (mn != (lo | (lo << 16)) || mx != (hi | (hi << 16))) atomicAdd(bad, 1u);
  }
}
template <int MODE>
void run(const char *name, unsigned *d) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  k<MODE><<<sms * 2, 1024>>>(d, 12345); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<sms * 2, 1024>>>(d, 12345); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double winst = (double)sms * 2 * 32 * ITER * 8 * ((MODE == 3) ? 1.5 : 2.0);
  printf("%-28s %.3f ms  %.2f warp-inst/ns (whole chip)\n", name, ms, winst / (ms * 1e6));
}
int main() {
  unsigned *d; cudaMalloc(&d, 148 * 4 * 1024 * 4);
  run<0>("VIMNMX.U16x2", d); run<1>("HMNMX2 (f16x2)", d); run<2>("VIMNMX + HMNMX2 mix", d); run<3>("VIMNMX + IMAD mix", d);
  unsigned *bad; cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
  check<<<4352, 256>>>(bad); unsigned h; cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost);
  printf("f16x2 min/max vs uint16 order on [0x0400,0x7C00]: %u mismatches\n", h);
  return 0;
}
