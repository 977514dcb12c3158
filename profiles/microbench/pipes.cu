// Micro-benchmark: which issue pipe do packed 16-bit min/max flavours use on sm_100a?
// Measures warp-instructions per cycle per SM for independent chains of
//   VIMNMX.U16x2 (__vminu2/__vmaxu2), HMNMX2 (__hmin2/__hmax2), LOP3, IMAD, PRMT
// and for mixes (if two ops are on different pipes the mix runs faster than either).
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define ITER 4096
template <int MODE>
__global__ void k(unsigned *out, unsigned seed) {
  unsigned a[8], b = seed + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = seed * (i + 3) + threadIdx.x * 7;
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) a[i] = __vminu2(a[i], b) + 0;                       // VIMNMX.U16x2
      if (MODE == 1) { __half2 h = __hmax2(*(__half2 *)&a[i], *(__half2 *)&b); a[i] = *(unsigned *)&h; }
      if (MODE == 2) { if (i & 1) a[i] = __vmaxu2(a[i], b); else { __half2 h = __hmax2(*(__half2 *)&a[i], *(__half2 *)&b); a[i] = *(unsigned *)&h; } }
      if (MODE == 3) a[i] = (a[i] ^ b) & 0x7fff7fffu;                    // LOP3
      if (MODE == 4) a[i] = a[i] * 3u + b;                               // IMAD
      if (MODE == 5) { if (i & 1) a[i] = __vmaxu2(a[i], b); else a[i] = a[i] * 3u + b; }  // VIMNMX + IMAD
      if (MODE == 6) a[i] = __byte_perm(a[i], b, 0x1032);                // PRMT
      if (MODE == 7) { if (i & 1) a[i] = (a[i] ^ b) & 0x7fff7fffu; else { __half2 h = __hmax2(*(__half2 *)&a[i], *(__half2 *)&b); a[i] = *(unsigned *)&h; } }
    }
    b += 0x00010001u * (it & 1);
  }
  unsigned r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char *name, unsigned *d) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  k<MODE><<<sms * 2, 1024>>>(d, 12345); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<sms * 2, 1024>>>(d, 12345); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double winst = (double)sms * 2 * 32 * ITER * 8;  // warp-instructions of the measured op
  printf("%-22s %.3f ms  %.2f warp-inst/cycle/SM (at %d MHz nominal)\n", name, ms, winst / (ms * 1e-3 * clk * 1e3) / sms, clk / 1000);
}
int main() {
  unsigned *d; cudaMalloc(&d, 148 * 4 * 1024 * 4);
  run<0>("VIMNMX.U16x2", d); run<1>("HMNMX2", d); run<2>("VIMNMX+HMNMX2 mix", d); run<3>("LOP3", d);
  run<4>("IMAD", d); run<5>("VIMNMX+IMAD mix", d); run<6>("PRMT", d); run<7>("LOP3+HMNMX2 mix", d);
  return 0;
}
