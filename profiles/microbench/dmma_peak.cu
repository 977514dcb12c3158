// FP64 tensor-core (DMMA) peak of one B200: a dependency-free stream of
// mma.sync.aligned.m8n8k4.row.col.f64 -- the instruction the score kernels
// (k_s2n_mma.cu, k_lm.cu) are built on -- with ACC independent accumulators per
// warp, swept over accumulators and warps per SM.  Prints TFLOP/s
// (2 * 8 * 8 * 4 flops per instruction per warp); the maximum is the
// denominator the score stages are stated against (profiles/README.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak.bin dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ACC>
__global__ void dmma_loop(double *out, int iters, double a0, double b0) {
  double c[ACC][2];
#pragma unroll
  for (int i = 0; i < ACC; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ACC; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;  // keep the chain alive
}

template <int ACC>
static double run(int sms, int warps, double *d_out) {
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dmma_loop<ACC><<<sms, warps * 32>>>(d_out, 64, 1.0, 1.0);  // warm-up
  cudaEventRecord(e0);
  dmma_loop<ACC><<<sms, warps * 32>>>(d_out, iters, 1.0, 1.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flops = 2.0 * 8 * 8 * 4 * (double)ACC * iters * warps * sms;
  return flops / (ms * 1e-3) / 1e12;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double *d_out;
  cudaMalloc(&d_out, 8);
  double best = 0;
  printf("{\"device\": \"%s\", \"sms\": %d, \"sweep\": [", p.name, p.multiProcessorCount);
  const int warps_list[] = {4, 8, 16, 32};
  bool first = true;
  for (int w : warps_list) {
    double r[4] = {run<1>(p.multiProcessorCount, w, d_out), run<2>(p.multiProcessorCount, w, d_out),
                   run<4>(p.multiProcessorCount, w, d_out), run<8>(p.multiProcessorCount, w, d_out)};
    const int acc[4] = {1, 2, 4, 8};
    for (int i = 0; i < 4; i++) {
      printf("%s{\"warps_per_sm\": %d, \"acc\": %d, \"tflops\": %.2f}", first ? "" : ", ", w, acc[i], r[i]);
      first = false;
      if (r[i] > best) best = r[i];
    }
  }
  printf("], \"dmma_fp64_peak_tflops\": %.2f}\n", best);
  return cudaGetLastError() != cudaSuccess;
}
