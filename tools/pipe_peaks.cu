// Measured arithmetic pipe peaks of the GPU this runs on: FFMA, DFMA and DMMA (mma.sync.m8n8k4.f64) in FMA/s.
// These are the denominators for the "fraction of the FP64 / FP32 pipe" numbers of the large-block kernels
// (profiles/r2_pipe_peaks.json).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipe_peaks.cu -o pipe_peaks.bin
#include <cuda_runtime.h>
#include <cstdio>

template <typename T, int CH>
__global__ void fma_kernel(T* out, int iters, T a, T b) {
  T acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = (T)(threadIdx.x + c);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = fma(acc[c], a, b);
  }
  T s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += acc[c];
  if (s == (T)12345.678) out[0] = s;
}

template <int CH>
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c0[CH], c1[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) { c0[c] = threadIdx.x + c; c1[c] = c; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0[c]), "+d"(c1[c]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += c0[c] + c1[c];
  if (s == 12345.678) out[0] = s;
}

template <typename F>
double time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, 64);
  const int blocks = sms * 8, threads = 256, iters = 20000;
  const double nthr = (double)blocks * threads;
  double t;
  t = time_ms([&] { fma_kernel<float, 8><<<blocks, threads>>>((float*)out, iters, 1.0001f, 0.5f); });
  const double ffma = nthr * iters * 8 / (t * 1e-3);
  t = time_ms([&] { fma_kernel<double, 8><<<blocks, threads>>>(out, iters, 1.0001, 0.5); });
  const double dfma = nthr * iters * 8 / (t * 1e-3);
  t = time_ms([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters, 1.0001, 0.5); });
  const double dmma = (nthr / 32) * iters * 8 * 256.0 / (t * 1e-3);
  t = time_ms([&] { dmma_kernel<2><<<blocks, threads>>>(out, iters * 4, 1.0001, 0.5); });
  const double dmma2 = (nthr / 32) * iters * 4 * 2 * 256.0 / (t * 1e-3);
  // one warp per SMSP, 4 independent accumulators: what a warp-per-node kernel can draw
  t = time_ms([&] { dmma_kernel<4><<<sms, 128>>>(out, iters * 4, 1.0001, 0.5); });
  const double dmma_1w = ((double)sms * 128 / 32) * iters * 4 * 4 * 256.0 / (t * 1e-3);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz\": %d, \"ffma_tfma_s\": %.3f, \"dfma_tfma_s\": %.3f, \"dmma_m8n8k4_tfma_s\": %.3f, "
         "\"dmma_2chains_tfma_s\": %.3f, \"dmma_1warp_per_smsp_4chains_tfma_s\": %.3f, \"note\": \"FMA/s in units of 1e12; flops = 2x\"}\n",
         p.name, sms, p.clockRate / 1000, ffma / 1e12, dfma / 1e12, dmma / 1e12, dmma2 / 1e12, dmma_1w / 1e12);
  return 0;
}
