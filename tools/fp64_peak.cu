// FP64 peak microbenchmark for the roofline denominators: DFMA issue rate with
// independent accumulators, and mma.sync.m8n8k4.f64 (DMMA).  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
// -o tools/fp64_peak tools/fp64_peak.cu ; run on the GPU box, prints one JSON line.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double *out, double a, double b, int iters) {
  double acc[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) acc[k] = threadIdx.x * 1e-9 + k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = fma(acc[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double *out, int iters) {
  double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[2]), "+d"(c0[3]) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[2]), "+d"(c1[3]) : "d"(a), "d"(b));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c0[2] + c0[3] + c1[0] + c1[1] + c1[2] + c1[3];
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int threads = 1024, blocks = sms * 2, iters = 20000;
  double *out;
  cudaMalloc(&out, sizeof(double) * threads * blocks);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms;
  double best_dfma = 0, best_dmma = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<8><<<blocks, threads>>>(out, 0.999999, 1e-9, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double tf = 2.0 * 8 * iters * (double)threads * blocks / (ms * 1e-3) / 1e12;
    if (tf > best_dfma) best_dfma = tf;
    cudaEventRecord(e0);
    dmma_kernel<<<blocks, threads>>>(out, iters / 4);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    // one m8n8k4 mma = 8*8*4*2 = 512 flop per warp
    tf = 512.0 * 16 * (iters / 4) * (double)(threads / 32) * blocks / (ms * 1e-3) / 1e12;
    if (tf > best_dmma) best_dmma = tf;
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("{\"fp64_dfma_tflops\": %.3f, \"fp64_dmma_tflops\": %.3f, \"sms\": %d, \"cuda_status\": \"%s\"}\n", best_dfma, best_dmma, sms,
         cudaGetErrorString(e));
  return e != cudaSuccess;
}
