// Dependent-issue latencies (cycles) of the instructions the fused PoseUKF kernel chains: DFMA, DMUL, DADD, DMMA m8n8k4,
// SHFL (64-bit = 2), LDS.64, MUFU.RSQ64H path.  One warp, clock64 around an unrolled dependent chain.
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
__global__ void lat(double *out, long long *cyc, double a, double b) {
  __shared__ double sm[64];
  sm[threadIdx.x] = threadIdx.x * 0.0;
  sm[threadIdx.x + 32] = 0.0;
  __syncwarp();
  double x = a + threadIdx.x * 1e-9;
  long long t0, t1;
  // DFMA
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) x = fma(x, b, a);
  t1 = clock64(); cyc[0] = t1 - t0;
  // DADD
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) x = x + a;
  t1 = clock64(); cyc[1] = t1 - t0;
  // SHFL double (2 x SHFL.IDX)
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
  t1 = clock64(); cyc[2] = t1 - t0;
  // DMMA chain through the accumulator
  double c0 = x, c1 = x;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  t1 = clock64(); cyc[3] = t1 - t0;
  // DMMA chain through the B operand
  double d0 = x, d1 = 0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(d0), "=d"(d1) : "d"(a), "d"(d0), "d"(0.0), "d"(0.0));
  }
  t1 = clock64(); cyc[4] = t1 - t0;
  // LDS pointer chase (index in a double)
  int idx = threadIdx.x;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) idx = (int)sm[idx] + (threadIdx.x & 31);
  t1 = clock64(); cyc[5] = t1 - t0;
  // independent DFMA throughput inside one warp (8 chains)
  double y[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) y[k] = x + k;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N / 8; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) y[k] = fma(y[k], b, a);
  t1 = clock64(); cyc[6] = t1 - t0;
  double s = x + c0 + c1 + d0 + d1 + idx;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += y[k];
  out[threadIdx.x] = s;
}
int main() {
  double *out; long long *cyc, h[8];
  cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8 * 8);
  lat<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999);
  lat<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999);
  cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
  const char *nm[] = {"dfma_dep", "dadd_dep", "shfl64_dep", "dmma_acc_dep", "dmma_b_dep", "lds_i2f_dep", "dfma_8chains_per_instr"};
  printf("{");
  for (int k = 0; k < 7; ++k) printf("\"%s\": %.2f%s", nm[k], (double)h[k] / N, k < 6 ? ", " : "");
  printf("}\n");
  return cudaDeviceSynchronize() != cudaSuccess;
}
