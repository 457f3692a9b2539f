// Micro-benchmark (bring-up tool): issue pace of the warp-level mma.sync.m16n8k16 (bf16, fp32 accumulate) on sm_100a,
// per SM sub-core, with 1..8 independent accumulator chains per warp and 1..4 warps per sub-core.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/ubench_hmma tests/ubench_hmma.cu && tests/ubench_hmma
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int CHAINS>
__global__ void k(int iters, long long* out, float* sink) {
  float d[CHAINS][4];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) d[c][i] = 0.0f;
  uint32_t a[4] = {0x3C003C00u + threadIdx.x, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u}, b0 = 0x3C003C00u, b1 = 0x3C003C00u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                   : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.0f;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) s += d[c][i];
  if (s == 123.456f) *sink = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int CHAINS>
void run(int warps, long long* d_out, float* d_sink) {
  const int iters = 4096;
  k<CHAINS><<<148, warps * 32>>>(iters, d_out, d_sink);
  k<CHAINS><<<148, warps * 32>>>(iters, d_out, d_sink);
  long long h = 0;
  cudaDeviceSynchronize();
  cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
  const double mmas_per_subcore = (double)iters * CHAINS * warps / 4.0;
  printf("warps/SM %2d chains %d: %7.2f cycles per MMA per sub-core (%6.1f dense bf16 TFLOP/s at 148 SMs x 1.9 GHz)\n", warps, CHAINS,
         h / mmas_per_subcore, 4096.0 * 4 * 148 * 1.9e9 / (h / mmas_per_subcore) / 1e12);
}

int main() {
  long long* d_out; float* d_sink;
  cudaMalloc(&d_out, 64); cudaMalloc(&d_sink, 4);
  for (int warps : {4, 8, 16}) { run<1>(warps, d_out, d_sink); run<2>(warps, d_out, d_sink); run<6>(warps, d_out, d_sink); run<8>(warps, d_out, d_sink); }
  return 0;
}
