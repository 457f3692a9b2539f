// Micro-benchmark (bring-up tool, not part of the product): how many cycles does one tcgen05.mma
// (cta_group::1, kind::f16, M = 128, K = 16, both operands in shared memory) take on this GPU as a function of
// N, of the swizzle mode and of the A tile's start row?  k_conv_tc's issue loop is reproduced without TMA
// traffic and without an epilogue, so the figure is the tensor pipe's own pace for the convolution's shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/ubench_umma tests/ubench_umma.cu && tests/ubench_umma
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void umma(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

template <int KC>
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  constexpr uint64_t layout = KC == 64 ? 2 : (KC == 32 ? 4 : 6);
  constexpr uint64_t sbo = (8 * KC * 2) >> 4;
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

// mode 0: tap shifts of a W = 7 board ((tap/3-1)*8 + tap%3-1 rows); mode 1: no shift (8-row aligned starts);
// mode 2: like 0 but every MMA of a k-block goes to ONE accumulator tile (MT = 1 behaviour, dependent chain)
template <int KC>
__global__ void __launch_bounds__(64, 1) k_ubench(int N, int MT, int supers, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int ROWB = KC * 2;
  const int a_rows = MT * 128 + 32;
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)a_rows * ROWB;                    // 1024-aligned: a_rows * ROWB is a multiple of 1024
  const int nb = 4;                                            // weight blocks cycled through
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + (size_t)nb * N * ROWB);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (a_rows + nb * N) * ROWB / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 32) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a_desc = make_desc<KC>(smem_u32(sA)), b_desc = make_desc<KC>(smem_u32(sB));
    const int n_kb = 9 * (64 / KC);                            // 64 input channels
    long long t0 = clock64();
    for (int it = 0; it < supers; ++it) {
      for (int kb = 0; kb < n_kb; ++kb) {
        const int tap = kb % 9;
        const int shift = mode == 1 ? 8 : 9 + (tap / 3 - 1) * 8 + (tap % 3 - 1);
        const uint64_t bd = b_desc + (uint64_t)((kb % nb) * N * ROWB >> 4);
        for (int t = 0; t < MT; ++t) {
          const uint64_t ad = a_desc + (uint64_t)(((uint32_t)(t * 128 + shift) * ROWB) >> 4);
          const uint32_t d = tmem + (uint32_t)((mode == 2 ? 0 : t) * N);
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) umma(d, ad + 2 * k, bd + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    const long long t_issue = clock64();
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@!p bra W_%=;\n}\n" ::"r"(smem_u32(bar)) : "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t_issue - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int KC>
void run(int N, int MT, int mode, int grid, long long* d_out) {
  const int supers = 16;
  const size_t smem = 200 * 1024;                              // one CTA per SM
  cudaFuncSetAttribute(k_ubench<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  for (int rep = 0; rep < 2; ++rep) k_ubench<KC><<<grid, 64, smem>>>(N, MT, supers, mode, d_out);
  long long h[2] = {0, 0};
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  const double n_mma = (double)supers * 9 * (64 / KC) * MT * (KC / 16);
  const double floor_cyc = 128.0 * N / 256.0, smem_cyc = (128.0 + N) * 32.0 / 128.0;
  printf("KC=%2d N=%3d MT=%d mode=%d grid=%3d: %6.1f cycles/MMA (issue loop alone %6.1f); tensor floor %5.1f, operand bytes/128 = %5.1f  %s\n",
         KC, N, MT, mode, grid, h[0] / n_mma, h[1] / n_mma, floor_cyc, smem_cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  for (int grid : {1, 148}) {
    for (int mode = 0; mode < 3; ++mode) {
      run<64>(64, 4, mode, grid, d_out);
      run<64>(128, 4, mode, grid, d_out);
      run<64>(256, 2, mode, grid, d_out);
    }
    run<64>(16, 8, 0, grid, d_out);
    run<64>(32, 8, 0, grid, d_out);
    run<32>(64, 4, 0, grid, d_out);
    run<16>(16, 8, 0, grid, d_out);
  }
  return 0;
}
