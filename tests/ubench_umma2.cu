// Micro-benchmark (bring-up tool, not part of the product): the pace of tcgen05.mma with cta_group::2 (a CTA pair on
// one TPC, M = 256 per instruction, the B operand split across the pair) against cta_group::1 at k_conv_tc's shapes.
// The convolution's issue loop is reproduced (9 taps x 64 input channels, shifted A start rows, MT accumulator tiles)
// with both operands in shared memory, no TMA and no epilogue.  Per instruction a cta_group::2 MMA does the work of
// TWO cta_group::1 MMAs (one per SM) while each SM reads 128 rows of A but only N/2 rows of B from its shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/ubench_umma2 tests/ubench_umma2.cu && tests/ubench_umma2
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int CTAS>
__device__ __forceinline__ void umma(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  if constexpr (CTAS == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xFFFFFFFF;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}
template <int KC>
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  constexpr uint64_t layout = KC == 64 ? 2 : (KC == 32 ? 4 : 6);
  constexpr uint64_t sbo = (8 * KC * 2) >> 4;
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .u32 spins;\nmov.u32 spins, 0;\nW_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\n"
      "add.u32 spins, spins, 1;\nsetp.lt.u32 p, spins, 0x04000000;\n@p bra W_%=;\ntrap;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// CTAS = 1: every CTA runs the loop on its own (cluster of 1).  CTAS = 2: clusters of two CTAs; rank 0 issues M = 256
// MMAs over both CTAs' A tiles and B halves, the accumulators land in both CTAs' TMEM, the commit is multicast.
template <int KC, int CTAS>
__global__ void __launch_bounds__(64, 1) k_ubench2(int N, int MT, int supers, int aligned, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int ROWB = KC * 2;
  const int a_rows = MT * 128 + 32;
  const int nb = 4, n_local = N / CTAS;                        // B rows held by this CTA
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)a_rows * ROWB;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + (size_t)nb * n_local * ROWB);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (a_rows + nb * n_local) * ROWB / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  const int warp = __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0);
  uint32_t rank = 0;
  if (CTAS == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    if (CTAS == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CTAS == 2) cg::this_cluster().sync(); else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xFFFFFFFFu, *slot, 0);
  if (warp == 1 && rank == 0 && elect_one_sync()) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (((128u * CTAS) >> 4) << 24);
    const uint64_t a_desc = make_desc<KC>(smem_u32(sA)), b_desc = make_desc<KC>(smem_u32(sB));
    const int n_kb = 9 * (64 / KC);
    const long long t0 = clock64();
    for (int it = 0; it < supers; ++it) {
      for (int kb = 0; kb < n_kb; ++kb) {
        const int tap = kb % 9;
        const int shift = aligned ? 8 : 9 + (tap / 3 - 1) * 8 + (tap % 3 - 1);
        const uint64_t bd = b_desc + (uint64_t)((kb % nb) * n_local * ROWB >> 4);
        for (int t = 0; t < MT; ++t) {
          const uint64_t ad = a_desc + (uint64_t)(((uint32_t)(t * 128 + shift) * ROWB) >> 4);
          const uint32_t d = tmem + (uint32_t)(t * N);
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) umma<CTAS>(d, ad + 2 * k, bd + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
      }
    }
    if (CTAS == 1)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
    const long long t_issue = clock64();
    bar_wait(bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t_issue - t0; }
  }
  if (CTAS == 2 && rank == 1 && threadIdx.x == 32) bar_wait(bar, 0);      // the peer's copy of the multicast commit
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CTAS == 2) cg::this_cluster().sync(); else __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CTAS == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int KC, int CTAS>
void run(int N, int MT, int aligned, int grid, long long* d_out) {
  const int supers = 16;
  const size_t smem = 200 * 1024;
  cudaFuncSetAttribute(k_ubench2<KC, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(grid); lc.blockDim = dim3(64); lc.dynamicSmemBytes = smem;
  cudaLaunchAttribute la[1];
  la[0].id = cudaLaunchAttributeClusterDimension;
  la[0].val.clusterDim.x = CTAS; la[0].val.clusterDim.y = 1; la[0].val.clusterDim.z = 1;
  lc.attrs = la; lc.numAttrs = 1;
  cudaError_t e = cudaSuccess;
  for (int rep = 0; rep < 2 && e == cudaSuccess; ++rep) e = cudaLaunchKernelEx(&lc, k_ubench2<KC, CTAS>, N, MT, supers, aligned, d_out);
  long long h[2] = {0, 0};
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  const double n_mma = (double)supers * 9 * (64 / KC) * MT * (KC / 16);
  const double per_sm = h[0] / n_mma;                          // one instruction = one 128-row MMA PER SM in both modes
  const double math = 128.0 * N / 256.0, opnd = (128.0 + (double)N / CTAS) * 32.0 / 128.0;
  printf("cta_group::%d KC=%2d N=%3d MT=%d %s grid=%3d: %6.1f cycles per (128 x N x 16 per SM) MMA (issue loop %6.1f); math floor %5.1f, "
         "operand fetch model %5.1f -> %.2f of the tensor peak  %s\n", CTAS, KC, N, MT, aligned ? "aligned" : "taps   ", grid, per_sm,
         h[1] / n_mma, math, opnd, math / per_sm, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  for (int grid : {2, 148}) {
    for (int aligned : {0, 1}) {
      run<64, 1>(64, 4, aligned, grid, d_out);
      run<64, 2>(64, 4, aligned, grid, d_out);
      run<64, 1>(128, 2, aligned, grid, d_out);
      run<64, 2>(128, 2, aligned, grid, d_out);
      run<64, 2>(256, 2, aligned, grid, d_out);
    }
    run<32, 1>(32, 8, 0, grid, d_out);
    run<32, 2>(32, 8, 0, grid, d_out);
    run<16, 1>(16, 8, 0, grid, d_out);
    run<16, 2>(16, 8, 0, grid, d_out);
  }
  return 0;
}
