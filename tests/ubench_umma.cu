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

// `copy` > 0: warp 0 meanwhile streams 16 KB bulk copies (global -> shared, the weight ring's traffic) into a separate
// shared-memory region, `copy` of them in flight, to see what concurrent TMA writes cost the MMA's operand reads.
// mode 0: tap shifts of a W = 7 board ((tap/3-1)*8 + tap%3-1 rows); mode 1: no shift (8-row aligned starts);
// mode 2: like 0 but every MMA of a k-block goes to ONE accumulator tile (MT = 1 behaviour, dependent chain)
// mode 3: like 0 plus the streaming-weights protocol per k-block: try_wait on the barrier committed 4 blocks ago
//         (always complete by then), tcgen05.fence::after_thread_sync, the MMAs, tcgen05.commit to the slot's barrier
template <int KC>
__global__ void __launch_bounds__(64 + 8 * 32, 1) k_ubench(int N, int MT, int supers, int mode, long long* out, const uint8_t* src, int copy) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int ROWB = KC * 2;
  const int a_rows = MT * 128 + 32;
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)a_rows * ROWB;                    // 1024-aligned: a_rows * ROWB is a multiple of 1024
  const int nb = 4;                                            // weight blocks cycled through
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + (size_t)nb * N * ROWB);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  uint64_t* cbar = bar + 2;                                    // [8] copy-slot barriers
  volatile int* stop = reinterpret_cast<volatile int*>(bar + 10);
  uint8_t* sC = reinterpret_cast<uint8_t*>(bar + 32);          // copy ring: `copy` slots of 16 KB (16-byte aligned)
  for (int i = threadIdx.x; i < (a_rows + nb * N) * ROWB / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  // warp index / TMEM base through a shuffle: known warp-uniform, so the issue loop lives in uniform registers (a
  // divergent `threadIdx.x == 32` loop needs five R2UR moves + ELECT per MMA and measures its own issue overhead)
  const int warp = __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(cbar + i)) : "memory");
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 12 + i)) : "memory");
    *stop = 0;
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
  const uint32_t tmem = __shfl_sync(0xFFFFFFFFu, *slot, 0);
  if (threadIdx.x == 0 && copy > 0) {
    long long n = 0;
    const long long c0 = clock64();
    for (int i = 0; i < copy; ++i) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(cbar + i)), "r"(16384) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(sC + i * 16384)), "l"(src + ((size_t)i * 16384) % (288 * 1024)), "r"(16384), "r"(smem_u32(cbar + i)) : "memory");
    }
    for (long long k = 0;; ++k) {
      const int i = (int)(k % copy);
      const uint32_t ph = (uint32_t)(k / copy) & 1;
      asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W_%=;\n}\n" ::"r"(smem_u32(cbar + i)), "r"(ph) : "memory");
      ++n;
      if (*stop) break;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(cbar + i)), "r"(16384) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(sC + i * 16384)), "l"(src + ((size_t)(k + copy) * 16384) % (288 * 1024)), "r"(16384), "r"(smem_u32(cbar + i)) : "memory");
    }
    // drain what is still in flight before the CTA may exit
    for (long long k2 = n; k2 < n + copy - 1; ++k2) {
      const int i = (int)(k2 % copy);
      const uint32_t ph = (uint32_t)(k2 / copy) & 1;
      asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W_%=;\n}\n" ::"r"(smem_u32(cbar + i)), "r"(ph) : "memory");
    }
    if (blockIdx.x == 0) { out[2] = n * 16384; out[3] = clock64() - c0; }
  }
  if (warp >= 2) {
    // epilogue stand-in: drain accumulator columns with tcgen05.ld (lane quarter = warp % 4) until the MMAs are done
    const uint32_t tb = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256u;     // the "other" accumulator stage
    uint32_t acc = 0;
    while (!*stop) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(tb + (uint32_t)c * 16u));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) acc ^= v[i];
      }
    }
    if (acc == 0x12345678u) out[7] = acc;
  }
  if (warp == 1 && elect_one_sync()) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a_desc = make_desc<KC>(smem_u32(sA)), b_desc = make_desc<KC>(smem_u32(sB));
    const int n_kb = 9 * (64 / KC);                            // 64 input channels
    long long t0 = clock64();
    uint32_t ring = 0;
    for (int it = 0; it < supers; ++it) {
      for (int kb = 0; kb < n_kb; ++kb) {
        const uint32_t rs = ring & 3;
        if (mode == 3) {
          if (ring >= 4)
            asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W_%=;\n}\n"
                         ::"r"(smem_u32(bar + 12 + rs)), "r"(((ring >> 2) - 1) & 1) : "memory");
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const int tap = kb % 9;
        const int shift = mode == 1 ? 8 : 9 + (tap / 3 - 1) * 8 + (tap % 3 - 1);
        const uint64_t bd = b_desc + (uint64_t)((kb % nb) * N * ROWB >> 4);
        for (int t = 0; t < MT; ++t) {
          const uint64_t ad = a_desc + (uint64_t)(((uint32_t)(t * 128 + shift) * ROWB) >> 4);
          const uint32_t d = tmem + (uint32_t)((mode == 2 ? 0 : t) * N);
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) umma(d, ad + 2 * k, bd + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        if (mode == 3)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + 12 + rs)) : "memory");
        ++ring;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    const long long t_issue = clock64();
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@!p bra W_%=;\n}\n" ::"r"(smem_u32(bar)) : "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t_issue - t0; }
    *stop = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int KC>
void run(int N, int MT, int mode, int grid, long long* d_out, const uint8_t* d_src = nullptr, int copy = 0, int readers = 0) {
  const int supers = 16;
  const size_t smem = 200 * 1024;                              // one CTA per SM
  cudaFuncSetAttribute(k_ubench<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  for (int rep = 0; rep < 2; ++rep) k_ubench<KC><<<grid, 64 + 32 * readers, smem>>>(N, MT, supers, mode, d_out, d_src, copy);
  long long h[4] = {0, 0, 0, 1};
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  const double n_mma = (double)supers * 9 * (64 / KC) * MT * (KC / 16);
  const double floor_cyc = 128.0 * N / 256.0, smem_cyc = (128.0 + N) * 32.0 / 128.0;
  printf("KC=%2d N=%3d MT=%d mode=%d grid=%3d: %6.1f cycles/MMA (issue loop alone %6.1f); tensor floor %5.1f, operand bytes/128 = %5.1f  %s\n",
         KC, N, MT, mode, grid, h[0] / n_mma, h[1] / n_mma, floor_cyc, smem_cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
  if (readers > 0) printf("      with %d warps draining the other accumulator stage through tcgen05.ld\n", readers);
  if (copy > 0) printf("      with %d x 16 KB bulk copies in flight into shared memory: %.1f B/cycle copied per SM\n", copy, (double)h[2] / (double)h[3]);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  uint8_t* d_src;
  cudaMalloc(&d_src, 512 * 1024);
  cudaMemset(d_src, 0, 512 * 1024);
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
  // the gomoku shape (N = 128, MT = 2) with the weight ring's traffic beside it
  for (int copy : {0, 1, 2, 4}) run<64>(128, 2, 0, 148, d_out, d_src, copy);
  for (int copy : {0, 2}) run<64>(64, 4, 0, 148, d_out, d_src, copy);
  // the streaming protocol (wait / fence / commit per k-block) at the gomoku and connect4 shapes
  run<64>(128, 2, 3, 148, d_out);
  run<64>(128, 2, 3, 148, d_out, d_src, 2, 8);
  run<64>(64, 4, 3, 148, d_out);
  // ... and with epilogue warps reading TMEM meanwhile (accumulators use columns [0, MT*N) <= 256 here)
  for (int readers : {4, 8}) {
    run<64>(64, 4, 0, 148, d_out, d_src, 0, readers);
    run<64>(128, 2, 0, 148, d_out, d_src, 0, readers);
    run<64>(128, 2, 0, 148, d_out, d_src, 2, readers);
  }
  return 0;
}
