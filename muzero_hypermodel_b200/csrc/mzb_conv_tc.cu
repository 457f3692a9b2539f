// K6: 3x3 convolution as an implicit GEMM on the 5th-generation tensor cores (tcgen05 / UMMA),
// bf16 operands, fp32 accumulators in TMEM, folded batch-norm + action-plane term + residual + ReLU fused
// into the epilogue.  Reference: conv3x3 + BatchNorm2d + relu / ResidualBlock (models.py:206-229),
// DynamicsNetwork's first convolution with the action plane (:363, :551-568).
//
// Formulation.  Activations live in the PADDED NHWC layout of mzb_resnet.cuh: flat rows of C channels,
// (H+1)*(W+1) rows per image with zero pad rows, so the 3x3 neighbour (dy,dx) of EVERY row is the row at
// flat offset dy*(W+1)+dx.  One CTA owns MT tiles of 128 consecutive rows (M = 128 per UMMA):
//   * TMA loads the rows [m0 - halo, m0 + MT*128 + halo) ONCE into shared memory (SWIZZLE_128B/64B/32B by
//     channel-chunk width) - the input is read from L2/HBM once, not once per tap;
//   * the A operand of tap (dy,dx) is the SAME shared-memory tile addressed through a UMMA descriptor whose
//     start address is shifted by (halo + dy*(W+1) + dx) rows.  The swizzle XOR is a function of the absolute
//     shared-memory address bits (measured: every row shift is exact with descriptor base_offset = 0 for
//     SWIZZLE_128B/64B/32B, tests/debug_conv_tc.py), so the im2col matrix is never materialised;
//   * the B operand (weights [C_out][9*C_in], K-major) stays resident in shared memory when it fits (loaded once
//     per CTA) and streams through a 4-stage TMA/mbarrier ring otherwise, each k-block feeding all MT accumulators;
//   * the kernel is persistent (one CTA per SM loops over super-tiles of MT tiles) and warp specialised:
//     warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM allocator, warps 2-9 = two epilogue
//     warpgroups (tcgen05.ld 32x32b, scale/shift/plane/residual/ReLU, bf16 pack, 16-byte stores), with the
//     activation stages and the TMEM accumulators double buffered.
// Rows that are pad positions compute garbage that is not stored (pads stay zero for the next layer) - or, in the
// stem's ZP instantiation, are stored as zeros.
#include <cuda.h>
#include <stdlib.h>

#include "mzb_resnet_model.h"

namespace {

constexpr int kStages = 4;
constexpr int kEpiWarpsWide = 8;                   // two epilogue warpgroups: layers with C_out >= 64
constexpr int kEpiWarpsNarrow = 16;                // four for narrow layers, whose pace the per-tile epilogue latency sets

struct TcArgs {
  int B, H, W, Cin, Cout, R_img, mt, n_chunks, relu, ncols, tail_rows, b_resident, zero_pads, plane_classes;
  long long rows_valid;                 // B * R_img
  long long rows_cover;                 // rows the super-tiles cover: rows_valid (+ the trailing halo when zero_pads)
  const float* scale; const float* shift; const float* plane; const float* plane_table;
  const __nv_bfloat16* residual; __nv_bfloat16* y;
  long long* debug;                     // optional timeline of CTA 0: [role][iteration][4] clock64 stamps
  // fused head projection (the heads' 1x1 convolutions, models.py:398-404, 447-456): proj_out[b][r][pos] =
  // sum_c y[b,pos,c] * proj_w[r][c] on the bf16-rounded outputs; bias is added by the head kernel
  const float* proj_w; float* proj_out; int proj_r, proj_hw;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // One asm statement (no C++ control flow: the compiler keeps the surrounding role loop warp-uniform, so
  // addresses and descriptors stay in uniform registers).  The spin is bounded: a protocol bug traps (launch
  // error) instead of hanging the GPU.
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .u32 spins;\n"
      "mov.u32 spins, 0;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "add.u32 spins, spins, 1;\n"
      "setp.lt.u32 p, spins, 0x08000000;\n"
      "@p bra WAIT_%=;\n"
      "trap;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// cta_group::2: ONE instruction issued by the leader CTA of a pair multiplies both CTAs' 128-row A tiles (M = 256) by a
// B tile of which each CTA's shared memory holds half the rows; the accumulators land in both CTAs' TMEM.
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of all MMAs issued so far -> the barrier at the same shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xFFFFFFFF;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO=1 | SBO>>4 | version 1 |
// base_offset 0 | layout type.  The start address may sit at ANY row of a TMA-written swizzled tile.
template <int KC>
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  constexpr uint64_t layout = KC == 64 ? 2 : (KC == 32 ? 4 : 6);         // SWIZZLE_128B / 64B / 32B
  constexpr uint64_t sbo = (8 * KC * 2) >> 4;                            // 8 rows of KC bf16
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

// Persistent, warp-specialised kernel.  A CTA owns a contiguous range of 128-row tiles and loops over it in super-tiles
// (up to MT tiles of 128 rows):
//   warp 0        TMA producer: activation rows of super-tile i+1 are loaded while i is multiplied (2 A stages);
//                 weights either stay resident in shared memory (loaded once) or stream through a ring
//   warp 1        MMA issuer; accumulators are double buffered in TMEM (2 x MT x C_out columns), so the MMAs of
//                 super-tile i+1 run while the epilogue drains super-tile i
//   warps 2..9    two epilogue warpgroups (TMEM lane quarter = warp % 4); work items (tile, 64-column group) are
//                 dealt round-robin to the groups
// ZP / PR: the rarely used epilogue features (pad-row zeroing of the stem; fused head projection with PR rows, a
// compile-time count so the dot products are straight-line code) are separate instantiations, so the common layer
// keeps its register allocation.  RES: weights resident in shared memory (true) or streamed through the ring (false) -
// compile-time because the MMA stream is issue-bound and every instruction between two bursts of MMAs shows.
// NC: compile-time channel-chunk count of the streamed-weights path (0 = runtime).  With NC = 2 (gomoku: 128 input
// channels) the 18 k-blocks of a super-tile are fully unrolled and the ring is 3 deep - a depth that divides 18 - so a
// block's ring slot (kb % 3) and barrier parity ((kb / 3) & 1) are compile-time constants: the slot counter, the
// divisions and the R2UR moves that sat between two bursts of MMAs (99 cycles per MMA in situ against 64 for the
// instruction stream alone) disappear.
// PAIR: two CTAs on one TPC form a cluster and share every MMA (cta_group::2, M = 256): each CTA loads its own
// super-tile of activations but only HALF the rows of every weight block, so an MMA fetches 128 + N/2 instead of
// 128 + N operand rows from each SM's shared memory - the fetch, not the math, bounds an N = 64 MMA (tests/ubench_umma2.cu:
// 43 against 53 cycles per 128 x 64 x 16 MMA per SM).  The leader CTA (cluster rank 0) issues; the peer's "activations
// landed" and both CTAs' "accumulator stage drained" events reach the leader's barriers through cluster-scope arrives,
// the MMA completion is committed to both CTAs' barriers by multicast.  Resident weights only (RES).
template <int KC, bool ZP, int PR, int kEpiWarps, bool RES, int NC = 0, bool PAIR = false>
__global__ void __launch_bounds__(64 + kEpiWarps * 32, 1)
k_conv_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAtail,
          const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
  constexpr int ROWB = KC * 2;                                           // bytes per shared-memory row
  constexpr int kThreads = 64 + kEpiWarps * 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // warp index / TMEM base are broadcast through a shuffle so the compiler knows they are warp-uniform and keeps the
  // role loops (addresses, descriptors, barrier phases) in uniform registers
  const int warp = __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr int RING = NC == 2 ? 3 : kStages;
  const int n_chunks = NC > 0 ? NC : a.n_chunks;
  const int MT = a.mt, N = a.Cout, halo = geo_halo(a.W);
  const int a_rows = MT * 128 + a.tail_rows;
  const uint32_t a_chunk_bytes = (uint32_t)a_rows * ROWB;
  const uint32_t a_stage_bytes = (uint32_t)n_chunks * a_chunk_bytes;
  uint32_t rank = 0;                                                     // cluster rank (PAIR): 0 = leader
  if (PAIR) {
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    // broadcast through a shuffle: the compiler then knows the rank is warp-uniform.  Without it the role branch on the
    // rank counts as divergent and ptxas wraps EVERY tcgen05.mma of the issue loop in an ELECT / BRA.U.ANY loop (7
    // instructions per MMA instead of back-to-back UTCHMMA)
    rank = __shfl_sync(0xFFFFFFFFu, rank, 0);
  }
  const int NB = PAIR ? N / 2 : N;                                       // weight rows this CTA holds of every k-block
  const uint32_t b_block_bytes = (uint32_t)NB * ROWB;
  const int NKB = 9 * n_chunks;
  const int b_slots = RES ? NKB : RING;
  uint8_t* sA = smem;
  uint8_t* sB = sA + 2 * (size_t)a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)b_slots * b_block_bytes);
  uint64_t* a_full = bars;                 // [2]
  uint64_t* a_empty = bars + 2;            // [2]
  uint64_t* acc_full = bars + 4;           // [2]
  uint64_t* acc_empty = bars + 6;          // [2]
  uint64_t* b_full = bars + 8;             // [kStages] (ring) / [0] = resident weights landed
  uint64_t* b_empty = bars + 8 + kStages;  // [kStages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + 2 * kStages);
  uint64_t* peer_a_full = bars + 10 + 2 * kStages;   // [2] (PAIR, leader): the peer CTA's activations of stage s landed
  float* s_scale = reinterpret_cast<float*>(bars + 12 + 2 * kStages);
  float* s_shift = s_scale + N;
  float* s_proj = s_shift + N;             // [PR][N]
  // The action plane's term acc += plane[b] * table[position][channel] (DynamicsNetwork's first convolution) from shared
  // memory.  Read from global memory, a warp's 32 rows (32 positions) made every float4 of the table an uncoalesced
  // request - 16 per row, each touching up to 32 lines: the layer took 110 us against 67 us for a plain one.  The table
  // only depends on which taps fall outside the board, i.e. on the border class of the position (top / inner / bottom
  // line x left / inner / right column): nine rows when the board has an interior, else one row per position.
  float* s_ptab = s_proj + PR * N;
  // Work split: every CTA owns one CONTIGUOUS range of 128-row tiles (sizes differ by at most one tile) and walks it in
  // super-tiles of up to MT tiles; only the range's last super-tile may be short.  A strided split in whole super-tiles
  // would leave a tail wave in which most SMs idle (connect4: 12.1 waves of work took 13).
  // PAIR: the range belongs to the cluster; per iteration the leader takes MT tiles and the peer the next MT (the peer
  // may get fewer, or none, in the last iteration - it still takes part in every barrier).
  const long long total_tiles = (a.rows_cover + 127) / 128;
  const long long n_owner = PAIR ? gridDim.x / 2 : gridDim.x, owner = PAIR ? blockIdx.x / 2 : blockIdx.x;
  const long long t_base = total_tiles / n_owner, t_rem = total_tiles % n_owner;
  const long long tile_begin = owner * t_base + (owner < t_rem ? owner : t_rem);
  const long long tile_end = tile_begin + t_base + (owner < t_rem ? 1 : 0);
  const int it_stride = PAIR ? 2 * MT : MT;
  const int n_it = (int)((tile_end - tile_begin + it_stride - 1) / it_stride);
  const long long tile_first = tile_begin + (long long)rank * MT;
  auto mt_of = [&](long long tile0) -> int { const long long r = tile_end - tile0; return r <= 0 ? 0 : (r < MT ? (int)r : MT); };

  // Programmatic dependent launch: the next kernel of the stream may be scheduled as soon as this grid's CTAs free
  // their SMs (its prologue - barrier setup, TMEM allocation, weight loads - then overlaps this grid's tail); every
  // thread that reads a predecessor's output executes griddepcontrol.wait first.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(a_full + i, 1); mbar_init(a_empty + i, 1); mbar_init(acc_full + i, 1);
      mbar_init(acc_empty + i, PAIR ? 2 * kEpiWarps : kEpiWarps);         // PAIR (leader): both CTAs' epilogue warps arrive here
      mbar_init(peer_a_full + i, 1);
    }
    for (int s = 0; s < kStages; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(a.ncols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(a.ncols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  for (int i = threadIdx.x; i < N; i += kThreads) { s_scale[i] = a.scale[i]; s_shift[i] = a.shift[i]; }
  if (PR > 0) for (int i = threadIdx.x; i < PR * N; i += kThreads) s_proj[i] = i < a.proj_r * N ? a.proj_w[i] : 0.0f;
  if (a.plane) {                           // weights, not activations: safe before griddepcontrol.wait
    const int rows = a.plane_classes ? 9 : a.H * a.W;
    for (int i = threadIdx.x; i < rows * N; i += kThreads) {
      const int r = i / N, c = i - r * N;
      int pos = r;
      if (a.plane_classes) {
        const int cy = r / 3, cx = r - cy * 3;
        pos = (cy == 0 ? 0 : (cy == 2 ? a.H - 1 : 1)) * a.W + (cx == 0 ? 0 : (cx == 2 ? a.W - 1 : 1));
      }
      s_ptab[i] = a.plane_table[(size_t)pos * N + c] * a.scale[c];        // the scale is folded into the weights, so into this term too
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) cluster_sync_all();                // both CTAs' barriers are initialised before any remote arrive / multicast
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = __shfl_sync(0xFFFFFFFFu, *tmem_slot, 0);

  if (warp == 0) {
    // ---------------- TMA producer.  The whole warp runs the (warp-uniform) control flow so that addresses and
    // descriptors live in uniform registers; one elected lane issues the copies.
    if (elect_one_sync()) {
      const bool leader = true;
      if (RES) {
        if (leader) mbar_expect_tx(b_full, (uint32_t)NKB * b_block_bytes);
        for (int kb = 0; kb < NKB; ++kb) {
          const int tap = kb / n_chunks, j = kb % n_chunks;
          if (leader) tma_load_2d(smem_u32(sB + (size_t)kb * b_block_bytes), &tmB, tap * a.Cin + j * KC, (int)rank * NB, b_full);
        }
      }
      asm volatile("griddepcontrol.wait;" ::: "memory");      // the activations come from the preceding kernel
      long long ring = 0;
      int it = 0;
      for (; it < n_it; ++it) {
        const long long tile0 = tile_first + (long long)it * it_stride;
        const int s = it & 1;
        const int mt_cur = mt_of(tile0);
        const long long p0 = clock64();
        if (it >= 2) mbar_wait(a_empty + s, ((it >> 1) - 1) & 1);
        if (a.debug && blockIdx.x == 0 && it < 32 && leader) { long long* d = a.debug + (3 * 32 + it) * 4; d[0] = p0; d[1] = clock64(); }
        const long long m0 = tile0 * 128;
        if (leader) mbar_expect_tx(a_full + s, mt_cur > 0 ? (uint32_t)n_chunks * (uint32_t)(mt_cur * 128 + a.tail_rows) * ROWB : 0u);
        for (int j = 0; j < n_chunks && mt_cur > 0; ++j) {
          uint8_t* dst = sA + (size_t)s * a_stage_bytes + (size_t)j * a_chunk_bytes;
          for (int box = 0; box < mt_cur; ++box)
            if (leader) tma_load_2d(smem_u32(dst + (size_t)box * 128 * ROWB), &tmA, j * KC, (int)(m0 + (long long)box * 128), a_full + s);
          if (leader) tma_load_2d(smem_u32(dst + (size_t)mt_cur * 128 * ROWB), &tmAtail, j * KC, (int)(m0 + (long long)mt_cur * 128), a_full + s);
        }
        if (!RES) {
          if constexpr (NC == 2) {
#pragma unroll
            for (int kb = 0; kb < 9 * NC; ++kb) {
              const int rs = kb % RING;                                   // 18 % 3 == 0: slot and parity do not depend on `it`
              if (it > 0 || kb >= RING) mbar_wait(b_empty + rs, (uint32_t)((kb / RING) + 1) & 1);
              if (leader) mbar_expect_tx(b_full + rs, b_block_bytes);
              if (leader) tma_load_2d(smem_u32(sB + (size_t)rs * b_block_bytes), &tmB, (kb / NC) * a.Cin + (kb % NC) * KC, 0, b_full + rs);
            }
          } else {
            for (int kb = 0; kb < NKB; ++kb, ++ring) {
              const int rs = (int)(ring % kStages);
              if (ring >= kStages) mbar_wait(b_empty + rs, (uint32_t)((ring / kStages) - 1) & 1);
              if (leader) mbar_expect_tx(b_full + rs, b_block_bytes);
              const int tap = kb / n_chunks, j = kb % n_chunks;
              if (leader) tma_load_2d(smem_u32(sB + (size_t)rs * b_block_bytes), &tmB, tap * a.Cin + j * KC, 0, b_full + rs);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: warp-uniform loop, tcgen05.mma / commit issued by one elected lane
    if (elect_one_sync()) {
      if (PAIR && rank != 0) {
        // peer CTA of a pair: forward "my activations of stage s (and my half of the weights) landed" to the leader,
        // which issues the MMAs for both CTAs
        if (RES) mbar_wait(b_full, 0);
        for (int it = 0; it < n_it; ++it) {
          const int s = it & 1;
          mbar_wait(a_full + s, (it >> 1) & 1);
          mbar_arrive_cluster(peer_a_full + s, 0);
        }
      } else {
      const bool leader = true;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (((PAIR ? 256u : 128u) >> 4) << 24);
      const uint64_t desc_hi = make_desc<KC>(0);
      const uint64_t b_base_desc = desc_hi | (uint64_t)((smem_u32(sB) >> 4) & 0x3FFF);
      if (RES) mbar_wait(b_full, 0);
      uint32_t ring = 0;
      int it = 0;
      for (; it < n_it; ++it) {
        const long long tile0 = tile_first + (long long)it * it_stride;
        const int s = it & 1;
        const int mt_cur = mt_of(tile0);                    // the leader's count; the peer's is never larger
        const long long c0 = clock64();
        mbar_wait(a_full + s, (it >> 1) & 1);
        if (PAIR) mbar_wait(peer_a_full + s, (it >> 1) & 1);
        const long long c1 = clock64();
        if (it >= 2) mbar_wait(acc_empty + s, ((it >> 1) - 1) & 1);
        const long long c2 = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // descriptors differ only in the 14-bit start-address field (16-byte units): one add per MMA
        const uint64_t a_stage_desc = desc_hi | (uint64_t)((smem_u32(sA + (size_t)s * a_stage_bytes) >> 4) & 0x3FFF);
        const uint32_t d_base = tmem_base + (uint32_t)(s * MT * N);
        // The k-block loop is unrolled over the 9 taps so that the tap's row shift is a compile-time expression and the
        // per-block bookkeeping stays a handful of uniform-datapath adds: descriptor arithmetic in vector registers
        // (a division by n_chunks, R2UR moves) sat between the MMA bursts and was NOT overlapped with them - the in-situ
        // cost per MMA was 73 cycles against the 51 the same instruction stream takes alone (tests/ubench_umma.cu).
        const uint32_t blk16 = b_block_bytes >> 4, chunk16 = a_chunk_bytes >> 4;
        const int pitch = geo_pitch(a.W);
        int kb = 0;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int shift = (tap / 3 - 1) * pitch + (tap % 3 - 1);
          const uint64_t a_tap_desc = a_stage_desc + (uint64_t)(((uint32_t)(halo + shift) * ROWB) >> 4);
          auto k_block = [&](const int j) {
            uint64_t bd;
            int rs = 0;
            if (RES) {
              bd = b_base_desc + (uint64_t)((uint32_t)kb * blk16);
            } else if constexpr (NC == 2) {
              const int kbc = tap * NC + j;                               // compile-time after unrolling
              rs = kbc % RING;
              mbar_wait(b_full + rs, (uint32_t)(kbc / RING) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              bd = b_base_desc + (uint64_t)((uint32_t)rs * blk16);
            } else {
              rs = (int)(ring % kStages);
              mbar_wait(b_full + rs, (ring / kStages) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              bd = b_base_desc + (uint64_t)((uint32_t)rs * blk16);
              ++ring;
            }
            uint64_t ad = a_tap_desc + (uint64_t)((uint32_t)j * chunk16);
            uint32_t d = d_base;
            for (int t = 0; t < mt_cur; ++t, ad += (128 * ROWB) >> 4, d += (uint32_t)N) {
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)
                if (leader) {
                  if (PAIR) umma_bf16_pair(d, ad + 2 * k, bd + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                  else umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
            }
            if (!RES && leader) umma_commit(b_empty + rs);
            ++kb;
          };
          if constexpr (NC > 0) {
#pragma unroll
            for (int j = 0; j < NC; ++j) k_block(j);
          } else {
            for (int j = 0; j < n_chunks; ++j) k_block(j);
          }
        }
        if (leader) {
          if (PAIR) { umma_commit_pair(a_empty + s); umma_commit_pair(acc_full + s); }      // both CTAs' barriers
          else {
            umma_commit(a_empty + s);            // activation stage reusable once these MMAs retire
            umma_commit(acc_full + s);           // accumulators of this super-tile complete
          }
        }
        if (a.debug && blockIdx.x == 0 && it < 32 && leader) { long long* d = a.debug + (0 * 32 + it) * 4; d[0] = c0; d[1] = c1; d[2] = c2; d[3] = clock64(); }
      }
      }
    }
  } else {
    // ---------------- epilogue warpgroups
    // A group of four warps (TMEM lane quarter = warp % 4, thread = row) drains one 128-row tile at a time in steps of
    // 32 columns.  The loop is software pipelined: the residual bytes of the NEXT step are requested before the current
    // step's accumulators are converted, so the global-memory latency (the epilogue's dominant cost: it, not the MMA
    // issue, set the kernel's period) hides behind the arithmetic of the step before.
    const int q = warp & 3, group = (warp - 2) >> 2;
    const int ncg = (N + 63) / 64;
    constexpr int NG = kEpiWarps / 4;
    const uint32_t s_scale_u32 = smem_u32(s_scale), s_shift_u32 = smem_u32(s_shift), s_proj_u32 = smem_u32(s_proj), s_ptab_u32 = smem_u32(s_ptab);
    const uint32_t R_img = (uint32_t)a.R_img, Wp = (uint32_t)geo_pitch(a.W);
    struct Item { long long row_off; int b, pos, cls, t, g0, gw; uint32_t m; bool valid; };
    auto get_item = [&](uint32_t m0, int item, Item& I) {
      I.t = item / ncg;
      I.g0 = (item - I.t * ncg) * 64;
      I.gw = N - I.g0 < 64 ? N - I.g0 : 64;                      // columns in this group (multiple of 16)
      I.m = m0 + (uint32_t)(I.t * 128 + q * 32 + lane);           // rows_cover < 2^31 (checked by the host)
      const uint32_t bb = I.m / R_img, rem = I.m - bb * R_img;
      const uint32_t yy = rem / Wp, xx = rem - yy * Wp;
      I.b = (int)bb;
      I.valid = (long long)I.m < a.rows_valid && geo_is_pixel((int)yy, (int)xx, a.W);
      I.row_off = ((long long)I.m + halo) * (long long)N;
      I.pos = (int)((yy - 1) * (uint32_t)a.W + xx);
      I.cls = a.plane_classes ? (int)((yy == 1u ? 0u : (yy == (uint32_t)a.H ? 2u : 1u)) * 3u + (xx == 0u ? 0u : (xx == (uint32_t)a.W - 1u ? 2u : 1u)))
                              : I.pos;
    };
    // 256-bit global accesses for the wide layers (measured: connect4 64 -> 64 channels 79.6 -> 71.6 us per layer in step,
    // gomoku 162 -> 154); the 16-channel stem layers of breakout were 10 % slower with them and keep the 128-bit form
    const bool wide_ls = N >= 64;
    auto load_res = [&](const Item& I, int hh, uint4 (&r)[4]) {
      if (I.valid && a.residual) {
        const uint4* rp = reinterpret_cast<const uint4*>(a.residual + I.row_off + I.g0 + hh * 32);
        const int cw = I.gw - hh * 32;
        // 32-byte requests (LDG.256): half the load instructions of the 16-byte form; a row-per-lane access touches 32
        // lines per instruction whatever its width, and every such wavefront is taken from the MMA's operand fetch
        if (!wide_ls) {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (i * 8 < cw) r[i] = __ldg(rp + i);
          return;
        }
#pragma unroll
        for (int i = 0; i < 4; i += 2)
          if (i * 8 < cw)
            asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[i].x), "=r"(r[i].y), "=r"(r[i].z), "=r"(r[i].w), "=r"(r[i + 1].x), "=r"(r[i + 1].y),
                           "=r"(r[i + 1].z), "=r"(r[i + 1].w) : "l"(rp + i));
      }
    };
    asm volatile("griddepcontrol.wait;" ::: "memory");        // residual / plane reads below: earlier kernels' outputs
    int it = 0, item_base = 0;
    for (; it < n_it; ++it) {
      const long long tile0 = tile_first + (long long)it * it_stride;
      const int s = it & 1;
      const int n_items = mt_of(tile0) * ncg;
      const uint32_t m0 = (uint32_t)(tile0 * 128);
      // items are dealt round-robin over ALL super-tiles (not restarted per super-tile)
      int item = (((group - item_base) % NG) + NG) % NG;
      item_base = (item_base + n_items) % NG;
      Item cur{}, nxt{};
      uint4 res[4], resn[4];
      bool have = item < n_items;
      if (have) {
        get_item(m0, item, cur);
        load_res(cur, 0, res);                                    // independent of the MMAs: in flight during the wait
      }
      const long long e0 = clock64();
      mbar_wait(acc_full + s, (it >> 1) & 1);
      const long long e1 = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      while (have) {
        const float pl = (cur.valid && a.plane) ? a.plane[cur.b] : 0.0f;
        const bool ptab = cur.valid && a.plane;
        const uint32_t ptab_u32 = s_ptab_u32 + (uint32_t)(cur.cls * N) * 4;
        float pacc[PR > 0 ? PR : 1];
#pragma unroll
        for (int r = 0; r < PR; ++r) pacc[r] = 0.0f;
        const int n_half = (cur.gw + 31) / 32;
        const int next_item = item + NG;
        for (int hh = 0; hh < n_half; ++hh) {
          const int c0 = cur.g0 + hh * 32, cw = cur.gw - hh * 32 < 32 ? cur.gw - hh * 32 : 32;   // 16 or 32 columns
          uint32_t v[32];
          const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * MT * N + cur.t * N + c0);
          tmem_ld16_nowait(tbase, v);
          if (cw > 16) tmem_ld16_nowait(tbase + 16, v + 16);
          // request the next step's residual row before touching this step's accumulators
          if (hh + 1 < n_half) {
            load_res(cur, hh + 1, resn);
          } else if (next_item < n_items) {
            get_item(m0, next_item, nxt);
            load_res(nxt, 0, resn);
          }
          tmem_wait_ld();
          if (!cur.valid) {
            // stem mode: the buffers change resolution between layers, so the pad rows (and the trailing halo) are
            // re-written as zeros by every layer instead of relying on a zeroed workspace
            if (ZP && (long long)cur.m < a.rows_cover) {
              uint4* op = reinterpret_cast<uint4*>(a.y + cur.row_off + c0);
#pragma unroll
              for (int i = 0; i < 4; ++i) if (i * 8 < cw) op[i] = make_uint4(0u, 0u, 0u, 0u);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              if (c * 16 >= cw) break;
              float f[16];
#pragma unroll
              for (int i4 = 0; i4 < 4; ++i4) {
                float4 sh;                                    // the batch-norm scale is folded into the bf16 weights (w_tc)
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sh.x), "=f"(sh.y), "=f"(sh.z), "=f"(sh.w)
                             : "r"(s_shift_u32 + (uint32_t)(c0 + c * 16 + i4 * 4) * 4));
                float4 pt = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (ptab)
                  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pt.x), "=f"(pt.y), "=f"(pt.z), "=f"(pt.w)
                               : "r"(ptab_u32 + (uint32_t)(c0 + c * 16 + i4 * 4) * 4));
                const float shv[4] = {sh.x, sh.y, sh.z, sh.w};
                const float ptv[4] = {pt.x, pt.y, pt.z, pt.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int i = i4 * 4 + j;
                  float acc = __uint_as_float(v[c * 16 + i]);
                  if (ptab) acc = fmaf(pl, ptv[j], acc);
                  f[i] = acc + shv[j];
                }
              }
              if (a.residual) {
                const uint32_t rw[8] = {res[2 * c].x, res[2 * c].y, res[2 * c].z, res[2 * c].w,
                                        res[2 * c + 1].x, res[2 * c + 1].y, res[2 * c + 1].z, res[2 * c + 1].w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  f[2 * i] += __uint_as_float(rw[i] << 16);
                  f[2 * i + 1] += __uint_as_float(rw[i] & 0xFFFF0000u);
                }
              }
              uint32_t o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float lo = f[2 * i], hi = f[2 * i + 1];
                if (a.relu) { lo = fmaxf(lo, 0.0f); hi = fmaxf(hi, 0.0f); }
                const __nv_bfloat162 pk = __floats2bfloat162_rn(lo, hi);
                o[i] = *reinterpret_cast<const uint32_t*>(&pk);
                if (PR > 0) { f[2 * i] = __uint_as_float(o[i] << 16); f[2 * i + 1] = __uint_as_float(o[i] & 0xFFFF0000u); }
              }
              if (wide_ls) {
                // one 32-byte store (STG.256) per 16 columns: a complete sector, half the store instructions
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(a.y + cur.row_off + c0 + c * 16),
                             "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
              } else {
                uint4* op = reinterpret_cast<uint4*>(a.y + cur.row_off + c0 + c * 16);
                op[0] = make_uint4(o[0], o[1], o[2], o[3]);
                op[1] = make_uint4(o[4], o[5], o[6], o[7]);
              }
              if (PR > 0) {
#pragma unroll
                for (int r = 0; r < PR; ++r) {
#pragma unroll
                  for (int i4 = 0; i4 < 4; ++i4) {
                    float4 ww;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(ww.x), "=f"(ww.y), "=f"(ww.z), "=f"(ww.w)
                                 : "r"(s_proj_u32 + (uint32_t)(r * N + c0 + c * 16 + i4 * 4) * 4));
                    pacc[r] = fmaf(f[4 * i4], ww.x, pacc[r]); pacc[r] = fmaf(f[4 * i4 + 1], ww.y, pacc[r]);
                    pacc[r] = fmaf(f[4 * i4 + 2], ww.z, pacc[r]); pacc[r] = fmaf(f[4 * i4 + 3], ww.w, pacc[r]);
                  }
                }
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) res[i] = resn[i];
        }
        if (PR > 0 && cur.valid) {
          float* po = a.proj_out + ((long long)cur.b * a.proj_r) * a.proj_hw + cur.pos;
#pragma unroll
          for (int r = 0; r < PR; ++r) {
            if (r < a.proj_r) {
              if (ncg == 1) po[(long long)r * a.proj_hw] = pacc[r];
              else atomicAdd(po + (long long)r * a.proj_hw, pacc[r]);     // two column groups: 0 + x + y, order-independent
            }
          }
        }
        item = next_item;
        have = item < n_items;
        cur = nxt;
      }
      // this warp is done reading the TMEM stage: release it to the MMA issuer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(acc_empty + s, 0);          // the leader's barrier counts both CTAs' epilogue warps
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(acc_empty + s)) : "memory");
      }
      if (a.debug && blockIdx.x == 0 && it < 32 && lane == 0 && q == 2) { long long* d = a.debug + ((1 + group) * 32 + it) * 4; d[0] = e0; d[1] = e1; d[2] = clock64(); d[3] = 0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) cluster_sync_all();                // no CTA of the pair leaves (or frees TMEM) while the other still works
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.ncols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.ncols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

bool make_map_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t row_bytes, uint32_t box_inner,
                 uint32_t box_outer, int kc) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {row_bytes};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// MZB_PDL=0 launches the convolutions without the programmatic-dependent-launch attribute (comparison knob)
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MZB_PDL"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v == 1;
}

int pick_kc(int cin) { return cin % 64 == 0 ? 64 : (cin % 32 == 0 ? 32 : 16); }

struct TcPlan { int kc, n_chunks, mt, tail_rows, b_resident, ncols; size_t smem; };

// ring depth of the streamed-weights path: 3 (divides the 18 k-blocks of a 128-channel layer: compile-time slots) or 4
int ring_depth(int kc, int n_chunks) { return (kc == 64 && n_chunks == 2) ? 3 : kStages; }

// Largest MT whose double-buffered activation stages, weights (resident if they fit, else a ring) and
// 2 x MT x C_out TMEM columns fit in one SM.
bool make_plan(int cin, int cout, int W, int plane_rows, TcPlan* out) {
  TcPlan p{};
  p.kc = pick_kc(cin);
  p.n_chunks = cin / p.kc;
  p.tail_rows = 2 * geo_halo(W) <= 32 ? 32 : 128;
  const size_t rowb = (size_t)p.kc * 2, limit = 225 * 1024;
  const size_t b_all = (size_t)9 * p.n_chunks * cout * rowb, b_ring = (size_t)ring_depth(p.kc, p.n_chunks) * cout * rowb;
  // barriers, scale/shift, <= 8 projection rows, the action plane's table (9 border classes or H*W positions)
  const size_t misc = 1024 + 512 + 8 * (size_t)cout + 32 * (size_t)cout + 4 * (size_t)cout * plane_rows;
  // narrow layers (C_out <= 32) are bound by per-super-tile latencies, not by the MMAs: give them more rows per step
  static int mt_narrow = -1;
  if (mt_narrow < 0) { const char* e = getenv("MZB_TC_MT_NARROW"); mt_narrow = e ? atoi(e) : 8; }
  for (int mt = cout <= 32 ? mt_narrow : 4; mt >= 1; mt >>= 1) {
    if (2 * mt * cout > 512) continue;
    const size_t a2 = 2 * (size_t)p.n_chunks * (mt * 128 + p.tail_rows) * rowb;
    if (a2 + b_all + misc <= limit) { p.mt = mt; p.b_resident = 1; p.smem = a2 + b_all + misc; break; }
    if (a2 + b_ring + misc <= limit) { p.mt = mt; p.b_resident = 0; p.smem = a2 + b_ring + misc; break; }
  }
  if (!p.mt) return false;
  p.ncols = 32;
  while (p.ncols < 2 * p.mt * cout) p.ncols <<= 1;
  if (out) *out = p;
  return true;
}

// Which layers run as CTA pairs.  The pair form is bit-identical to the single-CTA form; in step (connect4, 16,384 boards)
// it takes 60-62 us against 64-67 us for a PLAIN layer (the MMA stream is bound by its operand fetch, and a pair fetches
// 128 + 32 instead of 128 + 64 rows per SM and instruction), but 88 against 81 us with a residual input, 80 against 70 with
// the action plane and 96 / 111 against 90 / 107 with a head projection: those layers are paced by their epilogue, and a
// pair moves at the pace of the slower of its two epilogues.  Mode 1 (default) = plain layers only, 2 = every eligible
// layer, 0 = never (MZB_TC_PAIR; mzb_conv_tc_pair_enable() overrides: the tests compare the forms in one process).
int g_tc_pair = -1;                    // -1: follow MZB_TC_PAIR
int pair_mode() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MZB_TC_PAIR"); v = e ? atoi(e) : 1; }
  return g_tc_pair >= 0 ? g_tc_pair : v;
}

// cta_group::2 form: clusters of two CTAs (64 -> 64 channel layers with resident weights)
template <int PRV>
int launch_pair(unsigned grid, size_t smem, cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmAtail,
                const CUtensorMap& tmB, const TcArgs& a) {
  static bool configured = false;
  if (!configured) {
    MZB_CUDA(cudaFuncSetAttribute(k_conv_tc<64, false, PRV, kEpiWarpsWide, true, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(grid); lc.blockDim = dim3(64 + kEpiWarpsWide * 32); lc.dynamicSmemBytes = smem; lc.stream = stream;
  cudaLaunchAttribute la[2];
  la[0].id = cudaLaunchAttributeClusterDimension;
  la[0].val.clusterDim.x = 2; la[0].val.clusterDim.y = 1; la[0].val.clusterDim.z = 1;
  la[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  la[1].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = la; lc.numAttrs = pdl_enabled() ? 2 : 1;
  MZB_CUDA(cudaLaunchKernelEx(&lc, k_conv_tc<64, false, PRV, kEpiWarpsWide, true, 0, true>, tmA, tmAtail, tmB, a));
  return MZB_OK;
}

}  // namespace

static bool g_tc_enabled = true;
static long long* g_tc_debug = nullptr;
extern "C" void mzb_conv_tc_debug_buffer(long long* d_buf) { g_tc_debug = d_buf; }   // bring-up: 4*32*4 int64
extern "C" void mzb_conv_tc_enable(int on) { g_tc_enabled = on != 0; }
extern "C" void mzb_conv_tc_pair_enable(int mode) { g_tc_pair = mode; }
bool mzb_conv_tc_enabled() { return g_tc_enabled && encode_fn() != nullptr; }

bool mzb_conv_tc_supported(const ConvParams& cp, int H, int W, int cin_stride) {
  return g_tc_enabled && cp.stride == 1 && cp.cin == cin_stride && cp.cin % 16 == 0 && cp.cout % 16 == 0 && cp.cout >= 16 &&
         cp.cout <= 256 && W <= 61 && cp.w_tc != nullptr && encode_fn() != nullptr && make_plan(cp.cin, cp.cout, W, cp.extra_plane ? ((H >= 3 && W >= 3) ? 9 : H * W) : 0, nullptr);
}

int mzb_conv_tc_launch(int B, int H, int W, const ConvParams& cp, const __nv_bfloat16* x, const float* plane,
                       const __nv_bfloat16* residual, int relu, __nv_bfloat16* y, cudaStream_t stream, int zero_pads,
                       const TcProj* proj) {
  TcPlan p;
  MZB_CHECK_ARG(make_plan(cp.cin, cp.cout, W, cp.extra_plane ? ((H >= 3 && W >= 3) ? 9 : H * W) : 0, &p), "tensor-core convolution: no tile configuration fits");
  const Geo g{H, W, cp.cin, 1};
  const long long rows_total = geo_rows_total(g, B);
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  const long long total_tiles_h = (geo_rows_per_image(H, W) * (long long)B + (zero_pads ? geo_halo(W) : 0) + 127) / 128;
  // the pair form: 64 -> 64 channels, one 64-channel chunk, resident weights, enough tiles for every CTA of an even grid
  const int pmode = pair_mode();
  const bool plain_layer = residual == nullptr && !cp.extra_plane && !(proj && proj->r > 0);
  const bool pair = (pmode == 2 || (pmode == 1 && plain_layer)) && p.kc == 64 && p.n_chunks == 1 && cp.cout == 64 && p.b_resident && !zero_pads &&
                    n_sm % 2 == 0 && total_tiles_h >= 2ll * n_sm;
  CUtensorMap tmA, tmAtail, tmB;
  if (!make_map_2d(&tmA, x, (uint64_t)cp.cin, (uint64_t)rows_total, (uint64_t)cp.cin * 2, (uint32_t)p.kc, 128, p.kc) ||
      !make_map_2d(&tmAtail, x, (uint64_t)cp.cin, (uint64_t)rows_total, (uint64_t)cp.cin * 2, (uint32_t)p.kc,
                   (uint32_t)p.tail_rows, p.kc) ||
      !make_map_2d(&tmB, cp.w_tc, (uint64_t)9 * cp.cin, (uint64_t)cp.cout, (uint64_t)9 * cp.cin * 2, (uint32_t)p.kc,
                   (uint32_t)(pair ? cp.cout / 2 : cp.cout), p.kc)) {
    mzb_set_error("cuTensorMapEncodeTiled failed (C_in=%d C_out=%d rows=%lld)", cp.cin, cp.cout, rows_total);
    return MZB_ECUDA;
  }
  TcArgs a{};
  a.B = B; a.H = H; a.W = W; a.Cin = cp.cin; a.Cout = cp.cout; a.R_img = geo_rows_per_image(H, W); a.mt = p.mt; a.n_chunks = p.n_chunks;
  a.relu = relu; a.ncols = p.ncols; a.tail_rows = p.tail_rows; a.b_resident = p.b_resident;
  a.rows_valid = (long long)B * a.R_img;
  a.zero_pads = zero_pads;
  a.rows_cover = a.rows_valid + (zero_pads ? geo_halo(W) : 0);
  MZB_CHECK_ARG(a.rows_cover + 4 * 128 < (1ll << 31), "tensor-core convolution: %lld rows exceed the 32-bit row index", a.rows_cover);
  a.scale = cp.scale; a.shift = cp.shift; a.plane = cp.extra_plane ? plane : nullptr; a.plane_table = cp.plane_table;
  a.residual = residual; a.y = y;
  a.plane_classes = (H >= 3 && W >= 3) ? 1 : 0;
  a.debug = g_tc_debug;
  if (proj && proj->r > 0) {
    MZB_CHECK_ARG(proj->r <= 8 && proj->w && proj->out, "fused head projection: at most 8 rows");
    a.proj_w = proj->w; a.proj_out = proj->out; a.proj_r = proj->r; a.proj_hw = H * W;
    if (cp.cout > 64) MZB_CUDA(cudaMemsetAsync(proj->out, 0, sizeof(float) * (size_t)B * proj->r * H * W, stream));
  }
  const long long total_tiles = (a.rows_cover + 127) / 128;
  const unsigned grid = (unsigned)(total_tiles < n_sm ? total_tiles : n_sm);   // persistent: one CTA per SM
  const int pr = (a.proj_r + 1) / 2 * 2;              // instantiated projection heights: 2, 4, 6, 8
  static int narrow_on = -1;
  if (narrow_on < 0) { const char* e = getenv("MZB_TC_NARROW_EPI"); narrow_on = (e && atoi(e) == 0) ? 0 : 1; }
  const bool narrow = narrow_on && cp.cout <= 32;
  MZB_CHECK_ARG(!(zero_pads && pr), "pad zeroing and head projection are not combined");
  if (pair) {
    int rc;
    switch (pr) {
      case 0: rc = launch_pair<0>(grid, p.smem, stream, tmA, tmAtail, tmB, a); break;
      case 2: rc = launch_pair<2>(grid, p.smem, stream, tmA, tmAtail, tmB, a); break;
      case 4: rc = launch_pair<4>(grid, p.smem, stream, tmA, tmAtail, tmB, a); break;
      case 6: rc = launch_pair<6>(grid, p.smem, stream, tmA, tmAtail, tmB, a); break;
      default: rc = launch_pair<8>(grid, p.smem, stream, tmA, tmAtail, tmB, a); break;
    }
    if (rc) return rc;
    MZB_LAUNCH_CHECK();
    return MZB_OK;
  }
#define LAUNCH_RES(KCV, ZPV, PRV, EWV, RESV, NCV)                                                                           \
  {                                                                                                                 \
    static bool configured = false;                                                                                 \
    if (!configured) {                                                                                              \
      MZB_CUDA(cudaFuncSetAttribute(k_conv_tc<KCV, ZPV, PRV, EWV, RESV, NCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      configured = true;                                                                                            \
    }                                                                                                               \
    cudaLaunchConfig_t lc = {};                                                                                     \
    lc.gridDim = dim3(grid); lc.blockDim = dim3(64 + EWV * 32); lc.dynamicSmemBytes = p.smem; lc.stream = stream;   \
    cudaLaunchAttribute la[1];                                                                                      \
    la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                  \
    la[0].val.programmaticStreamSerializationAllowed = 1;                                                           \
    lc.attrs = la; lc.numAttrs = pdl_enabled() ? 1 : 0;                                                             \
    MZB_CUDA(cudaLaunchKernelEx(&lc, k_conv_tc<KCV, ZPV, PRV, EWV, RESV, NCV>, tmA, tmAtail, tmB, a));                    \
  }
#define LAUNCH_EW(KCV, ZPV, PRV, EWV)                                                                               \
  {                                                                                                                 \
    if (p.b_resident) LAUNCH_RES(KCV, ZPV, PRV, EWV, true, 0)                                                       \
    else if (KCV == 64 && p.n_chunks == 2 && !ZPV) LAUNCH_RES(KCV, ZPV, PRV, EWV, false, 2)                         \
    else LAUNCH_RES(KCV, ZPV, PRV, EWV, false, 0)                                                                   \
  }
#define LAUNCH_ONE(KCV, ZPV, PRV)                                                                                   \
  {                                                                                                                 \
    if constexpr (PRV == 0) {                                                                                       \
      if (narrow) LAUNCH_EW(KCV, ZPV, PRV, kEpiWarpsNarrow)                                                         \
      else LAUNCH_EW(KCV, ZPV, PRV, kEpiWarpsWide)                                                                  \
    } else LAUNCH_EW(KCV, ZPV, PRV, kEpiWarpsWide)                                                                  \
  }
#define LAUNCH_KC(KCV)                                                                                              \
  {                                                                                                                 \
    if (zero_pads) LAUNCH_ONE(KCV, true, 0)                                                                         \
    else if (pr == 0) LAUNCH_ONE(KCV, false, 0)                                                                     \
    else if (pr == 2) LAUNCH_ONE(KCV, false, 2)                                                                     \
    else if (pr == 4) LAUNCH_ONE(KCV, false, 4)                                                                     \
    else if (pr == 6) LAUNCH_ONE(KCV, false, 6)                                                                     \
    else LAUNCH_ONE(KCV, false, 8)                                                                                  \
  }
  if (p.kc == 64) LAUNCH_KC(64) else if (p.kc == 32) LAUNCH_KC(32) else LAUNCH_KC(16)
#undef LAUNCH_ONE
#undef LAUNCH_EW
#undef LAUNCH_RES
#undef LAUNCH_KC
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}
