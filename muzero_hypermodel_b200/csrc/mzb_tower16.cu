// K7n: recurrent inference of a NARROW residual MuZero network (16 channels, latent <= 48 positions - Breakout's
// 16 x 6 x 6 hidden state) as ONE kernel: dynamics convolution with the action plane, the dynamics tower, the reward
// head's 1x1 projection, per-channel min-max scaling + hidden-state store, the prediction tower and the value / policy
// 1x1 projections (models.py:363-404, 447-456, 551-595, 612-619).  Only the head mlps stay separate launches.
//
// Why not the tcgen05 kernel per layer (mzb_conv_tc.cu): a 16-channel 3x3 convolution is a 128 x 16 x 144 GEMM tile whose
// MMA needs 8 cycles of math and 41-57 of operand fetch (tests/ubench_umma*.cu), every layer is a latency-bound launch
// (9 convolutions + 3 heads + min-max + gather = 14 launches of 20-40 us per simulation for 16,384 games) and the whole
// activation set of an image is 1.1 KB.  Here a WARP owns an image: its activations live in shared memory in the padded
// layout of mzb_resnet.cuh (so a tap is a row shift), all layers' weights are resident in shared memory (9 x 4.6 KB),
// and each convolution is 9 taps x ceil(HW / 8) n-tiles of mma.sync.m16n8k16 (bf16 operands, fp32 accumulate) with the roles
// of mzb_stem16.cu: A = a tap's weights [cout][cin] (one ldmatrix.x4 per tap, used by the image's 6 n-tiles); B = 8 positions x 16
// input channels, one ldmatrix.x4 feeds two n-tiles; D = [cout][position], stored through stmatrix.trans, the residual read
// through ldmatrix.trans.  An n-tile is 8 CONSECUTIVE PADDED ROWS starting at the first pixel line (6 n-tiles span the 42
// rows of a 6 x 6 image; the zero-column rows inside are computed and stored as zeros): with 48-byte rows any 8 consecutive
// rows are conflict-free, whereas 8 consecutive POSITIONS are 8 of 9 consecutive rows and rows r and r + 8 share a bank
// group - every ldmatrix of the position-mapped form took 2 wavefronts per 8 x 8 matrix (ncu: 2,400 shared-memory load
// wavefronts per image where 1,300 were expected).  The kernel is bound by shared-memory bandwidth, so wavefronts per
// layer and image are what counts: 108 (activations) + 36 (weights) + 24 (epilogue).
// The warp-level MMA is the right size for a 36 x 16 x 144 product; nothing but the input state, the output state and
// the projection rows touches global memory.  Arithmetic per layer is the bf16 path's: fp32 accumulate, folded
// batch-norm, residual, ReLU, one bf16 rounding of the stored activation.
#include <cuda_bf16.h>

#include "mzb_resnet_model.h"

namespace {

constexpr int kRowB = 48;              // bytes per shared-memory row: 16 bf16 + 16 B pad -> conflict-free ldmatrix / 4-byte stores
constexpr int kTapB = 16 * kRowB;      // one tap of one convolution: 16 output channels x 16 input channels
constexpr int kConvB = 9 * kTapB;
constexpr int kMaxBlocks = 4;
constexpr int kWarps = 20;
constexpr int NP = 3;                  // pairs of n-tiles: up to 48 positions
constexpr int kTabP = 72;              // row length of the transposed action-plane table: 72 / 2 = 4 (mod 16) -> the 8-byte
                                       // loads of a half-warp (4 channels x 4 position pairs) hit 16 distinct bank pairs

struct T16Conv { const __nv_bfloat16* w; const float* scale; const float* shift; };
struct T16Args {
  int B, H, W, A, n_dyn, n_pred, n_conv;
  T16Conv conv[1 + 4 * kMaxBlocks];            // dyn_conv, dyn blocks (c1, c2)..., pred blocks (c1, c2)...
  const float* plane_table;                    // [H*W][16]
  const float* w_r; int r_r;                   // reward 1x1 [r_r][16]
  const float* w_vp; int r_vp;                 // value | policy 1x1 [r_vp][16]
  const void* state_in; int in_layout; long long in_row_stride; const int* in_slot; long long slot_stride;
  const int* action;
  void* state_out; int out_layout; long long out_row_stride, out_off;
  float* proj_r; float* proj_vp;               // [B][r * H*W] fp32, bias added by the head kernels
  int* counters;                               // [0] next image, [1] warps that have left (both zero between launches)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  // not volatile: the compiler may issue a tap's loads ahead of the previous tap's MMAs; the memory clobber keeps them
  // behind the __syncwarp() that publishes the previous layer's stores
  asm("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void stsm_x4_t(uint32_t addr, const uint32_t (&r)[4]) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};"
               :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// Per-lane addressing of one pair of n-tiles (16 padded rows): matrix lane >> 3 of an ldmatrix / stmatrix .x4 is
// (n-tile (lane >> 4), channel half (lane >> 3) & 1), its row is lane & 7.
struct Lane {
  uint32_t b_off[NP];       // byte offset of the row this lane addresses; rows past the last pixel line: a row of the zero
                            // line (reads give zeros or discarded sums, the masked stores write zeros onto zeros)
  uint32_t ok;              // bit 2 nt + e: row 8 nt + 2 (lane & 3) + e of the span is a pixel (else stored as zero)
  uint32_t a_off;           // weights: row (cout) and channel half inside a tap
};

// One 3x3 convolution (16 -> 16 channels) over the warp's image: in -> out, both padded shared-memory buffers.
template <bool PLANE, bool RES>
__device__ __forceinline__ void conv16(uint32_t in_u32, uint32_t out_u32, uint32_t res_u32, uint32_t w_u32, const float* sh, float pl,
                                       const float* ptab, const Lane& L, int pitch, int lane, int HP) {
  float acc[2 * NP][4];
#pragma unroll
  for (int j = 0; j < 2 * NP; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const uint32_t shift = (uint32_t)(((tap / 3 - 1) * pitch + (tap % 3 - 1)) * kRowB);
    uint32_t af[4];                                    // a tap's weights are used by this warp's 6 n-tiles and not again: no reason to
    ldsm_x4(w_u32 + (uint32_t)tap * kTapB + L.a_off, af);   // hold all nine taps (36 registers) - that is what limits the warps per SM
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      if (16 * p >= HP) continue;                      // warp-uniform: HP = H * pitch rows from the first to the last pixel
      uint32_t bf[4];
      ldsm_x4(in_u32 + L.b_off[p] + shift, bf);
      mma16816(acc[2 * p], af, bf[0], bf[1]);
      if (16 * p + 8 < HP) mma16816(acc[2 * p + 1], af, bf[2], bf[3]);
    }
  }
  const int c_lo = lane >> 2, q2 = (lane & 3) * 2;
  const float sh_lo = sh[c_lo], sh_hi = sh[c_lo + 8];  // the batch-norm scale is folded into the bf16 weights (ConvParams::w_tc), as in k_conv_tc
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    if (16 * p >= HP) continue;
    uint32_t rs[4] = {0u, 0u, 0u, 0u}, pk[4];
    if (RES) ldsm_x4_t(res_u32 + L.b_off[p], rs);
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int nt = 2 * p + jj;
      const bool k0 = (L.ok >> (2 * nt)) & 1u, k1 = (L.ok >> (2 * nt + 1)) & 1u;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v0 = acc[nt][2 * h], v1 = acc[nt][2 * h + 1];
        if (PLANE) {
          const float2 t = *reinterpret_cast<const float2*>(ptab + (c_lo + 8 * h) * kTabP + 8 * nt + q2);
          v0 = fmaf(pl, t.x, v0); v1 = fmaf(pl, t.y, v1);
        }
        v0 += h ? sh_hi : sh_lo; v1 += h ? sh_hi : sh_lo;
        if (RES) { v0 += bf_lo(rs[2 * jj + h]); v1 += bf_hi(rs[2 * jj + h]); }
        v0 = k0 ? fmaxf(v0, 0.0f) : 0.0f; v1 = k1 ? fmaxf(v1, 0.0f) : 0.0f;
        const __nv_bfloat162 pr = __floats2bfloat162_rn(v0, v1);
        pk[2 * jj + h] = *reinterpret_cast<const uint32_t*>(&pr);
      }
    }
    stsm_x4_t(out_u32 + L.b_off[p], pk);
  }
  __syncwarp();
}

// 1x1 projection of the warp's image onto r <= 16 rows, out[k][pos] = sum_c x[pos][c] * w[k][c], as one more MMA tap: the
// fp32 weights are split into two bf16 terms (w = hi + lo exactly to 2^-17 relative), the activations are bf16 already, so
// two MMAs per n-tile give the fp32 FMA chain's result to rounding level for 60 instructions instead of the 450 of a
// lane-per-position loop (which was a fifth of the kernel's issue slots and 9 % of its shared-memory wavefronts).
// s_pw: [hi | lo][16][kRowB] bf16 rows; s_pos: [48] position of padded row `rel` of the span, -1 for zero-column rows.
__device__ __forceinline__ void project(uint32_t buf_u32, uint32_t pw_u32, int r, float* out, int HW, const int* s_pos, const Lane& L,
                                        int lane, int HP) {
  uint32_t ah[4], al[4];
  ldsm_x4(pw_u32 + L.a_off, ah);
  ldsm_x4(pw_u32 + kTapB + L.a_off, al);
  const int c_lo = lane >> 2, q2 = (lane & 3) * 2;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    if (16 * p >= HP) continue;
    uint32_t bf[4];
    ldsm_x4(buf_u32 + L.b_off[p], bf);
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      if (16 * p + 8 * jj >= HP) continue;
      float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      mma16816(acc, al, bf[2 * jj], bf[2 * jj + 1]);
      mma16816(acc, ah, bf[2 * jj], bf[2 * jj + 1]);
      const int2 pos = *reinterpret_cast<const int2*>(s_pos + 16 * p + 8 * jj + q2);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = c_lo + 8 * h;
        if (k < r) {
          if (pos.x >= 0) out[(long long)k * HW + pos.x] = acc[2 * h];
          if (pos.y >= 0) out[(long long)k * HW + pos.y] = acc[2 * h + 1];
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kWarps * 32, 1) k_recurrent16(const T16Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = a.H, W = a.W, HW = H * W, pitch = geo_pitch(W), halo = geo_halo(W);
  const int rows = halo + geo_rows_per_image(H, W) + halo;
  uint8_t* s_w = smem;                                              // [n_conv][9][16][kRowB]
  float* s_sc = reinterpret_cast<float*>(s_w + (size_t)a.n_conv * kConvB);       // [n_conv][32]: scale | shift
  float* s_ptab = s_sc + a.n_conv * 32;                             // [16][kTabP]: action-plane table, channel-major
  int* s_row = reinterpret_cast<int*>(s_ptab + 16 * kTabP);           // [HW]: padded row of position p (no divisions in the loops)
  int* s_pos = s_row + ((HW + 1) & ~1);                               // [16 NP]: position of row `rel` of the span, -1 for a pad row
  uint8_t* s_pw = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_pos + 16 * NP) + 15) & ~(uintptr_t)15);
                                                                    // [reward | value+policy][hi | lo][16][kRowB]: projection weights
  uint8_t* s_act = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_pw + 4 * kTapB) + 127) & ~(uintptr_t)127);
  const int HP = H * pitch;                                           // padded rows from the first pixel to the end of the last line
  const size_t buf_bytes = (size_t)rows * kRowB;
  // ---- stage the weights: w_tc (scale folded in) [16 cout][9 taps][16 cin] -> [tap][cout] rows of kRowB bytes
  for (int i = threadIdx.x; i < a.n_conv * 9 * 16 * 2; i += blockDim.x) {
    const int half = i & 1, n = (i >> 1) & 15, tap = (i >> 5) % 9, ci = i / (9 * 32);
    const uint4 v = *reinterpret_cast<const uint4*>(a.conv[ci].w + (size_t)n * 144 + tap * 16 + half * 8);
    *reinterpret_cast<uint4*>(s_w + (size_t)ci * kConvB + tap * kTapB + n * kRowB + half * 16) = v;
  }
  for (int i = threadIdx.x; i < a.n_conv * 32; i += blockDim.x) {
    const int ci = i >> 5, k = i & 31;
    s_sc[i] = k < 16 ? a.conv[ci].scale[k] : a.conv[ci].shift[k - 16];
  }
  for (int i = threadIdx.x; i < 16 * kTabP; i += blockDim.x) {       // conv[0] = the dynamics convolution
    const int c = i / kTabP, rel = i - c * kTabP, y = rel / pitch, x = rel - y * pitch;       // indexed by the row of the span
    s_ptab[i] = (rel < HP && x < W) ? a.plane_table[(y * W + x) * 16 + c] * a.conv[0].scale[c] : 0.0f;
  }
  for (int i = threadIdx.x; i < 2 * 16 * 16; i += blockDim.x) {       // w = hi + lo, two bf16 terms; rows >= r are zero
    const int which = i >> 8, k = (i >> 4) & 15, c = i & 15;
    const int r = which ? a.r_vp : a.r_r;
    const float w = k < r ? (which ? a.w_vp : a.w_r)[k * 16 + c] : 0.0f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(w), lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    uint8_t* base = s_pw + (size_t)which * 2 * kTapB + k * kRowB + c * 2;
    *reinterpret_cast<__nv_bfloat16*>(base) = hi;
    *reinterpret_cast<__nv_bfloat16*>(base + kTapB) = lo;
  }
  for (int p = threadIdx.x; p < HW; p += blockDim.x) s_row[p] = halo + (p / W + 1) * pitch + p % W;
  for (int i = threadIdx.x; i < 16 * NP; i += blockDim.x) {
    const int y = i / pitch, x = i - y * pitch;
    s_pos[i] = (i < H * pitch && x < W) ? y * W + x : -1;
  }
  for (size_t i = threadIdx.x; i < kWarps * 2 * buf_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(s_act)[i] = make_uint4(0u, 0u, 0u, 0u);       // pad rows stay zero for the whole kernel
  __syncthreads();

  // the warp's two buffers: buf(i) by arithmetic (an indexed pointer array would live in local memory)
  uint8_t* const buf0 = s_act + (size_t)(warp * 2) * buf_bytes;
  auto buf = [&](int i) -> uint8_t* { return buf0 + (size_t)i * buf_bytes; };
  Lane L;
  {
    L.ok = 0u;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int rel = 16 * p + 8 * (lane >> 4) + (lane & 7);
      L.b_off[p] = (uint32_t)((rel < HP ? halo + pitch + rel : halo) * kRowB + ((lane >> 3) & 1) * 16);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r2 = 16 * p + 8 * (e >> 1) + (lane & 3) * 2 + (e & 1);
        if (r2 < HP && r2 % pitch < W) L.ok |= 1u << (4 * p + e);
      }
    }
    L.a_off = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * kRowB + (lane >> 4) * 16);
  }
  const uint32_t w_u32 = smem_u32(s_w);
  uint64_t pol_stream;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));

  // Images are handed out by a counter, not by a fixed stride: 16,384 images over 148 x 20 warps are 5.5 rounds, and with a fixed
  // assignment the launch lasts 6.  Which warp computes an image does not change its result.
  while (true) {
    int b = 0;
    if (lane == 0) b = atomicAdd(a.counters, 1);
    b = __shfl_sync(0xFFFFFFFFu, b, 0);
    if (b >= a.B) break;
    // ---- hidden state of the parent node -> buffer 0
    const long long in_off = (long long)b * a.in_row_stride + (a.in_slot ? (long long)a.in_slot[b] * a.slot_stride : 0);
    if (a.in_layout == 2) {
      // the hidden-state pool streams through L2 (read once per child, 38 MB per simulation in and out): evict_first keeps it
      // from displacing the tree records, whose dependent loads pace k_select / k_expand_backup (L2 hit rate 26 % / 36 %);
      // measured: k_expand_backup 6.2 -> 5.2 us, k_select unchanged (19 us: one partial wave whose length is the deepest walk)
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.state_in) + in_off);
      for (int i = lane; i < 2 * HW; i += 32) {
        const int p = i >> 1;
        uint4 v;
        asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + i), "l"(pol_stream));
        *reinterpret_cast<uint4*>(buf(0) + s_row[p] * kRowB + (i & 1) * 16) = v;
      }
    } else {                                                        // fp32 NCHW rows (per-row API)
      const float* src = reinterpret_cast<const float*>(a.state_in) + in_off;
      for (int i = lane; i < 16 * HW; i += 32) {
        const int c = i / HW, p = i - c * HW;
        *reinterpret_cast<__nv_bfloat16*>(buf(0) + s_row[p] * kRowB + c * 2) = __float2bfloat16_rn(src[i]);
      }
    }
    const float pl = __fdiv_rn((float)a.action[b], (float)a.A);     // action * ones / action_space_size (:553-568)
    __syncwarp();
    // ---- dynamics: conv + action plane, residual tower (models.py:377-387)
    int ci = 0;
    conv16<true, false>(smem_u32(buf(0)), smem_u32(buf(1)), 0u, w_u32, s_sc + 16, pl, s_ptab, L, pitch, lane, HP);
    ci = 1;
    int cur = 1;
    for (int k = 0; k < a.n_dyn; ++k, ci += 2) {
      const int t = 1 - cur, o = cur;                             // the block's second convolution writes over its input
      conv16<false, false>(smem_u32(buf(cur)), smem_u32(buf(t)), 0u, w_u32 + (uint32_t)ci * kConvB, s_sc + ci * 32 + 16, 0.0f,
                           nullptr, L, pitch, lane, HP);
      conv16<false, true>(smem_u32(buf(t)), smem_u32(buf(o)), smem_u32(buf(cur)), w_u32 + (uint32_t)(ci + 1) * kConvB,
                          s_sc + (ci + 1) * 32 + 16, 0.0f, nullptr, L, pitch, lane, HP);
      cur = o;
    }
    // ---- reward head projection on the UN-normalised next state (:388-391)
    if (a.proj_r) project(smem_u32(buf(cur)), smem_u32(s_pw), a.r_r, a.proj_r + (long long)b * a.r_r * HW, HW, s_pos, L, lane, HP);
    // ---- per-channel min-max scaling (:571-586) -> next buffer + the caller's hidden-state slot
    const int nx = cur;                                             // in place: a lane re-writes the elements it read
    {
      const int cp = lane & 7, pg = lane >> 3;                      // channel pair, position group (p = 4 i + pg)
      float lo0 = CUDART_INF_F, hi0 = -CUDART_INF_F, lo1 = CUDART_INF_F, hi1 = -CUDART_INF_F;
      for (int p = pg; p < HW; p += 4) {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(buf(cur) + s_row[p] * kRowB + cp * 4);
        lo0 = fminf(lo0, bf_lo(v)); hi0 = fmaxf(hi0, bf_lo(v)); lo1 = fminf(lo1, bf_hi(v)); hi1 = fmaxf(hi1, bf_hi(v));
      }
#pragma unroll
      for (int o = 8; o < 32; o <<= 1) {
        lo0 = fminf(lo0, __shfl_xor_sync(0xFFFFFFFFu, lo0, o)); hi0 = fmaxf(hi0, __shfl_xor_sync(0xFFFFFFFFu, hi0, o));
        lo1 = fminf(lo1, __shfl_xor_sync(0xFFFFFFFFu, lo1, o)); hi1 = fmaxf(hi1, __shfl_xor_sync(0xFFFFFFFFu, hi1, o));
      }
      float s0 = __fsub_rn(hi0, lo0), s1 = __fsub_rn(hi1, lo1);
      if (s0 < 1e-5f) s0 = __fadd_rn(s0, 1e-5f);
      if (s1 < 1e-5f) s1 = __fadd_rn(s1, 1e-5f);
      const float i0 = 1.0f / s0, i1 = 1.0f / s1;                   // one reciprocal per channel (as k_minmax_store_bf16)
      const long long out_base = (long long)b * a.out_row_stride + a.out_off;
      for (int p = pg; p < HW; p += 4) {
        const int row = s_row[p];
        const uint32_t v = *reinterpret_cast<const uint32_t*>(buf(cur) + row * kRowB + cp * 4);
        const __nv_bfloat162 pk = __floats2bfloat162_rn((bf_lo(v) - lo0) * i0, (bf_hi(v) - lo1) * i1);
        const uint32_t pw = *reinterpret_cast<const uint32_t*>(&pk);
        *reinterpret_cast<uint32_t*>(buf(nx) + row * kRowB + cp * 4) = pw;
        if (a.state_out) {
          if (a.out_layout == 2) {
            asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(reinterpret_cast<__nv_bfloat16*>(a.state_out) + out_base + (long long)p * 16 + cp * 2),
                         "r"(pw), "l"(pol_stream) : "memory");
          } else {                                                  // fp32 NCHW
            float* o = reinterpret_cast<float*>(a.state_out) + out_base;
            o[(long long)(cp * 2) * HW + p] = bf_lo(pw);
            o[(long long)(cp * 2 + 1) * HW + p] = bf_hi(pw);
          }
        }
      }
      __syncwarp();
    }
    // ---- prediction tower + value / policy projections (:447-456)
    cur = nx;
    if (a.proj_vp) {
      for (int k = 0; k < a.n_pred; ++k, ci += 2) {
        const int t = 1 - cur, o = cur;                             // the block's second convolution writes over its input
        conv16<false, false>(smem_u32(buf(cur)), smem_u32(buf(t)), 0u, w_u32 + (uint32_t)ci * kConvB, s_sc + ci * 32 + 16,
                             0.0f, nullptr, L, pitch, lane, HP);
        conv16<false, true>(smem_u32(buf(t)), smem_u32(buf(o)), smem_u32(buf(cur)), w_u32 + (uint32_t)(ci + 1) * kConvB,
                            s_sc + (ci + 1) * 32 + 16, 0.0f, nullptr, L, pitch, lane, HP);
        cur = o;
      }
      project(smem_u32(buf(cur)), smem_u32(s_pw) + 2 * kTapB, a.r_vp, a.proj_vp + (long long)b * a.r_vp * HW, HW, s_pos, L, lane, HP);
    }
    __syncwarp();
  }
  // a warp leaves after its failed fetch; the last one to leave (every fetch of the launch is behind it) zeroes both counters for
  // the next launch - no memset node per simulation in the captured search graph
  if (lane == 0) {
    __threadfence();
    if (atomicAdd(a.counters + 1, 1) == (int)(gridDim.x * kWarps) - 1) {
      a.counters[0] = 0;
      a.counters[1] = 0;
      __threadfence();
    }
  }
}

size_t tower16_smem(const mzb_resnet_model* m) {
  const int n_conv = 1 + 2 * (int)m->dyn_blocks.size() + 2 * (int)m->pred_blocks.size();
  const int HW = m->Hl * m->Wl, rows = 2 * geo_halo(m->Wl) + geo_rows_per_image(m->Hl, m->Wl);
  return 128 + (size_t)n_conv * kConvB + sizeof(float) * ((size_t)n_conv * 32 + (size_t)16 * kTabP + HW + 1 + 16 * NP) + 16 + 4 * kTapB +
         128 + (size_t)kWarps * 2 * rows * kRowB;
}

}  // namespace

static bool g_tower16_enabled = true;
extern "C" void mzb_tower16_enable(int on) { g_tower16_enabled = on != 0; }     // comparison / bring-up knob

bool mzb_tower16_supported(const mzb_resnet_model* m, int in_layout, int out_layout) {
  if (!g_tower16_enabled || !mzb_conv_tc_enabled() || m->precision != 1 || m->C != 16 || m->Hl * (m->Wl + 1) > 16 * NP) return false;
  if (m->dyn_blocks.size() > kMaxBlocks || m->pred_blocks.size() > kMaxBlocks) return false;
  if ((in_layout != 0 && in_layout != 2) || (out_layout != 0 && out_layout != 2)) return false;
  if (!m->dyn_conv.w_bf16 || !m->dyn_conv.plane_table || m->dyn_conv.cin != 16 || !m->pv_w) return false;
  if (m->reward.r > 16 || m->value.r + m->policy.r > 16) return false;      // the projections are one 16-row MMA tap each
  for (const auto* blocks : {&m->dyn_blocks, &m->pred_blocks})
    for (const Block& b : *blocks)
      if (!b.c1.w_bf16 || !b.c2.w_bf16 || b.c1.cin != 16 || b.c2.cin != 16) return false;
  return tower16_smem(m) <= 227 * 1024;
}

int mzb_tower16_recurrent(mzb_resnet_model* m, int B, const void* state_in, int in_layout, long long in_row_stride,
                          const int* in_slot, long long slot_stride, const int* action, void* state_out, int out_layout,
                          long long out_row_stride, long long out_off, float* proj_r, float* proj_vp, cudaStream_t s) {
  T16Args a{};
  a.B = B; a.H = m->Hl; a.W = m->Wl; a.A = m->A;
  a.n_dyn = (int)m->dyn_blocks.size(); a.n_pred = (int)m->pred_blocks.size();
  int k = 0;
  auto put = [&](const ConvParams& c) { a.conv[k++] = T16Conv{c.w_tc, c.scale, c.shift}; };
  put(m->dyn_conv);
  for (const Block& b : m->dyn_blocks) { put(b.c1); put(b.c2); }
  for (const Block& b : m->pred_blocks) { put(b.c1); put(b.c2); }
  a.n_conv = k;
  a.plane_table = m->dyn_conv.plane_table;
  a.w_r = m->reward.w1x1; a.r_r = m->reward.r;
  a.w_vp = m->pv_w; a.r_vp = m->value.r + m->policy.r;
  a.state_in = state_in; a.in_layout = in_layout; a.in_row_stride = in_row_stride; a.in_slot = in_slot; a.slot_stride = slot_stride;
  a.action = action;
  a.state_out = state_out; a.out_layout = out_layout; a.out_row_stride = out_row_stride; a.out_off = out_off;
  a.proj_r = proj_r; a.proj_vp = proj_vp;
  a.counters = m->t16_counters;
  const size_t smem = tower16_smem(m);
  static bool configured = false;
  if (!configured) {
    MZB_CUDA(cudaFuncSetAttribute(k_recurrent16, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  int grid = (B + kWarps - 1) / kWarps;
  if (grid > n_sm) grid = n_sm;                      // persistent: the weights are staged once per CTA
  k_recurrent16<<<grid, kWarps * 32, smem, s>>>(a);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}
