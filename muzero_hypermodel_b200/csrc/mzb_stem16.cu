// K7s: a residual tower over 16-channel rows with the IMAGE RESIDENT IN SHARED MEMORY - the three resolutions of
// Breakout's DownSample stem (models.py:226-275: resblocks1 at 48 x 48 x 8 in the pixel-pair form = 48 x 24 rows of 16
// channels, resblocks2 at 24 x 24 x 16, resblocks3 at 12 x 12 x 16).
//
// Layer by layer on k_conv_tc these 16 convolutions were 16 launches of ~ 263 us for 16,384 frames: each reads and
// writes a 300-600 MB activation tensor (HBM-bound at twice the traffic a fused block needs) through an MMA whose N = 16
// tile is 8 cycles of math behind 41-57 cycles of operand fetch.  Here a group of warps owns an image: it is copied into
// shared memory once (cp.async, padded layout of mzb_resnet.cuh with 48-byte rows), ALL residual blocks of the stage
// run on it there, and it is written back once, in place - 2 x the tensor per stage instead of 6 x per block.
//
// The convolution is an implicit GEMM on mma.sync.m16n8k16 with the roles chosen for shared-memory traffic:
//   A (16 x 16, row)  = one tap's weights [cout][cin] - the nine A fragments of a layer live in REGISTERS (36 per lane),
//                       loaded once per layer and warp;
//   B (16 x 8,  col)  = 8 consecutive padded rows x 16 input channels, shifted by the tap offset: one ldmatrix.x4 feeds two
//                       MMAs (16 rows);
//   D (16 x 8)        = [cout][row]: the epilogue adds the folded batch-norm shift (the scale is inside ConvParams::w_tc),
//                       the residual (ldmatrix.trans of the block's input, same addresses) and stores 16 rows x 16
//                       channels with ONE stmatrix.x4.trans.
// So a 16-row tile costs 9 ldmatrix + 18 MMAs (+ 2 for the epilogue) instead of the 12 + 18 per tile of the [row][cout]
// form.  Pad rows (the zero line / zero column of the layout) are computed like pixels and stored as zeros (byte mask).
// Arithmetic is the bf16 path's: bf16 operands, fp32 accumulate, one bf16 rounding of each stored activation.
#include <cuda_bf16.h>

#include <cstdlib>

#include "mzb_resnet_model.h"

namespace {

constexpr int kRowB = 48;              // bytes per shared-memory row: 16 bf16 + 16 B pad -> conflict-free ldmatrix / stmatrix
constexpr int kTapB = 16 * kRowB;
constexpr int kConvB = 9 * kTapB;
constexpr int kWarps = 16;
constexpr int kMaxConv = 8;

struct S16Args {
  int B, H, W, n_blocks;
  int R, T;                              // padded rows per image, 16-row tiles per image
  int gw, G, nbuf;                       // warps per image group, groups per CTA, buffers per group (3 = next image prefetched)
  int buf_bytes;
  const __nv_bfloat16* w[kMaxConv];      // w_tc [16 cout][9][16 cin]
  const float* shift[kMaxConv];
  __nv_bfloat16* x;                      // padded activations [halo + B * R + halo][16], updated in place
  // optional tail (pixel-pair stage only): DownSample.conv2, stride 2, on the resident image; then x is NOT written back
  const __nv_bfloat16* w_s2;             // [6][16][16] (pack_s2_mma) or NULL
  const float* shift_s2;
  __nv_bfloat16* y2;                     // padded output [halo + B * R2 + halo][16] of H/2 x W pixels
  int R2, T2;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void stsm_x4_t(uint32_t addr, const uint32_t (&r)[4]) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};"
               :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma1688(float (&d)[4], const uint32_t (&a)[2], uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(b0));
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void group_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory");
}

// One 3x3 convolution of the group's image: `in` -> `out` (shared-memory byte addresses of row 0 of the image region).
// RES: out also holds the block's input, which is added before the ReLU (read and written by the same lanes).
// PAIR: the pixel-pair stage (pack_conv_pair).  Its left / right taps are three-quarters zero: the left pair row feeds only its
// SECOND pixel (input channels 8-15) to the outputs' FIRST pixel, the right pair row only its first pixel (channels 0-7) to the
// outputs' second pixel - so those six taps are K = 8 MMAs on one 16-byte half of the rows (ldmatrix.x2: 2 wavefronts instead
// of 4 per tile and tap; 24 instead of 36 operand wavefronts per tile and layer).
template <bool RES, bool PAIR>
__device__ __forceinline__ void conv_image(uint32_t in, uint32_t out, uint32_t w_u32, const float* sh, const uint8_t* ok, int T,
                                           int pitch, int gwarp, int gw, int lane) {
  uint32_t af[9][4];                     // PAIR: taps 0 / 2 of a kernel line use af[tap][0..1] only
  {
    const uint32_t a_row = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * kRowB);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      if (PAIR && tap % 3 != 1) {
        uint32_t h2[2];
        ldsm_x2(w_u32 + (uint32_t)tap * kTapB + a_row + (tap % 3 == 0 ? 16u : 0u), h2);
        af[tap][0] = h2[0]; af[tap][1] = h2[1]; af[tap][2] = 0u; af[tap][3] = 0u;
      } else {
        ldsm_x4(w_u32 + (uint32_t)tap * kTapB + a_row + (uint32_t)((lane >> 4) * 16), af[tap]);
      }
    }
  }
  // rows of an ldmatrix.x2 (lanes 0-15; the other lanes pass valid, unused addresses): n-tile (lane >> 3) & 1, row lane & 7
  const uint32_t l2_off = (uint32_t)((8 * ((lane >> 3) & 1) + (lane & 7)) * kRowB);
  const float sh_lo = sh[lane >> 2], sh_hi = sh[(lane >> 2) + 8];
  const uint32_t l_off = (uint32_t)((8 * (lane >> 4) + (lane & 7)) * kRowB + ((lane >> 3) & 1) * 16);
  const int q2 = (lane & 3) * 2;
  for (int t = gwarp; t < T; t += gw) {
    const uint32_t tile = (uint32_t)(t * 16 * kRowB) + l_off;
    float acc[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int shift = (tap / 3 - 1) * pitch + (tap % 3 - 1);
      if (PAIR && tap % 3 != 1) {
        uint32_t b2[2];
        const uint32_t a2[2] = {af[tap][0], af[tap][1]};
        ldsm_x2(in + (uint32_t)(t * 16 * kRowB) + l2_off + (uint32_t)(shift * kRowB) + (tap % 3 == 0 ? 16u : 0u), b2);
        mma1688(acc[0], a2, b2[0]);
        mma1688(acc[1], a2, b2[1]);
      } else {
        uint32_t bf[4];
        ldsm_x4(in + tile + (uint32_t)(shift * kRowB), bf);
        mma16816(acc[0], af[tap], bf[0], bf[1]);
        mma16816(acc[1], af[tap], bf[2], bf[3]);
      }
    }
    uint32_t rs[4] = {0u, 0u, 0u, 0u};
    if (RES) ldsm_x4_t(out + tile, rs);
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint16_t m2 = *reinterpret_cast<const uint16_t*>(ok + t * 16 + 8 * j + q2);     // pad rows are stored as zeros
      const bool k0 = (m2 & 0xFF) != 0, k1 = (m2 >> 8) != 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v0 = acc[j][2 * h] + (h ? sh_hi : sh_lo), v1 = acc[j][2 * h + 1] + (h ? sh_hi : sh_lo);
        if (RES) { v0 += bf_lo(rs[2 * j + h]); v1 += bf_hi(rs[2 * j + h]); }
        v0 = k0 ? fmaxf(v0, 0.0f) : 0.0f; v1 = k1 ? fmaxf(v1, 0.0f) : 0.0f;
        const __nv_bfloat162 p = __floats2bfloat162_rn(v0, v1);
        pk[2 * j + h] = *reinterpret_cast<const uint32_t*>(&p);
      }
    }
    stsm_x4_t(out + tile, pk);
  }
}

// DownSample.conv2 (3x3, stride 2, padding 1, 8 -> 16 channels) from the pixel-pair image `in` to the H/2 x W image `out`
// (same pitch and halo): output row r2 = (yy, xx) reads, for ky = 0..2, the pair rows (2 (yy - 1) + ky, xx - 1) and
// (.., xx) - a per-lane gather, which ldmatrix takes as it is.  No batch-norm, no ReLU (models.py:236-255).
__device__ __forceinline__ void conv_s2_image(uint32_t in, uint32_t out, uint32_t w_u32, const float* sh, const uint8_t* ok, int T2, int R2,
                                              int W, int pitch, int halo, int gwarp, int gw, int lane) {
  uint32_t af[6][4];
  {
    const uint32_t a_off = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * kRowB + (lane >> 4) * 16);
#pragma unroll
    for (int tap = 0; tap < 6; ++tap) ldsm_x4(w_u32 + (uint32_t)tap * kTapB + a_off, af[tap]);
  }
  const float sh_lo = sh[lane >> 2], sh_hi = sh[(lane >> 2) + 8];
  const uint32_t half = (uint32_t)(((lane >> 3) & 1) * 16);
  const uint32_t l_off = (uint32_t)((8 * (lane >> 4) + (lane & 7)) * kRowB) + half;
  const int q2 = (lane & 3) * 2;
  for (int t = gwarp; t < T2; t += gw) {
    const int r2 = 16 * t + 8 * (lane >> 4) + (lane & 7), yy = r2 / pitch, xx = r2 - yy * pitch;
    // rows that are not pixels gather from the zero halo (their sums are discarded by the mask)
    const int src = (r2 < R2 && geo_is_pixel(yy, xx, W)) ? 2 * (yy - 1) * pitch + xx - 1 : -halo;
    const uint32_t base = in + (uint32_t)(src * kRowB) + half;
    float acc[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;
#pragma unroll
    for (int tap = 0; tap < 6; ++tap) {
      uint32_t bf[4];
      ldsm_x4(base + (uint32_t)(((tap >> 1) * pitch + (tap & 1)) * kRowB), bf);
      mma16816(acc[0], af[tap], bf[0], bf[1]);
      mma16816(acc[1], af[tap], bf[2], bf[3]);
    }
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint16_t m2 = *reinterpret_cast<const uint16_t*>(ok + t * 16 + 8 * j + q2);
      const bool k0 = (m2 & 0xFF) != 0, k1 = (m2 >> 8) != 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float v0 = k0 ? acc[j][2 * h] + (h ? sh_hi : sh_lo) : 0.0f, v1 = k1 ? acc[j][2 * h + 1] + (h ? sh_hi : sh_lo) : 0.0f;
        const __nv_bfloat162 p = __floats2bfloat162_rn(v0, v1);
        pk[2 * j + h] = *reinterpret_cast<const uint32_t*>(&p);
      }
    }
    stsm_x4_t(out + (uint32_t)(t * 16 * kRowB) + l_off, pk);
  }
}

__global__ void __launch_bounds__(kWarps * 32, 1) k_stem_tower16(const S16Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_conv = 2 * a.n_blocks, pitch = geo_pitch(a.W), halo = geo_halo(a.W);
  const bool tail = a.w_s2 != nullptr;
  const int n_w = n_conv + (tail ? 1 : 0);                               // weight slots (the tail's six taps fit one)
  uint8_t* s_w = smem;                                                   // [n_w][9][16][kRowB]
  float* s_sh = reinterpret_cast<float*>(s_w + (size_t)n_w * kConvB);     // [n_w][16]
  uint8_t* s_ok = reinterpret_cast<uint8_t*>(s_sh + n_w * 16);            // [16 T]: 1 = a pixel row, 0 = pad / beyond the image
  uint8_t* s_ok2 = s_ok + 16 * a.T;                                      // [16 T2]: the same for the tail's output geometry
  uint8_t* s_act = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_ok2 + 16 * a.T2) + 127) & ~(uintptr_t)127);
  for (int i = threadIdx.x; i < n_conv * 9 * 16 * 2; i += blockDim.x) {
    const int half = i & 1, n = (i >> 1) & 15, tap = (i >> 5) % 9, ci = i / (9 * 32);
    const uint4 v = *reinterpret_cast<const uint4*>(a.w[ci] + (size_t)n * 144 + tap * 16 + half * 8);
    *reinterpret_cast<uint4*>(s_w + (size_t)ci * kConvB + tap * kTapB + n * kRowB + half * 16) = v;
  }
  for (int i = threadIdx.x; i < n_conv * 16; i += blockDim.x) s_sh[i] = a.shift[i >> 4][i & 15];
  if (tail) {
    for (int i = threadIdx.x; i < 6 * 16 * 2; i += blockDim.x) {
      const int hf = i & 1, n = (i >> 1) & 15, tap = i >> 5;
      *reinterpret_cast<uint4*>(s_w + (size_t)n_conv * kConvB + tap * kTapB + n * kRowB + hf * 16) =
          *reinterpret_cast<const uint4*>(a.w_s2 + (size_t)(tap * 16 + n) * 16 + hf * 8);
    }
    for (int i = threadIdx.x; i < 16; i += blockDim.x) s_sh[n_conv * 16 + i] = a.shift_s2[i];
    for (int i = threadIdx.x; i < 16 * a.T2; i += blockDim.x) {
      const int yy = i / pitch, xx = i - yy * pitch;
      s_ok2[i] = (i < a.R2 && geo_is_pixel(yy, xx, a.W)) ? 1 : 0;
    }
  }
  for (int i = threadIdx.x; i < 16 * a.T; i += blockDim.x) {
    const int yy = i / pitch, xx = i - yy * pitch;
    s_ok[i] = (i < a.R && geo_is_pixel(yy, xx, a.W)) ? 1 : 0;
  }
  for (size_t i = threadIdx.x; i < (size_t)a.G * a.nbuf * a.buf_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(s_act)[i] = make_uint4(0u, 0u, 0u, 0u);      // halo rows stay zero for the whole kernel
  __syncthreads();

  const int grp = warp / a.gw, gwarp = warp - grp * a.gw, gthreads = a.gw * 32, gtid = gwarp * 32 + lane;
  const int bar = 1 + grp;
  uint8_t* const gbuf = s_act + (size_t)grp * a.nbuf * a.buf_bytes;
  const uint32_t img0 = smem_u32(gbuf) + (uint32_t)(halo * kRowB);        // row 0 of the image region in buffer 0
  const uint32_t w_u32 = smem_u32(s_w);
  const long long stride = (long long)gridDim.x * a.G;
  const int chunks = 2 * a.R;                                            // 16-byte pieces of an image in global memory
  auto issue_load = [&](long long b, int bi) {
    const uint4* src = reinterpret_cast<const uint4*>(a.x + ((long long)halo + b * a.R) * 16);
    const uint32_t dst = img0 + (uint32_t)bi * (uint32_t)a.buf_bytes;
    for (int i = gtid; i < chunks; i += gthreads) cp_async16(dst + (uint32_t)((i >> 1) * kRowB + (i & 1) * 16), src + i);
    cp_async_commit();
  };

  long long b = (long long)blockIdx.x * a.G + grp;
  int cur = 0;
  if (b < a.B) issue_load(b, 0);
  while (b < a.B) {
    const long long nb = b + stride;
    const bool pre = a.nbuf == 3 && nb < a.B;
    if (pre) { issue_load(nb, (cur + 2) % 3); cp_async_wait<1>(); } else cp_async_wait<0>();
    group_barrier(bar, gthreads);
    const int tb = a.nbuf == 3 ? (cur + 1) % 3 : 1 - cur;
    const uint32_t X = img0 + (uint32_t)cur * (uint32_t)a.buf_bytes, Tm = img0 + (uint32_t)tb * (uint32_t)a.buf_bytes;
    for (int k = 0; k < a.n_blocks; ++k) {
      if (tail) conv_image<false, true>(X, Tm, w_u32 + (uint32_t)(2 * k) * kConvB, s_sh + 2 * k * 16, s_ok, a.T, pitch, gwarp, a.gw, lane);
      else conv_image<false, false>(X, Tm, w_u32 + (uint32_t)(2 * k) * kConvB, s_sh + 2 * k * 16, s_ok, a.T, pitch, gwarp, a.gw, lane);
      group_barrier(bar, gthreads);
      if (tail) conv_image<true, true>(Tm, X, w_u32 + (uint32_t)(2 * k + 1) * kConvB, s_sh + (2 * k + 1) * 16, s_ok, a.T, pitch, gwarp, a.gw, lane);
      else conv_image<true, false>(Tm, X, w_u32 + (uint32_t)(2 * k + 1) * kConvB, s_sh + (2 * k + 1) * 16, s_ok, a.T, pitch, gwarp, a.gw, lane);
      group_barrier(bar, gthreads);
    }
    if (tail) {
      conv_s2_image(X, Tm, w_u32 + (uint32_t)n_conv * kConvB, s_sh + n_conv * 16, s_ok2, a.T2, a.R2, a.W, pitch, halo, gwarp, a.gw, lane);
      group_barrier(bar, gthreads);
      uint4* dst = reinterpret_cast<uint4*>(a.y2 + ((long long)halo + b * a.R2) * 16);
      const uint8_t* src = gbuf + (size_t)tb * a.buf_bytes + (size_t)halo * kRowB;
      for (int i = gtid; i < 2 * a.R2; i += gthreads) dst[i] = *reinterpret_cast<const uint4*>(src + (i >> 1) * kRowB + (i & 1) * 16);
      if (b == a.B - 1)                                                   // the trailing halo of the new geometry
        for (int i = gtid; i < 2 * halo; i += gthreads) dst[2 * a.R2 + i] = make_uint4(0u, 0u, 0u, 0u);
    } else {
      uint4* dst = reinterpret_cast<uint4*>(a.x + ((long long)halo + b * a.R) * 16);
      const uint8_t* src = gbuf + (size_t)cur * a.buf_bytes + (size_t)halo * kRowB;
      for (int i = gtid; i < chunks; i += gthreads) dst[i] = *reinterpret_cast<const uint4*>(src + (i >> 1) * kRowB + (i & 1) * 16);
    }
    group_barrier(bar, gthreads);                                         // the image's buffer is free again
    if (a.nbuf == 3) cur = (cur + 2) % 3;
    else if (nb < a.B) issue_load(nb, cur);
    b = nb;
  }
}

struct Plan { int gw, G, nbuf, buf_bytes, T, T2; size_t smem; };

// n_w: weight slots (convolutions of the blocks + 1 for the stride-2 tail)
bool make_plan(int H, int W, int n_w, bool tail, Plan* p) {
  const int R = geo_rows_per_image(H, W), T = (R + 15) / 16, halo = geo_halo(W);
  const int T2 = tail ? (geo_rows_per_image(H / 2, W) + 15) / 16 : 0;
  const int buf_bytes = (int)mzb_align_up((size_t)(2 * halo + 16 * T) * kRowB, 128);
  const size_t fixed = 128 + (size_t)n_w * kConvB + sizeof(float) * n_w * 16 + (size_t)16 * (T + T2) + 128;
  int gw = 2;                                                              // <= 8 groups: one named barrier each
  while (gw < kWarps && gw * 2 * 4 <= T) gw *= 2;                          // >= 4 tiles per warp and layer
  // tuning knobs (tests/profile_stem16.py), read once
  static const int env_gw = [] { const char* e = getenv("MZB_STEM16_GW"); return e ? atoi(e) : 0; }();
  static const int force_nbuf = [] { const char* e = getenv("MZB_STEM16_NBUF"); return e ? atoi(e) : 0; }();
  if (env_gw == 2 || env_gw == 4 || env_gw == 8 || env_gw == 16) gw = env_gw;
  for (; gw <= kWarps; gw *= 2) {
    const int G = kWarps / gw;
    for (int nbuf = 3; nbuf >= 2; --nbuf) {
      if (force_nbuf && nbuf != force_nbuf) continue;
      const size_t smem = fixed + (size_t)G * nbuf * buf_bytes;
      if (smem <= 227 * 1024) { *p = Plan{gw, G, nbuf, buf_bytes, T, T2, smem}; return true; }
    }
  }
  return false;
}

}  // namespace

static bool g_stem16_enabled = true;
extern "C" void mzb_stem16_enable(int on) { g_stem16_enabled = on != 0; }       // comparison / bring-up knob

bool mzb_stem16_supported(const std::vector<Block>& blocks, int H, int W, int C) {
  if (!g_stem16_enabled || C != 16 || blocks.empty() || 2 * blocks.size() > (size_t)kMaxConv) return false;
  for (const Block& b : blocks)
    for (const ConvParams* c : {&b.c1, &b.c2})
      if (!c->w_tc || c->cin != 16 || c->cout != 16 || c->extra_plane || c->stride != 1) return false;
  Plan p;
  return make_plan(H, W, 2 * (int)blocks.size() + 1, H % 2 == 0, &p);       // with room for the stride-2 tail
}

// In place: x holds the padded input of the tower on entry and its output on return (pad rows re-written as zeros).
int mzb_stem16_tower(const std::vector<Block>& blocks, int B, int H, int W, __nv_bfloat16* x, cudaStream_t s,
                     const __nv_bfloat16* s2_w, const float* s2_shift, __nv_bfloat16* y2) {
  Plan p;
  const bool tail = s2_w != nullptr;
  if (tail && (!s2_shift || !y2 || H % 2 != 0)) { mzb_set_error("stem16: stride-2 tail needs shift, output and an even height"); return MZB_EINVAL; }
  if (!make_plan(H, W, 2 * (int)blocks.size() + (tail ? 1 : 0), tail, &p)) { mzb_set_error("stem16: image %d x %d does not fit shared memory", H, W); return MZB_EUNSUPPORTED; }
  S16Args a{};
  a.B = B; a.H = H; a.W = W; a.n_blocks = (int)blocks.size();
  a.R = geo_rows_per_image(H, W); a.T = p.T; a.gw = p.gw; a.G = p.G; a.nbuf = p.nbuf; a.buf_bytes = p.buf_bytes;
  int k = 0;
  for (const Block& b : blocks) {
    a.w[k] = b.c1.w_tc; a.shift[k++] = b.c1.shift;
    a.w[k] = b.c2.w_tc; a.shift[k++] = b.c2.shift;
  }
  a.x = x;
  a.w_s2 = s2_w; a.shift_s2 = s2_shift; a.y2 = y2;
  a.R2 = tail ? geo_rows_per_image(H / 2, W) : 0; a.T2 = p.T2;
  static bool configured = false;
  if (!configured) {
    MZB_CUDA(cudaFuncSetAttribute(k_stem_tower16, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  int grid = (B + p.G - 1) / p.G;
  if (grid > n_sm) grid = n_sm;                      // persistent: the weights are staged once per CTA
  k_stem_tower16<<<grid, kWarps * 32, p.smem, s>>>(a);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}
