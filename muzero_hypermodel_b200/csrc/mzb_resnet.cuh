// Residual MuZero network (models.py:206-619) on device: layer program + parameters.
//
// Activations are NHWC ([B][H][W][C], channel fastest): a 3x3 convolution is an implicit GEMM with
// M = B*H*W output positions, N = C_out, K = 9*C_in (tap-major).  Two arithmetic paths share the layer
// program, the head / min-max / pooling kernels and the weights:
//   * fp32  (precision 0): CUDA-core direct convolution, float32 activations - the exact path parity is
//           pinned on (1e-4 vs torch's conv).
//   * bf16  (precision 1): tcgen05 implicit GEMM, bf16 operands staged by TMA, fp32 accumulation in TMEM,
//           folded batch-norm + residual + ReLU in the epilogue (mzb_conv_tc.cu).  Stated bf16 bound.
// Eval-mode batch-norm (self_play.py:29) is folded to per-channel scale/shift at set_weights time.
#pragma once
#include <cuda_bf16.h>

#include <vector>

#include "mzb_common.cuh"

// Activation geometry.  pad = 0: dense NHWC, row (b,y,x) = (b*H + y)*W + x.
// pad = 1 (tensor-core path): every image is stored as (H+1) lines of W+1 rows - one zero line above it and ONE zero
// column per line, which is the right neighbour of that line's last pixel and the left neighbour of the next line's
// first pixel - after a leading halo of W+2 zero rows, so the 3x3 neighbour (dy,dx) of ANY row is the row at flat
// offset dy*(W+1)+dx and out-of-image taps read zeros.  (H+1)(W+1) rows per image: 75 % useful for a 6x7 board, against
// 67 % with a zero column on both sides.  The pad rows are written once (workspace init) and never again, except in the
// DownSample stem, whose layers re-write them (see k_conv_s2).
struct Geo {
  int H, W, C, pad;
};
__host__ __device__ __forceinline__ int geo_pitch(int W) { return W + 1; }                 // rows per image line
__host__ __device__ __forceinline__ int geo_halo(int W) { return W + 2; }                  // >= the largest tap offset
__host__ __device__ __forceinline__ int geo_rows_per_image(int H, int W) { return (H + 1) * (W + 1); }
// row `rem` of an image -> line yy (0 = the zero line), column xx (W = the zero column); a pixel iff yy >= 1 && xx < W
__host__ __device__ __forceinline__ bool geo_is_pixel(int yy, int xx, int W) { return yy >= 1 && xx < W; }
__host__ __device__ __forceinline__ long long geo_row(const Geo& g, int b, int y, int x) {
  return g.pad ? (long long)geo_halo(g.W) + (long long)b * geo_rows_per_image(g.H, g.W) + (long long)(y + 1) * geo_pitch(g.W) + x
               : ((long long)b * g.H + y) * g.W + x;
}
__host__ __device__ __forceinline__ long long geo_rows_total(const Geo& g, long long B) {
  return g.pad ? 2ll * geo_halo(g.W) + B * geo_rows_per_image(g.H, g.W) : B * g.H * g.W;
}

struct ConvParams {
  int cin, cout, stride;          // 3x3, padding 1, no bias (models.py:206-209)
  int extra_plane;                // 1: one more input channel, constant per image (the action plane :553-568)
  float* w;                       // fp32 [9][cin + extra_plane][cout]
  __nv_bfloat16* w_bf16;          // bf16 [cout][9][cin]   (K-major B operand of the implicit GEMM), or NULL
  __nv_bfloat16* w_tc;            // bf16 [cout][9][cin] with the folded batch-norm SCALE multiplied in (fp32 product, one bf16
                                  // rounding): B operand of k_conv_tc, whose epilogue then only adds `shift`
  float* plane_table;             // fp32 [H*W][cout]: sum over in-bounds taps of the extra-plane weights
  float* scale;                   // folded batch-norm: y = conv * scale + shift   (identity when no bn)
  float* shift;
};

struct HeadParams {
  int cin, r, hw;                 // conv1x1 cin -> r channels (with bias), flattened r*hw (channel-major like .view)
  int n_fc;                       // Linear layers of the mlp (hidden ... out)
  int fc_in[4], fc_out[4];
  float* w1x1;                    // [r][cin]
  float* b1x1;                    // [r]
  float* fc_w[4];                 // [in][out]: transposed at set_weights so lanes = outputs read consecutive floats
  float* fc_b[4];
  int out;                        // logits width
  void* mma;                      // bf16 path: the mlp packed for the tensor-core head kernel (mzb_head_mma.cu), or NULL
};

struct Block { ConvParams c1, c2; };
