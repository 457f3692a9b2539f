// Residual MuZero network: model handle, weight packing, layer program, fp32 kernels (K6-K8 exact path).
// Reference: models.py:206-619.  See mzb_resnet.cuh for the data layout.
#include <math.h>
#include <string.h>

#include "mzb_fc.cuh"
#include "mzb_resnet.cuh"
#include "mzb_resnet_model.h"

namespace {

// ------------------------------------------------------------------------------------------ helpers
template <class T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <class T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// NCHW fp32 rows (optionally slot addressed) -> NHWC T in geometry g
template <class T>
__global__ void k_nchw_to_nhwc(const float* __restrict__ in, long long in_row_stride, const int* __restrict__ in_slot,
                               long long slot_stride, int B, Geo g, T* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = g.H * g.W;
  if (i >= (long long)B * HW * g.C) return;
  const int c = (int)(i % g.C);
  const int p = (int)((i / g.C) % HW);
  const int b = (int)(i / ((long long)g.C * HW));
  const float* src = in + b * in_row_stride + (in_slot ? (long long)in_slot[b] * slot_stride : 0);
  stf(out + geo_row(g, b, p / g.W, p % g.W) * g.C + c, src[(long long)c * HW + p]);
}

// dense NHWC T rows (slot addressed pool) -> NHWC T in geometry g, 16 bytes per thread (C*sizeof(T) % 16 == 0)
template <class T>
__global__ void k_gather_nhwc(const T* __restrict__ in, long long in_row_stride, const int* __restrict__ in_slot,
                              long long slot_stride, int B, Geo g, T* __restrict__ out) {
  constexpr int V = 16 / sizeof(T);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = g.H * g.W, cv = g.C / V;
  if (i >= (long long)B * HW * cv) return;
  const int c = (int)(i % cv) * V;
  const int p = (int)((i / cv) % HW);
  const int b = (int)(i / ((long long)cv * HW));
  const uint4 v = *reinterpret_cast<const uint4*>(in + b * in_row_stride + (in_slot ? (long long)in_slot[b] * slot_stride : 0) +
                                                  (long long)p * g.C + c);
  *reinterpret_cast<uint4*>(out + geo_row(g, b, p / g.W, p % g.W) * g.C + c) = v;
}

// 3x3 convolution, padding 1, stride s, + folded batch-norm + residual + ReLU.  Direct fp32-accumulate form:
// one thread per (position, output channel); x reads broadcast across the warp, weight reads coalesced.
template <class T>
__global__ void k_conv3x3_direct(const T* __restrict__ x, int B, Geo gi, ConvParams cp, const float* __restrict__ plane,
                                 const T* __restrict__ residual, int relu, Geo go, T* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * go.H * go.W * cp.cout) return;
  const int co = (int)(i % cp.cout);
  const int ox = (int)((i / cp.cout) % go.W);
  const int oy = (int)((i / ((long long)cp.cout * go.W)) % go.H);
  const int b = (int)(i / ((long long)cp.cout * go.W * go.H));
  const int cw = cp.cin + cp.extra_plane;
  float acc = 0.0f;
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * cp.stride + ky - 1;
    if (iy < 0 || iy >= gi.H) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * cp.stride + kx - 1;
      if (ix < 0 || ix >= gi.W) continue;
      const T* xp = x + geo_row(gi, b, iy, ix) * gi.C;
      const float* wp = cp.w + (size_t)(ky * 3 + kx) * cw * cp.cout + co;
      for (int c = 0; c < cp.cin; ++c) acc = fmaf(ldf(xp + c), wp[(size_t)c * cp.cout], acc);
      if (cp.extra_plane) acc = fmaf(plane[b], wp[(size_t)cp.cin * cp.cout], acc);
    }
  }
  float v = fmaf(acc, cp.scale[co], cp.shift[co]);
  const long long o = geo_row(go, b, oy, ox) * go.C + co;
  if (residual) v += ldf(residual + o);
  if (relu) v = fmaxf(v, 0.0f);
  stf(y + o, v);
}

// AvgPool2d(kernel 3, stride 2, padding 1), count_include_pad (models.py:257-262)
template <class T>
__global__ void k_avgpool(const T* __restrict__ x, int B, Geo gi, Geo go, T* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int C = go.C;
  if (i >= (long long)B * go.H * go.W * C) return;
  const int c = (int)(i % C);
  const int ox = (int)((i / C) % go.W);
  const int oy = (int)((i / ((long long)C * go.W)) % go.H);
  const int b = (int)(i / ((long long)C * go.W * go.H));
  float s = 0.0f;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      const int iy = oy * 2 + ky - 1, ix = ox * 2 + kx - 1;
      if (iy >= 0 && iy < gi.H && ix >= 0 && ix < gi.W) s += ldf(x + geo_row(gi, b, iy, ix) * C + c);
    }
  stf(y + geo_row(go, b, oy, ox) * C + c, s / 9.0f);
}

// ---- bf16 DownSample stem (tensor-core path).  Every stem buffer is in the PADDED layout and every writer also
// writes the pad rows and the trailing halo as zeros: the three rotating buffers change resolution from layer to
// layer, so (unlike the latent path) the pads cannot be left to a zeroed workspace.
//
// Stride-2 3x3 convolution, padding 1 (DownSample.conv1 / conv2, models.py:236-255).  Weights [9*cin][cout]
// (cout a multiple of 8) are broadcast from shared memory, 8 output channels per pass.
// OBS: the input is the caller's fp32 NCHW observation (no staging copy); else a padded NHWC bf16 tensor.
// PIXEL-PAIR layout of the first stage (C/2 channels at half resolution): two horizontally adjacent pixels share one
// row of 2*(C/2) channels, so an 8-channel stage still has the 16-channel rows the UMMA needs WITHOUT zero padding
// (half the bytes of a channel-padded layout); a 3x3 convolution over pixels is a 3x3 convolution over pairs with a
// structured-sparse weight matrix (pack_conv_pair).  out_pair: a thread owns pixel `po` of output pair-row m;
// in_pair: input pixel (y, x) is row (y, x >> 1), channels (x & 1) * cin ...
template <bool OBS>
__global__ void __launch_bounds__(128) k_conv_s2(const void* __restrict__ xin, int B, Geo gi, int cin, int in_pair,
                                                 const float* __restrict__ w, const float* __restrict__ scale,
                                                 const float* __restrict__ shift, int cout, Geo go, int out_pair,
                                                 __nv_bfloat16* __restrict__ y) {
  extern __shared__ float sw[];
  float* s_scale = sw + 9 * cin * cout;
  float* s_shift = s_scale + cout;
  for (int i = threadIdx.x; i < 9 * cin * cout; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < cout; i += blockDim.x) { s_scale[i] = scale[i]; s_shift[i] = shift[i]; }
  __syncthreads();
  const int Wp = geo_pitch(go.W), CO = go.C, ppr = out_pair ? 2 : 1;         // pixels per output row
  const long long R_img = geo_rows_per_image(go.H, go.W);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long m = t / ppr;
  const int po = (int)(t - m * ppr);
  if (m >= (long long)B * R_img + geo_halo(go.W)) return;
  const int b = (int)(m / R_img);
  const int rem = (int)(m - (long long)b * R_img);
  const int yy = rem / Wp, xx = rem - yy * Wp;
  const bool valid = b < B && geo_is_pixel(yy, xx, go.W);
  uint4* out = reinterpret_cast<uint4*>(y + (m + geo_halo(go.W)) * CO + po * cout);
  if (!valid) {
    for (int i = 0; i < cout / 8; ++i) out[i] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const int oy = yy - 1, ox = xx * ppr + po;
  for (int co0 = 0; co0 < cout; co0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int iy = 2 * oy + ky - 1, ix = 2 * ox + kx - 1;
        const float* wt = sw + (size_t)(ky * 3 + kx) * cin * cout + co0;
        if (OBS) {
          if (iy < 0 || iy >= gi.H || ix < 0 || ix >= gi.W) continue;
          const float* src = (const float*)xin + ((size_t)b * cin * gi.H + iy) * gi.W + ix;
          for (int c = 0; c < cin; ++c) {
            const float v = __bfloat162float(__float2bfloat16_rn(src[(size_t)c * gi.H * gi.W]));   // bf16 operands on this path
            const float4* w4 = reinterpret_cast<const float4*>(wt + (size_t)c * cout);
            const float4 w0 = w4[0], w1 = w4[1];
            acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
            acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
          }
        } else {
          // out-of-image taps land on zero pad rows of the padded input: no bounds checks
          const int ixr = in_pair ? (ix >> 1) : ix, coff = in_pair ? (ix & 1) * cin : 0;
          const uint4* src = reinterpret_cast<const uint4*>((const __nv_bfloat16*)xin + geo_row(gi, b, iy, ixr) * gi.C + coff);
          for (int c8 = 0; c8 < cin / 8; ++c8) {
            const uint4 raw = src[c8];
            const uint32_t r4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int h = 0; h < 8; ++h) {
              const float v = __uint_as_float((h & 1) ? (r4[h >> 1] & 0xFFFF0000u) : (r4[h >> 1] << 16));
              const float4* w4 = reinterpret_cast<const float4*>(wt + (size_t)(c8 * 8 + h) * cout);
              const float4 w0 = w4[0], w1 = w4[1];
              acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
              acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
            }
          }
        }
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 pk = __floats2bfloat162_rn(fmaf(acc[2 * i], s_scale[co0 + 2 * i], s_shift[co0 + 2 * i]),
                                                      fmaf(acc[2 * i + 1], s_scale[co0 + 2 * i + 1], s_shift[co0 + 2 * i + 1]));
      o[i] = *reinterpret_cast<const uint32_t*>(&pk);
    }
    out[co0 / 8] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// DownSample.conv1 specialised for the pixel-pair output (frame width a multiple of 4): a thread owns one output
// PAIR, i.e. the input columns 4*xp-1 .. 4*xp+3 of three frame lines, fetched as one aligned float4 plus one scalar per
// (line, plane) - consecutive lanes read consecutive 16-byte segments of the fp32 frame (the kernel is a 1.8 GB read
// of the observation batch; the generic thread-per-pixel form issued 27 scattered 4-byte loads per pixel).
__global__ void __launch_bounds__(128) k_stem_conv1_pair(const float* __restrict__ obs, int B, Geo gi, int cin,
                                                         const float* __restrict__ w, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, int cout, Geo go,
                                                         __nv_bfloat16* __restrict__ y) {
  extern __shared__ float sw[];
  float* s_scale = sw + 9 * cin * cout;
  float* s_shift = s_scale + cout;
  for (int i = threadIdx.x; i < 9 * cin * cout; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < cout; i += blockDim.x) { s_scale[i] = scale[i]; s_shift[i] = shift[i]; }
  __syncthreads();
  const int Wp = geo_pitch(go.W), CO = go.C;
  const long long R_img = geo_rows_per_image(go.H, go.W);
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= (long long)B * R_img + geo_halo(go.W)) return;
  const int b = (int)(m / R_img);
  const int rem = (int)(m - (long long)b * R_img);
  const int yy = rem / Wp, xx = rem - yy * Wp;
  uint4* out = reinterpret_cast<uint4*>(y + (m + geo_halo(go.W)) * CO);
  if (!(b < B && geo_is_pixel(yy, xx, go.W))) {
    for (int i = 0; i < CO / 8; ++i) out[i] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const int oy = yy - 1, xp = xx;
  auto bf = [](float v) { return __bfloat162float(__float2bfloat16_rn(v)); };      // bf16 operands on this path
  for (int co0 = 0; co0 < cout; co0 += 8) {
    // packed FFMA2 (two IEEE fp32 FMAs per instruction over adjacent output channels: the same chain per output as the
    // scalar form, half the FMA instructions - the kernel is issue-bound: 71 % of the issue slots at 37 % of the FMA pipe)
    float2 a0[4], a1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { a0[j] = make_float2(0.0f, 0.0f); a1[j] = make_float2(0.0f, 0.0f); }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = 2 * oy + ky - 1;
      if (iy < 0 || iy >= gi.H) continue;
      for (int c = 0; c < cin; ++c) {
        const float* line = obs + ((size_t)b * cin + c) * gi.H * gi.W + (size_t)iy * gi.W;
        const float4 qv = *reinterpret_cast<const float4*>(line + 4 * xp);
        const float in[5] = {xp > 0 ? bf(line[4 * xp - 1]) : 0.0f, bf(qv.x), bf(qv.y), bf(qv.z), bf(qv.w)};
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4* w4 = reinterpret_cast<const float4*>(sw + (size_t)((ky * 3 + kx) * cin + c) * cout + co0);
          const float4 w0 = w4[0], w1 = w4[1];
          const float2 wv[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
          const float2 x0 = make_float2(in[kx], in[kx]), x1 = make_float2(in[kx + 2], in[kx + 2]);
#pragma unroll
          for (int j = 0; j < 4; ++j) { a0[j] = __ffma2_rn(x0, wv[j], a0[j]); a1[j] = __ffma2_rn(x1, wv[j], a1[j]); }
        }
      }
    }
    uint32_t o0[4], o1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float sc0 = s_scale[co0 + 2 * i], sc1 = s_scale[co0 + 2 * i + 1], sh0 = s_shift[co0 + 2 * i], sh1 = s_shift[co0 + 2 * i + 1];
      const __nv_bfloat162 p0 = __floats2bfloat162_rn(fmaf(a0[i].x, sc0, sh0), fmaf(a0[i].y, sc1, sh1));
      const __nv_bfloat162 p1 = __floats2bfloat162_rn(fmaf(a1[i].x, sc0, sh0), fmaf(a1[i].y, sc1, sh1));
      o0[i] = *reinterpret_cast<const uint32_t*>(&p0);
      o1[i] = *reinterpret_cast<const uint32_t*>(&p1);
    }
    out[co0 / 8] = make_uint4(o0[0], o0[1], o0[2], o0[3]);
    out[(cout + co0) / 8] = make_uint4(o1[0], o1[1], o1[2], o1[3]);
  }
}

// AvgPool2d(kernel 3, stride 2, padding 1), count_include_pad, padded bf16 layouts on both sides (out-of-image taps
// read the zero pads); one thread per (output row, 8-channel chunk), pad rows written as zeros.
__global__ void __launch_bounds__(256) k_avgpool_pad(const __nv_bfloat16* __restrict__ x, int B, Geo gi, Geo go,
                                                     __nv_bfloat16* __restrict__ y) {
  const int cv = go.C / 8, Wp = geo_pitch(go.W);
  const long long R_img = geo_rows_per_image(go.H, go.W);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long m = i / cv;
  const int ch = (int)(i - m * cv);
  if (m >= (long long)B * R_img + geo_halo(go.W)) return;
  const int b = (int)(m / R_img);
  const int rem = (int)(m - (long long)b * R_img);
  const int yy = rem / Wp, xx = rem - yy * Wp;
  uint4* out = reinterpret_cast<uint4*>(y + (m + geo_halo(go.W)) * go.C) + ch;
  if (!(b < B && geo_is_pixel(yy, xx, go.W))) { *out = make_uint4(0u, 0u, 0u, 0u); return; }
  const int oy = yy - 1, ox = xx;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.0f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const uint4 raw = *(reinterpret_cast<const uint4*>(x + geo_row(gi, b, oy * 2 + ky - 1, ox * 2 + kx - 1) * gi.C) + ch);
      const uint32_t r4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int h = 0; h < 4; ++h) { s[2 * h] += __uint_as_float(r4[h] << 16); s[2 * h + 1] += __uint_as_float(r4[h] & 0xFFFF0000u); }
    }
  uint32_t o[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const __nv_bfloat162 pk = __floats2bfloat162_rn(s[2 * h] / 9.0f, s[2 * h + 1] / 9.0f);
    o[h] = *reinterpret_cast<const uint32_t*>(&pk);
  }
  *out = make_uint4(o[0], o[1], o[2], o[3]);
}

// Head: conv1x1(+bias) -> flatten (channel-major, like .view on NCHW) -> mlp -> logits -> decode.
// One WARP per image (grid-stride), 8 warps per block; the head's weights are staged once per block in shared
// memory (fc weights transposed [in][out] so lanes = outputs read conflict-free).  mode 0: support_to_scalar ->
// scalar_out; mode 1: softmax over legal -> priors_out.  Reductions are warp shuffles.
struct HeadSmem { int w1, b1, fw[4], fb[4], scratch, per_warp, total; };

__host__ __device__ inline HeadSmem head_smem(const HeadParams& hp, int warps) {
  HeadSmem L{};
  int off = 0;
  L.w1 = off; off += 16 * hp.cin;                  // padded to the largest R instantiation, zero filled
  L.b1 = off; off += hp.r;
  for (int l = 0; l < hp.n_fc; ++l) { L.fw[l] = off; off += hp.fc_in[l] * hp.fc_out[l]; L.fb[l] = off; off += hp.fc_out[l]; }
  off = (off + 3) & ~3;
  L.scratch = off;
  int widest = hp.r * hp.hw;
  for (int l = 0; l < hp.n_fc; ++l) widest = widest > hp.fc_out[l] ? widest : hp.fc_out[l];
  L.per_warp = 2 * ((widest + 3) & ~3);
  L.total = off + warps * L.per_warp;
  return L;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

template <class T, int R>
__global__ void __launch_bounds__(256) k_head(const T* __restrict__ x, int B, Geo g, HeadParams hp, int S, int mode,
                                              const uint8_t* __restrict__ legal, float* __restrict__ logits_out,
                                              float* __restrict__ scalar_out, float* __restrict__ priors_out,
                                              const float* __restrict__ proj, long long proj_stride, int proj_off) {
  extern __shared__ float sm[];
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const HeadSmem L = head_smem(hp, warps);
  for (int i = threadIdx.x; i < R * hp.cin; i += blockDim.x) sm[L.w1 + i] = i < hp.r * hp.cin ? hp.w1x1[i] : 0.0f;
  for (int i = threadIdx.x; i < hp.r; i += blockDim.x) sm[L.b1 + i] = hp.b1x1[i];
  for (int l = 0; l < hp.n_fc; ++l) {
    const int in = hp.fc_in[l], out = hp.fc_out[l];
    for (int i = threadIdx.x; i < in * out; i += blockDim.x) sm[L.fw[l] + i] = hp.fc_w[l][i];     // already [in][out]
    for (int i = threadIdx.x; i < out; i += blockDim.x) sm[L.fb[l] + i] = hp.fc_b[l][i];
  }
  __syncthreads();
  float* bufA = sm + L.scratch + warp * L.per_warp;
  float* bufB = bufA + L.per_warp / 2;
  const int n = hp.out;
  // conv1x1 mapping: a group of LPP lanes shares one position, each lane owns 16-byte channel chunks
  constexpr int V = 16 / (int)sizeof(T);
  const int cpr = hp.cin / V;                               // 16-byte chunks per row
  const int LPP = cpr < 32 ? cpr : 32;                      // lanes per position (power of two)
  const int PPI = 32 / LPP;                                 // positions per warp iteration
  const int sub = lane % LPP;
  for (int b = blockIdx.x * warps + warp; b < B; b += gridDim.x * warps) {
    if (proj) {
      // the 1x1 convolution was computed by the producing convolution's epilogue (mzb_conv_tc.cu): add the bias
      const float* pr = proj + (long long)b * proj_stride + proj_off;
      for (int rr = 0; rr < hp.r; ++rr) {
        const float bias = sm[L.b1 + rr];
        for (int i = lane; i < hp.hw; i += 32) bufA[rr * hp.hw + i] = pr[rr * hp.hw + i] + bias;
      }
    } else
    for (int p0 = 0; p0 < hp.hw; p0 += PPI) {
      const int p = p0 + lane / LPP;
      const bool valid = p < hp.hw;
      float acc[R];
#pragma unroll
      for (int rr = 0; rr < R; ++rr) acc[rr] = 0.0f;
      if (valid) {
        const T* row = x + geo_row(g, b, p / g.W, p % g.W) * hp.cin;
        for (int ch = sub; ch < cpr; ch += LPP) {
          const uint4 raw = *reinterpret_cast<const uint4*>(row + ch * V);
          float xv[V];
          if (sizeof(T) == 2) {
            const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { xv[(2 * j) % V] = __uint_as_float(w4[j] << 16); xv[(2 * j + 1) % V] = __uint_as_float(w4[j] & 0xFFFF0000u); }
          } else {
            xv[0] = __uint_as_float(raw.x); xv[1] = __uint_as_float(raw.y); xv[2 % V] = __uint_as_float(raw.z); xv[3 % V] = __uint_as_float(raw.w);
          }
          const float* w = sm + L.w1 + ch * V;
#pragma unroll
          for (int rr = 0; rr < R; ++rr) {
#pragma unroll
            for (int j = 0; j < V; ++j) acc[rr] = fmaf(xv[j], w[rr * hp.cin + j], acc[rr]);
          }
        }
      }
      for (int o = LPP >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int rr = 0; rr < R; ++rr) acc[rr] += __shfl_xor_sync(0xFFFFFFFFu, acc[rr], o);
      }
      if (valid && sub == 0) {
#pragma unroll
        for (int rr = 0; rr < R; ++rr) if (rr < hp.r) bufA[rr * hp.hw + p] = acc[rr] + sm[L.b1 + rr];
      }
    }
    __syncwarp();
    float* in = bufA;
    float* out = bufB;
    for (int l = 0; l < hp.n_fc; ++l) {
      const int ni = hp.fc_in[l], no = hp.fc_out[l];
      const float* w = sm + L.fw[l];
      for (int o = lane; o < no; o += 32) {
        // four interleaved partial sums: the dot products are short dependent FMA chains otherwise
        float a0 = sm[L.fb[l] + o], a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        int k = 0;
        const float* wk = w + o;
        for (; k + 4 <= ni; k += 4, wk += 4 * no) {
          const float4 x4 = *reinterpret_cast<const float4*>(in + k);      // one broadcast load for four inputs
          a0 = fmaf(x4.x, wk[0], a0);
          a1 = fmaf(x4.y, wk[no], a1);
          a2 = fmaf(x4.z, wk[2 * no], a2);
          a3 = fmaf(x4.w, wk[3 * no], a3);
        }
        for (; k < ni; ++k) a0 = fmaf(in[k], w[k * no + o], a0);
        const float acc = (a0 + a1) + (a2 + a3);
        out[o] = l < hp.n_fc - 1 ? elu_f32(acc) : acc;
      }
      __syncwarp();
      float* t = in; in = out; out = t;
    }
    if (logits_out) for (int o = lane; o < n; o += 32) logits_out[(long long)b * n + o] = in[o];
    if (mode == 0 && scalar_out) {
      // support_to_scalar (models.py:641-662): softmax expectation over [-S..S], inverse transform
      float m = -CUDART_INF_F;
      for (int i = lane; i < n; i += 32) m = fmaxf(m, in[i]);
      m = warp_max(m);
      float se = 0.0f, sx = 0.0f;
      for (int i = lane; i < n; i += 32) { const float e = softmax_exp(in[i], m); se += e; sx = fmaf((float)(i - S), e, sx); }
      se = warp_sum(se); sx = warp_sum(sx);
      if (lane == 0) scalar_out[b] = inverse_value_transform(__fdiv_rn(sx, se));
    } else if (mode == 1 && priors_out) {
      const uint8_t* lg = legal ? legal + (long long)b * n : nullptr;
      float m = -CUDART_INF_F;
      for (int a = lane; a < n; a += 32) if (!lg || lg[a]) m = fmaxf(m, in[a]);
      m = warp_max(m);
      float se = 0.0f;
      for (int a = lane; a < n; a += 32) { const float e = (!lg || lg[a]) ? softmax_exp(in[a], m) : 0.0f; out[a] = e; se += e; }
      se = warp_sum(se);
      for (int a = lane; a < n; a += 32) priors_out[(long long)b * n + a] = __fdiv_rn(out[a], se);
    }
    __syncwarp();
  }
}

// Fused: per-(image, channel) min-max scaling (models.py:525-549) + normalised state written both to the next
// layer's input buffer and to the caller's state rows.  One warp per image, lanes = channel pairs, so every
// position is one coalesced row read; C must be even and <= 512.
template <class T> struct Pair;
template <> struct Pair<float> {
  static __device__ __forceinline__ float2 ld(const float* p) { return *reinterpret_cast<const float2*>(p); }
  static __device__ __forceinline__ void st(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
};
template <> struct Pair<__nv_bfloat16> {
  static __device__ __forceinline__ float2 ld(const __nv_bfloat16* p) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float2 v) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y); }
};

template <class T, int MAXP>
__global__ void __launch_bounds__(256) k_minmax_store(const T* __restrict__ x, int B, Geo g, T* __restrict__ y, int layout,
                                                      void* __restrict__ state, long long row_stride, long long off) {
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = g.C, HW = g.H * g.W, pairs = C / 2;       // MAXP = channel pairs per lane (C <= 64 * MAXP)
  for (int b = blockIdx.x * warps + warp; b < B; b += gridDim.x * warps) {
    float lo[2 * MAXP], hi[2 * MAXP];
#pragma unroll
    for (int k = 0; k < 2 * MAXP; ++k) { lo[k] = CUDART_INF_F; hi[k] = -CUDART_INF_F; }
    // pass 1: rows of one board line are contiguous; several independent row loads in flight per lane
    for (int yy = 0; yy < g.H; ++yy) {
      const T* line = x + geo_row(g, b, yy, 0) * C;
#pragma unroll 4
      for (int xx = 0; xx < g.W; ++xx) {
#pragma unroll
        for (int k = 0; k < MAXP; ++k) {
          const int c2 = lane + 32 * k;
          if (c2 < pairs) {
            const float2 v = Pair<T>::ld(line + (long long)xx * C + 2 * c2);
            lo[2 * k] = fminf(lo[2 * k], v.x); hi[2 * k] = fmaxf(hi[2 * k], v.x);
            lo[2 * k + 1] = fminf(lo[2 * k + 1], v.y); hi[2 * k + 1] = fmaxf(hi[2 * k + 1], v.y);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 2 * MAXP; ++k) {
      float scale = __fsub_rn(hi[k], lo[k]);
      if (scale < 1e-5f) scale = __fadd_rn(scale, 1e-5f);
      hi[k] = scale;
    }
    for (int yy = 0; yy < g.H; ++yy) {
      const long long lo_off = geo_row(g, b, yy, 0) * C;
#pragma unroll 4
      for (int xx = 0; xx < g.W; ++xx) {
        const long long ro = lo_off + (long long)xx * C;
        const int p = yy * g.W + xx;
#pragma unroll
        for (int k = 0; k < MAXP; ++k) {
          const int c2 = lane + 32 * k;
          if (c2 < pairs) {
            float2 v = Pair<T>::ld(x + ro + 2 * c2);
            v.x = __fdiv_rn(__fsub_rn(v.x, lo[2 * k]), hi[2 * k]);
            v.y = __fdiv_rn(__fsub_rn(v.y, lo[2 * k + 1]), hi[2 * k + 1]);
            Pair<T>::st(y + ro + 2 * c2, v);
            if (state) {
              const int c = 2 * c2;
              if (layout == 0) {
                float* o = (float*)state + b * row_stride + off;
                o[(long long)c * HW + p] = sizeof(T) == 2 ? __bfloat162float(__float2bfloat16_rn(v.x)) : v.x;
                o[(long long)(c + 1) * HW + p] = sizeof(T) == 2 ? __bfloat162float(__float2bfloat16_rn(v.y)) : v.y;
              } else if (layout == 1) {
                *reinterpret_cast<float2*>((float*)state + b * row_stride + off + (long long)p * C + c) = v;
              } else {
                *reinterpret_cast<__nv_bfloat162*>((__nv_bfloat16*)state + b * row_stride + off + (long long)p * C + c) =
                    __floats2bfloat162_rn(v.x, v.y);
              }
            }
          }
        }
      }
    }
  }
}

// bf16 single-pass variant: one warp per image, lane = (row group, 16-byte channel chunk); the image's rows stay in
// registers between the min/max reduction (shuffles across the row groups) and the rescale, so the activation is read
// once with all loads in flight.  The rescale multiplies by one reciprocal per channel (the fp32 path divides per
// element like the reference; here the bf16 rounding of the result is 2^-9, the reciprocal's error 2^-23).  C in {8,16,...,256} (C/8 a power of two).
template <int MAXIT>
__global__ void __launch_bounds__(256) k_minmax_store_bf16(const __nv_bfloat16* __restrict__ x, int B, Geo g,
                                                           __nv_bfloat16* __restrict__ y, int layout, void* __restrict__ state,
                                                           long long row_stride, long long off) {
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = g.C, HW = g.H * g.W, LPR = C / 8, RPI = 32 / LPR, sub = lane % LPR, rg = lane / LPR;
  const int nit = (HW + RPI - 1) / RPI;
  for (int b = blockIdx.x * warps + warp; b < B; b += gridDim.x * warps) {
    uint4 v[MAXIT];
    long long ro[MAXIT];                                  // element offset of this lane's chunk, per iteration
    float lo[8], hi[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { lo[k] = CUDART_INF_F; hi[k] = -CUDART_INF_F; }
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int p = it * RPI + rg;
      if (it < nit && p < HW) {
        ro[it] = geo_row(g, b, p / g.W, p % g.W) * C + sub * 8;
        v[it] = *reinterpret_cast<const uint4*>(x + ro[it]);
      }
    }
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int p = it * RPI + rg;
      if (it < nit && p < HW) {
        const uint32_t r4[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float a = __uint_as_float(r4[h] << 16), c = __uint_as_float(r4[h] & 0xFFFF0000u);
          lo[2 * h] = fminf(lo[2 * h], a); hi[2 * h] = fmaxf(hi[2 * h], a);
          lo[2 * h + 1] = fminf(lo[2 * h + 1], c); hi[2 * h + 1] = fmaxf(hi[2 * h + 1], c);
        }
      }
    }
    for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        lo[k] = fminf(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], o));
        hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], o));
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float scale = __fsub_rn(hi[k], lo[k]);
      if (scale < 1e-5f) scale = __fadd_rn(scale, 1e-5f);
      hi[k] = 1.0f / scale;                               // bf16 path: one reciprocal per channel, not a division per element
    }
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int p = it * RPI + rg;
      if (it < nit && p < HW) {
        const uint32_t r4[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
        float f[8];
        uint32_t o4[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          f[2 * h] = (__uint_as_float(r4[h] << 16) - lo[2 * h]) * hi[2 * h];
          f[2 * h + 1] = (__uint_as_float(r4[h] & 0xFFFF0000u) - lo[2 * h + 1]) * hi[2 * h + 1];
          const __nv_bfloat162 pk = __floats2bfloat162_rn(f[2 * h], f[2 * h + 1]);
          o4[h] = *reinterpret_cast<const uint32_t*>(&pk);
        }
        const uint4 packed = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        *reinterpret_cast<uint4*>(y + ro[it]) = packed;
        if (state) {
          const int c = sub * 8;
          if (layout == 2) {
            *reinterpret_cast<uint4*>((__nv_bfloat16*)state + b * row_stride + off + (long long)p * C + c) = packed;
          } else if (layout == 1) {
            float4* o = reinterpret_cast<float4*>((float*)state + b * row_stride + off + (long long)p * C + c);
            o[0] = make_float4(f[0], f[1], f[2], f[3]); o[1] = make_float4(f[4], f[5], f[6], f[7]);
          } else {
            float* o = (float*)state + b * row_stride + off;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              o[(long long)(c + 2 * h) * HW + p] = __uint_as_float(o4[h] << 16);
              o[(long long)(c + 2 * h + 1) * HW + p] = __uint_as_float(o4[h] & 0xFFFF0000u);
            }
          }
        }
      }
    }
  }
}

// Block-per-image variant for large images or small batches (gomoku: 121 positions x 128 channels, 1,024 images - one
// warp per image would leave most of the GPU idle): 256 threads = (row group, 16-byte channel chunk), rows in registers,
// min/max reduced with shuffles inside a warp and through shared memory across the warps.  Same arithmetic.
template <int MAXIT>
__global__ void __launch_bounds__(256) k_minmax_store_bf16_block(const __nv_bfloat16* __restrict__ x, int B, Geo g,
                                                                 __nv_bfloat16* __restrict__ y, int layout,
                                                                 void* __restrict__ state, long long row_stride, long long off) {
  __shared__ float s_lo[8][256], s_hi[8][256];          // [warp][channel]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = g.C, HW = g.H * g.W, LPR = C / 8, RPB = 256 / LPR, sub = threadIdx.x % LPR, rg = threadIdx.x / LPR;
  const int nit = (HW + RPB - 1) / RPB;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    uint4 v[MAXIT];
    long long ro[MAXIT];
    float lo[8], hi[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { lo[k] = CUDART_INF_F; hi[k] = -CUDART_INF_F; }
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int p = it * RPB + rg;
      if (it < nit && p < HW) {
        ro[it] = geo_row(g, b, p / g.W, p % g.W) * C + sub * 8;
        v[it] = *reinterpret_cast<const uint4*>(x + ro[it]);
      }
    }
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int p = it * RPB + rg;
      if (it < nit && p < HW) {
        const uint32_t r4[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float a = __uint_as_float(r4[h] << 16), c = __uint_as_float(r4[h] & 0xFFFF0000u);
          lo[2 * h] = fminf(lo[2 * h], a); hi[2 * h] = fmaxf(hi[2 * h], a);
          lo[2 * h + 1] = fminf(lo[2 * h + 1], c); hi[2 * h + 1] = fmaxf(hi[2 * h + 1], c);
        }
      }
    }
    for (int o = LPR; o < 32; o <<= 1) {                 // row groups that share a warp (LPR < 32)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        lo[k] = fminf(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], o));
        hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], o));
      }
    }
    __syncthreads();                                     // previous image's readers are done with the scratch
    if (lane < LPR) {                                   // LPR <= 32 divides 32: every warp holds all channel chunks
#pragma unroll
      for (int k = 0; k < 8; ++k) { s_lo[warp][sub * 8 + k] = lo[k]; s_hi[warp][sub * 8 + k] = hi[k]; }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float l = CUDART_INF_F, h = -CUDART_INF_F;
      for (int w2 = 0; w2 < 8; ++w2) { l = fminf(l, s_lo[w2][sub * 8 + k]); h = fmaxf(h, s_hi[w2][sub * 8 + k]); }
      float scale = __fsub_rn(h, l);
      if (scale < 1e-5f) scale = __fadd_rn(scale, 1e-5f);
      lo[k] = l; hi[k] = 1.0f / scale;
    }
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int p = it * RPB + rg;
      if (it < nit && p < HW) {
        const uint32_t r4[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
        float f[8];
        uint32_t o4[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          f[2 * h] = (__uint_as_float(r4[h] << 16) - lo[2 * h]) * hi[2 * h];
          f[2 * h + 1] = (__uint_as_float(r4[h] & 0xFFFF0000u) - lo[2 * h + 1]) * hi[2 * h + 1];
          const __nv_bfloat162 pk = __floats2bfloat162_rn(f[2 * h], f[2 * h + 1]);
          o4[h] = *reinterpret_cast<const uint32_t*>(&pk);
        }
        const uint4 packed = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        *reinterpret_cast<uint4*>(y + ro[it]) = packed;
        if (state) {
          const int c = sub * 8;
          if (layout == 2) {
            *reinterpret_cast<uint4*>((__nv_bfloat16*)state + b * row_stride + off + (long long)p * C + c) = packed;
          } else if (layout == 1) {
            float4* o = reinterpret_cast<float4*>((float*)state + b * row_stride + off + (long long)p * C + c);
            o[0] = make_float4(f[0], f[1], f[2], f[3]); o[1] = make_float4(f[4], f[5], f[6], f[7]);
          } else {
            float* o = (float*)state + b * row_stride + off;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              o[(long long)(c + 2 * h) * HW + p] = __uint_as_float(o4[h] << 16);
              o[(long long)(c + 2 * h + 1) * HW + p] = __uint_as_float(o4[h] & 0xFFFF0000u);
            }
          }
        }
      }
    }
  }
}

__global__ void k_zero_reward(int B, int S, float* __restrict__ logits, float* __restrict__ scalar) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int full = 2 * S + 1;
  if (logits) for (int i = 0; i < full; ++i) logits[(long long)b * full + i] = (i == S) ? 0.0f : -CUDART_INF_F;
  if (scalar) {
    float e[64];
    scalar[b] = support_to_scalar_dev([=](int i) { return i == S ? 0.0f : -CUDART_INF_F; },
                                      [&](int i) -> float& { return e[i < 64 ? i : 63]; }, S);
  }
}

__global__ void k_action_plane(const int* __restrict__ action, int B, int A, float* __restrict__ plane) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) plane[b] = __fdiv_rn((float)action[b], (float)A);          // action * ones / action_space_size (:553-568)
}

inline unsigned nblk(long long n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

// ------------------------------------------------------------------------------------------ model build
static void* dev_alloc(mzb_resnet_model* m, size_t bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 4) != cudaSuccess) return nullptr;
  cudaMemset(p, 0, bytes ? bytes : 4);
  m->allocs.push_back(p);
  return p;
}

static bool init_conv(mzb_resnet_model* m, ConvParams& c, int cin, int cout, int stride, int extra, int hw) {
  c.cin = cin; c.cout = cout; c.stride = stride; c.extra_plane = extra;
  c.w = (float*)dev_alloc(m, sizeof(float) * 9 * (size_t)(cin + extra) * cout);
  c.scale = (float*)dev_alloc(m, sizeof(float) * cout);
  c.shift = (float*)dev_alloc(m, sizeof(float) * cout);
  c.w_bf16 = (__nv_bfloat16*)dev_alloc(m, sizeof(__nv_bfloat16) * 9 * (size_t)cin * cout);
  c.w_tc = (__nv_bfloat16*)dev_alloc(m, sizeof(__nv_bfloat16) * 9 * (size_t)cin * cout);
  c.plane_table = extra ? (float*)dev_alloc(m, sizeof(float) * (size_t)hw * cout) : nullptr;
  return c.w && c.scale && c.shift && c.w_bf16 && c.w_tc && (!extra || c.plane_table);
}

static bool init_head(mzb_resnet_model* m, HeadParams& h, int cin, int r, int hw, const int32_t* hidden, int n_hidden, int out) {
  h.cin = cin; h.r = r; h.hw = hw; h.out = out; h.n_fc = n_hidden + 1; h.mma = nullptr;
  h.w1x1 = (float*)dev_alloc(m, sizeof(float) * (size_t)r * cin);
  h.b1x1 = (float*)dev_alloc(m, sizeof(float) * r);
  int cur = r * hw;
  for (int l = 0; l < h.n_fc; ++l) {
    const int o = l < n_hidden ? hidden[l] : out;
    if (o <= 0 || o > 128) return false;
    h.fc_in[l] = cur; h.fc_out[l] = o;
    h.fc_w[l] = (float*)dev_alloc(m, sizeof(float) * (size_t)cur * o);
    h.fc_b[l] = (float*)dev_alloc(m, sizeof(float) * o);
    if (!h.fc_w[l] || !h.fc_b[l]) return false;
    cur = o;
  }
  return h.w1x1 && h.b1x1;
}

static int count_tensors(const mzb_resnet_model* m) {
  const int blk = 2 * 5;                                  // conv + 4 bn tensors, twice
  int n = 0;
  if (m->downsample) n += 1 + 2 * blk + 1 + 3 * blk + 3 * blk;
  n += 5 + m->blocks * blk;                               // representation conv+bn (present even when unused), resblocks
  n += 5 + m->blocks * blk + 2 + 2 * m->reward.n_fc;      // dynamics
  n += m->blocks * blk + 2 + 2 + 2 * m->value.n_fc + 2 * m->policy.n_fc;
  return n;
}

extern "C" {

int mzb_resnet_create(mzb_resnet_model** out, const mzb_resnet_config* c) {
  MZB_CHECK_ARG(out && c, "NULL argument");
  *out = nullptr;
  MZB_CHECK_ARG(c->obs_channels > 0 && c->height > 0 && c->width > 0 && c->n_actions > 0 && c->blocks >= 0 &&
                c->channels > 0 && c->support_size > 0 && c->support_size <= 31, "resnet config out of range");
  MZB_CHECK_ARG(c->downsample == 0 || c->downsample == 1, "downsample must be 0 (False) or 1 (\"resnet\")");
  MZB_CHECK_ARG(c->n_fc_reward >= 0 && c->n_fc_reward <= 3 && c->n_fc_value >= 0 && c->n_fc_value <= 3 &&
                c->n_fc_policy >= 0 && c->n_fc_policy <= 3, "at most 3 hidden layers per head mlp");
  mzb_resnet_model* m = new mzb_resnet_model();
  m->cfg = *c;
  m->Cobs = c->obs_channels; m->H = c->height; m->W = c->width; m->A = c->n_actions; m->S = c->support_size;
  m->full = 2 * c->support_size + 1; m->blocks = c->blocks; m->C = c->channels; m->downsample = c->downsample;
  m->precision = c->precision;
  m->Hl = c->downsample ? (c->height + 15) / 16 : c->height;
  m->Wl = c->downsample ? (c->width + 15) / 16 : c->width;
  const int hw = m->Hl * m->Wl, C = m->C;
  bool ok = true;
  if (m->downsample) {
    ok = ok && init_conv(m, m->ds_conv1, m->Cobs, C / 2, 2, 0, 0) && init_conv(m, m->ds_conv2, C / 2, C, 2, 0, 0);
    m->ds1.resize(2); m->ds2.resize(3); m->ds3.resize(3);
    for (auto& b : m->ds1) ok = ok && init_conv(m, b.c1, C / 2, C / 2, 1, 0, 0) && init_conv(m, b.c2, C / 2, C / 2, 1, 0, 0);
    for (auto& b : m->ds2) ok = ok && init_conv(m, b.c1, C, C, 1, 0, 0) && init_conv(m, b.c2, C, C, 1, 0, 0);
    for (auto& b : m->ds3) ok = ok && init_conv(m, b.c1, C, C, 1, 0, 0) && init_conv(m, b.c2, C, C, 1, 0, 0);
    // bf16: the whole stem runs in the padded layout - stride-2 layers on a dedicated kernel, every residual block
    // on the tcgen05 convolution (resblocks1 with its C/2 channels zero-padded to a multiple of 16)
    const int w1 = (m->W - 1) / 2 + 1;
    if (m->precision == 1 && C % 16 == 0 && (C / 2) % 8 == 0 && C <= 128 && w1 % 2 == 0 && w1 / 2 <= 61) {
      m->stem_tc = 1;
      m->stem_cp1 = C;                       // pixel-pair rows: 2 x (C/2) channels
      m->ds1_tc.resize(2);
      for (auto& b : m->ds1_tc)
        ok = ok && init_conv(m, b.c1, m->stem_cp1, m->stem_cp1, 1, 0, 0) && init_conv(m, b.c2, m->stem_cp1, m->stem_cp1, 1, 0, 0);
      if (C == 16) {
        m->ds_conv2_mma = (__nv_bfloat16*)dev_alloc(m, sizeof(__nv_bfloat16) * 6 * C * 16);
        ok = ok && m->ds_conv2_mma;
      }
    }
  }
  ok = ok && init_conv(m, m->rep_conv, m->Cobs, C, 1, 0, 0) && init_conv(m, m->dyn_conv, C, C, 1, 1, hw);
  m->rep_blocks.resize(m->blocks); m->dyn_blocks.resize(m->blocks); m->pred_blocks.resize(m->blocks);
  for (auto* v : {&m->rep_blocks, &m->dyn_blocks, &m->pred_blocks})
    for (auto& b : *v) ok = ok && init_conv(m, b.c1, C, C, 1, 0, 0) && init_conv(m, b.c2, C, C, 1, 0, 0);
  ok = ok && init_head(m, m->reward, C, c->reduced_channels_reward, hw, c->fc_reward, c->n_fc_reward, m->full) &&
       init_head(m, m->value, C, c->reduced_channels_value, hw, c->fc_value, c->n_fc_value, m->full) &&
       init_head(m, m->policy, C, c->reduced_channels_policy, hw, c->fc_policy, c->n_fc_policy, m->A);
  m->pv_w = (float*)dev_alloc(m, sizeof(float) * (size_t)(m->value.r + m->policy.r) * C);
  m->t16_counters = (int*)dev_alloc(m, 256);              // work / exit counters of k_recurrent16 (zero between launches)
  ok = ok && m->pv_w && m->t16_counters;
  if (!ok) {
    mzb_resnet_destroy(m);
    mzb_set_error("resnet create: allocation failed or head mlp wider than 128");
    return MZB_ECUDA;
  }
  m->n_tensors = count_tensors(m);
  *out = m;
  return MZB_OK;
}

int mzb_resnet_destroy(mzb_resnet_model* m) {
  if (!m) return MZB_OK;
  mzb_search_graph_forget(m);
  for (HeadParams* h : {&m->reward, &m->value, &m->policy}) if (h->mma) { mzb_head_mma_free(h->mma); h->mma = nullptr; }
  for (void* p : m->allocs) cudaFree(p);
  delete m;
  return MZB_OK;
}

int mzb_resnet_num_tensors(const mzb_resnet_model* m) { return m ? m->n_tensors : 0; }

int mzb_resnet_latent_dims(const mzb_resnet_model* m, int32_t* C, int32_t* H, int32_t* W) {
  MZB_CHECK_ARG(m, "model is NULL");
  if (C) *C = m->C;
  if (H) *H = m->Hl;
  if (W) *W = m->Wl;
  return MZB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ weights
namespace {

struct Cursor {
  const float* const* t; const int64_t* numel; int n; int i; bool bad;
  const float* take(int64_t expect) {
    if (i >= n || numel[i] != expect || !t[i]) { bad = true; ++i; return nullptr; }
    return t[i++];
  }
};

bool upload(void* dst, const void* src, size_t bytes) { return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess; }

// conv weight [cout][cin(+extra)][3][3] -> fp32 [tap][cin+extra][cout], bf16 [cout][tap][cin]; plane table.
// `c` may be wider than the source tensor (src_cin x src_cout): the extra channels get zero weights, scale and shift.
bool pack_conv(ConvParams& c, const float* w, int src_cin, int src_cout, const std::vector<float>& scale_src,
               const std::vector<float>& shift_src, int Hl, int Wl) {
  const int cw = c.cin + c.extra_plane, src_cw = src_cin + c.extra_plane;
  std::vector<float> scale(c.cout, 0.0f), shift(c.cout, 0.0f);
  for (int o = 0; o < src_cout; ++o) { scale[o] = scale_src[o]; shift[o] = shift_src[o]; }
  std::vector<float> wp((size_t)9 * cw * c.cout, 0.0f);
  std::vector<__nv_bfloat16> wb((size_t)9 * c.cin * c.cout, __float2bfloat16(0.0f)), wt(wb);
  for (int o = 0; o < src_cout; ++o)
    for (int ci = 0; ci < src_cw; ++ci)
      for (int tap = 0; tap < 9; ++tap) {
        const float v = w[((size_t)o * src_cw + ci) * 9 + tap];
        const int cd = ci < src_cin ? ci : c.cin;                  // the extra plane stays the last input channel
        wp[((size_t)tap * cw + cd) * c.cout + o] = v;
        if (ci < src_cin) {
          wb[((size_t)o * 9 + tap) * c.cin + ci] = __float2bfloat16(v);
          wt[((size_t)o * 9 + tap) * c.cin + ci] = __float2bfloat16(v * scale[o]);
        }
      }
  bool ok = upload(c.w, wp.data(), wp.size() * 4) && upload(c.w_bf16, wb.data(), wb.size() * 2) && upload(c.w_tc, wt.data(), wt.size() * 2) &&
            upload(c.scale, scale.data(), scale.size() * 4) && upload(c.shift, shift.data(), shift.size() * 4);
  if (c.extra_plane) {
    std::vector<float> tab((size_t)Hl * Wl * c.cout, 0.0f);
    for (int y = 0; y < Hl; ++y)
      for (int x = 0; x < Wl; ++x)
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx) {
            const int iy = y + ky - 1, ix = x + kx - 1;
            if (iy < 0 || iy >= Hl || ix < 0 || ix >= Wl) continue;
            for (int o = 0; o < src_cout; ++o) tab[((size_t)y * Wl + x) * c.cout + o] += w[((size_t)o * src_cw + src_cin) * 9 + ky * 3 + kx];
          }
    ok = ok && upload(c.plane_table, tab.data(), tab.size() * 4);
  }
  return ok;
}

// Pixel-pair form of a C/2 -> C/2 3x3 convolution (see k_conv_s2): `c` has 2*src channels on both sides; output
// channel po*src + co of pair tap kxp reads input channel pi*src + ci with the pixel tap dx = 2*(kxp-1) + pi - po + 1.
bool pack_conv_pair(ConvParams& c, const float* w, int src, const std::vector<float>& scale_src, const std::vector<float>& shift_src) {
  const int C2 = 2 * src;
  std::vector<float> scale(C2), shift(C2);
  for (int po = 0; po < 2; ++po)
    for (int o = 0; o < src; ++o) { scale[po * src + o] = scale_src[o]; shift[po * src + o] = shift_src[o]; }
  std::vector<float> wp((size_t)9 * C2 * C2, 0.0f);
  std::vector<__nv_bfloat16> wb((size_t)9 * C2 * C2, __float2bfloat16(0.0f)), wt(wb);
  for (int po = 0; po < 2; ++po)
    for (int pi = 0; pi < 2; ++pi)
      for (int kxp = 0; kxp < 3; ++kxp) {
        const int dx = 2 * (kxp - 1) + pi - po + 1;
        if (dx < 0 || dx > 2) continue;
        for (int ky = 0; ky < 3; ++ky)
          for (int o = 0; o < src; ++o)
            for (int ci = 0; ci < src; ++ci) {
              const float v = w[((size_t)o * src + ci) * 9 + ky * 3 + dx];
              const int tap = ky * 3 + kxp, oc = po * src + o, ic = pi * src + ci;
              wp[((size_t)tap * C2 + ic) * C2 + oc] = v;
              wb[((size_t)oc * 9 + tap) * C2 + ic] = __float2bfloat16(v);
              wt[((size_t)oc * 9 + tap) * C2 + ic] = __float2bfloat16(v * scale[oc]);
            }
      }
  return upload(c.w, wp.data(), wp.size() * 4) && upload(c.w_bf16, wb.data(), wb.size() * 2) && upload(c.w_tc, wt.data(), wt.size() * 2) &&
         upload(c.scale, scale.data(), scale.size() * 4) && upload(c.shift, shift.data(), shift.size() * 4);
}

// DownSample.conv2 on pixel-pair input rows (16 = 2 pixels x 8 channels) as six K = 16 taps: tap 2 ky + 0 reads the pair
// row LEFT of the output pixel (its second pixel is input column 2 ox - 1: kx = 0; its first pixel is not a tap), tap
// 2 ky + 1 the pair row AT the output pixel (columns 2 ox, 2 ox + 1: kx = 1, 2).  [6][cout][16] bf16, scale folded in.
bool pack_s2_mma(__nv_bfloat16* dst, const float* w, int cin, int cout, const std::vector<float>& scale) {
  std::vector<__nv_bfloat16> wb((size_t)6 * cout * 2 * cin, __float2bfloat16(0.0f));
  for (int ky = 0; ky < 3; ++ky)
    for (int o = 0; o < cout; ++o)
      for (int ci = 0; ci < cin; ++ci) {
        auto src = [&](int kx) { return w[((size_t)o * cin + ci) * 9 + ky * 3 + kx] * scale[o]; };
        wb[((size_t)(2 * ky) * cout + o) * 2 * cin + cin + ci] = __float2bfloat16(src(0));
        wb[((size_t)(2 * ky + 1) * cout + o) * 2 * cin + ci] = __float2bfloat16(src(1));
        wb[((size_t)(2 * ky + 1) * cout + o) * 2 * cin + cin + ci] = __float2bfloat16(src(2));
      }
  return upload(dst, wb.data(), wb.size() * 2);
}

bool load_conv(Cursor& cur, ConvParams& c, bool with_bn, int Hl, int Wl, ConvParams* wide = nullptr, __nv_bfloat16* s2_mma = nullptr) {
  const int cw = c.cin + c.extra_plane;
  const float* w = cur.take((int64_t)c.cout * cw * 9);
  std::vector<float> scale(c.cout, 1.0f), shift(c.cout, 0.0f);
  if (with_bn) {
    const float* g = cur.take(c.cout); const float* b = cur.take(c.cout);
    const float* mu = cur.take(c.cout); const float* var = cur.take(c.cout);
    if (cur.bad) return false;
    for (int o = 0; o < c.cout; ++o) {
      scale[o] = g[o] / sqrtf(var[o] + 1e-5f);
      shift[o] = b[o] - mu[o] * scale[o];
    }
  }
  if (cur.bad) return false;
  return pack_conv(c, w, c.cin, c.cout, scale, shift, Hl, Wl) && (!wide || pack_conv_pair(*wide, w, c.cin, scale, shift)) &&
         (!s2_mma || pack_s2_mma(s2_mma, w, c.cin, c.cout, scale));
}

bool load_block(Cursor& cur, Block& b) { return load_conv(cur, b.c1, true, 0, 0) && load_conv(cur, b.c2, true, 0, 0); }

bool load_1x1(Cursor& cur, HeadParams& h) {
  const float* w = cur.take((int64_t)h.r * h.cin); const float* b = cur.take(h.r);
  if (cur.bad) return false;
  return upload(h.w1x1, w, sizeof(float) * h.r * h.cin) && upload(h.b1x1, b, sizeof(float) * h.r);
}
bool load_fc(Cursor& cur, HeadParams& h, mzb_resnet_model* m) {
  std::vector<std::vector<float>> w_host, b_host;
  for (int l = 0; l < h.n_fc; ++l) {
    const float* w = cur.take((int64_t)h.fc_in[l] * h.fc_out[l]); const float* b = cur.take(h.fc_out[l]);
    if (cur.bad) return false;
    std::vector<float> wt((size_t)h.fc_in[l] * h.fc_out[l]);                    // torch [out][in] -> [in][out]
    for (int o = 0; o < h.fc_out[l]; ++o)
      for (int k = 0; k < h.fc_in[l]; ++k) wt[(size_t)k * h.fc_out[l] + o] = w[(size_t)o * h.fc_in[l] + k];
    if (!upload(h.fc_w[l], wt.data(), sizeof(float) * wt.size()) || !upload(h.fc_b[l], b, sizeof(float) * h.fc_out[l])) return false;
    w_host.push_back(std::move(wt));
    b_host.emplace_back(b, b + h.fc_out[l]);
  }
  if (m->precision == 1) {                                // bf16 path: the mlp also goes to the tensor-core head kernel
    void* pk = h.mma;
    if (mzb_head_mma_pack(m, h, w_host, b_host, &pk)) h.mma = pk; else { mzb_head_mma_free(pk); h.mma = nullptr; }
  }
  return true;
}

}  // namespace

extern "C" int mzb_resnet_set_weights(mzb_resnet_model* m, const float* const* h_tensors, const int64_t* h_numel,
                                      int n_tensors, void* stream) {
  MZB_CHECK_ARG(m && h_tensors && h_numel, "NULL argument");
  MZB_CHECK_ARG(n_tensors == m->n_tensors, "expected %d weight tensors (state_dict order, num_batches_tracked skipped), got %d",
                m->n_tensors, n_tensors);
  MZB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  Cursor cur{h_tensors, h_numel, n_tensors, 0, false};
  bool ok = true;
  // state_dict order = module registration order (models.py:300-429)
  if (m->downsample) {
    ok = ok && load_conv(cur, m->ds_conv1, false, 0, 0);
    for (size_t i = 0; i < m->ds1.size(); ++i) {
      Block* wide = m->stem_tc ? &m->ds1_tc[i] : nullptr;
      ok = ok && load_conv(cur, m->ds1[i].c1, true, 0, 0, wide ? &wide->c1 : nullptr) &&
           load_conv(cur, m->ds1[i].c2, true, 0, 0, wide ? &wide->c2 : nullptr);
    }
    ok = ok && load_conv(cur, m->ds_conv2, false, 0, 0, nullptr, m->ds_conv2_mma);
    for (auto& b : m->ds2) ok = ok && load_block(cur, b);
    for (auto& b : m->ds3) ok = ok && load_block(cur, b);
  }
  ok = ok && load_conv(cur, m->rep_conv, true, 0, 0);
  for (auto& b : m->rep_blocks) ok = ok && load_block(cur, b);
  ok = ok && load_conv(cur, m->dyn_conv, true, m->Hl, m->Wl);
  for (auto& b : m->dyn_blocks) ok = ok && load_block(cur, b);
  ok = ok && load_1x1(cur, m->reward) && load_fc(cur, m->reward, m);
  for (auto& b : m->pred_blocks) ok = ok && load_block(cur, b);
  ok = ok && load_1x1(cur, m->value) && load_1x1(cur, m->policy) && load_fc(cur, m->value, m) && load_fc(cur, m->policy, m);
  ok = ok && cudaMemcpy(m->pv_w, m->value.w1x1, sizeof(float) * m->value.r * m->C, cudaMemcpyDeviceToDevice) == cudaSuccess &&
       cudaMemcpy(m->pv_w + (size_t)m->value.r * m->C, m->policy.w1x1, sizeof(float) * m->policy.r * m->C, cudaMemcpyDeviceToDevice) == cudaSuccess;
  if (!ok || cur.bad || cur.i != n_tensors) {
    mzb_set_error("resnet set_weights: tensor %d has an unexpected size (or upload failed)", cur.i - 1);
    return MZB_EINVAL;
  }
  MZB_CUDA(cudaDeviceSynchronize());
  return MZB_OK;
}

// ------------------------------------------------------------------------------------------ forward
namespace {

struct Runner {
  mzb_resnet_model* m; int B; cudaStream_t s; uint8_t* ws; size_t ws_bytes; size_t act_bytes;
  int rc = MZB_OK;
  int zero_pads = 0;                 // stem mode: every layer re-writes its pad rows (see k_conv_s2)
  long long cap_B = 0;               // batch the workspace was sized for (fixes the offsets inside it)
  bool proj_fused = false;           // the last conv() computed the requested head projection in its epilogue
  float* proj() { return reinterpret_cast<float*>(ws + 3 * act_bytes + mzb_align_up((size_t)cap_B * 4, 256)); }
  template <class T> T* buf(int i) { return reinterpret_cast<T*>(ws + (size_t)i * act_bytes); }
  float* plane() { return reinterpret_cast<float*>(ws + 3 * act_bytes); }
};

template <class T>
void conv(Runner& r, const T* x, Geo gi, const ConvParams& cp, const float* plane, const T* res, int relu, Geo go, T* y,
          const TcProj* proj = nullptr) {
  r.proj_fused = false;
  if (r.rc) return;
  if (sizeof(T) == 2 && gi.pad && go.pad && mzb_conv_tc_supported(cp, gi.H, gi.W, gi.C)) {
    r.rc = mzb_conv_tc_launch(r.B, gi.H, gi.W, cp, (const __nv_bfloat16*)x, plane, (const __nv_bfloat16*)res, relu,
                              (__nv_bfloat16*)y, r.s, r.zero_pads, proj);
    r.proj_fused = proj != nullptr && proj->r > 0;
    return;
  }
  if (r.zero_pads) { mzb_set_error("bf16 stem: layer %dx%d C %d->%d is not supported by the tensor-core kernel", gi.H, gi.W, cp.cin, cp.cout); r.rc = MZB_EUNSUPPORTED; return; }
  const long long n = (long long)r.B * go.H * go.W * cp.cout;
  k_conv3x3_direct<T><<<nblk(n, 128), 128, 0, r.s>>>(x, r.B, gi, cp, plane, res, relu, go, y);
  mzb_count_launch();
}

// residual tower: x -> blocks; uses the three rotating buffers, returns the buffer index holding the result
// `proj`: head projection fused into the tower's last convolution (r.proj_fused tells whether it happened)
template <class T>
int tower(Runner& r, std::vector<Block>& blocks, Geo g, int cur, const TcProj* proj = nullptr) {
  bool fused = false;
  for (size_t i = 0; i < blocks.size(); ++i) {
    Block& b = blocks[i];
    const int t = (cur + 1) % 3, o = (cur + 2) % 3;
    conv<T>(r, r.buf<T>(cur), g, b.c1, nullptr, nullptr, 1, g, r.buf<T>(t));
    conv<T>(r, r.buf<T>(t), g, b.c2, nullptr, r.buf<T>(cur), 1, g, r.buf<T>(o), i + 1 == blocks.size() ? proj : nullptr);
    fused = r.proj_fused;
    cur = o;
  }
  r.proj_fused = fused;
  return cur;
}

template <class T>
void head(Runner& r, const T* x, Geo g, const HeadParams& hp, int mode, const uint8_t* legal, float* logits,
          float* scalar, float* priors, const float* proj = nullptr, long long proj_stride = 0, int proj_off = 0) {
  if (r.rc || (!logits && !scalar && !priors)) return;
  if (sizeof(T) == 2 && hp.mma && proj && mzb_conv_tc_enabled()) {     // bf16 path: mlp on the tensor cores
    r.rc = mzb_head_mma_launch(hp.mma, proj, proj_stride, proj_off, r.B, r.m->S, mode, legal, logits, scalar, priors, r.s);
    return;
  }
  const int warps = 8;
  const size_t smem = sizeof(float) * (size_t)head_smem(hp, warps).total;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(k_head<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_head<T, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_head<T, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    configured = true;
  }
  const int V = 16 / (int)sizeof(T), cpr = hp.cin / V;
  if (smem > 200 * 1024 || hp.r > 16 || hp.cin % V != 0 || (cpr & (cpr - 1)) != 0) {
    mzb_set_error("head shape unsupported (smem %zu bytes, r=%d, cin=%d)", smem, hp.r, hp.cin);
    r.rc = MZB_EUNSUPPORTED;
    return;
  }
  int grid = (r.B + warps - 1) / warps;
  if (grid > 148 * 4) grid = 148 * 4;                 // persistent: weights are staged once per block
  if (hp.r <= 2) k_head<T, 2><<<grid, warps * 32, smem, r.s>>>(x, r.B, g, hp, r.m->S, mode, legal, logits, scalar, priors, proj, proj_stride, proj_off);
  else if (hp.r <= 4) k_head<T, 4><<<grid, warps * 32, smem, r.s>>>(x, r.B, g, hp, r.m->S, mode, legal, logits, scalar, priors, proj, proj_stride, proj_off);
  else k_head<T, 16><<<grid, warps * 32, smem, r.s>>>(x, r.B, g, hp, r.m->S, mode, legal, logits, scalar, priors, proj, proj_stride, proj_off);
  mzb_count_launch();
}

struct Outputs {
  void* state; int layout; long long row_stride, off;
  float* value_logits; float* reward_logits; float* policy_logits; float* value; float* reward; float* priors;
};

// min-max scaling of buffer `cur` into buffer `nx` + the caller's state rows, one fused launch
template <class T>
void minmax_store(Runner& r, int cur, int nx, Geo g, const Outputs& o) {
  if (r.rc) return;
  if (g.C % 2 != 0 || g.C > 512) { mzb_set_error("min-max kernel needs an even channel count <= 512"); r.rc = MZB_EUNSUPPORTED; return; }
  const int grid = (r.B + 7) / 8;                   // one image per warp
  if (sizeof(T) == 2 && g.C % 8 == 0 && g.C <= 256 && ((g.C / 8) & (g.C / 8 - 1)) == 0) {
    const int nit = (g.H * g.W + 32 / (g.C / 8) - 1) / (32 / (g.C / 8));
    const __nv_bfloat16* x = (const __nv_bfloat16*)r.buf<T>(cur);
    __nv_bfloat16* y = (__nv_bfloat16*)r.buf<T>(nx);
    if (nit <= 4) { k_minmax_store_bf16<4><<<grid, 256, 0, r.s>>>(x, r.B, g, y, o.layout, o.state, o.row_stride, o.off); mzb_count_launch(); return; }
    if (nit <= 12) { k_minmax_store_bf16<12><<<grid, 256, 0, r.s>>>(x, r.B, g, y, o.layout, o.state, o.row_stride, o.off); mzb_count_launch(); return; }
    if (nit <= 20) { k_minmax_store_bf16<20><<<grid, 256, 0, r.s>>>(x, r.B, g, y, o.layout, o.state, o.row_stride, o.off); mzb_count_launch(); return; }
    const int lpr = g.C / 8, rpb = 256 / lpr;                       // block per image: rows per block iteration
    if (lpr <= 32 && (g.H * g.W + rpb - 1) / rpb <= 8) {
      k_minmax_store_bf16_block<8><<<r.B, 256, 0, r.s>>>(x, r.B, g, y, o.layout, o.state, o.row_stride, o.off);
      mzb_count_launch();
      return;
    }
  }
  if (g.C <= 64) k_minmax_store<T, 1><<<grid, 256, 0, r.s>>>(r.buf<T>(cur), r.B, g, r.buf<T>(nx), o.layout, o.state, o.row_stride, o.off);
  else if (g.C <= 128) k_minmax_store<T, 2><<<grid, 256, 0, r.s>>>(r.buf<T>(cur), r.B, g, r.buf<T>(nx), o.layout, o.state, o.row_stride, o.off);
  else k_minmax_store<T, 8><<<grid, 256, 0, r.s>>>(r.buf<T>(cur), r.B, g, r.buf<T>(nx), o.layout, o.state, o.row_stride, o.off);
  mzb_count_launch();
}

template <class T>
void prediction_and_state(Runner& r, int cur, Geo g, const Outputs& o, const uint8_t* legal) {
  mzb_resnet_model* m = r.m;
  if (o.value_logits || o.policy_logits || o.value || o.priors) {
    const int rv = m->value.r, rp = m->policy.r, hw = g.H * g.W;
    const TcProj pj{m->pv_w, r.proj(), rv + rp};
    // narrow layers (C < 64) are epilogue-bound: there the heads keep their own (cheap) 1x1 convolution
    const int p = tower<T>(r, m->pred_blocks, g, cur, (sizeof(T) == 2 && m->pv_w && rv + rp <= 8 && m->C >= 64) ? &pj : nullptr);
    const float* pr = r.proj_fused ? r.proj() : nullptr;
    if (sizeof(T) == 2 && pr && m->value.mma && m->policy.mma && mzb_conv_tc_enabled() && !r.rc) {   // both head mlps as one launch
      const MmaHeadCall calls[2] = {
          {m->value.mma, pr, (long long)(rv + rp) * hw, 0, 0, nullptr, o.value_logits, o.value, nullptr},
          {m->policy.mma, pr, (long long)(rv + rp) * hw, rv * hw, 1, legal, o.policy_logits, nullptr, o.priors}};
      r.rc = mzb_head_mma_launch_n(2, calls, r.B, m->S, r.s);
      return;
    }
    head<T>(r, r.buf<T>(p), g, m->value, 0, nullptr, o.value_logits, o.value, nullptr, pr, (long long)(rv + rp) * hw, 0);
    head<T>(r, r.buf<T>(p), g, m->policy, 1, legal, o.policy_logits, nullptr, o.priors, pr, (long long)(rv + rp) * hw, rv * hw);
  }
}

// the latent-resolution geometry: padded on the tensor-core path
template <class T> Geo latent_geo(const mzb_resnet_model* m) { return Geo{m->Hl, m->Wl, m->C, sizeof(T) == 2 ? 1 : 0}; }

// DownSample.forward (models.py:264-275) on the bf16 path: padded layouts throughout, see k_conv_s2.
// Returns the buffer holding the latent tensor in the padded latent geometry; all three buffers end with clean pads.
int stem_tc(Runner& r, const float* obs) {
  typedef __nv_bfloat16 T;
  mzb_resnet_model* m = r.m;
  const int B = r.B, C = m->C, Cp1 = m->stem_cp1;
  const Geo g0{m->H, m->W, m->Cobs, 0};
  const int h1 = (g0.H - 1) / 2 + 1, w1 = (g0.W - 1) / 2 + 1;
  const Geo g1{h1, w1 / 2, Cp1, 1};                         // pixel pairs: w1/2 rows of 2*(C/2) channels per line
  const Geo g2{(h1 - 1) / 2 + 1, (w1 - 1) / 2 + 1, C, 1};
  const Geo g3{(g2.H - 1) / 2 + 1, (g2.W - 1) / 2 + 1, C, 1};
  const Geo gl{m->Hl, m->Wl, C, 1};
  if ((g3.H - 1) / 2 + 1 != m->Hl || (g3.W - 1) / 2 + 1 != m->Wl) { mzb_set_error("latent size mismatch"); r.rc = MZB_EINVAL; return 0; }
  auto out_rows = [&](const Geo& g) { return (long long)B * geo_rows_per_image(g.H, g.W) + geo_halo(g.W); };
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(k_conv_s2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(k_conv_s2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(k_stem_conv1_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    configured = true;
  }
  const size_t sm1 = sizeof(float) * ((size_t)9 * m->Cobs * (C / 2) + C), sm2 = sizeof(float) * ((size_t)9 * (C / 2) * C + 2 * C);
  if (sm1 > 96 * 1024 || sm2 > 96 * 1024) { mzb_set_error("bf16 stem: stride-2 weights exceed 96 KiB of shared memory"); r.rc = MZB_EUNSUPPORTED; return 0; }
  r.zero_pads = 1;
  // The leading halo is never written by a layer, and the buffers change geometry (halo g1 >= g2 >= g3 >= latent): rows that
  // are halo at this call's first resolutions held image rows of the previous call's later ones, and the (-1, -1) tap of image
  // 0's first pixel reads the halo's last row.  Clear it once per call (found by the run-to-run determinism test).
  { const size_t lead = (size_t)std::max(geo_halo(g1.W) * g1.C, geo_halo(g2.W) * g2.C) * sizeof(T);
    for (int i = 0; i < 3; ++i) cudaMemsetAsync(r.buf<T>(i), 0, lead, r.s); }
  if (g0.W % 4 == 0 && (reinterpret_cast<uintptr_t>(obs) & 15) == 0)
    k_stem_conv1_pair<<<nblk(out_rows(g1), 128), 128, sm1, r.s>>>(obs, B, g0, m->Cobs, m->ds_conv1.w, m->ds_conv1.scale,
                                                                m->ds_conv1.shift, C / 2, g1, r.buf<T>(1));
  else
    k_conv_s2<true><<<nblk(2 * out_rows(g1), 128), 128, sm1, r.s>>>(obs, B, g0, m->Cobs, 0, m->ds_conv1.w, m->ds_conv1.scale,
                                                                  m->ds_conv1.shift, C / 2, g1, 1, r.buf<T>(1));
  mzb_count_launch();
  // a stage's blocks: one launch with the image resident in shared memory (mzb_stem16.cu, in place), else layer by layer
  auto stage = [&](std::vector<Block>& blocks, const Geo& g, int at) {
    if (r.rc || !mzb_stem16_supported(blocks, g.H, g.W, g.C)) return tower<T>(r, blocks, g, at);
    r.rc = mzb_stem16_tower(blocks, B, g.H, g.W, r.buf<T>(at), r.s);
    return at;
  };
  int cur = 1;
  if (!r.rc && m->ds_conv2_mma && mzb_stem16_supported(m->ds1_tc, g1.H, g1.W, g1.C) && g2.H * 2 == g1.H && g2.W == g1.W && C == 16) {
    // resblocks1 AND DownSample.conv2 on the image while it is resident in shared memory: its 604 MB (16,384 frames) are
    // neither written nor read back
    r.rc = mzb_stem16_tower(m->ds1_tc, B, g1.H, g1.W, r.buf<T>(1), r.s, m->ds_conv2_mma, m->ds_conv2.shift, r.buf<T>(2));
    cur = 2;
  } else {
    cur = stage(m->ds1_tc, g1, 1);
    const int nx = (cur + 1) % 3;
    k_conv_s2<false><<<nblk(out_rows(g2), 128), 128, sm2, r.s>>>(r.buf<T>(cur), B, g1, C / 2, 1, m->ds_conv2.w, m->ds_conv2.scale,
                                                               m->ds_conv2.shift, C, g2, 0, r.buf<T>(nx));
    mzb_count_launch(); cur = nx; }
  // (the two average pools as tails of the stage launches were measured: + 165 / + 29 us in the stage kernels against 139 / 36 us
  // for k_avgpool_pad - a serial, latency-bound phase per image group costs what the separate DRAM-bound launch costs)
  cur = stage(m->ds2, g2, cur);
  { const int nx = (cur + 1) % 3;
    k_avgpool_pad<<<nblk(out_rows(g3) * (C / 8), 256), 256, 0, r.s>>>(r.buf<T>(cur), B, g2, g3, r.buf<T>(nx));
    mzb_count_launch(); cur = nx; }
  cur = stage(m->ds3, g3, cur);
  { const int nx = (cur + 1) % 3;
    k_avgpool_pad<<<nblk(out_rows(gl) * (C / 8), 256), 256, 0, r.s>>>(r.buf<T>(cur), B, g3, gl, r.buf<T>(nx));
    mzb_count_launch(); cur = nx; }
  r.zero_pads = 0;
  // the latent-resolution layers never write pads: clear the latent region of the two other buffers
  const size_t latent_bytes = (size_t)geo_rows_total(gl, B) * C * sizeof(T);
  cudaMemsetAsync(r.buf<T>((cur + 1) % 3), 0, latent_bytes, r.s);
  cudaMemsetAsync(r.buf<T>((cur + 2) % 3), 0, latent_bytes, r.s);
  return cur;
}

template <class T>
int run_initial(Runner& r, const float* obs, const uint8_t* legal, const Outputs& o) {
  mzb_resnet_model* m = r.m;
  const int B = r.B, C = m->C;
  const Geo gl = latent_geo<T>(m);
  int cur = 0;
  if (m->downsample && sizeof(T) == 2 && m->stem_tc && mzb_conv_tc_enabled()) {
    cur = stem_tc(r, obs);
    if (r.rc) return r.rc;
  } else if (m->downsample) {                                        // DownSample.forward (models.py:264-275)
    Geo g0{m->H, m->W, m->Cobs, 0};
    {
      const long long n = (long long)B * g0.H * g0.W * g0.C;
      k_nchw_to_nhwc<T><<<nblk(n, 256), 256, 0, r.s>>>(obs, (long long)m->Cobs * g0.H * g0.W, nullptr, 0, B, g0, r.buf<T>(0));
      mzb_count_launch();
    }
    Geo g1{(g0.H - 1) / 2 + 1, (g0.W - 1) / 2 + 1, C / 2, 0};
    conv<T>(r, r.buf<T>(0), g0, m->ds_conv1, nullptr, nullptr, 0, g1, r.buf<T>(1));
    cur = tower<T>(r, m->ds1, g1, 1);
    Geo g2{(g1.H - 1) / 2 + 1, (g1.W - 1) / 2 + 1, C, 0};
    { const int nx = (cur + 1) % 3; conv<T>(r, r.buf<T>(cur), g1, m->ds_conv2, nullptr, nullptr, 0, g2, r.buf<T>(nx)); cur = nx; }
    cur = tower<T>(r, m->ds2, g2, cur);
    Geo g3{(g2.H - 1) / 2 + 1, (g2.W - 1) / 2 + 1, C, 0};
    { const int nx = (cur + 1) % 3;
      k_avgpool<T><<<nblk((long long)B * g3.H * g3.W * C, 256), 256, 0, r.s>>>(r.buf<T>(cur), B, g2, g3, r.buf<T>(nx));
      mzb_count_launch(); cur = nx; }
    cur = tower<T>(r, m->ds3, g3, cur);
    if ((g3.H - 1) / 2 + 1 != m->Hl || (g3.W - 1) / 2 + 1 != m->Wl) { mzb_set_error("latent size mismatch"); return MZB_EINVAL; }
    // the last pooling writes the latent tensor densely; the padded path then restores the zero pads the stem
    // overwrote (the buffers were used at other resolutions) and scatters the latent into the padded layout
    Geo gd{m->Hl, m->Wl, C, 0};
    { const int nx = (cur + 1) % 3;
      k_avgpool<T><<<nblk((long long)B * gd.H * gd.W * C, 256), 256, 0, r.s>>>(r.buf<T>(cur), B, g3, gd, r.buf<T>(nx));
      mzb_count_launch(); cur = nx; }
    if (gl.pad) {
      T* dense = r.buf<T>(cur);
      const int a = (cur + 1) % 3, b2 = (cur + 2) % 3;
      cudaMemsetAsync(r.buf<T>(a), 0, r.act_bytes, r.s);
      cudaMemsetAsync(r.buf<T>(b2), 0, r.act_bytes, r.s);
      const long long n = (long long)B * gd.H * gd.W * C;
      k_gather_nhwc<T><<<nblk(n / (16 / sizeof(T)), 256), 256, 0, r.s>>>(dense, (long long)gd.H * gd.W * C, nullptr, 0, B, gl, r.buf<T>(a));
      mzb_count_launch();
      cudaMemsetAsync(dense, 0, r.act_bytes, r.s);
      cur = a;
    }
  } else {
    // observation planes NCHW fp32 -> NHWC staged in buffer 0 (Cobs channels); on the padded path the staging area
    // is zeroed again afterwards because buffer 0 is reused with C channels and must keep zero pad rows
    Geo g0{m->H, m->W, m->Cobs, gl.pad};
    const long long n = (long long)B * g0.H * g0.W * g0.C;
    if (gl.pad) cudaMemsetAsync(r.buf<T>(0), 0, (size_t)geo_rows_total(g0, B) * g0.C * sizeof(T), r.s);
    k_nchw_to_nhwc<T><<<nblk(n, 256), 256, 0, r.s>>>(obs, (long long)m->Cobs * g0.H * g0.W, nullptr, 0, B, g0, r.buf<T>(0));
    mzb_count_launch();
    conv<T>(r, r.buf<T>(0), g0, m->rep_conv, nullptr, nullptr, 1, gl, r.buf<T>(1));
    if (gl.pad) cudaMemsetAsync(r.buf<T>(0), 0, (size_t)geo_rows_total(g0, B) * g0.C * sizeof(T), r.s);
    cur = 1;
  }
  cur = tower<T>(r, m->rep_blocks, gl, cur);
  const int nx = (cur + 1) % 3;
  minmax_store<T>(r, cur, nx, gl, o);
  if (o.reward_logits || o.reward) { k_zero_reward<<<nblk(B, 128), 128, 0, r.s>>>(B, m->S, o.reward_logits, o.reward); mzb_count_launch(); }
  prediction_and_state<T>(r, nx, gl, o, legal);
  return r.rc;
}

template <class T>
int run_recurrent(Runner& r, const void* state_in, int in_layout, long long in_row_stride, const int* in_slot,
                  long long slot_stride, const int* action, const Outputs& o) {
  mzb_resnet_model* m = r.m;
  const int B = r.B, C = m->C;
  const Geo gl = latent_geo<T>(m);
  if (sizeof(T) == 2 && mzb_tower16_supported(m, in_layout, o.layout)) {
    // narrow network (Breakout: 16 x 6 x 6): the whole recurrent inference is one warp-per-image kernel; only the
    // head mlps remain, fed by its projection rows
    const int hw = gl.H * gl.W, rr = m->reward.r, rvp = m->value.r + m->policy.r;
    float* pr = r.proj();
    float* pvp = pr + (size_t)r.cap_B * rr * hw;
    const bool want_reward = o.reward_logits || o.reward;
    const bool want_pred = o.value_logits || o.policy_logits || o.value || o.priors;
    r.rc = mzb_tower16_recurrent(m, B, state_in, in_layout, in_row_stride, in_slot, slot_stride, action, o.state, o.layout,
                                 o.row_stride, o.off, want_reward ? pr : nullptr, want_pred ? pvp : nullptr, r.s);
    if (r.rc) return r.rc;
    if (want_reward && want_pred && m->reward.mma && m->value.mma && m->policy.mma) {        // the three head mlps as one launch
      const MmaHeadCall calls[3] = {
          {m->reward.mma, pr, (long long)rr * hw, 0, 0, nullptr, o.reward_logits, o.reward, nullptr},
          {m->value.mma, pvp, (long long)rvp * hw, 0, 0, nullptr, o.value_logits, o.value, nullptr},
          {m->policy.mma, pvp, (long long)rvp * hw, m->value.r * hw, 1, nullptr, o.policy_logits, nullptr, o.priors}};
      return r.rc = mzb_head_mma_launch_n(3, calls, B, m->S, r.s);
    }
    head<T>(r, r.buf<T>(0), gl, m->reward, 0, nullptr, o.reward_logits, o.reward, nullptr, pr, (long long)rr * hw, 0);
    if (want_pred) {
      head<T>(r, r.buf<T>(0), gl, m->value, 0, nullptr, o.value_logits, o.value, nullptr, pvp, (long long)rvp * hw, 0);
      head<T>(r, r.buf<T>(0), gl, m->policy, 1, nullptr, o.policy_logits, nullptr, o.priors, pvp, (long long)rvp * hw, m->value.r * hw);
    }
    return r.rc;
  }
  const long long n = (long long)B * gl.H * gl.W * C;
  if (in_layout == 0) {
    k_nchw_to_nhwc<T><<<nblk(n, 256), 256, 0, r.s>>>((const float*)state_in, in_row_stride, in_slot, slot_stride, B, gl, r.buf<T>(0));
  } else if ((in_layout == 1 && sizeof(T) == 4) || (in_layout == 2 && sizeof(T) == 2)) {
    if ((C * sizeof(T)) % 16 != 0) { mzb_set_error("channels * element size must be a multiple of 16 bytes"); return MZB_EUNSUPPORTED; }
    k_gather_nhwc<T><<<nblk(n / (16 / sizeof(T)), 256), 256, 0, r.s>>>((const T*)state_in, in_row_stride, in_slot, slot_stride, B, gl, r.buf<T>(0));
  } else {
    mzb_set_error("state layout %d does not match the model precision", in_layout);
    return MZB_EINVAL;
  }
  mzb_count_launch();
  k_action_plane<<<nblk(B, 128), 128, 0, r.s>>>(action, B, m->A, r.plane());
  mzb_count_launch();
  conv<T>(r, r.buf<T>(0), gl, m->dyn_conv, r.plane(), nullptr, 1, gl, r.buf<T>(1));           // DynamicsNetwork.forward :377-387
  const bool want_reward = o.reward_logits || o.reward;
  const TcProj pj{m->reward.w1x1, r.proj(), m->reward.r};
  int cur = tower<T>(r, m->dyn_blocks, gl, 1, (sizeof(T) == 2 && want_reward && m->reward.r <= 8 && m->C >= 64) ? &pj : nullptr);
  head<T>(r, r.buf<T>(cur), gl, m->reward, 0, nullptr, o.reward_logits, o.reward, nullptr,    // reward on the un-normalised state
          r.proj_fused ? r.proj() : nullptr, (long long)m->reward.r * gl.H * gl.W, 0);
  const int nx = (cur + 1) % 3;
  minmax_store<T>(r, cur, nx, gl, o);
  prediction_and_state<T>(r, nx, gl, o, nullptr);
  return r.rc;
}

size_t act_bytes_for(const mzb_resnet_model* m, long long B) {
  long long per = (long long)m->H * m->W * m->Cobs;
  if (m->downsample) {
    const long long h1 = (m->H - 1) / 2 + 1, w1 = (m->W - 1) / 2 + 1;
    per = std::max(per, h1 * w1 * (long long)(m->C / 2));
    const long long h2 = (h1 - 1) / 2 + 1, w2 = (w1 - 1) / 2 + 1;
    per = std::max(per, h2 * w2 * (long long)m->C);
  } else {
    per = std::max(per, (long long)(m->H + 1) * (m->W + 2) * m->Cobs);
  }
  per = std::max(per, (long long)(m->Hl + 1) * (m->Wl + 2) * m->C);
  long long halo = 2ll * (m->Wl + 3) * std::max(m->C, m->Cobs) + 2ll * (m->W + 3) * m->Cobs;
  if (m->stem_tc) {                                        // padded bf16 stem stages (2 bytes per element: half of `per`)
    const long long h1 = (m->H - 1) / 2 + 1, w1 = (m->W - 1) / 2 + 1, h2 = (h1 - 1) / 2 + 1, w2 = (w1 - 1) / 2 + 1;
    per = std::max(per, ((h1 + 1) * (w1 / 2 + 2) * m->stem_cp1 + 1) / 2);
    per = std::max(per, ((h2 + 1) * (w2 + 2) * m->C + 1) / 2);
    halo += (w1 + 3) * (long long)std::max(m->stem_cp1, m->C);
  }
  return mzb_align_up((size_t)((per * B + halo) * 4 + 4096), 1024);      // sized for fp32; bf16 uses half
}

size_t ws_bytes_for(const mzb_resnet_model* m, long long B) {
  // projection rows of the heads' 1x1 convolutions: reward AND value|policy regions (the one-kernel recurrent
  // inference of narrow networks fills both before any head runs)
  const size_t proj_rows = (size_t)(m->reward.r + m->value.r + m->policy.r) * m->Hl * m->Wl;
  return 3 * act_bytes_for(m, B) + mzb_align_up((size_t)B * 4, 256) + mzb_align_up((size_t)B * proj_rows * 4, 256) + 1024;
}

// The buffer offsets inside a workspace must not depend on the batch of the call (the zero pad rows of the padded
// layout sit at fixed places): derive them from the workspace CAPACITY = the largest batch it was sized for.
size_t act_bytes_of_workspace(const mzb_resnet_model* m, size_t workspace_bytes, long long* cap_B) {
  long long lo = 1, hi = 1ll << 24;
  while (lo < hi) {
    const long long mid = (lo + hi + 1) / 2;
    if (ws_bytes_for(m, mid) <= workspace_bytes) lo = mid; else hi = mid - 1;
  }
  *cap_B = lo;
  return act_bytes_for(m, lo);
}

}  // namespace

extern "C" {

size_t mzb_resnet_workspace_bytes(const mzb_resnet_model* m, int64_t max_batch) {
  if (!m || max_batch <= 0) return 0;
  return ws_bytes_for(m, max_batch);
}

/* Zero a freshly allocated workspace: the padded activation layout relies on zero pad rows that no kernel writes. */
int mzb_resnet_workspace_init(const mzb_resnet_model* m, void* d_workspace, size_t workspace_bytes, void* stream) {
  MZB_CHECK_ARG(m && d_workspace, "NULL argument");
  MZB_CUDA(cudaMemsetAsync(d_workspace, 0, workspace_bytes, (cudaStream_t)stream));
  return MZB_OK;
}

int mzb_resnet_initial(mzb_resnet_model* m, int64_t B, const float* d_obs, const uint8_t* d_legal, void* d_workspace,
                       size_t workspace_bytes, void* d_state_out, int state_layout, int64_t out_row_stride,
                       int64_t out_offset, float* d_value_logits, float* d_reward_logits, float* d_policy_logits,
                       float* d_value, float* d_reward, float* d_priors, void* stream) {
  MZB_CHECK_ARG(m && d_obs && d_workspace, "NULL argument");
  MZB_CHECK_ARG(B > 0 && B < (1ll << 24), "batch out of range");
  MZB_CHECK_ARG(workspace_bytes >= mzb_resnet_workspace_bytes(m, B), "workspace too small");
  MZB_CHECK_ARG(state_layout >= 0 && state_layout <= 2, "bad state layout");
  long long cap_B = 0;
  const size_t act_bytes = act_bytes_of_workspace(m, workspace_bytes, &cap_B);
  Runner r{m, (int)B, (cudaStream_t)stream, (uint8_t*)d_workspace, workspace_bytes, act_bytes};
  r.cap_B = cap_B;
  Outputs o{d_state_out, state_layout, out_row_stride, out_offset, d_value_logits, d_reward_logits, d_policy_logits,
            d_value, d_reward, d_priors};
  int rc = m->precision == 1 ? run_initial<__nv_bfloat16>(r, d_obs, d_legal, o) : run_initial<float>(r, d_obs, d_legal, o);
  if (rc) return rc;
  MZB_CUDA(cudaGetLastError());
  return MZB_OK;
}

int mzb_resnet_recurrent(mzb_resnet_model* m, int64_t B, const void* d_state_in, int in_layout, int64_t in_row_stride,
                         const int32_t* d_in_slot, int64_t slot_stride, const int32_t* d_action, void* d_workspace,
                         size_t workspace_bytes, void* d_state_out, int out_layout, int64_t out_row_stride,
                         int64_t out_offset, float* d_value_logits, float* d_reward_logits, float* d_policy_logits,
                         float* d_value, float* d_reward, float* d_priors, void* stream) {
  MZB_CHECK_ARG(m && d_state_in && d_action && d_workspace, "NULL argument");
  MZB_CHECK_ARG(B > 0 && B < (1ll << 24), "batch out of range");
  MZB_CHECK_ARG(workspace_bytes >= mzb_resnet_workspace_bytes(m, B), "workspace too small");
  MZB_CHECK_ARG(in_layout >= 0 && in_layout <= 2 && out_layout >= 0 && out_layout <= 2, "bad state layout");
  long long cap_B = 0;
  const size_t act_bytes = act_bytes_of_workspace(m, workspace_bytes, &cap_B);
  Runner r{m, (int)B, (cudaStream_t)stream, (uint8_t*)d_workspace, workspace_bytes, act_bytes};
  r.cap_B = cap_B;
  Outputs o{d_state_out, out_layout, out_row_stride, out_offset, d_value_logits, d_reward_logits, d_policy_logits,
            d_value, d_reward, d_priors};
  int rc = m->precision == 1
               ? run_recurrent<__nv_bfloat16>(r, d_state_in, in_layout, in_row_stride, d_in_slot, slot_stride, d_action, o)
               : run_recurrent<float>(r, d_state_in, in_layout, in_row_stride, d_in_slot, slot_stride, d_action, o);
  if (rc) return rc;
  MZB_CUDA(cudaGetLastError());
  return MZB_OK;
}

/* Measurement hook (bench.py roofline line): the first residual-block convolution of the representation tower
 * (C -> C channels, latent resolution, batch B, + folded batch-norm + ReLU) launched `iters` times on the workspace's
 * activation buffers - the dominant kernel of a resnet search, timed alone. */
int mzb_resnet_conv_probe(mzb_resnet_model* m, int64_t B, void* d_workspace, size_t workspace_bytes, int32_t iters,
                          void* stream) {
  MZB_CHECK_ARG(m && d_workspace && iters > 0, "NULL argument");
  MZB_CHECK_ARG(!m->rep_blocks.empty(), "model has no residual blocks");
  MZB_CHECK_ARG(B > 0 && workspace_bytes >= mzb_resnet_workspace_bytes(m, B), "workspace too small");
  long long cap_B = 0;
  const size_t act_bytes = act_bytes_of_workspace(m, workspace_bytes, &cap_B);
  Runner r{m, (int)B, (cudaStream_t)stream, (uint8_t*)d_workspace, workspace_bytes, act_bytes};
  r.cap_B = cap_B;
  for (int i = 0; i < iters && !r.rc; ++i) {
    if (m->precision == 1) {
      const Geo g = latent_geo<__nv_bfloat16>(m);
      conv<__nv_bfloat16>(r, r.buf<__nv_bfloat16>(i & 1), g, m->rep_blocks[0].c1, nullptr, nullptr, 1, g, r.buf<__nv_bfloat16>(2));
    } else {
      const Geo g = latent_geo<float>(m);
      conv<float>(r, r.buf<float>(i & 1), g, m->rep_blocks[0].c1, nullptr, nullptr, 1, g, r.buf<float>(2));
    }
  }
  return r.rc;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ debug / self-test
// One 3x3 convolution (no batch-norm, no ReLU) through both bf16 kernels on host data: dense NHWC float input
// [B,H,W,Cin], torch-layout weights [Cout][Cin][3][3]; outputs dense NHWC float.  Used by tests to localise
// descriptor / swizzle errors of the tensor-core kernel.
extern "C" int mzb_debug_conv3x3(int B, int H, int W, int Cin, int Cout, const float* h_x, const float* h_w,
                                 float* h_y_direct, float* h_y_tc) {
  mzb_resnet_model m{};
  ConvParams cp{};
  if (!init_conv(&m, cp, Cin, Cout, 1, 0, 0)) return MZB_ECUDA;
  std::vector<float> ones(Cout, 1.0f), zeros(Cout, 0.0f);
  const float* tensors[5] = {h_w, ones.data(), zeros.data(), zeros.data(), ones.data()};
  const int64_t numel[5] = {(int64_t)Cout * Cin * 9, Cout, Cout, Cout, Cout};
  Cursor cur{tensors, numel, 1, 0, false};
  if (!load_conv(cur, cp, false, 0, 0)) return MZB_EINVAL;
  const Geo g{H, W, Cin, 1}, go{H, W, Cout, 1};
  const size_t in_elems = (size_t)geo_rows_total(g, B) * Cin, out_elems = (size_t)geo_rows_total(go, B) * Cout;
  __nv_bfloat16 *dx = nullptr, *dy = nullptr;
  cudaMalloc(&dx, in_elems * 2); cudaMalloc(&dy, out_elems * 2);
  cudaMemset(dx, 0, in_elems * 2);
  std::vector<__nv_bfloat16> hx(in_elems, __float2bfloat16(0.0f));
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        for (int c = 0; c < Cin; ++c)
          hx[(size_t)geo_row(g, b, y, x) * Cin + c] = __float2bfloat16(h_x[(((size_t)b * H + y) * W + x) * Cin + c]);
  cudaMemcpy(dx, hx.data(), in_elems * 2, cudaMemcpyHostToDevice);
  std::vector<__nv_bfloat16> hy(out_elems);
  int rc = MZB_OK;
  for (int pass = 0; pass < 2 && rc == MZB_OK; ++pass) {
    cudaMemset(dy, 0, out_elems * 2);
    float* dst = pass == 0 ? h_y_direct : h_y_tc;
    if (pass == 0) {
      const long long n = (long long)B * H * W * Cout;
      k_conv3x3_direct<__nv_bfloat16><<<nblk(n, 128), 128>>>(dx, B, g, cp, nullptr, nullptr, 0, go, dy);
    } else {
      if (!mzb_conv_tc_supported(cp, H, W, Cin)) { mzb_set_error("shape not supported by the tensor-core kernel"); rc = MZB_EUNSUPPORTED; break; }
      rc = mzb_conv_tc_launch(B, H, W, cp, dx, nullptr, nullptr, 0, dy, 0);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { mzb_set_error("debug conv: %s", cudaGetErrorString(cudaGetLastError())); rc = MZB_ECUDA; break; }
    cudaMemcpy(hy.data(), dy, out_elems * 2, cudaMemcpyDeviceToHost);
    for (int b = 0; b < B; ++b)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
          for (int c = 0; c < Cout; ++c)
            dst[(((size_t)b * H + y) * W + x) * Cout + c] = __bfloat162float(hy[(size_t)geo_row(go, b, y, x) * Cout + c]);
  }
  cudaFree(dx); cudaFree(dy);
  for (void* p : m.allocs) cudaFree(p);
  return rc;
}
