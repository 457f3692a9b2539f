// Head mlp() on the tensor cores (bf16 path): the fc layers that follow the heads' 1x1 convolutions
// (models.py:398-404 reward, :447-456 value / policy; mlp :626-638) for 16 images per warp with
// mma.sync.m16n8k16 (bf16 operands, fp32 accumulate), then support_to_scalar / the legal-action softmax.
//
// Why not the warp-per-image kernel (k_head): ncu showed it issue-bound - 1,500 instructions per image for a
// 84 -> 64 -> 21 mlp, i.e. ~7 instructions per multiply-accumulate-lane; one m16n8k16 does 2,048 of them.  The
// GEMMs are far too small for tcgen05 (M = 16 images per warp, N <= 128), which is what the legacy warp-level MMA
// is still good for.
//
// Data flow per warp: A fragments of the first layer come straight from the fp32 projection buffer
// [B][r*hw] (+ the 1x1 bias per channel), later layers re-pack the previous accumulators (the m16n8 C layout of two
// adjacent column tiles IS the m16k16 A layout: no shuffles); weights are staged once per block in shared memory as
// bf16 [N][K_pad] with K_pad = 8 (mod 16) so the B-fragment loads are bank-conflict free; the last layer's logits
// go through a per-warp shared tile to one lane per image for the decode.
#include <cuda_bf16.h>

#include "mzb_fc.cuh"
#include "mzb_resnet_model.h"

namespace {

struct MmaLayer { int K, N, Kp, Np, w_off, b_off; };      // Kp: multiple of 16 (+8 row stride), Np: multiple of 8/16
struct MmaHead {
  int n_fc, r, hw, out;
  MmaLayer l[4];
  const __nv_bfloat16* w[4];      // [Np][Ks] bf16, Ks = Kp + 8
  const float* b[4];              // [N]
  const float* b1x1;              // [r]
  int smem_w_bytes, logit_stride, b0k_off;
};

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&p);
}

// One head's work: blockIdx.y selects it, so the reward / value / policy heads of one inference are ONE launch (each was a
// ~10 us latency-bound launch: 30 us of a 230 us Breakout simulation).
struct HeadJob {
  const float* proj; long long proj_stride; int proj_off; int mode;
  const uint8_t* legal; float* logits; float* scalar; float* priors;
  MmaHead hp;
};
constexpr int kMaxJobs = 3;
struct HeadJobs { HeadJob j[kMaxJobs]; };

// MAXN: widest layer output (64 or 128): bounds the accumulator / fragment register arrays
// MINB = 3 (<= 85 registers, 8 bytes of spills): three resident CTAs per SM.  The three heads of 16,384 images are 384 CTAs, which
// at two per SM (96 registers) ran as 1.3 waves - the merged launch took as long as two of the separate ones (ncu:
// launch__waves_per_multiprocessor 1.30).  Launches that fit one wave at two CTAs per SM keep the 96-register form.
template <int MAXN, int MINB = 0>
__global__ void __launch_bounds__(256, MINB) k_head_mma(const __grid_constant__ HeadJobs jobs, int B, int S) {
  extern __shared__ __align__(16) uint8_t smem[];
  const HeadJob& job = jobs.j[blockIdx.y];
  const MmaHead& hp = job.hp;
  const float* __restrict__ proj = job.proj;
  const long long proj_stride = job.proj_stride;
  const int proj_off = job.proj_off, mode = job.mode;
  const uint8_t* __restrict__ legal = job.legal;
  float* __restrict__ logits_out = job.logits;
  float* __restrict__ scalar_out = job.scalar;
  float* __restrict__ priors_out = job.priors;
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- stage weights (bf16, padded rows) and biases
  {
    const uint4* src;
    for (int l = 0; l < hp.n_fc; ++l) {
      const int bytes = hp.l[l].Np * (hp.l[l].Kp + 8) * 2;
      src = reinterpret_cast<const uint4*>(hp.w[l]);
      uint4* dst = reinterpret_cast<uint4*>(smem + hp.l[l].w_off);
      for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) dst[i] = src[i];
      float* bd = reinterpret_cast<float*>(smem + hp.l[l].b_off);
      for (int i = threadIdx.x; i < hp.l[l].Np; i += blockDim.x) bd[i] = i < hp.l[l].N ? hp.b[l][i] : 0.0f;
    }
  }
  float* b0k = reinterpret_cast<float*>(smem + hp.b0k_off);                  // 1x1 bias of the channel of flat index k
  for (int k = threadIdx.x; k < hp.l[0].Kp; k += blockDim.x) b0k[k] = k < hp.l[0].K ? hp.b1x1[k / hp.hw] : 0.0f;
  __syncthreads();
  float* tile = reinterpret_cast<float*>(smem + hp.smem_w_bytes) + (size_t)warp * 16 * hp.logit_stride;
  const int g = lane >> 2, t = lane & 3;
  const int K0 = hp.l[0].K;
  for (int img0 = (blockIdx.x * warps + warp) * 16; img0 < B; img0 += gridDim.x * warps * 16) {
    const int r0 = img0 + g, r1 = img0 + g + 8;
    const float* p0 = proj + (long long)(r0 < B ? r0 : B - 1) * proj_stride + proj_off;
    const float* p1 = proj + (long long)(r1 < B ? r1 : B - 1) * proj_stride + proj_off;
    uint32_t afrag[MAXN / 16][4];
    float acc[MAXN / 8][4];
    for (int l = 0; l < hp.n_fc; ++l) {
      const MmaLayer L = hp.l[l];
      const int nt = L.Np / 8, ks = L.Kp / 16, Ks = L.Kp + 8;
      const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(smem + L.w_off);
      const float* bias = reinterpret_cast<const float*>(smem + L.b_off);
#pragma unroll
      for (int j = 0; j < MAXN / 8; ++j) {
        if (j < nt) { acc[j][0] = acc[j][2] = bias[8 * j + 2 * t]; acc[j][1] = acc[j][3] = bias[8 * j + 2 * t + 1]; }
      }
      if (l == 0) {
        // first layer: K can be large (r*hw up to 484) - A fragments are formed on the fly from the projection rows
        for (int s = 0; s < ks; ++s) {
          uint32_t a[4];
#pragma unroll
          for (int h = 0; h < 2; ++h) {                     // h: k offset 0 / 8 inside the 16-wide step
            const int k = 16 * s + 8 * h + 2 * t;
            float x00 = 0.0f, x01 = 0.0f, x10 = 0.0f, x11 = 0.0f;
            if (k < K0) { x00 = p0[k] + b0k[k]; x10 = p1[k] + b0k[k]; }
            if (k + 1 < K0) { x01 = p0[k + 1] + b0k[k + 1]; x11 = p1[k + 1] + b0k[k + 1]; }
            a[2 * h] = pack_bf16(x00, x01);
            a[2 * h + 1] = pack_bf16(x10, x11);
          }
#pragma unroll
          for (int j = 0; j < MAXN / 8; ++j) {
            if (j < nt) {
              const uint32_t* wr = reinterpret_cast<const uint32_t*>(W + (size_t)(8 * j + g) * Ks + 16 * s + 2 * t);
              mma16816(acc[j], a, wr[0], wr[4]);
            }
          }
        }
      } else {
#pragma unroll
        for (int s = 0; s < MAXN / 16; ++s) {
          if (s < ks) {
#pragma unroll
            for (int j = 0; j < MAXN / 8; ++j) {
              if (j < nt) {
                const uint32_t* wr = reinterpret_cast<const uint32_t*>(W + (size_t)(8 * j + g) * Ks + 16 * s + 2 * t);
                mma16816(acc[j], afrag[s], wr[0], wr[4]);
              }
            }
          }
        }
      }
      if (l + 1 < hp.n_fc) {
        // ELU, then the accumulators of column tiles (2s, 2s+1) are the A fragment of k-step s of the next layer
#pragma unroll
        for (int s = 0; s < MAXN / 16; ++s) {
          if (2 * s < nt) {
            float e[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) { e[i] = elu_f32(acc[2 * s][i]); e[4 + i] = (2 * s + 1 < nt) ? elu_f32(acc[2 * s + 1][i]) : 0.0f; }
            afrag[s][0] = pack_bf16(e[0], e[1]); afrag[s][1] = pack_bf16(e[2], e[3]);
            afrag[s][2] = pack_bf16(e[4], e[5]); afrag[s][3] = pack_bf16(e[6], e[7]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < MAXN / 8; ++j) {
          if (j < nt) {
            const int c = 8 * j + 2 * t;
            tile[g * hp.logit_stride + c] = acc[j][0]; tile[g * hp.logit_stride + c + 1] = acc[j][1];
            tile[(g + 8) * hp.logit_stride + c] = acc[j][2]; tile[(g + 8) * hp.logit_stride + c + 1] = acc[j][3];
          }
        }
      }
    }
    __syncwarp();
    // ---- decode: two lanes per image (halves of the logits), combined with one shuffle
    const int img = img0 + (lane >> 1), half = lane & 1, n = hp.out;
    const float* in = tile + (lane >> 1) * hp.logit_stride;
    const bool ok = img < B;
    if (logits_out && ok) for (int o = half; o < n; o += 2) logits_out[(long long)img * n + o] = in[o];
    if (mode == 0 && scalar_out) {
      float m = -CUDART_INF_F;
      for (int i = half; i < n; i += 2) m = fmaxf(m, in[i]);
      m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, 1));
      float se = 0.0f, sx = 0.0f;
      for (int i = half; i < n; i += 2) { const float e = softmax_exp(in[i], m); se += e; sx = fmaf((float)(i - S), e, sx); }
      se += __shfl_xor_sync(0xFFFFFFFFu, se, 1); sx += __shfl_xor_sync(0xFFFFFFFFu, sx, 1);
      if (ok && half == 0) scalar_out[img] = inverse_value_transform(__fdiv_rn(sx, se));
    } else if (mode == 1 && priors_out) {
      const uint8_t* lg = (legal && ok) ? legal + (long long)img * n : nullptr;
      float m = -CUDART_INF_F;
      for (int a = half; a < n; a += 2) if (!lg || lg[a]) m = fmaxf(m, in[a]);
      m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, 1));
      float se = 0.0f;
      for (int a = half; a < n; a += 2) if (!lg || lg[a]) se += softmax_exp(in[a], m);
      se += __shfl_xor_sync(0xFFFFFFFFu, se, 1);
      if (ok) for (int a = half; a < n; a += 2) priors_out[(long long)img * n + a] = (!lg || lg[a]) ? __fdiv_rn(softmax_exp(in[a], m), se) : 0.0f;
    }
    __syncwarp();
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------ host side
struct MmaHeadPack {
  MmaHead h;
  bool ok = false;
};

static int roundup(int x, int m) { return (x + m - 1) / m * m; }

// Packs the fc weights of `hp` (device fp32 [in][out], see load_fc) as bf16 [Np][Kp + 8]; returns false when the shape
// is outside the kernel's limits (then the caller keeps k_head).
bool mzb_head_mma_pack(mzb_resnet_model* m, const HeadParams& hp, const std::vector<std::vector<float>>& w_host,
                       const std::vector<std::vector<float>>& b_host, void** opaque) {
  (void)b_host;
  // a pack made by an earlier set_weights is reused (same shapes: the layer buffers are re-uploaded in place)
  MmaHeadPack* prev = static_cast<MmaHeadPack*>(*opaque);
  auto* pk = prev ? prev : new MmaHeadPack();
  MmaHead& h = pk->h;
  h.n_fc = hp.n_fc; h.r = hp.r; h.hw = hp.hw; h.out = hp.out; h.b1x1 = hp.b1x1;
  int off = 0;
  bool ok = hp.n_fc >= 1 && hp.n_fc <= 4 && hp.out <= 128;
  for (int l = 0; l < hp.n_fc && ok; ++l) {
    MmaLayer& L = h.l[l];
    L.K = hp.fc_in[l]; L.N = hp.fc_out[l];
    L.Kp = roundup(L.K, 16);
    L.Np = l + 1 < hp.n_fc ? roundup(L.N, 16) : roundup(L.N, 8);
    if (L.Np > 128 || (l > 0 && L.Kp > 128)) { ok = false; break; }
    if (l > 0 && L.Kp != h.l[l - 1].Np) { ok = false; break; }
    const int Ks = L.Kp + 8;
    std::vector<__nv_bfloat16> wb((size_t)L.Np * Ks, __float2bfloat16(0.0f));
    for (int n = 0; n < L.N; ++n)
      for (int k = 0; k < L.K; ++k) wb[(size_t)n * Ks + k] = __float2bfloat16(w_host[l][(size_t)k * L.N + n]);   // w_host is [in][out]
    void* d = prev ? (void*)h.w[l] : nullptr;
    if (!d) {
      if (cudaMalloc(&d, wb.size() * 2) != cudaSuccess) { ok = false; break; }
      m->allocs.push_back(d);
    }
    cudaMemcpy(d, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice);
    h.w[l] = (const __nv_bfloat16*)d;
    h.b[l] = hp.fc_b[l];
    L.w_off = off; off += roundup(L.Np * Ks * 2, 16);
    L.b_off = off; off += roundup(L.Np * 4, 16);
  }
  h.b0k_off = off; off += roundup(h.l[0].Kp * 4, 16);
  h.smem_w_bytes = off;
  h.logit_stride = roundup(hp.out, 8) + 4;
  pk->ok = ok && (size_t)off + 8 * 16 * h.logit_stride * 4 <= 200 * 1024;
  *opaque = pk;
  return pk->ok;
}

void mzb_head_mma_free(void* opaque) { delete static_cast<MmaHeadPack*>(opaque); }

int mzb_head_mma_launch_n(int n, const MmaHeadCall* calls, int B, int S, cudaStream_t stream) {
  if (n < 1 || n > kMaxJobs) { mzb_set_error("tensor-core head: %d jobs in one launch", n); return MZB_EINVAL; }
  HeadJobs jobs{};
  int maxn = 0;
  size_t smem = 0;
  for (int i = 0; i < n; ++i) {
    auto* pk = static_cast<MmaHeadPack*>(calls[i].opaque);
    if (!pk || !pk->ok) { mzb_set_error("tensor-core head: shape not packed"); return MZB_EUNSUPPORTED; }
    const MmaHead& h = pk->h;
    for (int l = 0; l < h.n_fc; ++l) maxn = maxn > h.l[l].Np ? maxn : h.l[l].Np;
    smem = std::max(smem, (size_t)h.smem_w_bytes + (size_t)8 * 16 * h.logit_stride * 4);
    jobs.j[i] = HeadJob{calls[i].proj, calls[i].proj_stride, calls[i].proj_off, calls[i].mode, calls[i].legal,
                        calls[i].logits, calls[i].scalar, calls[i].priors, h};
  }
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(k_head_mma<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_head_mma<64, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_head_mma<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    configured = true;
  }
  int gx = (B + 8 * 16 - 1) / (8 * 16);
  if (gx > 148 * 2) gx = 148 * 2;
  const dim3 grid(gx, n);
  if (maxn <= 64 && gx * n > 148 * 2 && smem * 3 <= 200 * 1024) k_head_mma<64, 3><<<grid, 256, smem, stream>>>(jobs, B, S);
  else if (maxn <= 64) k_head_mma<64><<<grid, 256, smem, stream>>>(jobs, B, S);
  else k_head_mma<128><<<grid, 256, smem, stream>>>(jobs, B, S);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

int mzb_head_mma_launch(void* opaque, const float* proj, long long proj_stride, int proj_off, int B, int S, int mode,
                        const uint8_t* legal, float* logits, float* scalar, float* priors, cudaStream_t stream) {
  const MmaHeadCall c{opaque, proj, proj_stride, proj_off, mode, legal, logits, scalar, priors};
  return mzb_head_mma_launch_n(1, &c, B, S, stream);
}
