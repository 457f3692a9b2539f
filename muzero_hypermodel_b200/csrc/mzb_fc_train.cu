// Training step of the fully-connected MuZero network as ONE kernel: the unrolled forward
// (initial inference + K recurrent steps), the categorical cross-entropies, the backward pass through time and the
// reduction of the batch's gradients into the trainer's flat gradient bucket.
// Reference: trainer.py:124-255 (update_weights: unroll, per-step losses, gradient hooks, PER weights, priorities),
// :267-284 (loss_function), models.py:128-195 (network), :138-145 (min-max scaling), :626-638 (mlp: Linear + ELU).
//
// Mapping.  The unrolled graph of a batch is B independent chains of K+1 tiny dense layers (cartpole: 128 samples x
// 11 steps x ~1.4 k MACs): as library launches it is several hundred kernels per step and launch-bound (3.7 ms per step
// even replayed from a CUDA graph).  Here ONE WARP owns one sample: lanes = output units in the forward pass and in
// the weight-gradient update, lanes = input units for the data gradient.  Per CTA the weights are staged once in shared
// memory, TRANSPOSED with an odd row length ([in + 1][out | 1], the last row is the bias), so that "lane = output"
// (forward, weight gradients) and "lane = input" (data gradients) are both conflict-free.  Every warp keeps the
// activations of all K+1 steps and a PRIVATE gradient image of the whole network in shared memory - no atomics.  When a
// CTA's warps are done their images are summed in warp order; a second, tiny launch sums the per-CTA partial sums in CTA
// order: the result does not depend on scheduling (deterministic, run to run and rank to rank).
//
// Gradient flow, as autograd sees the reference's graph:
//   total = mean_b w_b * (value_loss_weight * sum_i CE(value_i) + sum_{i>=1} CE(reward_i) + sum_i CE(policy_i));
//   the losses of steps i >= 1 carry a 1 / gradient_scale_b hook; the hidden state returned by every recurrent step
//   carries a 0.5 hook, which scales EVERYTHING that flows into it (its value / policy heads and the next step);
//   the reward head reads the un-normalised next state, the other heads the min-max scaled one;
//   min / max route their gradient to the arg-min / arg-max element (torch.min / max over a dimension).
#include <math_constants.h>

#include "mzb_fc.cuh"

namespace {

constexpr int kMaxLayers = 4;
enum { NET_REP = 0, NET_DYN = 1, NET_REW = 2, NET_POL = 3, NET_VAL = 4 };

struct Lyr {
  int in, out, outp, wt, off_out;     // outp = out | 1; wt: offset of the transposed image; off_out: slot in a step record
  long long w, b;                     // offsets of weight [out][in] and bias [out] in the flat parameter bucket
};
struct TD {
  int obs, enc, A, full, S, K1, B;
  int n[5];
  Lyr l[5][kMaxLayers];
  int PT;                             // floats of the transposed image
  long long P;                        // floats of the flat bucket
  int rec, rec_s, rec_mm;             // step record: layer outputs | normalised state [enc] | arg-min, arg-max, scale
  int maxw;                           // widest vector (scratch size)
  float vlw, alpha;
};

__device__ __forceinline__ float elu_train(float x) { return x > 0.0f ? x : expf(x) - 1.0f; }     // ATen's ELU: exp(x) - 1

// y = W x + b (+ W[:, n_dense + hot] for a one-hot tail), lanes over outputs
__device__ __forceinline__ void lin_fwd(const float* __restrict__ Wt, const Lyr& L, const float* x, int n_dense, int hot,
                                        float* y, bool elu, int lane) {
  const float* w = Wt + L.wt;
  for (int o = lane; o < L.out; o += 32) {
    float acc = w[L.in * L.outp + o];
#pragma unroll 8
    for (int i = 0; i < n_dense; ++i) acc = fmaf(x[i], w[i * L.outp + o], acc);      // unrolled: the loads of 8 terms in flight
    if (hot >= 0) acc += w[(n_dense + hot) * L.outp + o];
    y[o] = elu ? elu_train(acc) : acc;
  }
  __syncwarp();
}

// gW += dy x^T (bias row included), dx = W^T dy for the dense inputs (dx may be NULL)
__device__ __forceinline__ void lin_bwd(const float* __restrict__ Wt, float* gW, const Lyr& L, const float* dy, const float* x,
                                        int n_dense, int hot, float* dx, int lane) {
  float* g = gW + L.wt;
  for (int o = lane; o < L.out; o += 32) {
    const float d = dy[o];
#pragma unroll 8
    for (int i = 0; i < n_dense; ++i) g[i * L.outp + o] = fmaf(x[i], d, g[i * L.outp + o]);
    if (hot >= 0) g[(n_dense + hot) * L.outp + o] += d;
    g[L.in * L.outp + o] += d;
  }
  if (dx) {
    const float* w = Wt + L.wt;
    for (int i = lane; i < n_dense; i += 32) {
      float acc = 0.0f;
#pragma unroll 8
      for (int o = 0; o < L.out; ++o) acc = fmaf(w[i * L.outp + o], dy[o], acc);
      dx[i] = acc;
    }
  }
  __syncwarp();
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, s));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, s);
  return v;
}

// forward of an mlp: outputs of every layer into the step record; returns the last layer's output
__device__ __forceinline__ const float* mlp_fwd(const TD& d, int net, const float* Wt, const float* x, int n_dense, int hot,
                                                float* rec, int lane) {
  const float* in = x;
  float* out = nullptr;
  for (int l = 0; l < d.n[net]; ++l) {
    const Lyr& L = d.l[net][l];
    out = rec + L.off_out;
    lin_fwd(Wt, L, in, l == 0 ? n_dense : L.in, l == 0 ? hot : -1, out, l + 1 < d.n[net], lane);
    in = out;
  }
  return out;
}

// backward of an mlp: `bufA` holds dL/d(last output) on entry; returns the buffer holding dL/d(dense inputs) (or NULL)
__device__ __forceinline__ float* mlp_bwd(const TD& d, int net, const float* Wt, float* gW, const float* x, int n_dense, int hot,
                                          const float* rec, float* bufA, float* bufB, bool want_dx, int lane) {
  for (int l = d.n[net] - 1; l >= 0; --l) {
    const Lyr& L = d.l[net][l];
    const float* in = l == 0 ? x : rec + d.l[net][l - 1].off_out;
    const bool need = l > 0 || want_dx;
    lin_bwd(Wt, gW, L, bufA, in, l == 0 ? n_dense : L.in, l == 0 ? hot : -1, need ? bufB : nullptr, lane);
    if (l > 0) {
      for (int i = lane; i < L.in; i += 32) { const float y = in[i]; bufB[i] *= y > 0.0f ? 1.0f : y + 1.0f; }   // ELU'(x) from its output
      __syncwarp();
    }
    float* t = bufA; bufA = bufB; bufB = t;
  }
  return want_dx ? bufA : nullptr;
}

// cross-entropy of `n` logits z against the soft targets t: returns the loss, writes coef * dL/dz into dz
__device__ __forceinline__ float ce_bwd(const float* z, const float* __restrict__ t, int n, float coef, float* dz, int lane) {
  float m = -CUDART_INF_F;
  for (int k = lane; k < n; k += 32) m = fmaxf(m, z[k]);
  m = warp_max(m);
  float se = 0.0f, st = 0.0f, stz = 0.0f;
  for (int k = lane; k < n; k += 32) { se += expf(z[k] - m); const float tk = t[k]; st += tk; stz += tk * (z[k] - m); }
  se = warp_sum(se); st = warp_sum(st); stz = warp_sum(stz);
  const float lse = logf(se);
  for (int k = lane; k < n; k += 32) dz[k] = coef * (expf(z[k] - m) / se * st - t[k]);
  __syncwarp();
  return st * lse - stz;                                       // -sum_k t_k (z_k - m - lse)
}

// expectation of the categorical support -> scalar (models.py:641-662), all lanes return it
__device__ __forceinline__ float support_scalar(const float* z, int S, int lane) {
  const int n = 2 * S + 1;
  float m = -CUDART_INF_F;
  for (int k = lane; k < n; k += 32) m = fmaxf(m, z[k]);
  m = warp_max(m);
  float se = 0.0f, num = 0.0f;
  for (int k = lane; k < n; k += 32) { const float e = expf(z[k] - m); se += e; num += e * (float)(k - S); }
  se = warp_sum(se); num = warp_sum(num);
  return inverse_value_transform(num / se);
}

struct TrainIO {
  const float* params; const float* obs; const long long* action; const float* tv_sup; const float* tr_sup; const float* tp;
  const float* tv_scalar; const float* weight; const float* gscale;
  float* grad; float* losses; float* priorities; float* loss; float* partial; unsigned int* unused;
};

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_fc_train(const TD d, const TrainIO io) {
  extern __shared__ float4 smem4[];
  float* Wt = reinterpret_cast<float*>(smem4);
  const int per_warp = d.PT + d.K1 * d.rec + d.obs + 3 * d.maxw + d.enc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* mine = Wt + d.PT + (size_t)warp * per_warp;
  float* gW = mine;                                   // private gradient image
  float* recs = gW + d.PT;                            // [K1][rec]
  float* xin = recs + d.K1 * d.rec;                   // observation
  float* bufA = xin + d.obs;
  float* bufB = bufA + d.maxw;
  float* bufC = bufB + d.maxw;
  float* ds = bufC + d.maxw;                          // dL/d(state) carried backwards through time

  // ---- stage the weights (transposed, bias as the last row); clear the gradient images
  for (int net = 0; net < 5; ++net)
    for (int l = 0; l < d.n[net]; ++l) {
      const Lyr& L = d.l[net][l];
      for (int e = threadIdx.x; e < L.in * L.out; e += WARPS * 32) {
        const int o = e / L.in, i = e - o * L.in;
        Wt[L.wt + i * L.outp + o] = io.params[L.w + e];
      }
      for (int o = threadIdx.x; o < L.out; o += WARPS * 32) Wt[L.wt + L.in * L.outp + o] = io.params[L.b + o];
    }
  for (int e = lane; e < d.PT; e += 32) gW[e] = 0.0f;
  __syncthreads();

  const int b = blockIdx.x * WARPS + warp;
  if (b < d.B) {
    const int K1 = d.K1, enc = d.enc, full = d.full, A = d.A;
    for (int i = lane; i < d.obs; i += 32) xin[i] = io.obs[(size_t)b * d.obs + i];
    __syncwarp();
    // ---------------- forward
    for (int step = 0; step < K1; ++step) {
      float* rec = recs + step * d.rec;
      const float* raw;
      if (step == 0) {
        raw = mlp_fwd(d, NET_REP, Wt, xin, d.obs, -1, rec, lane);
      } else {
        const int a = (int)io.action[(size_t)b * K1 + step];
        raw = mlp_fwd(d, NET_DYN, Wt, rec - d.rec + d.rec_s, enc, a, rec, lane);
        mlp_fwd(d, NET_REW, Wt, raw, enc, -1, rec, lane);              // on the un-normalised next state
      }
      // min-max scaling (models.py:138-145); arg-min / arg-max = first occurrence, as torch.min / max return it
      float lo = CUDART_INF_F, hi = -CUDART_INF_F;
      int ilo = 0, ihi = 0;
      for (int i = 0; i < enc; ++i) {
        const float v = raw[i];
        if (v < lo) { lo = v; ilo = i; }
        if (v > hi) { hi = v; ihi = i; }
      }
      float scale = hi - lo;
      if (scale < 1e-5f) scale += 1e-5f;
      float* s = rec + d.rec_s;
      for (int i = lane; i < enc; i += 32) s[i] = (raw[i] - lo) / scale;
      if (lane == 0) { rec[d.rec_mm] = __int_as_float(ilo); rec[d.rec_mm + 1] = __int_as_float(ihi); rec[d.rec_mm + 2] = scale; }
      __syncwarp();
      mlp_fwd(d, NET_POL, Wt, s, enc, -1, rec, lane);
      mlp_fwd(d, NET_VAL, Wt, s, enc, -1, rec, lane);
    }
    // ---------------- losses + backward through time
    const float coef_b = (io.weight ? io.weight[b] : 1.0f) / (float)d.B;
    const float inv_g = 1.0f / io.gscale[(size_t)b * K1 + K1 - 1];       // the reference's hooks all bind the LAST column
    float vl = 0.0f, rl = 0.0f, pl = 0.0f;
    for (int i = lane; i < enc; i += 32) ds[i] = 0.0f;
    __syncwarp();
    for (int step = K1 - 1; step >= 0; --step) {
      float* rec = recs + step * d.rec;
      const float* s = rec + d.rec_s;
      const float sc = step == 0 ? 1.0f : inv_g;
      const size_t bi = (size_t)b * K1 + step;
      // value head
      {
        const float* z = rec + d.l[NET_VAL][d.n[NET_VAL] - 1].off_out;
        const float pred = support_scalar(z, d.S, lane);
        if (lane == 0) io.priorities[bi] = powf(fabsf(pred - io.tv_scalar[bi]), d.alpha);
        vl += ce_bwd(z, io.tv_sup + bi * full, full, coef_b * d.vlw * sc, bufA, lane);
        const float* dx = mlp_bwd(d, NET_VAL, Wt, gW, s, enc, -1, rec, bufA, bufB, true, lane);
        for (int i = lane; i < enc; i += 32) ds[i] += dx[i];
        __syncwarp();
      }
      // policy head
      {
        const float* z = rec + d.l[NET_POL][d.n[NET_POL] - 1].off_out;
        pl += ce_bwd(z, io.tp + bi * A, A, coef_b * sc, bufA, lane);
        const float* dx = mlp_bwd(d, NET_POL, Wt, gW, s, enc, -1, rec, bufA, bufB, true, lane);
        for (int i = lane; i < enc; i += 32) ds[i] += dx[i];
        __syncwarp();
      }
      // the hook on the hidden state a recurrent step returned
      const float hook = step == 0 ? 1.0f : 0.5f;
      // min-max backward: y_j = (x_j - lo) / scale, scale = hi - lo
      const int ilo = __float_as_int(rec[d.rec_mm]), ihi = __float_as_int(rec[d.rec_mm + 1]);
      const float scale = rec[d.rec_mm + 2];
      float sum_dy = 0.0f, sum_dyy = 0.0f;
      for (int i = lane; i < enc; i += 32) { const float g = ds[i] * hook; sum_dy += g; sum_dyy += g * s[i]; }
      sum_dy = warp_sum(sum_dy); sum_dyy = warp_sum(sum_dyy);
      const float d_lo = (sum_dyy - sum_dy) / scale, d_hi = -sum_dyy / scale;
      for (int i = lane; i < enc; i += 32) {
        float g = ds[i] * hook / scale;
        if (i == ilo) g += d_lo;
        if (i == ihi) g += d_hi;
        bufC[i] = g;                                                    // dL/d(raw state)
      }
      __syncwarp();
      if (step > 0) {
        // reward head (un-normalised state): its data gradient joins dL/d(raw)
        const float* z = rec + d.l[NET_REW][d.n[NET_REW] - 1].off_out;
        const float* raw = rec + d.l[NET_DYN][d.n[NET_DYN] - 1].off_out;
        rl += ce_bwd(z, io.tr_sup + bi * full, full, coef_b * sc, bufA, lane);
        const float* dx = mlp_bwd(d, NET_REW, Wt, gW, raw, enc, -1, rec, bufA, bufB, true, lane);
        for (int i = lane; i < enc; i += 32) bufC[i] += dx[i];
        __syncwarp();
        // dynamics: input = previous normalised state + one-hot action
        const int a = (int)io.action[bi];
        for (int i = lane; i < enc; i += 32) bufA[i] = bufC[i];
        __syncwarp();
        const float* dprev = mlp_bwd(d, NET_DYN, Wt, gW, rec - d.rec + d.rec_s, enc, a, rec, bufA, bufB, true, lane);
        for (int i = lane; i < enc; i += 32) ds[i] = dprev[i];
        __syncwarp();
      } else {
        for (int i = lane; i < enc; i += 32) bufA[i] = bufC[i];
        __syncwarp();
        mlp_bwd(d, NET_REP, Wt, gW, xin, d.obs, -1, rec, bufA, bufB, false, lane);
      }
    }
    if (lane == 0) {
      io.losses[b] = vl; io.losses[d.B + b] = rl; io.losses[2 * d.B + b] = pl;
    }
  }
  __syncthreads();
  // ---------------- CTA partial sum (warp order), written in the flat bucket's layout
  float* part = io.partial + (size_t)blockIdx.x * d.P;
  for (int net = 0; net < 5; ++net)
    for (int l = 0; l < d.n[net]; ++l) {
      const Lyr& L = d.l[net][l];
      for (int e = threadIdx.x; e < L.in * L.out + L.out; e += WARPS * 32) {
        int src; long long dst;
        if (e < L.in * L.out) { const int o = e / L.in, i = e - o * L.in; src = L.wt + i * L.outp + o; dst = L.w + e; }
        else { const int o = e - L.in * L.out; src = L.wt + L.in * L.outp + o; dst = L.b + o; }
        float acc = 0.0f;
        for (int w = 0; w < WARPS; ++w) acc += Wt[d.PT + (size_t)w * per_warp + src];
        part[dst] = acc;
      }
    }
}

// Second (tiny) launch: gradient element p = sum over the CTAs' partial sums IN CTA ORDER (deterministic), eight loads in
// flight at a time; thread 0 forms the batch objective.  A "last CTA reduces" epilogue in the first kernel serialised
// 32 dependent L2 round trips per element on one CTA and took longer than the training step itself.
__global__ void k_fc_train_reduce(const TD d, const TrainIO io, int n_cta) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < d.P) {
    float acc = 0.0f;
    for (int c0 = 0; c0 < n_cta; c0 += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = c0 + j < n_cta ? io.partial[(size_t)(c0 + j) * d.P + p] : 0.0f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += v[j];
    }
    io.grad[p] = acc;
  }
  if (p == 0) {
    float total = 0.0f;
    for (int i = 0; i < d.B; ++i) {
      float l = io.losses[i] * d.vlw + io.losses[d.B + i] + io.losses[2 * d.B + i];
      if (io.weight) l *= io.weight[i];
      total += l;
    }
    io.loss[0] = total / (float)d.B;
  }
}

constexpr int kTrainWarps = 4;

bool build_td(const mzb_fc_train_desc* h, int B, int K1, float vlw, float alpha, TD* out) {
  TD d{};
  d.obs = h->obs_dim; d.enc = h->encoding_size; d.A = h->n_actions; d.S = h->support_size; d.full = 2 * h->support_size + 1;
  d.K1 = K1; d.B = B; d.vlw = vlw; d.alpha = alpha;
  int wt = 0, maxw = d.obs > d.enc ? d.obs : d.enc;
  long long P = 0;
  int sum_out[5] = {0, 0, 0, 0, 0};
  for (int net = 0; net < 5; ++net) {
    d.n[net] = h->n_layers[net];
    if (d.n[net] < 1 || d.n[net] > kMaxLayers) return false;
    for (int l = 0; l < d.n[net]; ++l) {
      Lyr& L = d.l[net][l];
      L.in = h->in[net][l]; L.out = h->out[net][l]; L.outp = L.out | 1; L.wt = wt;
      L.w = h->w_off[net][l]; L.b = h->b_off[net][l];
      wt += (L.in + 1) * L.outp;
      sum_out[net] += L.out;
      if (L.out > maxw) maxw = L.out;
      if (L.in > maxw) maxw = L.in;
      const long long end_w = L.w + (long long)L.in * L.out, end_b = L.b + L.out;
      P = end_w > P ? end_w : P;
      P = end_b > P ? end_b : P;
    }
  }
  // step record: slot of the representation / dynamics outputs (never both in one step), then reward, policy, value
  const int slot0 = sum_out[NET_REP] > sum_out[NET_DYN] ? sum_out[NET_REP] : sum_out[NET_DYN];
  int base[5] = {0, 0, slot0, slot0 + sum_out[NET_REW], slot0 + sum_out[NET_REW] + sum_out[NET_POL]};
  for (int net = 0; net < 5; ++net) {
    int off = base[net];
    for (int l = 0; l < d.n[net]; ++l) { d.l[net][l].off_out = off; off += d.l[net][l].out; }
  }
  d.rec_s = slot0 + sum_out[NET_REW] + sum_out[NET_POL] + sum_out[NET_VAL];
  d.rec_mm = d.rec_s + d.enc;
  d.rec = (d.rec_mm + 3 + 3) & ~3;
  d.PT = (wt + 3) & ~3;
  d.P = P;
  d.maxw = (maxw + 3) & ~3;
  *out = d;
  return true;
}

size_t train_smem(const TD& d, int warps) {
  return sizeof(float) * ((size_t)d.PT + (size_t)warps * ((size_t)d.PT + (size_t)d.K1 * d.rec + d.obs + 3 * d.maxw + d.enc)) + 16;
}

}  // namespace

extern "C" {

int64_t mzb_fc_train_workspace_bytes(const mzb_fc_train_desc* desc, int32_t batch, int32_t unroll_plus_1) {
  TD d;
  if (!desc || !build_td(desc, batch, unroll_plus_1, 1.0f, 1.0f, &d)) return -1;
  const int ctas = (batch + kTrainWarps - 1) / kTrainWarps;
  return (int64_t)(sizeof(float) * (size_t)ctas * (size_t)d.P + 256);
}

int mzb_fc_train_fits(const mzb_fc_train_desc* desc, int32_t batch, int32_t unroll_plus_1) {
  TD d;
  if (!desc || !build_td(desc, batch, unroll_plus_1, 1.0f, 1.0f, &d)) return 0;
  return train_smem(d, kTrainWarps) <= 226 * 1024 ? 1 : 0;
}

int mzb_fc_train_grad(const mzb_fc_train_desc* desc, const float* d_params, int64_t n_params, int32_t batch, int32_t unroll_plus_1,
                      const float* d_obs, const int64_t* d_action, const float* d_target_value_support,
                      const float* d_target_reward_support, const float* d_target_policy, const float* d_target_value_scalar,
                      const float* d_weight, const float* d_gradient_scale, double value_loss_weight, double per_alpha,
                      float* d_grad, float* d_losses, float* d_priorities, float* d_loss, void* d_workspace,
                      int64_t workspace_bytes, void* stream) {
  MZB_CHECK_ARG(desc && d_params && d_obs && d_action && d_target_value_support && d_target_reward_support && d_target_policy &&
                    d_target_value_scalar && d_gradient_scale && d_grad && d_losses && d_priorities && d_loss && d_workspace,
                "NULL argument");
  MZB_CHECK_ARG(batch > 0 && unroll_plus_1 > 0, "batch %d, unroll steps + 1 = %d", batch, unroll_plus_1);
  TD d;
  MZB_CHECK_ARG(build_td(desc, batch, unroll_plus_1, (float)value_loss_weight, (float)per_alpha, &d), "bad layer table");
  MZB_CHECK_ARG(d.P == n_params, "layer table covers %lld parameters, the bucket holds %lld", d.P, (long long)n_params);
  const size_t smem = train_smem(d, kTrainWarps);
  MZB_CHECK_ARG(smem <= 226 * 1024, "network / unroll too large for the one-kernel training step (%zu bytes of shared memory)", smem);
  const int ctas = (batch + kTrainWarps - 1) / kTrainWarps;
  const size_t need = sizeof(float) * (size_t)ctas * (size_t)d.P + 256;
  MZB_CHECK_ARG((size_t)workspace_bytes >= need, "workspace holds %lld bytes, %zu needed", (long long)workspace_bytes, need);
  static bool configured = false;
  if (!configured) {
    MZB_CUDA(cudaFuncSetAttribute(k_fc_train<kTrainWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    configured = true;
  }
  TrainIO io{d_params, d_obs, (const long long*)d_action, d_target_value_support, d_target_reward_support, d_target_policy,
             d_target_value_scalar, d_weight, d_gradient_scale, d_grad, d_losses, d_priorities, d_loss,
             reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(d_workspace) + 256), reinterpret_cast<unsigned int*>(d_workspace)};
  k_fc_train<kTrainWarps><<<ctas, kTrainWarps * 32, smem, (cudaStream_t)stream>>>(d, io);
  MZB_LAUNCH_CHECK();
  k_fc_train_reduce<<<(unsigned)((d.P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d, io, ctas);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

}  // extern "C"
