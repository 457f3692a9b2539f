// Fully-connected MuZero network on device: packed-weight description and the float32 arithmetic
// shared by the batched inference kernels (mzb_fc.cu) and the fused search kernels (mzb_fused.cu).
// Reference: models.py:80-195 (network), :626-638 (mlp), :641-662 (support_to_scalar).
//
// Every dot product is  acc = bias; for i ascending: acc = fmaf(x[i], W[o][i], acc)  in BOTH kernel
// families, so the batched and the fused path produce bit-identical network outputs.
#pragma once
#include "mzb_common.cuh"

#define MZB_FC_MAX_LAYERS 4

struct FcLayer {
  int in, out, outp;   // outp = out rounded up to 4 (zero padded)
  int w_off, b_off;    // float offsets into the pack: Wt [in][outp] (transposed), b [outp]
};
struct FcNet {
  int n;
  FcLayer l[MZB_FC_MAX_LAYERS];
};
struct FcDesc {
  int obs_dim, enc, A, S, full;
  int pack_floats, max_width;
  FcNet rep, dyn, rew, pol, val;
};

struct mzb_fc_model {
  mzb_fc_config cfg;
  FcDesc d;
  float* d_pack;
  int n_tensors;
  int rows_per_block;
  size_t smem_bytes;
};

__device__ __forceinline__ float elu_f32(float x) {
  // ATen's ELU evaluates exp(x) - 1 for x <= 0 (alpha = 1).  Evaluated unconditionally on min(x, 0) and selected:
  // a thread-per-game warp diverges on the sign of every hidden unit, and a branch around expf costs more than the
  // eight instructions it skips (same value bit for bit: exp(0) - 1 = 0 is never selected for x > 0).
  const float e = __fsub_rn(expf(fminf(x, 0.0f)), 1.0f);
  return x > 0.0f ? x : e;
}

// Inverse of the value transform h(x) (models.py:656-661), in torch's float32 operation order.
// Ill-conditioned by construction (subtracts 1 from sqrt(1 + 0.004 t)): keep the order, forbid FMA.
__device__ __forceinline__ float inverse_value_transform(float x) {
  float t = __fadd_rn(__fadd_rn(fabsf(x), 1.0f), 0.001f);
  t = __fmul_rn(0.004f, t);
  t = __fsqrt_rn(__fadd_rn(1.0f, t));
  t = __fdiv_rn(__fsub_rn(t, 1.0f), 0.002f);
  t = __fsub_rn(__fmul_rn(t, t), 1.0f);
  const float sgn = (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f);
  return __fmul_rn(sgn, t);
}

// h(x) = sign(x)(sqrt(|x|+1)-1) + 0.001 x  (models.py:671)
__device__ __forceinline__ float value_transform(float x) {
  const float sgn = (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f);
  const float r = __fsub_rn(__fsqrt_rn(__fadd_rn(fabsf(x), 1.0f)), 1.0f);
  return __fadd_rn(__fmul_rn(sgn, r), __fmul_rn(0.001f, x));
}

// support_to_scalar over `full` logits read through `logit(i)`; `e(i)` is scratch for the exponentials.
// softmax -> expectation over [-S..S] (both summed in index order) -> inverse transform.
template <class Logit, class Scratch>
__device__ __forceinline__ float support_to_scalar_dev(Logit logit, Scratch e, int S) {
  const int full = 2 * S + 1;
  float m = -CUDART_INF_F;
  for (int i = 0; i < full; ++i) m = fmaxf(m, logit(i));
  float sum = 0.0f;
  for (int i = 0; i < full; ++i) {
    const float v = softmax_exp(logit(i), m);
    e(i) = v;
    sum = __fadd_rn(sum, v);
  }
  // expectation = (sum_i support_i * e_i) / (sum_i e_i): one division instead of one per bin
  float num = 0.0f;
  for (int i = 0; i < full; ++i) num = fmaf((float)(i - S), e(i), num);
  return inverse_value_transform(__fdiv_rn(num, sum));
}

// mzb_fc.cu, internal: inference on hidden states kept in the tree store's blocked slots (mzb_tree.cuh)
int mzb_fc_initial_tree(mzb_fc_model* m, int64_t B, const float* d_obs, const uint8_t* d_legal, float* d_hidden, int S1,
                        int out_slot, float* d_value, float* d_reward, float* d_priors, void* stream);
int mzb_fc_recurrent_tree(mzb_fc_model* m, int64_t B, float* d_hidden, int S1, const int32_t* d_in_slot,
                          const int32_t* d_action, int out_slot, float* d_value, float* d_reward, float* d_priors,
                          void* stream);
