// Batched fully-connected MuZero inference (K4 recurrent, K5 initial) + support codec kernels.
// One thread owns one game row; the packed weights sit in shared memory (every lane of a warp
// reads the same weight -> one broadcast LDS.128 feeds four FMAs), activations in a transposed
// shared-memory tile act[i][thread] (conflict-free).  Any mlp() depth up to 4 Linear layers.
// Reference: models.py:128-195.
#include <vector>

#include "mzb_fc.cuh"

namespace {

struct RowIO {
  // state_in(row) = in + row*in_row_stride + (in_slot ? in_slot[row] : 0) * slot_stride
  const float* in; long long in_row_stride; const int* in_slot; long long slot_stride;
  float* out; long long out_row_stride; long long out_off;
  // blocked rows (the tree store's hidden slots, mzb_tree.cuh): when *_blk != 0 the row term becomes
  // (row >> 5) * blk + (row & 31) * row_stride
  long long in_blk, out_blk;
  __device__ __forceinline__ long long in_row(long long row) const {
    return in_blk ? (row >> 5) * in_blk + (row & 31) * in_row_stride : row * in_row_stride;
  }
  __device__ __forceinline__ long long out_row(long long row) const {
    return out_blk ? (row >> 5) * out_blk + (row & 31) * out_row_stride : row * out_row_stride;
  }
  float* value_logits; float* reward_logits; float* policy_logits;
  float* value; float* reward; float* priors;
};

// One Linear(+ELU) layer for this thread's row. x(i) reads input i; `hot` >= 0 adds the weight row of a
// one-hot input placed after the `n_direct` dense inputs (the action of the dynamics net, models.py:149-156).
template <class In>
__device__ __forceinline__ void fc_layer(const FcLayer& L, const float* __restrict__ pack, In x, int n_direct, int hot,
                                         float* __restrict__ out, int tid, int NT, bool elu) {
  for (int o = 0; o < L.outp; o += 4) {
    float4 acc = *reinterpret_cast<const float4*>(pack + L.b_off + o);
    const float* w = pack + L.w_off + o;
    for (int i = 0; i < n_direct; ++i) {
      const float xi = x(i);
      const float4 wv = *reinterpret_cast<const float4*>(w + (size_t)i * L.outp);
      acc.x = fmaf(xi, wv.x, acc.x); acc.y = fmaf(xi, wv.y, acc.y);
      acc.z = fmaf(xi, wv.z, acc.z); acc.w = fmaf(xi, wv.w, acc.w);
    }
    if (hot >= 0) {
      const float4 wv = *reinterpret_cast<const float4*>(w + (size_t)(n_direct + hot) * L.outp);
      acc.x = __fadd_rn(acc.x, wv.x); acc.y = __fadd_rn(acc.y, wv.y);
      acc.z = __fadd_rn(acc.z, wv.z); acc.w = __fadd_rn(acc.w, wv.w);
    }
    if (elu) { acc.x = elu_f32(acc.x); acc.y = elu_f32(acc.y); acc.z = elu_f32(acc.z); acc.w = elu_f32(acc.w); }
    out[(size_t)(o + 0) * NT + tid] = acc.x;
    if (o + 1 < L.out) out[(size_t)(o + 1) * NT + tid] = acc.y;
    if (o + 2 < L.out) out[(size_t)(o + 2) * NT + tid] = acc.z;
    if (o + 3 < L.out) out[(size_t)(o + 3) * NT + tid] = acc.w;
  }
}

// Whole mlp(): input through x / hot, result column returned (one of the two scratch tiles).
template <class In>
__device__ __forceinline__ float* fc_mlp(const FcNet& net, const float* pack, In x, int n_direct, int hot, float* t0,
                                         float* t1, int tid, int NT) {
  float* cur = t0;
  float* nxt = t1;
  fc_layer(net.l[0], pack, x, n_direct, hot, cur, tid, NT, net.n > 1);
  for (int k = 1; k < net.n; ++k) {
    const float* src = cur;
    fc_layer(net.l[k], pack, [=](int i) { return src[(size_t)i * NT + tid]; }, net.l[k].in, -1, nxt, tid, NT,
             k < net.n - 1);
    float* s = cur; cur = nxt; nxt = s;
  }
  return cur;
}

__device__ __forceinline__ void minmax_normalize(float* col, int n, int tid, int NT) {
  float lo = CUDART_INF_F, hi = -CUDART_INF_F;
  for (int i = 0; i < n; ++i) { const float v = col[(size_t)i * NT + tid]; lo = fminf(lo, v); hi = fmaxf(hi, v); }
  float scale = __fsub_rn(hi, lo);
  if (scale < 1e-5f) scale = __fadd_rn(scale, 1e-5f);            // models.py:141 / :164
  for (int i = 0; i < n; ++i) col[(size_t)i * NT + tid] = __fdiv_rn(__fsub_rn(col[(size_t)i * NT + tid], lo), scale);
}

// prediction heads + outputs for a normalised state held in column `st`
__device__ __forceinline__ void heads(const FcDesc& d, const float* pack, const float* st, float* t0, float* t1,
                                      int tid, int NT, long long row, const RowIO& io, const uint8_t* legal) {
  auto sx = [=](int i) { return st[(size_t)i * NT + tid]; };
  // policy (models.py:128-131)
  float* pl = fc_mlp(d.pol, pack, sx, d.enc, -1, t0, t1, tid, NT);
  if (io.policy_logits) for (int a = 0; a < d.A; ++a) io.policy_logits[row * d.A + a] = pl[(size_t)a * NT + tid];
  if (io.priors) {
    // Node.expand softmax (self_play.py:459-461): over the legal actions at the root, all actions below
    const uint8_t* lg = legal ? legal + row * d.A : nullptr;
    float m = -CUDART_INF_F;
    for (int a = 0; a < d.A; ++a) if (!lg || lg[a]) m = fmaxf(m, pl[(size_t)a * NT + tid]);
    float sum = 0.0f;
    for (int a = 0; a < d.A; ++a) {
      const float e = (!lg || lg[a]) ? softmax_exp(pl[(size_t)a * NT + tid], m) : 0.0f;
      pl[(size_t)a * NT + tid] = e;
      sum = __fadd_rn(sum, e);
    }
    for (int a = 0; a < d.A; ++a) io.priors[row * d.A + a] = __fdiv_rn(pl[(size_t)a * NT + tid], sum);
  }
  // value
  float* vl = fc_mlp(d.val, pack, sx, d.enc, -1, t0, t1, tid, NT);
  if (io.value_logits) for (int i = 0; i < d.full; ++i) io.value_logits[row * d.full + i] = vl[(size_t)i * NT + tid];
  if (io.value) {
    float* scratch = (vl == t0) ? t1 : t0;
    io.value[row] = support_to_scalar_dev([=](int i) { return vl[(size_t)i * NT + tid]; },
                                          [=](int i) -> float& { return scratch[(size_t)i * NT + tid]; }, d.S);
  }
}

__global__ void k_fc_initial(FcDesc d, const float* __restrict__ gpack, int B, const float* __restrict__ obs,
                             const uint8_t* __restrict__ legal, RowIO io) {
  extern __shared__ float4 smem4[];
  float* pack = reinterpret_cast<float*>(smem4);
  const int NT = blockDim.x, tid = threadIdx.x;
  for (int i = tid; i < d.pack_floats / 4; i += NT) smem4[i] = reinterpret_cast<const float4*>(gpack)[i];
  float* st = pack + d.pack_floats;
  float* t0 = st + (size_t)d.max_width * NT;
  float* t1 = t0 + (size_t)d.max_width * NT;
  __syncthreads();
  const long long row = (long long)blockIdx.x * NT + tid;
  if (row >= B) return;
  const float* ob = obs + row * d.obs_dim;
  // representation (models.py:133-145)
  float* enc = fc_mlp(d.rep, pack, [=](int i) { return ob[i]; }, d.obs_dim, -1, t0, t1, tid, NT);
  for (int i = 0; i < d.enc; ++i) st[(size_t)i * NT + tid] = enc[(size_t)i * NT + tid];
  minmax_normalize(st, d.enc, tid, NT);
  if (io.out) {
    float* o = io.out + io.out_row(row) + io.out_off;
    for (int i = 0; i < d.enc; ++i) o[i] = st[(size_t)i * NT + tid];
  }
  // reward := log(one-hot(centre)) (models.py:176-183)
  if (io.reward_logits)
    for (int i = 0; i < d.full; ++i) io.reward_logits[row * d.full + i] = (i == d.S) ? 0.0f : -CUDART_INF_F;
  if (io.reward) {
    const int S = d.S;
    io.reward[row] = support_to_scalar_dev([=](int i) { return i == S ? 0.0f : -CUDART_INF_F; },
                                           [=](int i) -> float& { return t0[(size_t)i * NT + tid]; }, d.S);
  }
  heads(d, pack, st, t0, t1, tid, NT, row, io, legal);
}

__global__ void k_fc_recurrent(FcDesc d, const float* __restrict__ gpack, int B, const int* __restrict__ action,
                               RowIO io) {
  extern __shared__ float4 smem4[];
  float* pack = reinterpret_cast<float*>(smem4);
  const int NT = blockDim.x, tid = threadIdx.x;
  for (int i = tid; i < d.pack_floats / 4; i += NT) smem4[i] = reinterpret_cast<const float4*>(gpack)[i];
  float* st = pack + d.pack_floats;
  float* t0 = st + (size_t)d.max_width * NT;
  float* t1 = t0 + (size_t)d.max_width * NT;
  __syncthreads();
  const long long row = (long long)blockIdx.x * NT + tid;
  if (row >= B) return;
  const float* sin = io.in + io.in_row(row) + (io.in_slot ? (long long)io.in_slot[row] * io.slot_stride : 0);
  const int a = action[row];
  // dynamics (models.py:147-170): state ++ one-hot(action) -> next state
  float* nx = fc_mlp(d.dyn, pack, [=](int i) { return sin[i]; }, d.enc, a, t0, t1, tid, NT);
  for (int i = 0; i < d.enc; ++i) st[(size_t)i * NT + tid] = nx[(size_t)i * NT + tid];
  // reward head on the UN-normalised next state (:159)
  auto sx = [=](int i) { return st[(size_t)i * NT + tid]; };
  float* rl = fc_mlp(d.rew, pack, sx, d.enc, -1, t0, t1, tid, NT);
  if (io.reward_logits) for (int i = 0; i < d.full; ++i) io.reward_logits[row * d.full + i] = rl[(size_t)i * NT + tid];
  if (io.reward) {
    float* scratch = (rl == t0) ? t1 : t0;
    io.reward[row] = support_to_scalar_dev([=](int i) { return rl[(size_t)i * NT + tid]; },
                                           [=](int i) -> float& { return scratch[(size_t)i * NT + tid]; }, d.S);
  }
  minmax_normalize(st, d.enc, tid, NT);
  if (io.out) {
    float* o = io.out + io.out_row(row) + io.out_off;
    for (int i = 0; i < d.enc; ++i) o[i] = st[(size_t)i * NT + tid];
  }
  heads(d, pack, st, t0, t1, tid, NT, row, io, nullptr);
}

__global__ void k_support_to_scalar(const float* __restrict__ logits, long long B, int S, float* __restrict__ out) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= B) return;
  const int full = 2 * S + 1;
  const float* l = logits + row * full;
  float m = -CUDART_INF_F;
  for (int i = 0; i < full; ++i) m = fmaxf(m, l[i]);
  float sum = 0.0f;
  for (int i = 0; i < full; ++i) sum = __fadd_rn(sum, softmax_exp(l[i], m));
  float num = 0.0f;
  for (int i = 0; i < full; ++i) num = fmaf((float)(i - S), softmax_exp(l[i], m), num);
  out[row] = inverse_value_transform(__fdiv_rn(num, sum));
}

// scalar_to_support (models.py:665-685): h-transform, clamp, two-hot on floor / floor+1
__global__ void k_scalar_to_support(const float* __restrict__ x, long long n, int S, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int full = 2 * S + 1;
  float v = value_transform(x[i]);
  v = fminf(fmaxf(v, (float)-S), (float)S);
  const float fl = floorf(v);
  const float prob = __fsub_rn(v, fl);
  float* o = out + i * full;
  for (int k = 0; k < full; ++k) o[k] = 0.0f;
  const int lo = (int)fl + S;
  o[lo] = __fsub_rn(1.0f, prob);
  if (lo + 1 <= 2 * S) o[lo + 1] = prob;       // upper index beyond the support is dropped (:682-684)
}

bool add_net(FcNet& net, int in, const int32_t* hidden, int n_hidden, int out, int& off, int& maxw) {
  if (n_hidden < 0 || n_hidden + 1 > MZB_FC_MAX_LAYERS) return false;
  net.n = n_hidden + 1;
  int cur = in;
  for (int k = 0; k < net.n; ++k) {
    const int o = k < n_hidden ? hidden[k] : out;
    if (o <= 0 || cur <= 0) return false;
    FcLayer& L = net.l[k];
    L.in = cur; L.out = o; L.outp = (o + 3) / 4 * 4;
    L.w_off = off; off += L.in * L.outp;
    L.b_off = off; off += L.outp;
    if (L.outp > maxw) maxw = L.outp;
    cur = o;
  }
  return true;
}

}  // namespace

extern "C" {

int mzb_fc_create(mzb_fc_model** out, const mzb_fc_config* c) {
  MZB_CHECK_ARG(out && c, "NULL argument");
  *out = nullptr;
  MZB_CHECK_ARG(c->obs_dim > 0 && c->encoding_size > 0 && c->n_actions > 0 && c->support_size > 0,
                "fc config: non-positive dimension");
  mzb_fc_model* m = new mzb_fc_model();
  m->cfg = *c;
  FcDesc& d = m->d;
  d.obs_dim = c->obs_dim; d.enc = c->encoding_size; d.A = c->n_actions; d.S = c->support_size; d.full = 2 * c->support_size + 1;
  int off = 0, maxw = (d.enc + 3) / 4 * 4;
  bool ok = add_net(d.rep, d.obs_dim, c->rep, c->n_rep, d.enc, off, maxw) &&
            add_net(d.dyn, d.enc + d.A, c->dyn, c->n_dyn, d.enc, off, maxw) &&
            add_net(d.rew, d.enc, c->rew, c->n_rew, d.full, off, maxw) &&
            add_net(d.pol, d.enc, c->pol, c->n_pol, d.A, off, maxw) &&
            add_net(d.val, d.enc, c->val, c->n_val, d.full, off, maxw);
  if (!ok) {
    delete m;
    mzb_set_error("fc config: an mlp has more than %d Linear layers or a non-positive width", MZB_FC_MAX_LAYERS);
    return MZB_EINVAL;
  }
  d.pack_floats = off;
  d.max_width = maxw;
  m->n_tensors = 2 * (d.rep.n + d.dyn.n + d.rew.n + d.pol.n + d.val.n);
  // rows per block: the largest of 128/64/32 whose tiles fit in shared memory next to the weights
  int device = 0, max_smem = 0;
  cudaGetDevice(&device);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  if (max_smem <= 0) max_smem = 227 * 1024;
  m->rows_per_block = 0;
  for (int nt = 128; nt >= 32; nt >>= 1) {
    const size_t need = ((size_t)d.pack_floats + 3ull * d.max_width * nt) * sizeof(float);
    if (need <= (size_t)max_smem) { m->rows_per_block = nt; m->smem_bytes = need; break; }
  }
  if (!m->rows_per_block) {
    delete m;
    mzb_set_error("fc model does not fit in shared memory (%d weights, width %d)", off, maxw);
    return MZB_EUNSUPPORTED;
  }
  cudaError_t e = cudaMalloc(&m->d_pack, sizeof(float) * (size_t)d.pack_floats);
  if (e == cudaSuccess) e = cudaMemset(m->d_pack, 0, sizeof(float) * (size_t)d.pack_floats);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_fc_initial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem_bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_fc_recurrent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem_bytes);
  if (e != cudaSuccess) {
    mzb_set_error("fc create: %s", cudaGetErrorString(e));
    delete m;
    return MZB_ECUDA;
  }
  *out = m;
  return MZB_OK;
}

int mzb_fc_destroy(mzb_fc_model* m) {
  if (!m) return MZB_OK;
  cudaFree(m->d_pack);
  delete m;
  return MZB_OK;
}

int mzb_fc_num_tensors(const mzb_fc_model* m) { return m ? m->n_tensors : 0; }
int mzb_fc_pack_floats(const mzb_fc_model* m) { return m ? m->d.pack_floats : 0; }
const float* mzb_fc_pack_ptr(const mzb_fc_model* m) { return m ? m->d_pack : nullptr; }

int mzb_fc_set_weights(mzb_fc_model* m, const float* const* h_tensors, int n_tensors, void* stream) {
  MZB_CHECK_ARG(m && h_tensors, "NULL argument");
  MZB_CHECK_ARG(n_tensors == m->n_tensors, "expected %d weight tensors (W,b per Linear), got %d", m->n_tensors, n_tensors);
  std::vector<float> pack((size_t)m->d.pack_floats, 0.0f);
  const FcNet* nets[5] = {&m->d.rep, &m->d.dyn, &m->d.rew, &m->d.pol, &m->d.val};   // state_dict order, models.py:98-126
  int t = 0;
  for (const FcNet* net : nets) {
    for (int k = 0; k < net->n; ++k) {
      const FcLayer& L = net->l[k];
      const float* W = h_tensors[t++];     // [out][in] row-major (torch.nn.Linear.weight)
      const float* b = h_tensors[t++];
      MZB_CHECK_ARG(W && b, "weight tensor %d is NULL", t - 2);
      for (int o = 0; o < L.out; ++o) {
        for (int i = 0; i < L.in; ++i) pack[(size_t)L.w_off + (size_t)i * L.outp + o] = W[(size_t)o * L.in + i];
        pack[(size_t)L.b_off + o] = b[o];
      }
    }
  }
  cudaStream_t s = (cudaStream_t)stream;
  MZB_CUDA(cudaMemcpyAsync(m->d_pack, pack.data(), sizeof(float) * pack.size(), cudaMemcpyHostToDevice, s));
  MZB_CUDA(cudaStreamSynchronize(s));      // the staging vector dies with this call
  return MZB_OK;
}

int mzb_fc_initial(mzb_fc_model* m, int64_t B, const float* d_obs, const uint8_t* d_legal, float* d_state_out,
                   int64_t out_row_stride, int64_t out_offset, float* d_value_logits, float* d_reward_logits,
                   float* d_policy_logits, float* d_value, float* d_reward, float* d_priors, void* stream) {
  MZB_CHECK_ARG(m && d_obs, "NULL argument");
  MZB_CHECK_ARG(B > 0 && B < (1ll << 31), "batch out of range: %lld", (long long)B);
  RowIO io{};
  io.out = d_state_out; io.out_row_stride = out_row_stride; io.out_off = out_offset;
  io.value_logits = d_value_logits; io.reward_logits = d_reward_logits; io.policy_logits = d_policy_logits;
  io.value = d_value; io.reward = d_reward; io.priors = d_priors;
  const int nt = m->rows_per_block;
  k_fc_initial<<<(unsigned)((B + nt - 1) / nt), nt, m->smem_bytes, (cudaStream_t)stream>>>(m->d, m->d_pack, (int)B, d_obs,
                                                                                            d_legal, io);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

int mzb_fc_recurrent(mzb_fc_model* m, int64_t B, const float* d_state_in, int64_t in_row_stride,
                     const int32_t* d_in_slot, int64_t slot_stride, const int32_t* d_action, float* d_state_out,
                     int64_t out_row_stride, int64_t out_offset, float* d_value_logits, float* d_reward_logits,
                     float* d_policy_logits, float* d_value, float* d_reward, float* d_priors, void* stream) {
  MZB_CHECK_ARG(m && d_state_in && d_action, "NULL argument");
  MZB_CHECK_ARG(B > 0 && B < (1ll << 31), "batch out of range: %lld", (long long)B);
  RowIO io{};
  io.in = d_state_in; io.in_row_stride = in_row_stride; io.in_slot = d_in_slot; io.slot_stride = slot_stride;
  io.out = d_state_out; io.out_row_stride = out_row_stride; io.out_off = out_offset;
  io.value_logits = d_value_logits; io.reward_logits = d_reward_logits; io.policy_logits = d_policy_logits;
  io.value = d_value; io.reward = d_reward; io.priors = d_priors;
  const int nt = m->rows_per_block;
  k_fc_recurrent<<<(unsigned)((B + nt - 1) / nt), nt, m->smem_bytes, (cudaStream_t)stream>>>(m->d, m->d_pack, (int)B,
                                                                                              d_action, io);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

__global__ void k_u8_to_unit_float(const uint8_t* __restrict__ in, long long n, float* __restrict__ out) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (i + 16 <= n) {
    const uint4 v = *reinterpret_cast<const uint4*>(in + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      reinterpret_cast<float4*>(out + i)[k] = make_float4(__fdiv_rn((float)(w[k] & 255u), 255.0f), __fdiv_rn((float)((w[k] >> 8) & 255u), 255.0f),
                                                          __fdiv_rn((float)((w[k] >> 16) & 255u), 255.0f), __fdiv_rn((float)(w[k] >> 24), 255.0f));
  } else {
    for (long long k = i; k < n; ++k) out[k] = __fdiv_rn((float)in[k], 255.0f);
  }
}

int mzb_u8_to_unit_float(const uint8_t* d_in, int64_t n, float* d_out, void* stream) {
  MZB_CHECK_ARG(d_in && d_out && n > 0, "bad argument");
  MZB_CHECK_ARG(((uintptr_t)d_in & 15) == 0 && ((uintptr_t)d_out & 15) == 0, "buffers must be 16-byte aligned");
  const long long threads = (n + 15) / 16;
  k_u8_to_unit_float<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_in, n, d_out);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

}  // extern "C"

// Internal (mzb_search_fc's modular path): hidden states read from / written to the tree store's blocked slots
// [G/32][S1][32][H] - row g, slot n lives at ((g >> 5) * S1 * 32 + n * 32 + (g & 31)) * H.
int mzb_fc_initial_tree(mzb_fc_model* m, int64_t B, const float* d_obs, const uint8_t* d_legal, float* d_hidden, int S1,
                        int out_slot, float* d_value, float* d_reward, float* d_priors, void* stream) {
  MZB_CHECK_ARG(m && d_obs && d_hidden, "NULL argument");
  RowIO io{};
  const long long H = m->d.enc;
  io.out = d_hidden; io.out_row_stride = H; io.out_blk = (long long)S1 * 32 * H; io.out_off = (long long)out_slot * 32 * H;
  io.value = d_value; io.reward = d_reward; io.priors = d_priors;
  const int nt = m->rows_per_block;
  k_fc_initial<<<(unsigned)((B + nt - 1) / nt), nt, m->smem_bytes, (cudaStream_t)stream>>>(m->d, m->d_pack, (int)B, d_obs,
                                                                                            d_legal, io);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

int mzb_fc_recurrent_tree(mzb_fc_model* m, int64_t B, float* d_hidden, int S1, const int32_t* d_in_slot,
                          const int32_t* d_action, int out_slot, float* d_value, float* d_reward, float* d_priors,
                          void* stream) {
  MZB_CHECK_ARG(m && d_hidden && d_in_slot && d_action, "NULL argument");
  RowIO io{};
  const long long H = m->d.enc;
  io.in = d_hidden; io.in_row_stride = H; io.in_blk = (long long)S1 * 32 * H; io.in_slot = d_in_slot; io.slot_stride = 32 * H;
  io.out = d_hidden; io.out_row_stride = H; io.out_blk = io.in_blk; io.out_off = (long long)out_slot * 32 * H;
  io.value = d_value; io.reward = d_reward; io.priors = d_priors;
  const int nt = m->rows_per_block;
  k_fc_recurrent<<<(unsigned)((B + nt - 1) / nt), nt, m->smem_bytes, (cudaStream_t)stream>>>(m->d, m->d_pack, (int)B,
                                                                                              d_action, io);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

extern "C" {

int mzb_support_to_scalar(const float* d_logits, int64_t B, int support_size, float* d_out, void* stream) {
  MZB_CHECK_ARG(d_logits && d_out && B > 0 && support_size > 0, "bad argument");
  k_support_to_scalar<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_logits, B, support_size, d_out);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

int mzb_scalar_to_support(const float* d_x, int64_t n, int support_size, float* d_out, void* stream) {
  MZB_CHECK_ARG(d_x && d_out && n > 0 && support_size > 0, "bad argument");
  k_scalar_to_support<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_x, n, support_size, d_out);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

}  // extern "C"
