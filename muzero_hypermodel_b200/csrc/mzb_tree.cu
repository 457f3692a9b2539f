// Modular MCTS tree kernels: root_init (K0), select (K1), expand_backup (K3), root_stats.
// A group of LPG lanes (2..32, power of two) owns one game; the lanes of a group stride over the
// A edges of a node (coalesced loads of one contiguous record) and reduce with warp shuffles.
// Reference semantics: self_play.py:261-431 (MCTS), :434-477 (Node), :551-568 (MinMaxStats).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "mzb_tree.cuh"

namespace {

constexpr int kThreads = 128;

template <int LPG>
__device__ __forceinline__ unsigned group_mask() {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned base = lane & ~(unsigned)(LPG - 1);
  return (LPG == 32 ? 0xFFFFFFFFu : ((1u << LPG) - 1u)) << base;
}

template <int LPG>
__device__ __forceinline__ double group_max(double v, unsigned mask) {
#pragma unroll
  for (int o = LPG / 2; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(mask, v, o);
    v = other > v ? other : v;
  }
  return v;
}
template <int LPG>
__device__ __forceinline__ float group_maxf(float v, unsigned mask) {
#pragma unroll
  for (int o = LPG / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(mask, v, o));
  return v;
}
template <int LPG>
__device__ __forceinline__ int group_sum(int v, unsigned mask) {
#pragma unroll
  for (int o = LPG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}
template <int LPG>
__device__ __forceinline__ unsigned group_ballot(bool p, unsigned mask) {
  const unsigned base = (threadIdx.x & 31u) & ~(unsigned)(LPG - 1);
  const unsigned b = __ballot_sync(mask, p);
  return LPG == 32 ? b : ((b >> base) & ((1u << LPG) - 1u));
}

// float32 softmax over the marked entries, summed in ACTION ORDER (the fused thread-per-game kernels
// do the same, so both paths produce identical priors).  `dst` receives e/sum (0 where !use).
template <int LPG>
__device__ __forceinline__ void group_softmax(const float* __restrict__ logits, const int* __restrict__ child_or_null,
                                              float* dst, int A, int lane, unsigned mask) {
  float m = -CUDART_INF_F;
  for (int a = lane; a < A; a += LPG) {
    const bool use = !child_or_null || child_or_null[a] != MZB_CHILD_ILLEGAL;
    if (use) m = fmaxf(m, logits[a]);
  }
  m = group_maxf<LPG>(m, mask);
  for (int a = lane; a < A; a += LPG) {
    const bool use = !child_or_null || child_or_null[a] != MZB_CHILD_ILLEGAL;
    dst[a] = use ? softmax_exp(logits[a], m) : 0.0f;
  }
  __syncwarp(mask);
  float sum = 0.0f;
  for (int a = 0; a < A; ++a) sum = __fadd_rn(sum, dst[a]);     // every lane, same order
  __syncwarp(mask);
  for (int a = lane; a < A; a += LPG) dst[a] = __fdiv_rn(dst[a], sum);
  __syncwarp(mask);
}

// ------------------------------------------------------------------------------------------ K0
template <int LPG>
__global__ void __launch_bounds__(kThreads) k_root_init(TreeView t, const float* __restrict__ reward,
                                                        const float* __restrict__ policy, int is_logits,
                                                        const uint8_t* __restrict__ legal,
                                                        const int8_t* __restrict__ to_play,
                                                        const double* __restrict__ noise, double alpha, double frac,
                                                        const uint32_t* __restrict__ slot,
                                                        const uint32_t* __restrict__ step) {
  const int g = blockIdx.x * (kThreads / LPG) + threadIdx.x / LPG;
  if (g >= t.G) return;
  const int lane = threadIdx.x % LPG;
  const unsigned mask = group_mask<LPG>();
  const int A = t.A;
  uint8_t* r = t.rec(g, 0);
  double* vs = t.value_sum(r);
  float* pr = t.prior(r);
  int* vi = t.visit(r);
  float* rw = t.reward(r);
  int* ch = t.child(r);
  const uint32_t my_slot = slot ? slot[g] : (uint32_t)g;
  const uint32_t my_step = step ? step[g] : 0u;
  for (int a = lane; a < A; a += LPG) {
    vs[a] = 0.0;
    vi[a] = 0;
    rw[a] = 0.0f;
    ch[a] = (!legal || legal[(size_t)g * A + a]) ? MZB_CHILD_NONE : MZB_CHILD_ILLEGAL;
  }
  __syncwarp(mask);
  const float* pol = policy + (size_t)g * A;
  if (is_logits) {
    group_softmax<LPG>(pol, ch, pr, A, lane, mask);
  } else {
    for (int a = lane; a < A; a += LPG) pr[a] = ch[a] == MZB_CHILD_ILLEGAL ? 0.0f : pol[a];
    __syncwarp(mask);
  }
  double* rp = t.root_prior + (size_t)g * A;
  if (frac > 0.0) {
    const double keep = __dsub_rn(1.0, frac);
    if (noise) {
      for (int a = lane; a < A; a += LPG)
        rp[a] = ch[a] == MZB_CHILD_ILLEGAL ? 0.0
                                           : __dadd_rn(__dmul_rn((double)pr[a], keep), __dmul_rn(noise[(size_t)g * A + a], frac));
    } else {
      // device-generated Dirichlet(alpha) over the legal actions: normalised Gamma draws
      for (int a = lane; a < A; a += LPG)
        rp[a] = ch[a] == MZB_CHILD_ILLEGAL ? 0.0 : gamma_sample(t.key, my_slot, my_step, (uint32_t)a, alpha);
      __syncwarp(mask);
      double sum = 0.0;
      for (int a = 0; a < A; ++a) sum = __dadd_rn(sum, rp[a]);
      __syncwarp(mask);
      for (int a = lane; a < A; a += LPG)
        rp[a] = ch[a] == MZB_CHILD_ILLEGAL
                    ? 0.0
                    : __dadd_rn(__dmul_rn((double)pr[a], keep), __dmul_rn(__ddiv_rn(rp[a], sum), frac));
    }
  } else {
    for (int a = lane; a < A; a += LPG) rp[a] = (double)pr[a];
  }
  if (lane == 0) {
    t.root_value_sum[g] = 0.0;
    t.vmin[g] = CUDART_INF;
    t.vmax[g] = -CUDART_INF;
    t.root_reward[g] = reward ? reward[g] : 0.0f;
    t.root_visit[g] = 0;
    t.path_len[g] = 0;
    t.max_depth[g] = 0;
    t.sims_done[g] = 0;
    t.slot[g] = my_slot;
    t.step[g] = my_step;
    t.to_play[g] = to_play ? to_play[g] : (int8_t)0;
  }
}

// ------------------------------------------------------------------------------------------ K1
template <int LPG>
__global__ void __launch_bounds__(kThreads) k_select(TreeView t, int* __restrict__ out_parent,
                                                     int* __restrict__ out_action, int* __restrict__ out_depth) {
  const int g = blockIdx.x * (kThreads / LPG) + threadIdx.x / LPG;
  if (g >= t.G) return;
  const int lane = threadIdx.x % LPG;
  const unsigned mask = group_mask<LPG>();
  const int A = t.A;
  const bool two = t.P == 2;
  const double vmin = t.vmin[g], vmax = t.vmax[g];
  const double* rp = t.root_prior + (size_t)g * A;
  const int sim = t.sims_done[g];
  uint32_t* path = t.path + (size_t)g * (t.S + 1);
  int node = 0, N = t.root_visit[g], depth = 0, action = 0;
  while (true) {
    uint8_t* r = t.rec(g, node);
    const double* vs = t.value_sum(r);
    const float* pr = t.prior(r);
    const int* vi = t.visit(r);
    const float* rw = t.reward(r);
    const int* ch = t.child(r);
    const double pbc0 = t.log_lut[N];
    const double sqrtN = __dsqrt_rn((double)N);
    // pass 1: lane-local best score, number of local maxima, first local argmax
    double best = -CUDART_INF;
    int n_best = 0, first = -1;
    for (int a = lane; a < A; a += LPG) {
      if (ch[a] == MZB_CHILD_ILLEGAL) continue;
      const double p = node == 0 ? rp[a] : (double)pr[a];
      const double s = ucb_score(pbc0, sqrtN, vi[a], p, vs[a], (double)rw[a], t.discount, two, vmin, vmax);
      if (s > best || first < 0) {
        best = s; n_best = 1; first = a;
      } else if (s == best) {
        ++n_best;
      }
    }
    const double gbest = group_max<LPG>(first < 0 ? -CUDART_INF : best, mask);
    const bool mine = first >= 0 && best == gbest;
    const int ties = group_sum<LPG>(mine ? n_best : 0, mask);
    if (ties == 1) {
      const unsigned who = group_ballot<LPG>(mine, mask);
      const int src = (threadIdx.x & 31 & ~(LPG - 1)) + (__ffs(who) - 1);
      action = __shfl_sync(mask, first, src);
    } else {
      // rare (first simulation at the root, symmetric priors): numpy.random.choice over the tied
      // maxima in child order (:372-378), draw replaced by the counter-based rule.
      uint32_t pick = rng_tie_index(t.key, t.slot[g], t.step[g], (uint32_t)sim, (uint32_t)depth, (uint32_t)ties);
      action = 0;
      for (int base = 0; base < A; base += LPG) {
        const int a = base + lane;
        bool tie = false;
        if (a < A && ch[a] != MZB_CHILD_ILLEGAL) {
          const double p = node == 0 ? rp[a] : (double)pr[a];
          tie = ucb_score(pbc0, sqrtN, vi[a], p, vs[a], (double)rw[a], t.discount, two, vmin, vmax) == gbest;
        }
        const unsigned m = group_ballot<LPG>(tie, mask);
        const uint32_t c = __popc(m);
        if (pick < c) {
          action = base + (int)__fns(m, 0, (int)pick + 1);
          break;
        }
        pick -= c;
      }
    }
    if (lane == 0) path[depth] = ((uint32_t)node << 16) | (uint32_t)action;
    ++depth;
    const int next = ch[action];
    if (next < 0) break;
    N = vi[action];
    node = next;
  }
  if (lane == 0) {
    atomicAdd(t.counters, (unsigned long long)depth);
    atomicAdd(t.counters + 1, 1ull);
    t.path_len[g] = depth;
    if (out_parent) out_parent[g] = node;
    if (out_action) out_action[g] = action;
    if (out_depth) out_depth[g] = depth;
  }
}

// ------------------------------------------------------------------------------------------ K3
template <int LPG>
__global__ void __launch_bounds__(kThreads) k_expand_backup(TreeView t, const float* __restrict__ value,
                                                            const float* __restrict__ reward,
                                                            const float* __restrict__ policy, int is_logits) {
  const int g = blockIdx.x * (kThreads / LPG) + threadIdx.x / LPG;
  if (g >= t.G) return;
  const int lane = threadIdx.x % LPG;
  const unsigned mask = group_mask<LPG>();
  const int A = t.A;
  const bool two = t.P == 2;
  const int L = t.path_len[g];
  const int fresh = t.sims_done[g] + 1;
  if (fresh > t.S || L <= 0) return;                // tree full / no pending select: nothing to do
  const uint32_t* path = t.path + (size_t)g * (t.S + 1);
  // --- expand: new node record + link from the parent edge (Node.expand :452-466)
  uint8_t* r = t.rec(g, fresh);
  double* vs = t.value_sum(r);
  float* pr = t.prior(r);
  int* vi = t.visit(r);
  float* rw = t.reward(r);
  int* ch = t.child(r);
  for (int a = lane; a < A; a += LPG) {
    vs[a] = 0.0; vi[a] = 0; rw[a] = 0.0f; ch[a] = MZB_CHILD_NONE;
  }
  const float* pol = policy + (size_t)g * A;
  if (is_logits) {
    group_softmax<LPG>(pol, nullptr, pr, A, lane, mask);
  } else {
    for (int a = lane; a < A; a += LPG) pr[a] = pol[a];
  }
  const float leaf_reward = reward[g];
  // --- backup along the path, leaf first (backpropagate :407-431).  Lane k of each chunk owns edge
  // (chunk_base + k): it loads that edge's statistics up front (independent, coalesced-ish loads),
  // then the float64 recurrence runs serially through shuffles so the rounding order is the reference's.
  double val = (double)value[g];
  double vmin = t.vmin[g], vmax = t.vmax[g];
  for (int hi = L; hi > 0; hi -= LPG) {
    const int lo = hi - LPG > 0 ? hi - LPG : 0;
    const int k = lo + lane;                        // my edge index in the path
    double e_vs = 0.0; int e_vi = 0; double e_rw = 0.0;
    double* p_vs = nullptr; int* p_vi = nullptr;
    if (k < hi) {
      const uint32_t pe = path[k];
      const int pn = pe >> 16, pa = pe & 0xFFFF;
      uint8_t* er = t.rec(g, pn);
      p_vs = t.value_sum(er) + pa;
      p_vi = t.visit(er) + pa;
      e_vs = *p_vs;
      e_vi = *p_vi;
      if (k == L - 1) {                             // the edge that now leads to the fresh node
        t.reward(er)[pa] = leaf_reward;
        t.child(er)[pa] = fresh;
        e_rw = (double)leaf_reward;
      } else {
        e_rw = (double)t.reward(er)[pa];
      }
    }
    for (int j = hi - 1; j >= lo; --j) {
      const int src = (threadIdx.x & 31 & ~(LPG - 1)) + (j - lo);
      double s_vs = __shfl_sync(mask, e_vs, src);
      int s_vi = __shfl_sync(mask, e_vi, src);
      const double s_rw = __shfl_sync(mask, e_rw, src);
      // node reached by edge j sits at depth j+1; the leaf at depth L
      const bool same = ((L - (j + 1)) & 1) == 0;
      backup_step(s_vs, s_vi, s_rw, val, t.discount, two, same, vmin, vmax);
      if (k == j) { e_vs = s_vs; e_vi = s_vi; }
    }
    if (k < hi) { *p_vs = e_vs; *p_vi = e_vi; }
  }
  if (lane == 0) {
    double rvs = t.root_value_sum[g];
    int rvi = t.root_visit[g];
    backup_step(rvs, rvi, (double)t.root_reward[g], val, t.discount, two, (L & 1) == 0, vmin, vmax);
    t.root_value_sum[g] = rvs;
    t.root_visit[g] = rvi;
    t.vmin[g] = vmin;
    t.vmax[g] = vmax;
    t.sims_done[g] = fresh;
    if (L > t.max_depth[g]) t.max_depth[g] = L;
    t.path_len[g] = 0;
  }
}

// ------------------------------------------------------------------------------------------ stats
__global__ void k_root_stats(TreeView t, int* __restrict__ visits, double* __restrict__ root_value,
                             int* __restrict__ max_depth, double* __restrict__ cvs, float* __restrict__ crw,
                             double* __restrict__ cpr, double* __restrict__ minmax) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)t.G * t.A;
  if (i < total) {
    const int g = (int)(i / t.A), a = (int)(i % t.A);
    uint8_t* r = t.rec(g, 0);
    const bool legal = t.child(r)[a] != MZB_CHILD_ILLEGAL;
    if (visits) visits[i] = legal ? t.visit(r)[a] : 0;
    if (cvs) cvs[i] = legal ? t.value_sum(r)[a] : 0.0;
    if (crw) crw[i] = legal ? t.reward(r)[a] : 0.0f;
    if (cpr) cpr[i] = legal ? t.root_prior[i] : 0.0;
  }
  if (i < (size_t)t.G) {
    const int g = (int)i;
    const int n = t.root_visit[g];
    if (root_value) root_value[g] = n > 0 ? __ddiv_rn(t.root_value_sum[g], (double)n) : 0.0;   // Node.value :446-449
    if (max_depth) max_depth[g] = t.max_depth[g];
    if (minmax) { minmax[2 * g] = t.vmin[g]; minmax[2 * g + 1] = t.vmax[g]; }
  }
}

int pick_lpg(int A) {
  int l = 2;
  while (l < A && l < 32) l <<= 1;
  return l;
}

struct Offsets {
  size_t nodes, root_prior, path, rvs, vmin, vmax, rrew, rvis, plen, mdep, sdone, slot, step, toplay, hidden;
  size_t tmp_parent, tmp_action, tmp_value, tmp_reward, tmp_priors, counters, total;
};

Offsets layout(const mzb_tree_config& c) {
  Offsets o;
  const size_t G = c.n_games, A = c.n_actions, S1 = (size_t)c.num_simulations + 1;
  const size_t Gp = (G + 31) / 32 * 32;          // node records and hidden slots are blocked by 32 games
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t at = off; off = mzb_align_up(off + bytes, 256); return at; };
  o.nodes = take(Gp * S1 * 24 * A);
  o.root_prior = take(G * A * 8);
  o.path = take(G * S1 * 4);
  o.rvs = take(G * 8); o.vmin = take(G * 8); o.vmax = take(G * 8);
  o.rrew = take(G * 4); o.rvis = take(G * 4); o.plen = take(G * 4); o.mdep = take(G * 4); o.sdone = take(G * 4);
  o.slot = take(G * 4); o.step = take(G * 4); o.toplay = take(G);
  o.hidden = take(Gp * S1 * (size_t)c.hidden_floats * 4);
  o.tmp_parent = take(G * 4); o.tmp_action = take(G * 4); o.tmp_value = take(G * 4); o.tmp_reward = take(G * 4);
  o.tmp_priors = take(G * A * 4);
  o.counters = take(64);
  o.total = off;
  return o;
}

int validate(const mzb_tree_config* c) {
  MZB_CHECK_ARG(c, "config is NULL");
  MZB_CHECK_ARG(c->n_games > 0, "n_games must be positive, got %d", c->n_games);
  MZB_CHECK_ARG(c->n_actions > 0 && c->n_actions <= 65535, "n_actions out of range: %d", c->n_actions);
  MZB_CHECK_ARG(c->num_simulations > 0 && c->num_simulations <= 65534, "num_simulations out of range: %d",
                c->num_simulations);
  MZB_CHECK_ARG(c->hidden_floats >= 0, "hidden_floats negative");
  if (c->n_players != 1 && c->n_players != 2) {
    mzb_set_error("More than two player mode not implemented.");        // self_play.py:431
    return MZB_EUNSUPPORTED;
  }
  return MZB_OK;
}

}  // namespace

#define DISPATCH_LPG(lpg, KERNEL, grid, stream, ...)                                              \
  switch (lpg) {                                                                                  \
    case 2: KERNEL<2><<<grid, kThreads, 0, stream>>>(__VA_ARGS__); break;                         \
    case 4: KERNEL<4><<<grid, kThreads, 0, stream>>>(__VA_ARGS__); break;                         \
    case 8: KERNEL<8><<<grid, kThreads, 0, stream>>>(__VA_ARGS__); break;                         \
    case 16: KERNEL<16><<<grid, kThreads, 0, stream>>>(__VA_ARGS__); break;                       \
    default: KERNEL<32><<<grid, kThreads, 0, stream>>>(__VA_ARGS__); break;                       \
  }

extern "C" {

size_t mzb_tree_workspace_bytes(const mzb_tree_config* cfg) {
  if (validate(cfg) != MZB_OK) return 0;
  return layout(*cfg).total;
}

int mzb_tree_create(mzb_tree** out, const mzb_tree_config* cfg, void* d_workspace, size_t workspace_bytes,
                    const double* h_log_lut) {
  MZB_CHECK_ARG(out, "out is NULL");
  *out = nullptr;
  int rc = validate(cfg);
  if (rc != MZB_OK) return rc;
  const Offsets o = layout(*cfg);
  MZB_CHECK_ARG(d_workspace, "workspace is NULL");
  MZB_CHECK_ARG(((uintptr_t)d_workspace & 255) == 0, "workspace must be 256-byte aligned");
  MZB_CHECK_ARG(workspace_bytes >= o.total, "workspace too small: %zu < %zu", workspace_bytes, o.total);
  mzb_tree* t = new mzb_tree();
  t->cfg = *cfg;
  t->bytes = o.total;
  t->lpg = pick_lpg(cfg->n_actions);
  const int S1 = cfg->num_simulations + 1;
  std::vector<double> lut(S1);
  for (int n = 0; n < S1; ++n)
    lut[n] = h_log_lut ? h_log_lut[n] : log(((double)n + cfg->pb_c_base + 1.0) / cfg->pb_c_base) + cfg->pb_c_init;
  cudaError_t e = cudaMalloc(&t->d_log_lut, sizeof(double) * S1);
  if (e == cudaSuccess) e = cudaMemcpy(t->d_log_lut, lut.data(), sizeof(double) * S1, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    mzb_set_error("log LUT upload: %s", cudaGetErrorString(e));
    delete t;
    return MZB_ECUDA;
  }
  uint8_t* w = (uint8_t*)d_workspace;
  TreeView& v = t->v;
  v.G = cfg->n_games; v.A = cfg->n_actions; v.S = cfg->num_simulations; v.P = cfg->n_players; v.H = cfg->hidden_floats;
  v.discount = cfg->discount;
  v.nodes = w + o.nodes;
  v.rec_bytes = 24 * (size_t)cfg->n_actions;
  v.blk_stride = v.rec_bytes * S1 * 32;
  v.root_prior = (double*)(w + o.root_prior);
  v.path = (uint32_t*)(w + o.path);
  v.root_value_sum = (double*)(w + o.rvs);
  v.vmin = (double*)(w + o.vmin);
  v.vmax = (double*)(w + o.vmax);
  v.root_reward = (float*)(w + o.rrew);
  v.root_visit = (int*)(w + o.rvis);
  v.path_len = (int*)(w + o.plen);
  v.max_depth = (int*)(w + o.mdep);
  v.sims_done = (int*)(w + o.sdone);
  v.slot = (uint32_t*)(w + o.slot);
  v.step = (uint32_t*)(w + o.step);
  v.to_play = (int8_t*)(w + o.toplay);
  v.hidden = cfg->hidden_floats ? (float*)(w + o.hidden) : nullptr;
  v.log_lut = t->d_log_lut;
  t->tmp_parent = (int*)(w + o.tmp_parent); t->tmp_action = (int*)(w + o.tmp_action);
  t->tmp_value = (float*)(w + o.tmp_value); t->tmp_reward = (float*)(w + o.tmp_reward);
  t->tmp_priors = (float*)(w + o.tmp_priors);
  v.key = rng_key(cfg->seed);
  v.counters = (unsigned long long*)(w + o.counters);
  cudaMemset(v.counters, 0, 64);
  *out = t;
  return MZB_OK;
}

int mzb_tree_destroy(mzb_tree* t) {
  if (!t) return MZB_OK;
  mzb_search_graph_forget(t);
  cudaFree(t->d_log_lut);
  delete t;
  return MZB_OK;
}

float* mzb_tree_hidden_ptr(mzb_tree* t) { return t ? t->v.hidden : nullptr; }

int mzb_tree_root_init(mzb_tree* t, const float* d_reward, const float* d_policy, int policy_is_logits,
                       const uint8_t* d_legal, const int8_t* d_to_play, const double* d_noise, double alpha,
                       double frac, const uint32_t* d_slot, const uint32_t* d_step, void* stream) {
  MZB_CHECK_ARG(t && d_policy, "tree / policy is NULL");
  MZB_CHECK_ARG(frac >= 0.0 && frac <= 1.0, "exploration fraction out of [0,1]: %f", frac);
  MZB_CHECK_ARG(d_noise || frac == 0.0 || alpha > 0.0, "dirichlet alpha must be positive");
  const int gpb = kThreads / t->lpg;
  const int grid = (t->v.G + gpb - 1) / gpb;
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_LPG(t->lpg, k_root_init, grid, s, t->v, d_reward, d_policy, policy_is_logits, d_legal, d_to_play, d_noise,
               alpha, frac, d_slot, d_step);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

int mzb_tree_select(mzb_tree* t, int32_t* d_parent_slot, int32_t* d_action, int32_t* d_depth, void* stream) {
  MZB_CHECK_ARG(t, "tree is NULL");
  const int gpb = kThreads / t->lpg;
  const int grid = (t->v.G + gpb - 1) / gpb;
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_LPG(t->lpg, k_select, grid, s, t->v, d_parent_slot, d_action, d_depth);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

int mzb_tree_expand_backup(mzb_tree* t, const float* d_value, const float* d_reward, const float* d_policy,
                           int policy_is_logits, void* stream) {
  MZB_CHECK_ARG(t && d_value && d_reward && d_policy, "NULL argument");
  const int gpb = kThreads / t->lpg;
  const int grid = (t->v.G + gpb - 1) / gpb;
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_LPG(t->lpg, k_expand_backup, grid, s, t->v, d_value, d_reward, d_policy, policy_is_logits);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

int mzb_tree_root_stats(mzb_tree* t, int32_t* d_visits, double* d_root_value, int32_t* d_max_depth,
                        double* d_child_value_sum, float* d_child_reward, double* d_child_prior, double* d_minmax,
                        void* stream) {
  MZB_CHECK_ARG(t, "tree is NULL");
  const size_t total = (size_t)t->v.G * t->v.A;
  const int threads = 256;
  const int grid = (int)((total + threads - 1) / threads);
  k_root_stats<<<grid, threads, 0, (cudaStream_t)stream>>>(t->v, d_visits, d_root_value, d_max_depth,
                                                           d_child_value_sum, d_child_reward, d_child_prior, d_minmax);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

int mzb_tree_counters_sync(mzb_tree* t, uint64_t* h_counters2, int reset, void* stream) {
  MZB_CHECK_ARG(t && h_counters2, "NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  MZB_CUDA(cudaMemcpyAsync(h_counters2, t->v.counters, 16, cudaMemcpyDeviceToHost, s));
  if (reset) MZB_CUDA(cudaMemsetAsync(t->v.counters, 0, 16, s));
  MZB_CUDA(cudaStreamSynchronize(s));
  return MZB_OK;
}

// Test hook: the hoisted-reciprocal division of mzb_common.cuh against the compiler's div.rn.f64, element-wise.
static __global__ void k_debug_ddiv_rcp(const double* __restrict__ a, const double* __restrict__ b, long long n,
                                        double* __restrict__ fast, double* __restrict__ ref) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double y = rcp_divisor_ok(b[i]) ? rcp_refined(b[i]) : 0.0;
  double f = rcp_divisor_ok(b[i]) ? ddiv_rcp(a[i], b[i], y) : __ddiv_rn(a[i], b[i]);
  const double r = __ddiv_rn(a[i], b[i]);
  // the branch-free form (ddiv_rcp_nb): wherever it declares its fast path valid the quotient must be the same bits;
  // a mismatch is reported as a quotient that cannot equal the reference (bits flipped)
  if (rcp_divisor_ok(b[i])) {
    bool ok;
    const double q = ddiv_rcp_nb(a[i], b[i], y, ok);
    if (ok && __double_as_longlong(q) != __double_as_longlong(r)) f = __longlong_as_double(~__double_as_longlong(r));
  }
  fast[i] = f;
  ref[i] = r;
}

int mzb_debug_ddiv_rcp(const double* d_a, const double* d_b, int64_t n, double* d_fast, double* d_ref, void* stream) {
  MZB_CHECK_ARG(d_a && d_b && d_fast && d_ref && n > 0, "bad argument");
  k_debug_ddiv_rcp<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_a, d_b, n, d_fast, d_ref);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

int mzb_tree_export_game_sync(mzb_tree* t, int32_t game, double* h_value_sum, float* h_prior, int32_t* h_visit,
                              float* h_reward, int32_t* h_child, double* h_root_prior, double* h_scalars,
                              void* stream) {
  MZB_CHECK_ARG(t, "tree is NULL");
  MZB_CHECK_ARG(game >= 0 && game < t->v.G, "game index %d out of range", game);
  cudaStream_t s = (cudaStream_t)stream;
  const TreeView& v = t->v;
  const int S1 = v.S + 1, A = v.A;
  std::vector<uint8_t> buf(v.rec_bytes * S1);      // the game's records are 32 records apart (blocked layout)
  MZB_CUDA(cudaMemcpy2DAsync(buf.data(), v.rec_bytes, v.nodes + v.rec_index(game, 0) * v.rec_bytes, 32 * v.rec_bytes,
                             v.rec_bytes, S1, cudaMemcpyDeviceToHost, s));
  double rvs; float rrw; int rvi, sd;
  MZB_CUDA(cudaMemcpyAsync(&rvs, v.root_value_sum + game, 8, cudaMemcpyDeviceToHost, s));
  MZB_CUDA(cudaMemcpyAsync(&rrw, v.root_reward + game, 4, cudaMemcpyDeviceToHost, s));
  MZB_CUDA(cudaMemcpyAsync(&rvi, v.root_visit + game, 4, cudaMemcpyDeviceToHost, s));
  MZB_CUDA(cudaMemcpyAsync(&sd, v.sims_done + game, 4, cudaMemcpyDeviceToHost, s));
  if (h_root_prior) MZB_CUDA(cudaMemcpyAsync(h_root_prior, v.root_prior + (size_t)game * A, 8 * (size_t)A, cudaMemcpyDeviceToHost, s));
  MZB_CUDA(cudaStreamSynchronize(s));
  for (int n = 0; n < S1; ++n) {
    const uint8_t* r = buf.data() + (size_t)n * v.rec_bytes;
    if (h_value_sum) memcpy(h_value_sum + (size_t)n * A, r, 8 * (size_t)A);
    if (h_prior) memcpy(h_prior + (size_t)n * A, r + 8 * (size_t)A, 4 * (size_t)A);
    if (h_visit) memcpy(h_visit + (size_t)n * A, r + 12 * (size_t)A, 4 * (size_t)A);
    if (h_reward) memcpy(h_reward + (size_t)n * A, r + 16 * (size_t)A, 4 * (size_t)A);
    if (h_child) memcpy(h_child + (size_t)n * A, r + 20 * (size_t)A, 4 * (size_t)A);
  }
  if (h_scalars) { h_scalars[0] = rvi; h_scalars[1] = rvs; h_scalars[2] = rrw; h_scalars[3] = sd + 1; }
  return MZB_OK;
}

}  // extern "C"
