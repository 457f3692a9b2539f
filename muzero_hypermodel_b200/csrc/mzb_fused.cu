// K12: whole-search kernel for fully-connected MuZero configs - one launch runs MCTS.run
// (self_play.py:261-362) for every game: initial inference, root expansion + exploration noise,
// num_simulations x {pUCT select walk, recurrent inference, expand, value backup}.
//
// One THREAD owns one game.  The network weights (6-16 KB) live in shared memory and every lane of a
// warp reads the same weight, so a broadcast LDS.128 feeds four FMAs; activations stay in registers
// (layer widths are template constants, loops fully unrolled).  The tree is the same HBM store the
// modular kernels use (mzb_tree.cuh): a thread reads one node's edges as one contiguous record.
// The float32/float64 arithmetic is shared with the batched kernels (mzb_fc.cuh, mzb_common.cuh), so
// this path and the modular path (K5, K0, K1+K4+K3 per simulation) give bit-identical trees.
#include <stdlib.h>

#include "mzb_fc.cuh"
#include "mzb_tree.cuh"

namespace {

__host__ __device__ constexpr int pad4(int x) { return (x + 3) / 4 * 4; }

// ---- one Linear layer, compile-time shape.  w: Wt [*][OUTP] in shared memory, b: [OUTP].
template <int NIN, int OUT, int OUTP>
__device__ __forceinline__ void lin(const float* __restrict__ w, const float* __restrict__ b, const float (&x)[NIN],
                                    int hot, float (&y)[OUT], bool elu) {
#pragma unroll
  for (int o = 0; o < OUT; ++o) y[o] = b[o];
#pragma unroll
  for (int i = 0; i < NIN; ++i) {
#pragma unroll
    for (int o = 0; o < OUT; ++o) y[o] = fmaf(x[i], w[i * OUTP + o], y[o]);
  }
  if (hot >= 0) {
    const float* wr = w + (NIN + hot) * OUTP;
#pragma unroll
    for (int o = 0; o < OUT; ++o) y[o] = __fadd_rn(y[o], wr[o]);
  }
  if (elu) {
#pragma unroll
    for (int o = 0; o < OUT; ++o) y[o] = elu_f32(y[o]);
  }
}

// ---- mlp(): IN inputs (NDIRECT dense + optional one-hot tail), one optional hidden layer H, OUT outputs.
template <int NDIRECT, int IN, int H, int OUT>
struct Mlp {
  static constexpr int OUT0 = H > 0 ? H : OUT;
  static constexpr int SIZE0 = IN * pad4(OUT0) + pad4(OUT0);
  static constexpr int SIZE = SIZE0 + (H > 0 ? H * pad4(OUT) + pad4(OUT) : 0);
  __device__ __forceinline__ static void run(const float* __restrict__ p, const float (&x)[NDIRECT], int hot,
                                             float (&y)[OUT]) {
    if constexpr (H > 0) {
      float h[H];
      lin<NDIRECT, H, pad4(H)>(p, p + IN * pad4(H), x, hot, h, true);
      lin<H, OUT, pad4(OUT)>(p + SIZE0, p + SIZE0 + H * pad4(OUT), h, -1, y, false);
    } else {
      lin<NDIRECT, OUT, pad4(OUT)>(p, p + IN * pad4(OUT), x, hot, y, false);
    }
  }
};

template <int OBS_, int ENC_, int A_, int SUP_, int REP_H, int DYN_H, int REW_H, int VAL_H, int POL_H>
struct Shape {
  static constexpr int OBS = OBS_, ENC = ENC_, A = A_, SUP = SUP_, FULL = 2 * SUP_ + 1;
  using Rep = Mlp<OBS, OBS, REP_H, ENC>;
  using Dyn = Mlp<ENC, ENC + A, DYN_H, ENC>;
  using Rew = Mlp<ENC, ENC, REW_H, FULL>;
  using Pol = Mlp<ENC, ENC, POL_H, A>;
  using Val = Mlp<ENC, ENC, VAL_H, FULL>;
  // pack order = add_net order in mzb_fc.cu: rep, dyn, rew, pol, val
  static constexpr int OFF_REP = 0;
  static constexpr int OFF_DYN = OFF_REP + Rep::SIZE;
  static constexpr int OFF_REW = OFF_DYN + Dyn::SIZE;
  static constexpr int OFF_POL = OFF_REW + Rew::SIZE;
  static constexpr int OFF_VAL = OFF_POL + Pol::SIZE;
  static constexpr int PACK = OFF_VAL + Val::SIZE;
  static bool matches(const FcDesc& d) {
    auto one = [](const FcNet& n, int h) { return h > 0 ? (n.n == 2 && n.l[0].out == h) : n.n == 1; };
    return d.obs_dim == OBS && d.enc == ENC && d.A == A && d.S == SUP && d.pack_floats == PACK && one(d.rep, REP_H) &&
           one(d.dyn, DYN_H) && one(d.rew, REW_H) && one(d.val, VAL_H) && one(d.pol, POL_H);
  }
};

template <int N>
__device__ __forceinline__ void minmax_regs(float (&s)[N]) {
  float lo = CUDART_INF_F, hi = -CUDART_INF_F;
#pragma unroll
  for (int i = 0; i < N; ++i) { lo = fminf(lo, s[i]); hi = fmaxf(hi, s[i]); }
  float scale = __fsub_rn(hi, lo);
  if (scale < 1e-5f) scale = __fadd_rn(scale, 1e-5f);
#pragma unroll
  for (int i = 0; i < N; ++i) s[i] = __fdiv_rn(__fsub_rn(s[i], lo), scale);
}

template <int SUP>
__device__ __forceinline__ float s2s_regs(const float (&l)[2 * SUP + 1]) {
  constexpr int FULL = 2 * SUP + 1;
  float e[FULL];
  float m = -CUDART_INF_F;
#pragma unroll
  for (int i = 0; i < FULL; ++i) m = fmaxf(m, l[i]);
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < FULL; ++i) { e[i] = softmax_exp(l[i], m); sum = __fadd_rn(sum, e[i]); }
  float num = 0.0f;
#pragma unroll
  for (int i = 0; i < FULL; ++i) num = fmaf((float)(i - SUP), e[i], num);
  return inverse_value_transform(__fdiv_rn(num, sum));
}

// softmax over the legal actions (all when legal == NULL), summed in action order; 0 elsewhere
template <int A>
__device__ __forceinline__ void priors_regs(const float (&logit)[A], const uint8_t* legal, float (&p)[A]) {
  float m = -CUDART_INF_F;
#pragma unroll
  for (int a = 0; a < A; ++a) if (!legal || legal[a]) m = fmaxf(m, logit[a]);
  float sum = 0.0f;
#pragma unroll
  for (int a = 0; a < A; ++a) { p[a] = (!legal || legal[a]) ? softmax_exp(logit[a], m) : 0.0f; sum = __fadd_rn(sum, p[a]); }
#pragma unroll
  for (int a = 0; a < A; ++a) p[a] = __fdiv_rn(p[a], sum);
}

// ---- vectorised tree accesses.  One thread owns one game, so every global access of a warp touches 32 different
// lines: the load/store unit then needs one pass per lane and REQUEST COUNT, not bytes, is what the kernel pays for
// (the scalar form issued ~100 requests per simulation and kept the L1 pipe ~70 % busy).  A node record is 24*A
// contiguous bytes (mzb_tree.cuh), 16-byte aligned when A is even and 8-byte aligned otherwise; hidden states are
// ENC contiguous floats.
template <int A>
struct Rec {
  static constexpr int WORDS = 6 * A;                    // value_sum f64[A] | prior f32[A] | visit i32[A] | reward f32[A] | child i32[A]
  uint32_t w[WORDS];
  __device__ __forceinline__ void load(const uint8_t* r) {
    if constexpr (A % 2 == 0) {
      const uint4* p = reinterpret_cast<const uint4*>(r);
#pragma unroll
      for (int i = 0; i < WORDS / 4; ++i) { const uint4 q = p[i]; w[4 * i] = q.x; w[4 * i + 1] = q.y; w[4 * i + 2] = q.z; w[4 * i + 3] = q.w; }
    } else {
      const uint2* p = reinterpret_cast<const uint2*>(r);
#pragma unroll
      for (int i = 0; i < WORDS / 2; ++i) { const uint2 q = p[i]; w[2 * i] = q.x; w[2 * i + 1] = q.y; }
    }
  }
  __device__ __forceinline__ void store(uint8_t* r) const {
    if constexpr (A % 2 == 0) {
      uint4* p = reinterpret_cast<uint4*>(r);
#pragma unroll
      for (int i = 0; i < WORDS / 4; ++i) p[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    } else {
      uint2* p = reinterpret_cast<uint2*>(r);
#pragma unroll
      for (int i = 0; i < WORDS / 2; ++i) p[i] = make_uint2(w[2 * i], w[2 * i + 1]);
    }
  }
  __device__ __forceinline__ double vs(int a) const { return __hiloint2double((int)w[2 * a + 1], (int)w[2 * a]); }
  __device__ __forceinline__ float pr(int a) const { return __uint_as_float(w[2 * A + a]); }
  __device__ __forceinline__ int vi(int a) const { return (int)w[3 * A + a]; }
  __device__ __forceinline__ float rw(int a) const { return __uint_as_float(w[4 * A + a]); }
  __device__ __forceinline__ int ch(int a) const { return (int)w[5 * A + a]; }
  __device__ __forceinline__ void set(int a, double value_sum, float prior, int visit, float reward, int child) {
    w[2 * a] = (uint32_t)__double2loint(value_sum); w[2 * a + 1] = (uint32_t)__double2hiint(value_sum);
    w[2 * A + a] = __float_as_uint(prior); w[3 * A + a] = (uint32_t)visit; w[4 * A + a] = __float_as_uint(reward);
    w[5 * A + a] = (uint32_t)child;
  }
};

template <int N>
__device__ __forceinline__ void load_floats(const float* __restrict__ p, float (&x)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) { const float4 q = reinterpret_cast<const float4*>(p)[i]; x[4 * i] = q.x; x[4 * i + 1] = q.y; x[4 * i + 2] = q.z; x[4 * i + 3] = q.w; }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = p[i];
  }
}
template <int N>
__device__ __forceinline__ void store_floats(float* __restrict__ p, const float (&x)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = x[i];
  }
}

struct SearchIO {
  const float* obs; const uint8_t* legal; const int8_t* to_play; const double* noise;
  double alpha, frac;
  const uint32_t* slot; const uint32_t* step;
  int num_sims;
  int* visits; double* root_value; float* root_pred_value; int* max_depth;
};

// THREADS per block, PHASE_SYNC: a block barrier before the network phase of every simulation re-aligns the
// block's warps so they fetch the (fully unrolled, ~100 KB) network code together; PB_LUT: the exploration factor
// pb(N, n) = (log((N+base+1)/base)+init) * (sqrt(N)/(n+1)) comes from a shared-memory table built with the same
// three float64 operations (bit-identical), removing a double division and square root per child and level.
template <class SH, int THREADS, bool PHASE_SYNC, bool PB_LUT>
__global__ void __launch_bounds__(THREADS, 512 / THREADS) k_search_fc(TreeView t, const float* __restrict__ gpack, SearchIO io) {
  constexpr int A = SH::A, ENC = SH::ENC, FULL = SH::FULL;
  extern __shared__ float4 smem4[];
  float* pack = reinterpret_cast<float*>(smem4);
  double* lut = reinterpret_cast<double*>(pack + SH::PACK);
  const int S1 = io.num_sims + 1;
  double* pbt = lut + ((S1 + 1) & ~1);                     // [S1][S1] when PB_LUT
  for (int i = threadIdx.x; i < SH::PACK / 4; i += THREADS) smem4[i] = reinterpret_cast<const float4*>(gpack)[i];
  for (int i = threadIdx.x; i < S1; i += THREADS) lut[i] = t.log_lut[i];
  if (PB_LUT) {
    for (int i = threadIdx.x; i < S1 * S1; i += THREADS) {
      const int N = i / S1, n = i % S1;
      pbt[i] = ucb_pb(t.log_lut[N], __dsqrt_rn((double)N), n);
    }
  }
  __syncthreads();
  const int g_raw = blockIdx.x * THREADS + threadIdx.x;
  const bool active = g_raw < t.G;
  if (!PHASE_SYNC && !active) return;
  const int g = active ? g_raw : 0;                        // idle threads of the last block only take part in barriers
  const bool two = t.P == 2;
  const size_t G = (size_t)t.G;
  const uint32_t my_slot = io.slot ? io.slot[g] : (uint32_t)g;
  const uint32_t my_step = io.step ? io.step[g] : 0u;
  const uint8_t* legal = io.legal ? io.legal + (size_t)g * A : nullptr;
  float* hid = t.hidden + (size_t)g * (t.S + 1) * ENC;

  // ---------------- initial inference (models.py:172-190) + root expansion (self_play.py:292-314)
  double rp[A];                       // root priors, float64 after the noise mix
  float root_reward = 0.0f;
  if (active) {
    float ob[SH::OBS];
    load_floats<SH::OBS>(io.obs + (size_t)g * SH::OBS, ob);
    float st[ENC];
    SH::Rep::run(pack + SH::OFF_REP, ob, -1, st);
    minmax_regs(st);
    store_floats<ENC>(hid, st);
    float pl[A], pri[A];
    SH::Pol::run(pack + SH::OFF_POL, st, -1, pl);
    priors_regs<A>(pl, legal, pri);
    if (io.root_pred_value) {
      float vl[FULL];
      SH::Val::run(pack + SH::OFF_VAL, st, -1, vl);
      io.root_pred_value[g] = s2s_regs<SH::SUP>(vl);
    }
    float zl[FULL];
#pragma unroll
    for (int i = 0; i < FULL; ++i) zl[i] = (i == SH::SUP) ? 0.0f : -CUDART_INF_F;
    root_reward = s2s_regs<SH::SUP>(zl);
    // root record
    {
      Rec<A> rr;
#pragma unroll
      for (int a = 0; a < A; ++a) {
        const bool ok = !legal || legal[a];
        rr.set(a, 0.0, ok ? pri[a] : 0.0f, 0, 0.0f, ok ? MZB_CHILD_NONE : MZB_CHILD_ILLEGAL);
      }
      rr.store(t.rec(g, 0));
    }
    if (io.frac > 0.0) {
      const double keep = __dsub_rn(1.0, io.frac);
      double nz[A];
      if (io.noise) {
#pragma unroll
        for (int a = 0; a < A; ++a) nz[a] = io.noise[(size_t)g * A + a];
      } else {
        double sum = 0.0;
#pragma unroll
        for (int a = 0; a < A; ++a) {
          nz[a] = (!legal || legal[a]) ? gamma_sample(t.key, my_slot, my_step, (uint32_t)a, io.alpha) : 0.0;
          sum = __dadd_rn(sum, nz[a]);
        }
#pragma unroll
        for (int a = 0; a < A; ++a) nz[a] = __ddiv_rn(nz[a], sum);
      }
#pragma unroll
      for (int a = 0; a < A; ++a)
        rp[a] = (!legal || legal[a]) ? __dadd_rn(__dmul_rn((double)pri[a], keep), __dmul_rn(nz[a], io.frac)) : 0.0;
    } else {
#pragma unroll
      for (int a = 0; a < A; ++a) rp[a] = (!legal || legal[a]) ? (double)pri[a] : 0.0;
    }
#pragma unroll
    for (int a = 0; a < A; ++a) t.root_prior[(size_t)g * A + a] = rp[a];
  }

  int root_visit = 0, max_depth = 0;
  unsigned int depth_sum = 0;
  double root_vs = 0.0, vmin = CUDART_INF, vmax = -CUDART_INF;
  uint32_t* path = t.path + g;                       // transposed use of the path buffer: path[depth * G]
  // PB_LUT kernels (num_simulations <= 63) keep the search path AND the statistics of the edges they walked in
  // per-thread local memory (L1-resident), so the backup issues no dependent global loads.
  constexpr int LP = PB_LUT ? 64 : 1;
  uint32_t lp_edge[LP]; double lp_vs[LP]; int lp_vi[LP]; float lp_rw[LP];

  for (int sim = 0; sim < io.num_sims; ++sim) {
    // ---------------- select walk (self_play.py:326-335, 364-405)
    int node = 0, N = root_visit, depth = 0, action = 0;
    while (active) {
      uint8_t* r = t.rec(g, node);
      double vs[A]; float pr[A], rw[A]; int vi[A], ch[A];
      {
        Rec<A> rr;
        rr.load(r);
#pragma unroll
        for (int a = 0; a < A; ++a) { vs[a] = rr.vs(a); pr[a] = rr.pr(a); vi[a] = rr.vi(a); rw[a] = rr.rw(a); ch[a] = rr.ch(a); }
      }
      const double pbc0 = PB_LUT ? 0.0 : lut[N];
      const double sqrtN = PB_LUT ? 0.0 : __dsqrt_rn((double)N);
      const double* pbrow = pbt + N * S1;
      double sc[A];
      double best = -CUDART_INF;
      int n_best = 0;
      action = -1;
#pragma unroll
      for (int a = 0; a < A; ++a) {
        if (ch[a] == MZB_CHILD_ILLEGAL) { sc[a] = -CUDART_INF; continue; }
        const double p = node == 0 ? rp[a] : (double)pr[a];
        const double pb = PB_LUT ? pbrow[vi[a]] : ucb_pb(pbc0, sqrtN, vi[a]);
        sc[a] = ucb_score_pb(pb, vi[a], p, vs[a], (double)rw[a], t.discount, two, vmin, vmax);
        if (sc[a] > best || action < 0) { best = sc[a]; n_best = 1; action = a; }
        else if (sc[a] == best) ++n_best;
      }
      if (n_best > 1) {
        int pick = (int)rng_tie_index(t.key, my_slot, my_step, (uint32_t)sim, (uint32_t)depth, (uint32_t)n_best);
#pragma unroll
        for (int a = 0; a < A; ++a) {
          if (ch[a] != MZB_CHILD_ILLEGAL && sc[a] == best) {
            if (pick == 0) action = a;
            --pick;
          }
        }
      }
      int next = MZB_CHILD_NONE, nv = 0;
      double nvs = 0.0; float nrw = 0.0f;
#pragma unroll
      for (int a = 0; a < A; ++a) if (a == action) { next = ch[a]; nv = vi[a]; nvs = vs[a]; nrw = rw[a]; }
      if (PB_LUT) {
        lp_edge[depth] = ((uint32_t)node << 16) | (uint32_t)action; lp_vs[depth] = nvs; lp_vi[depth] = nv; lp_rw[depth] = nrw;
      } else {
        path[(size_t)depth * G] = ((uint32_t)node << 16) | (uint32_t)action;
      }
      ++depth;
      if (next < 0) break;
      N = nv;
      node = next;
    }
    const int L = depth, fresh = sim + 1;
    if (PHASE_SYNC) __syncthreads();          // warps of the block enter the unrolled network code together
    if (!active) continue;

    // ---------------- recurrent inference on the parent's hidden state (models.py:192-195)
    float value, reward, pri[A];
    {
      float st[ENC];
      load_floats<ENC>(hid + (size_t)node * ENC, st);
      float nx[ENC];
      SH::Dyn::run(pack + SH::OFF_DYN, st, action, nx);
      {
        float rl[FULL];
        SH::Rew::run(pack + SH::OFF_REW, nx, -1, rl);
        reward = s2s_regs<SH::SUP>(rl);
      }
      minmax_regs(nx);
      store_floats<ENC>(hid + (size_t)fresh * ENC, nx);
      {
        float pl[A];
        SH::Pol::run(pack + SH::OFF_POL, nx, -1, pl);
        priors_regs<A>(pl, nullptr, pri);
      }
      {
        float vl[FULL];
        SH::Val::run(pack + SH::OFF_VAL, nx, -1, vl);
        value = s2s_regs<SH::SUP>(vl);
      }
    }

    // ---------------- expand (self_play.py:346-352)
    {
      Rec<A> rr;
#pragma unroll
      for (int a = 0; a < A; ++a) rr.set(a, 0.0, pri[a], 0, 0.0f, MZB_CHILD_NONE);
      rr.store(t.rec(g, fresh));
      uint8_t* pr_ = t.rec(g, node);
      t.reward(pr_)[action] = reward;
      t.child(pr_)[action] = fresh;
    }

    // ---------------- backup (self_play.py:407-431), leaf first
    double val = (double)value;
    for (int k = L - 1; k >= 0; --k) {
      const uint32_t pe = PB_LUT ? lp_edge[k] : path[(size_t)k * G];
      uint8_t* er = t.rec(g, (int)(pe >> 16));
      const int pa = (int)(pe & 0xFFFFu);
      double e_vs = PB_LUT ? lp_vs[k] : t.value_sum(er)[pa];
      int e_vi = PB_LUT ? lp_vi[k] : t.visit(er)[pa];
      const double e_rw = (k == L - 1) ? (double)reward : (PB_LUT ? (double)lp_rw[k] : (double)t.reward(er)[pa]);
      backup_step(e_vs, e_vi, e_rw, val, t.discount, two, ((L - (k + 1)) & 1) == 0, vmin, vmax);
      t.value_sum(er)[pa] = e_vs;
      t.visit(er)[pa] = e_vi;
    }
    backup_step(root_vs, root_visit, (double)root_reward, val, t.discount, two, (L & 1) == 0, vmin, vmax);
    max_depth = L > max_depth ? L : max_depth;
    depth_sum += (unsigned int)L;
  }
  atomicAdd(t.counters, (unsigned long long)depth_sum);
  atomicAdd(t.counters + 1, (unsigned long long)io.num_sims);

  if (!active) return;
  // ---------------- publish the per-game scalars (same fields the modular kernels keep)
  t.root_value_sum[g] = root_vs;
  t.vmin[g] = vmin;
  t.vmax[g] = vmax;
  t.root_reward[g] = root_reward;
  t.root_visit[g] = root_visit;
  t.path_len[g] = 0;
  t.max_depth[g] = max_depth;
  t.sims_done[g] = io.num_sims;
  t.slot[g] = my_slot;
  t.step[g] = my_step;
  t.to_play[g] = io.to_play ? io.to_play[g] : (int8_t)0;
  if (io.visits) {
    uint8_t* r = t.rec(g, 0);
#pragma unroll
    for (int a = 0; a < A; ++a) io.visits[(size_t)g * A + a] = t.child(r)[a] == MZB_CHILD_ILLEGAL ? 0 : t.visit(r)[a];
  }
  if (io.root_value) io.root_value[g] = root_visit > 0 ? __ddiv_rn(root_vs, (double)root_visit) : 0.0;
  if (io.max_depth) io.max_depth[g] = max_depth;
}

using CartpoleShape = Shape<4, 8, 2, 10, 0, 16, 16, 16, 16>;       // games/cartpole.py:21-71
using TicTacToeFcShape = Shape<27, 32, 9, 10, 0, 16, 16, 0, 0>;    // games/tictactoe.py:20-70, network="fullyconnected"

template <class SH, int THREADS, bool PHASE_SYNC, bool PB_LUT>
int launch_variant(mzb_tree* t, mzb_fc_model* m, const SearchIO& io, cudaStream_t s) {
  const int S1 = io.num_sims + 1;
  const size_t smem = sizeof(float) * SH::PACK + sizeof(double) * (size_t)(((S1 + 1) & ~1) + (PB_LUT ? S1 * S1 : 0));
  static bool configured = false;
  if (!configured) {
    MZB_CUDA(cudaFuncSetAttribute(k_search_fc<SH, THREADS, PHASE_SYNC, PB_LUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  MZB_CHECK_ARG(smem <= 96 * 1024, "fused search: shared memory %zu > 96 KiB", smem);
  const int grid = (t->v.G + THREADS - 1) / THREADS;
  k_search_fc<SH, THREADS, PHASE_SYNC, PB_LUT><<<grid, THREADS, smem, s>>>(t->v, m->d_pack, io);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

// Kernel variant: MZB_FUSED_VARIANT = threads(128|256) + 1000*phase_sync + 10000*pb_lut (tuning knob; default below)
int fused_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MZB_FUSED_VARIANT");
    v = e ? atoi(e) : 11256;
  }
  return v;
}

template <class SH>
int launch_fused(mzb_tree* t, mzb_fc_model* m, const SearchIO& io, cudaStream_t s) {
  int v = fused_variant();
  const bool lut_ok = io.num_sims <= 63;                 // (S+1)^2 doubles must fit next to the weights
  const int threads = v % 1000;
  const bool sync = (v / 1000) % 10 != 0;
  const bool lut = (v / 10000) % 10 != 0 && lut_ok;
  if (threads == 256) {
    if (sync) return lut ? launch_variant<SH, 256, true, true>(t, m, io, s) : launch_variant<SH, 256, true, false>(t, m, io, s);
    return lut ? launch_variant<SH, 256, false, true>(t, m, io, s) : launch_variant<SH, 256, false, false>(t, m, io, s);
  }
  if (sync) return lut ? launch_variant<SH, 128, true, true>(t, m, io, s) : launch_variant<SH, 128, true, false>(t, m, io, s);
  return lut ? launch_variant<SH, 128, false, true>(t, m, io, s) : launch_variant<SH, 128, false, false>(t, m, io, s);
}

}  // namespace

extern "C" {

int mzb_search_fc_is_fused(const mzb_fc_model* m) {
  if (!m) return 0;
  return (CartpoleShape::matches(m->d) || TicTacToeFcShape::matches(m->d)) ? 1 : 0;
}

int mzb_search_fc(mzb_tree* t, mzb_fc_model* m, const float* d_obs, const uint8_t* d_legal, const int8_t* d_to_play,
                  const double* d_noise, double alpha, double frac, const uint32_t* d_slot, const uint32_t* d_step,
                  int32_t num_simulations, int allow_fused, int32_t* d_visits, double* d_root_value,
                  float* d_root_predicted_value, int32_t* d_max_depth, void* stream) {
  MZB_CHECK_ARG(t && m && d_obs, "NULL argument");
  MZB_CHECK_ARG(t->v.A == m->d.A, "tree has %d actions, network %d", t->v.A, m->d.A);
  MZB_CHECK_ARG(t->v.H == m->d.enc, "tree hidden slots hold %d floats, network encoding_size is %d", t->v.H, m->d.enc);
  MZB_CHECK_ARG(num_simulations > 0 && num_simulations <= t->v.S, "num_simulations %d outside 1..%d", num_simulations,
                t->v.S);
  MZB_CHECK_ARG(frac >= 0.0 && frac <= 1.0, "exploration fraction out of [0,1]: %f", frac);
  cudaStream_t s = (cudaStream_t)stream;
  const int G = t->v.G;
  if (allow_fused) {
    SearchIO io{d_obs, d_legal, d_to_play, d_noise, alpha, frac, d_slot, d_step, num_simulations,
                d_visits, d_root_value, d_root_predicted_value, d_max_depth};
    if (CartpoleShape::matches(m->d)) return launch_fused<CartpoleShape>(t, m, io, s);
    if (TicTacToeFcShape::matches(m->d)) return launch_fused<TicTacToeFcShape>(t, m, io, s);
  }
  // modular path: K5, K0, then (K1, K4, K3) per simulation - any FC shape
  const int64_t hs = (int64_t)(t->v.S + 1) * t->v.H;
  int rc = mzb_fc_initial(m, G, d_obs, d_legal, t->v.hidden, hs, 0, nullptr, nullptr, nullptr, d_root_predicted_value,
                          t->tmp_reward, t->tmp_priors, s);
  if (rc) return rc;
  rc = mzb_tree_root_init(t, t->tmp_reward, t->tmp_priors, 0, d_legal, d_to_play, d_noise, alpha, frac, d_slot, d_step, s);
  if (rc) return rc;
  for (int sim = 0; sim < num_simulations; ++sim) {
    rc = mzb_tree_select(t, t->tmp_parent, t->tmp_action, nullptr, s);
    if (rc) return rc;
    rc = mzb_fc_recurrent(m, G, t->v.hidden, hs, t->tmp_parent, t->v.H, t->tmp_action, t->v.hidden, hs,
                          (int64_t)(sim + 1) * t->v.H, nullptr, nullptr, nullptr, t->tmp_value, t->tmp_reward,
                          t->tmp_priors, s);
    if (rc) return rc;
    rc = mzb_tree_expand_backup(t, t->tmp_value, t->tmp_reward, t->tmp_priors, 0, s);
    if (rc) return rc;
  }
  if (d_visits || d_root_value || d_max_depth)
    return mzb_tree_root_stats(t, d_visits, d_root_value, d_max_depth, nullptr, nullptr, nullptr, nullptr, s);
  return MZB_OK;
}

}  // extern "C"
