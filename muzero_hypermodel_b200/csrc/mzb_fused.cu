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

#define MZB_FUSED_DEFAULT_EXP 213639 // FFMA2 network, root record in registers, search path in shared memory, barrier per 4 warps, branch-free scores, reward/value heads share one copy of the code, L2 policies (records evict_last, hidden states evict_first)

namespace {

__host__ __device__ constexpr int pad4(int x) { return (x + 3) / 4 * 4; }

// ---- one Linear layer, compile-time shape.  w: Wt [*][OUTP] in shared memory, b: [OUTP].
// F2: the same fmaf chain per output, issued as packed FFMA2 (fma.rn.f32x2: two IEEE fp32 FMAs per instruction on
// sm_100) over adjacent outputs - bit-identical to the scalar form, half the FMA instructions.
// FE: ELU through ex2.approx (max(x, __expf(min(x, 0)) - 1): five instructions instead of eleven)
template <int NIN, int OUT, int OUTP, bool F2 = false, bool FE = false>
__device__ __forceinline__ void lin(const float* __restrict__ w, const float* __restrict__ b, const float (&x)[NIN],
                                    int hot, float (&y)[OUT], bool elu) {
  if constexpr (F2) {
    constexpr int NP = OUT / 2;
    float2 acc[NP > 0 ? NP : 1];
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[p] = *reinterpret_cast<const float2*>(b + 2 * p);
    float last = (OUT & 1) ? b[OUT - 1] : 0.0f;
#pragma unroll
    for (int i = 0; i < NIN; ++i) {
      const float2 xx = make_float2(x[i], x[i]);
#pragma unroll
      for (int p = 0; p < NP; ++p) acc[p] = __ffma2_rn(xx, *reinterpret_cast<const float2*>(w + i * OUTP + 2 * p), acc[p]);
      if (OUT & 1) last = fmaf(x[i], w[i * OUTP + OUT - 1], last);
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) { y[2 * p] = acc[p].x; y[2 * p + 1] = acc[p].y; }
    if (OUT & 1) y[OUT - 1] = last;
  } else {
#pragma unroll
    for (int o = 0; o < OUT; ++o) y[o] = b[o];
#pragma unroll
    for (int i = 0; i < NIN; ++i) {
#pragma unroll
      for (int o = 0; o < OUT; ++o) y[o] = fmaf(x[i], w[i * OUTP + o], y[o]);
    }
  }
  if (hot >= 0) {
    const float* wr = w + (NIN + hot) * OUTP;
#pragma unroll
    for (int o = 0; o < OUT; ++o) y[o] = __fadd_rn(y[o], wr[o]);
  }
  if (elu) {
#pragma unroll
    for (int o = 0; o < OUT; ++o) y[o] = FE ? fmaxf(y[o], __fsub_rn(__expf(fminf(y[o], 0.0f)), 1.0f)) : elu_f32(y[o]);
  }
}

// ---- mlp(): IN inputs (NDIRECT dense + optional one-hot tail), one optional hidden layer H, OUT outputs.
template <int NDIRECT, int IN, int H, int OUT>
struct Mlp {
  static constexpr int OUT0 = H > 0 ? H : OUT;
  static constexpr int SIZE0 = IN * pad4(OUT0) + pad4(OUT0);
  static constexpr int SIZE = SIZE0 + (H > 0 ? H * pad4(OUT) + pad4(OUT) : 0);
  template <bool F2 = false, bool FE = false>
  __device__ __forceinline__ static void run(const float* __restrict__ p, const float (&x)[NDIRECT], int hot,
                                             float (&y)[OUT]) {
    if constexpr (H > 0) {
      float h[H];
      lin<NDIRECT, H, pad4(H), F2, FE>(p, p + IN * pad4(H), x, hot, h, true);
      lin<H, OUT, pad4(OUT), F2, FE>(p + SIZE0, p + SIZE0 + H * pad4(OUT), h, -1, y, false);
    } else {
      lin<NDIRECT, OUT, pad4(OUT), F2, FE>(p, p + IN * pad4(OUT), x, hot, y, false);
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
  for (int i = 0; i < N; ++i) {
    // the minimum itself gives 0 / scale: a zero numerator fails the division's fast-path check (FCHK) and some lane of
    // a warp holds the minimum at almost every i, so the ~100-instruction slow path ran 2.5 times per simulation;
    // divide scale / scale there instead and select the exact +0
    const float d = __fsub_rn(s[i], lo);
    const float q = __fdiv_rn(d == 0.0f ? scale : d, scale);
    s[i] = d == 0.0f ? 0.0f : q;
  }
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
  // same, with an L2 cache policy (createpolicy): the records are what the walk re-reads, so they are kept (evict_last)
  // against the hidden states, which are read at most A times and are written / read with evict_first
  __device__ __forceinline__ void load_hint(const uint8_t* r, uint64_t pol) {
    static_assert(A % 2 == 0, "vector form");
#pragma unroll
    for (int i = 0; i < WORDS / 4; ++i)
      asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                   : "=r"(w[4 * i]), "=r"(w[4 * i + 1]), "=r"(w[4 * i + 2]), "=r"(w[4 * i + 3]) : "l"(r + 16 * i), "l"(pol));
  }
  __device__ __forceinline__ void store_hint(uint8_t* r, uint64_t pol) const {
#pragma unroll
    for (int i = 0; i < WORDS / 4; ++i)
      asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(r + 16 * i), "r"(w[4 * i]), "r"(w[4 * i + 1]),
                   "r"(w[4 * i + 2]), "r"(w[4 * i + 3]), "l"(pol) : "memory");
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
__device__ __forceinline__ void load_floats_hint(const float* __restrict__ p, float (&x)[N], uint64_t pol) {
  static_assert(N % 4 == 0, "vector form");
#pragma unroll
  for (int i = 0; i < N / 4; ++i)
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(x[4 * i]), "=f"(x[4 * i + 1]), "=f"(x[4 * i + 2]), "=f"(x[4 * i + 3]) : "l"(p + 4 * i), "l"(pol));
}
template <int N>
__device__ __forceinline__ void store_floats_hint(float* __restrict__ p, const float (&x)[N], uint64_t pol) {
  static_assert(N % 4 == 0, "vector form");
#pragma unroll
  for (int i = 0; i < N / 4; ++i)
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p + 4 * i), "f"(x[4 * i]), "f"(x[4 * i + 1]),
                 "f"(x[4 * i + 2]), "f"(x[4 * i + 3]), "l"(pol) : "memory");
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
// EXP: experiment / tuning flags (MZB_FUSED_EXP): 1 = packed FFMA2 network, 2 = root record in registers (A <= 4),
// 4/8/16/32 = timing-only diagnostics (blocked layout, aliased trees, no network, no walk) - results are NOT valid.
// 4 = the search path's edge statistics in shared memory for the first 12 levels (deeper levels: local memory).
enum { X_F2 = 1, X_ROOTREG = 2, X_SPATH = 4, X_ALIAS = 8, X_NONET = 16, X_NOWALK = 32, X_SYNC2 = 64, X_HALFBAR = 128, X_RCP = 256, X_BF = 512, X_FELU = 1024, X_SCHEDBAR = 2048, X_BAR2 = 4096, X_P1 = 8192, X_HEADLOOP = 16384, X_UNPEEL = 32768, X_L2HINT = 65536, X_L2HINT2 = 131072 };
__host__ __device__ constexpr int smem_path_depth(int blocks_per_sm) { return blocks_per_sm >= 3 ? 8 : 12; }
__host__ __device__ constexpr int smem_path_depth(int blocks_per_sm, int threads) {
  // 16 bytes per level and thread next to ~18 KB of weights + tables per CTA: 12 levels up to 640 threads per SM
  return threads * blocks_per_sm > 768 ? 8 : (threads * blocks_per_sm > 640 ? 10 : 12);
}

template <class SH, int THREADS, bool PHASE_SYNC, bool PB_LUT, int EXP, int MINB = 512 / THREADS>
__global__ void __launch_bounds__(THREADS, MINB) k_search_fc(TreeView t, const float* __restrict__ gpack, SearchIO io) {
  constexpr int A = SH::A, ENC = SH::ENC, FULL = SH::FULL;
  constexpr bool F2 = (EXP & X_F2) != 0;
  constexpr bool ROOTREG = (EXP & X_ROOTREG) != 0 && A <= 4;
  constexpr bool FE = (EXP & X_FELU) != 0;
  constexpr bool L2H = (EXP & X_L2HINT) != 0 && A % 2 == 0 && ENC % 4 == 0;
  uint64_t pol_keep = 0, pol_stream = 0;
  if (L2H) {
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  }
  extern __shared__ float4 smem4[];
  float* pack = reinterpret_cast<float*>(smem4);
  double* lut = reinterpret_cast<double*>(pack + SH::PACK);
  const int S1 = io.num_sims + 1;
  // pb(N, n) only for n <= N (a child is never visited more often than its parent): triangular table, row N at
  // N (N + 1) / 2 - half the shared memory of the square one
  double* pbt = lut + ((S1 + 1) & ~1);
  const int TRI = (S1 * (S1 + 1) / 2 + 1) & ~1;     // even: what follows stays 16-byte aligned
  // SPATH: [PD][THREADS] value sums f64 | rewards f32 | (node << 16 | action << 8... ) see path_put
  constexpr int PD = ((EXP & X_SPATH) != 0 && PB_LUT) ? smem_path_depth(MINB, THREADS) : 0;
  double* sp_vs = pbt + (PB_LUT ? TRI : 0);
  float* sp_rw = reinterpret_cast<float*>(sp_vs + PD * THREADS);
  uint32_t* sp_ev = reinterpret_cast<uint32_t*>(sp_rw + PD * THREADS);
  // RCP: refined reciprocals of the visit counts 1..S (the divisor of value_sum / visit_count)
  constexpr bool RCP = (EXP & X_RCP) != 0 && PB_LUT;
  // BF: branch-free scores - every division of a level goes through a hoisted reciprocal with the validity checks
  // OR-ed into one flag (one rare branch per level to the exact form), so the children's float64 chains interleave
  constexpr bool BF = (EXP & X_BF) != 0 && PB_LUT;
  double* rcpn = reinterpret_cast<double*>(sp_ev + PD * THREADS);
  double2* rcpn2 = reinterpret_cast<double2*>(sp_ev + PD * THREADS);       // {refined 1 / n, (double)n}
  if (RCP) for (int i = threadIdx.x; i < S1; i += THREADS) rcpn[i] = i > 0 ? rcp_refined((double)i) : 0.0;
  if (BF) for (int i = threadIdx.x; i <= S1; i += THREADS) rcpn2[i] = make_double2(i > 0 ? rcp_refined((double)i) : 0.0, (double)i);
  for (int i = threadIdx.x; i < SH::PACK / 4; i += THREADS) smem4[i] = reinterpret_cast<const float4*>(gpack)[i];
  for (int i = threadIdx.x; i < S1; i += THREADS) lut[i] = t.log_lut[i];
  if (PB_LUT) {
    for (int i = threadIdx.x; i < S1 * S1; i += THREADS) {
      const int N = i / S1, n = i % S1;
      if (n <= N) pbt[N * (N + 1) / 2 + n] = ucb_pb(t.log_lut[N], __dsqrt_rn((double)N), n);
    }
  }
  __syncthreads();
  const int g_raw = blockIdx.x * THREADS + threadIdx.x;
  const bool active = g_raw < t.G;
  if (!PHASE_SYNC && !active) return;
  const int g = active ? g_raw : 0;                        // idle threads of the last block only take part in barriers
  const bool two = (EXP & X_P1) ? false : t.P == 2;        // X_P1: single-player instantiation (the launcher checks t.P)
  const size_t G = (size_t)t.G;
  const uint32_t my_slot = io.slot ? io.slot[g] : (uint32_t)g;
  const uint32_t my_step = io.step ? io.step[g] : 0u;
  const uint8_t* legal = io.legal ? io.legal + (size_t)g * A : nullptr;
  const int gm = (EXP & X_ALIAS) ? (g & 8191) : g;
  // per-thread bases of the blocked store (mzb_tree.cuh): node n of this game lives n * 32 records further on, so
  // an address is one IMAD with a compile-time stride instead of a 64-bit multiply by runtime sizes
  constexpr size_t RB = 24 * (size_t)A;
  uint8_t* const rec_base = t.nodes + t.rec_index(gm, 0) * RB;
  float* const hid_base = t.hidden + t.rec_index(gm, 0) * ENC;
  auto rec_of = [&](int n) -> uint8_t* { return rec_base + (size_t)n * (32 * RB); };
  auto hid_of = [&](int n) -> float* { return hid_base + (size_t)n * (32 * ENC); };

  // ---------------- initial inference (models.py:172-190) + root expansion (self_play.py:292-314)
  double rp[A];                       // root priors, float64 after the noise mix
  float root_reward = 0.0f;
  Rec<ROOTREG ? A : 1> root;          // ROOTREG: the root's record lives in registers for the whole search
  if (active) {
    float ob[SH::OBS];
    load_floats<SH::OBS>(io.obs + (size_t)g * SH::OBS, ob);
    float st[ENC];
    SH::Rep::template run<F2, FE>(pack + SH::OFF_REP, ob, -1, st);
    minmax_regs(st);
    store_floats<ENC>(hid_of(0), st);
    float pl[A], pri[A];
    SH::Pol::template run<F2, FE>(pack + SH::OFF_POL, st, -1, pl);
    priors_regs<A>(pl, legal, pri);
    if (io.root_pred_value) {
      float vl[FULL];
      SH::Val::template run<F2, FE>(pack + SH::OFF_VAL, st, -1, vl);
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
      if constexpr (ROOTREG) {
#pragma unroll
        for (int i = 0; i < Rec<A>::WORDS; ++i) root.w[i] = rr.w[i];
      } else {
        rr.store(rec_of(0));
      }
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

  // One level of the select walk (self_play.py:364-405) on the record `rr` of `node`: scores, arg-max with the
  // reference's tie rule, path entry; returns the chosen child (< 0: leaf reached) and leaves the action in `action`.
  int action = 0, N = 0;
  double mm_den = 0.0, mm_y = 0.0;          // RCP: vmax - vmin of this simulation and its refined reciprocal
  bool mm_ok = false;
  auto walk_level = [&](const Rec<A>& rr, const bool root_level, const int node, const int depth, const int sim) -> int {
    double vs[A]; float pr[A], rw[A]; int vi[A], ch[A];
#pragma unroll
    for (int a = 0; a < A; ++a) { vs[a] = rr.vs(a); pr[a] = rr.pr(a); vi[a] = rr.vi(a); rw[a] = rr.rw(a); ch[a] = rr.ch(a); }
    const double pbc0 = PB_LUT ? 0.0 : lut[N];
    const double sqrtN = PB_LUT ? 0.0 : __dsqrt_rn((double)N);
    const double* pbrow = pbt + ((N * (N + 1)) >> 1);
    double sc[A];
    double best = -CUDART_INF;
    int n_best = 0;
    action = -1;
    if constexpr (BF) {
      const bool norm = vmax > vmin;
      bool bad = norm && !mm_ok;
#pragma unroll
      for (int a = 0; a < A; ++a) {
        const int n = vi[a];
        const double2 yn = rcpn2[n];
        const double p = root_level ? rp[a] : (double)pr[a];
        const double s = __dmul_rn(pbrow[n], p);
        bool okv, okq;
        double v = ddiv_rcp_nb(vs[a], yn.y, yn.x, okv);
        if (two) v = -v;
        const double q = __dadd_rn((double)rw[a], __dmul_rn(t.discount, v));
        const double qn = __dsub_rn(q, vmin);
        const double u = ddiv_rcp_nb(qn, mm_den, mm_y, okq);
        okq = okq || qn == 0.0;
        const bool has = n > 0;
        sc[a] = has ? __dadd_rn(s, norm ? u : q) : s;
        bad = bad || (has && !(okv && (!norm || okq)));
      }
      if (bad) {
#pragma unroll
        for (int a = 0; a < A; ++a)
          sc[a] = ucb_score_pb(pbrow[vi[a]], vi[a], root_level ? rp[a] : (double)pr[a], vs[a], (double)rw[a], t.discount, two, vmin, vmax);
      }
    }
#pragma unroll
    for (int a = 0; a < A; ++a) {
      if ((root_level || !BF) && ch[a] == MZB_CHILD_ILLEGAL) { sc[a] = -CUDART_INF; continue; }   // illegal edges exist at the root only
      if constexpr (BF) {
        if (sc[a] > best || action < 0) { best = sc[a]; n_best = 1; action = a; }
        else if (sc[a] == best) ++n_best;
        continue;
      }
      const double p = root_level ? rp[a] : (double)pr[a];
      const double pb = PB_LUT ? pbrow[vi[a]] : ucb_pb(pbc0, sqrtN, vi[a]);
      if constexpr (RCP)
        sc[a] = ucb_score_pb_rcp(pb, vi[a], p, vs[a], (double)rw[a], t.discount, two, vmin, vmax, rcpn[vi[a]], mm_den, mm_y, mm_ok);
      else
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
      // node <= 63 and action < 2^8 here (PB_LUT: num_simulations <= 63): edge and visit count share one word
      if (PD > 0 && depth < PD) {
        sp_ev[depth * THREADS + threadIdx.x] = ((uint32_t)node << 24) | ((uint32_t)action << 16) | (uint32_t)nv;
        sp_vs[depth * THREADS + threadIdx.x] = nvs; sp_rw[depth * THREADS + threadIdx.x] = nrw;
      } else {
        lp_edge[depth] = ((uint32_t)node << 16) | (uint32_t)action; lp_vs[depth] = nvs; lp_vi[depth] = nv; lp_rw[depth] = nrw;
      }
    } else {
      path[(size_t)depth * G] = ((uint32_t)node << 16) | (uint32_t)action;
    }
    N = nv;
    return next;
  };

  for (int sim = 0; sim < io.num_sims; ++sim) {
    // ---------------- select walk (self_play.py:326-335): the root level straight from registers (ROOTREG), then
    // one dependent record load per level
    int node = 0, depth = 0;
    N = root_visit;
    if ((RCP || BF) && vmax > vmin) {
      mm_den = __dsub_rn(vmax, vmin);
      mm_ok = rcp_divisor_ok(mm_den);
      mm_y = mm_ok ? rcp_refined(mm_den) : 0.0;
    }
    if (active) {
      int next;
      if constexpr (ROOTREG && (EXP & X_UNPEEL) != 0) {
        // one copy of the level code: the root's record comes from registers, the others from the store
        Rec<A> rr;
#pragma unroll
        for (int i = 0; i < Rec<A>::WORDS; ++i) rr.w[i] = root.w[i];
        bool at_root = true;
        for (;;) {
          next = walk_level(rr, at_root, node, depth, sim);
          ++depth;
          if (next < 0) break;
          node = next;
          rr.load(rec_of(node));
          at_root = false;
        }
      } else if constexpr (ROOTREG) {
        Rec<A> rr;
#pragma unroll
        for (int i = 0; i < Rec<A>::WORDS; ++i) rr.w[i] = root.w[i];
        next = walk_level(rr, true, 0, 0, sim);
      } else {
        Rec<A> rr;
        rr.load(rec_of(0));
        next = walk_level(rr, true, 0, 0, sim);
      }
      if constexpr (!(ROOTREG && (EXP & X_UNPEEL) != 0)) depth = 1;
      if (EXP & X_NOWALK) next = -1;
      while (next >= 0) {
        node = next;
        Rec<A> rr;
        if constexpr (L2H) rr.load_hint(rec_of(node), pol_keep); else rr.load(rec_of(node));
        next = walk_level(rr, false, node, depth, sim);
        ++depth;
      }
    }
    const int L = depth, fresh = sim + 1;
    // the parent's hidden state is requested before the phase barrier, so its latency overlaps the wait
    float st[ENC];
    if (active) { if constexpr (L2H) load_floats_hint<ENC>(hid_of(node), st, pol_stream); else load_floats<ENC>(hid_of(node), st); }
    if (PHASE_SYNC) {                         // warps of the block enter the unrolled network code together
      if constexpr ((EXP & X_BAR2) != 0) {
        asm volatile("bar.sync %0, 64;" ::"r"(1 + (int)(threadIdx.x >> 6)) : "memory");       // per pair of consecutive warps
      } else if constexpr ((EXP & X_SCHEDBAR) != 0) {
        // the warps that share a scheduler (warp id mod 4) and hence an L0 instruction cache enter the network together
        asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)((threadIdx.x >> 5) & 3)), "n"(THREADS / 4) : "memory");
      } else if constexpr ((EXP & X_HALFBAR) != 0) {
        asm volatile("bar.sync %0, 128;" ::"r"(1 + (int)(threadIdx.x >> 7)) : "memory");      // per group of four warps
      } else if constexpr ((EXP & X_SYNC2) != 0) {
        if (sim & 1) __syncthreads();
      } else {
        __syncthreads();
      }
    }
    if (!active) continue;

    // ---------------- recurrent inference on the parent's hidden state (models.py:192-195)
    float value, reward, pri[A];
    if constexpr ((EXP & X_NONET) != 0) {
      value = st[0] + 0.5f; reward = 1.0f;
#pragma unroll
      for (int a = 0; a < A; ++a) pri[a] = 1.0f / A;
      store_floats<ENC>(hid_of(fresh), st);
    } else {
      float nx[ENC];
      SH::Dyn::template run<F2, FE>(pack + SH::OFF_DYN, st, action, nx);
      if constexpr ((EXP & X_HEADLOOP) != 0 && SH::Rew::SIZE == SH::Val::SIZE) {
        // reward head (on the raw next state) and value head (on the normalised one) share one copy of the code
        reward = 0.0f; value = 0.0f;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          float hl[FULL];
          SH::Rew::template run<F2, FE>(pack + (h ? SH::OFF_VAL : SH::OFF_REW), nx, -1, hl);
          const float o = s2s_regs<SH::SUP>(hl);
          if (h == 0) {
            reward = o;
            minmax_regs(nx);
            if constexpr (L2H) store_floats_hint<ENC>(hid_of(fresh), nx, pol_stream); else store_floats<ENC>(hid_of(fresh), nx);
            float pl[A];
            SH::Pol::template run<F2, FE>(pack + SH::OFF_POL, nx, -1, pl);
            priors_regs<A>(pl, nullptr, pri);
          } else {
            value = o;
          }
        }
      } else {
      {
        float rl[FULL];
        SH::Rew::template run<F2, FE>(pack + SH::OFF_REW, nx, -1, rl);
        reward = s2s_regs<SH::SUP>(rl);
      }
      minmax_regs(nx);
      store_floats<ENC>(hid_of(fresh), nx);
      {
        float pl[A];
        SH::Pol::template run<F2, FE>(pack + SH::OFF_POL, nx, -1, pl);
        priors_regs<A>(pl, nullptr, pri);
      }
      {
        float vl[FULL];
        SH::Val::template run<F2, FE>(pack + SH::OFF_VAL, nx, -1, vl);
        value = s2s_regs<SH::SUP>(vl);
      }
      }
    }

    // ---------------- expand (self_play.py:346-352)
    {
      Rec<A> rr;
#pragma unroll
      for (int a = 0; a < A; ++a) rr.set(a, 0.0, pri[a], 0, 0.0f, MZB_CHILD_NONE);
      if constexpr (L2H) rr.store_hint(rec_of(fresh), pol_keep); else rr.store(rec_of(fresh));
      bool in_regs = false;
      if constexpr (ROOTREG) {
        if (node == 0) {
          in_regs = true;
#pragma unroll
          for (int a = 0; a < A; ++a)
            if (a == action) { root.w[4 * A + a] = __float_as_uint(reward); root.w[5 * A + a] = (uint32_t)fresh; }
        }
      }
      if (!in_regs) {
        uint8_t* pr_ = rec_of(node);
        if constexpr (L2H && (EXP & X_L2HINT2) != 0) {
          asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(pr_ + 16 * A + 4 * action), "f"(reward), "l"(pol_keep) : "memory");
          asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(pr_ + 20 * A + 4 * action), "r"(fresh), "l"(pol_keep) : "memory");
        } else {
          t.reward(pr_)[action] = reward;
          t.child(pr_)[action] = fresh;
        }
      }
    }

    // ---------------- backup (self_play.py:407-431), leaf first
    double val = (double)value;
    if constexpr (BF && ROOTREG && PD > 0) {
      // Levels L-1 .. 1 are edges of non-root nodes (always in the HBM store), level 0 is the root's edge (always in
      // registers): peeled, so neither carries the other's branch; field offsets are compile-time.
      for (int k = L - 1; k >= 1; --k) {
        int pn, pa, e_vi; double e_vs; float l_rw;
        if (k < PD) {
          const uint32_t ev = sp_ev[k * THREADS + threadIdx.x];
          pn = (int)(ev >> 24); pa = (int)((ev >> 16) & 0xFFu); e_vi = (int)(ev & 0xFFFFu);
          e_vs = sp_vs[k * THREADS + threadIdx.x]; l_rw = sp_rw[k * THREADS + threadIdx.x];
        } else {
          const uint32_t pe = lp_edge[k];
          pn = (int)(pe >> 16); pa = (int)(pe & 0xFFFFu);
          e_vs = lp_vs[k]; e_vi = lp_vi[k]; l_rw = lp_rw[k];
        }
        const double e_rw = (k == L - 1) ? (double)reward : (double)l_rw;
        backup_step_rcp(e_vs, e_vi, e_rw, val, t.discount, two, ((L - (k + 1)) & 1) == 0, vmin, vmax, rcpn2);
        uint8_t* er = rec_of(pn);
        if constexpr (L2H && (EXP & X_L2HINT2) != 0) {
          asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(er + 8 * pa), "d"(e_vs), "l"(pol_keep) : "memory");
          asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(er + 12 * A + 4 * pa), "r"(e_vi), "l"(pol_keep) : "memory");
        } else {
          reinterpret_cast<double*>(er)[pa] = e_vs;
          reinterpret_cast<int*>(er + 12 * A)[pa] = e_vi;
        }
      }
      {
        const uint32_t ev = sp_ev[threadIdx.x];
        const int pa = (int)((ev >> 16) & 0xFFu);
        int e_vi = (int)(ev & 0xFFFFu);
        double e_vs = sp_vs[threadIdx.x];
        const double e_rw = (L == 1) ? (double)reward : (double)sp_rw[threadIdx.x];
        backup_step_rcp(e_vs, e_vi, e_rw, val, t.discount, two, ((L - 1) & 1) == 0, vmin, vmax, rcpn2);
#pragma unroll
        for (int a = 0; a < A; ++a)
          if (a == pa) {
            root.w[2 * a] = (uint32_t)__double2loint(e_vs); root.w[2 * a + 1] = (uint32_t)__double2hiint(e_vs);
            root.w[3 * A + a] = (uint32_t)e_vi;
          }
      }
    } else
    for (int k = L - 1; k >= 0; --k) {
      int pn, pa, e_vi; double e_vs; float l_rw = 0.0f;
      if (PD > 0 && k < PD) {
        const uint32_t ev = sp_ev[k * THREADS + threadIdx.x];
        pn = (int)(ev >> 24); pa = (int)((ev >> 16) & 0xFFu); e_vi = (int)(ev & 0xFFFFu);
        e_vs = sp_vs[k * THREADS + threadIdx.x]; l_rw = sp_rw[k * THREADS + threadIdx.x];
      } else {
        const uint32_t pe = PB_LUT ? lp_edge[k] : path[(size_t)k * G];
        pn = (int)(pe >> 16); pa = (int)(pe & 0xFFFFu);
        if (PB_LUT) { e_vs = lp_vs[k]; e_vi = lp_vi[k]; l_rw = lp_rw[k]; }
      }
      uint8_t* er = rec_of(pn);
      if (!PB_LUT) { e_vs = t.value_sum(er)[pa]; e_vi = t.visit(er)[pa]; l_rw = t.reward(er)[pa]; }
      const double e_rw = (k == L - 1) ? (double)reward : (double)l_rw;
      if constexpr (BF) backup_step_rcp(e_vs, e_vi, e_rw, val, t.discount, two, ((L - (k + 1)) & 1) == 0, vmin, vmax, rcpn2);
      else backup_step(e_vs, e_vi, e_rw, val, t.discount, two, ((L - (k + 1)) & 1) == 0, vmin, vmax);
      bool in_regs = false;
      if constexpr (ROOTREG) {
        if (pn == 0) {
          in_regs = true;
#pragma unroll
          for (int a = 0; a < A; ++a)
            if (a == pa) {
              root.w[2 * a] = (uint32_t)__double2loint(e_vs); root.w[2 * a + 1] = (uint32_t)__double2hiint(e_vs);
              root.w[3 * A + a] = (uint32_t)e_vi;
            }
        }
      }
      if (!in_regs) {
        t.value_sum(er)[pa] = e_vs;
        t.visit(er)[pa] = e_vi;
      }
    }
    if constexpr (BF) backup_step_rcp(root_vs, root_visit, (double)root_reward, val, t.discount, two, (L & 1) == 0, vmin, vmax, rcpn2);
    else backup_step(root_vs, root_visit, (double)root_reward, val, t.discount, two, (L & 1) == 0, vmin, vmax);
    max_depth = L > max_depth ? L : max_depth;
    depth_sum += (unsigned int)L;
  }
  atomicAdd(t.counters, (unsigned long long)depth_sum);
  atomicAdd(t.counters + 1, (unsigned long long)io.num_sims);

  if (!active) return;
  if constexpr (ROOTREG) {
    Rec<A> rr;
#pragma unroll
    for (int i = 0; i < Rec<A>::WORDS; ++i) rr.w[i] = root.w[i];
    rr.store(rec_of(0));
  }
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
#pragma unroll
    for (int a = 0; a < A; ++a) {
      int c, v;
      if constexpr (ROOTREG) { c = (int)root.w[5 * A + a]; v = (int)root.w[3 * A + a]; }
      else { uint8_t* r = rec_of(0); c = t.child(r)[a]; v = t.visit(r)[a]; }
      io.visits[(size_t)g * A + a] = c == MZB_CHILD_ILLEGAL ? 0 : v;
    }
  }
  if (io.root_value) io.root_value[g] = root_visit > 0 ? __ddiv_rn(root_vs, (double)root_visit) : 0.0;
  if (io.max_depth) io.max_depth[g] = max_depth;
}

using CartpoleShape = Shape<4, 8, 2, 10, 0, 16, 16, 16, 16>;       // games/cartpole.py:21-71
using TicTacToeFcShape = Shape<27, 32, 9, 10, 0, 16, 16, 0, 0>;    // games/tictactoe.py:20-70, network="fullyconnected"

template <class SH, bool PB_LUT, int EXP, int THREADS = 256, bool PHASE_SYNC = true, int MINB = 512 / THREADS>
int launch_variant(mzb_tree* t, mzb_fc_model* m, const SearchIO& io, cudaStream_t s) {
  const int S1 = io.num_sims + 1;
  const size_t smem = sizeof(float) * SH::PACK + sizeof(double) * (size_t)(((S1 + 1) & ~1) + (PB_LUT ? ((S1 * (S1 + 1) / 2 + 1) & ~1) : 0)) +
                      (((EXP & X_SPATH) != 0 && PB_LUT) ? (size_t)smem_path_depth(MINB, THREADS) * THREADS * 16 : 0) +
                      (((EXP & X_RCP) != 0 && PB_LUT) ? sizeof(double) * (size_t)S1 : 0) +
                      (((EXP & X_BF) != 0 && PB_LUT) ? sizeof(double2) * (size_t)(S1 + 1) : 0);
  static bool configured = false;
  if (!configured) {
    MZB_CUDA(cudaFuncSetAttribute(k_search_fc<SH, THREADS, PHASE_SYNC, PB_LUT, EXP, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  MZB_CHECK_ARG(smem <= 160 * 1024, "fused search: shared memory %zu > 160 KiB", smem);
  const int grid = (t->v.G + THREADS - 1) / THREADS;
  k_search_fc<SH, THREADS, PHASE_SYNC, PB_LUT, EXP, MINB><<<grid, THREADS, smem, s>>>(t->v, m->d_pack, io);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

// MZB_FUSED_EXP: experiment flags of k_search_fc (see the enum); default = the shipped configuration
int fused_exp() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MZB_FUSED_EXP");
    v = e ? atoi(e) : MZB_FUSED_DEFAULT_EXP;
  }
  return v;
}

template <class SH, bool TUNE>
int launch_fused(mzb_tree* t, mzb_fc_model* m, const SearchIO& io, cudaStream_t s) {
  const bool lut = io.num_sims <= 63;                 // (S+1)^2 doubles must fit next to the weights
  if (!lut) return launch_variant<SH, false, MZB_FUSED_DEFAULT_EXP>(t, m, io, s);
  // small batches (tictactoe at 4,096 games = 16 CTAs of 256): one warp per CTA spreads the games over all SMs
  // (fewer games per warp - 16 / 8 / 4 active lanes, i.e. 256 / 512 / 1,024 one-warp CTAs for 4,096 games - is slower:
  // 0.45 / 0.59 / 1.10 ms against 0.47 ms per tictactoe search; several unsynchronised warps per SM thrash the
  // instruction caches on the unrolled network code, like the kernel without its phase barrier)
  if (t->v.G < 148 * 128) return launch_variant<SH, true, MZB_FUSED_DEFAULT_EXP, 32, false, 2>(t, m, io, s);
  if constexpr (TUNE) {
    switch (fused_exp()) {
      case 0: return launch_variant<SH, true, 0>(t, m, io, s);
      case 7: return launch_variant<SH, true, 7>(t, m, io, s);
      case 23: return launch_variant<SH, true, 23>(t, m, io, s);
      case 263: return launch_variant<SH, true, 263>(t, m, io, s);
      case 135: return launch_variant<SH, true, 135>(t, m, io, s);
      case 647: return launch_variant<SH, true, 647>(t, m, io, s);                  // + branch-free scores: 4.20 -> 3.64 ms
      case 17031: return launch_variant<SH, true, 17031>(t, m, io, s);              // reward and value heads as one two-trip loop: 3.57 -> 3.46 ms
      case 82567: return launch_variant<SH, true, 82567>(t, m, io, s);              // + L2 policies: records evict_last, hidden states evict_first
      case 213639: return launch_variant<SH, true, 213639>(t, m, io, s);            // + the backup's / expand's partial record stores evict_last too: 3.46 -> 3.34 ms
      // (swept at 606,208 games, 6.58 ms: evict_last fraction 1.0 / 0.75 / 0.5 / 0.25 makes no difference; hidden states
      // evict_normal or evict_unchanged instead of evict_first: 6.67 - the streaming of the hidden states is what pays)
      // (the loop body is ~47 KB of SASS against a 32 KB L1.5 instruction cache: code size shows).  On top of it: root
      // level not peeled 3.48; ex2-based ELU 3.40 (not adopted: changes the network's rounding); barrier per pair of
      // consecutive warps 3.68 (on 647); parent's hidden state requested before the barrier: no change
      // measured on top of 647 and dropped (303,104 games, 3.64 ms): ex2-based ELU 3.61 (-380 instructions per simulation
      // buy 0.7 %: not issue-bound); barrier per scheduler (warps w, w+4) 3.72; one 512-thread CTA per SM 3.76;
      // prefetch.global.L1 of both children's records 3.69; compile-time single player 3.67
      case 2007: return launch_variant<SH, true, 7, 128, true, 5>(t, m, io, s);     // 5 x 128 threads per SM (96 registers, no spills): slower
      case 1135: return launch_variant<SH, true, 135, 256, true, 1>(t, m, io, s);   // 8 warps per SM: 6.22 ms
      case 3135: return launch_variant<SH, true, 135, 128, true, 3>(t, m, io, s);   // 12 warps per SM: 5.22 ms (16: 4.23)
      case 5135: return launch_variant<SH, true, 135, 128, true, 5>(t, m, io, s);   // 20 warps per SM (96 registers)
      case 6135: return launch_variant<SH, true, 135, 128, true, 6>(t, m, io, s);   // 24 warps per SM (80 registers)
      case 19: return launch_variant<SH, true, 19>(t, m, io, s);
      case 35: return launch_variant<SH, true, 35>(t, m, io, s);
      default: break;
    }
  }
  return launch_variant<SH, true, MZB_FUSED_DEFAULT_EXP>(t, m, io, s);
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
    if (CartpoleShape::matches(m->d)) return launch_fused<CartpoleShape, true>(t, m, io, s);
    if (TicTacToeFcShape::matches(m->d)) return launch_fused<TicTacToeFcShape, false>(t, m, io, s);
  }
  // modular path: K5, K0, then (K1, K4, K3) per simulation - any FC shape
  const int S1 = t->v.S + 1;
  int rc = mzb_fc_initial_tree(m, G, d_obs, d_legal, t->v.hidden, S1, 0, d_root_predicted_value, t->tmp_reward,
                               t->tmp_priors, s);
  if (rc) return rc;
  rc = mzb_tree_root_init(t, t->tmp_reward, t->tmp_priors, 0, d_legal, d_to_play, d_noise, alpha, frac, d_slot, d_step, s);
  if (rc) return rc;
  for (int sim = 0; sim < num_simulations; ++sim) {
    rc = mzb_tree_select(t, t->tmp_parent, t->tmp_action, nullptr, s);
    if (rc) return rc;
    rc = mzb_fc_recurrent_tree(m, G, t->v.hidden, S1, t->tmp_parent, t->tmp_action, sim + 1, t->tmp_value, t->tmp_reward,
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
