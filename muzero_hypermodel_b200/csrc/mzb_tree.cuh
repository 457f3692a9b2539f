// Tree store layout + device routines shared by the modular tree kernels (mzb_tree.cu).
//
// HBM layout (one workspace, caller-owned), for G games, A actions, S simulations:
//   nodes      [G/32][S+1][32] records of 24*A bytes:  value_sum f64[A] | prior f32[A] | visit i32[A] |
//                                                      reward f32[A]   | child i32[A]
//              NODE-MAJOR inside a block of 32 consecutive games (one warp of the thread-per-game kernels): the 32
//              records a warp touches at the SAME node index - the root every simulation, the fresh node of
//              simulation k, the depth-1 nodes most of the time - are one contiguous 32*24*A-byte run instead of
//              32 lines scattered at a game-sized stride (cartpole whole-search kernel: 5.06 -> 4.62 ms).
//              record (g, n) describes the A edges leaving node n; edge (n, a) carries what the
//              reference keeps on the child Node (self_play.py:434-442).  child = slot of the
//              expanded child, -1 = not expanded, -2 = illegal at the root.
//   root_prior [G][A] f64   root priors after the Dirichlet mix are true float64 (:474-477)
//   path       [G][S+1] u32 (node << 16 | action) edges of the current search path
//   scalars    SoA per game: root value_sum/min/max f64, root reward f32, counters i32, rng u32
//   hidden     [G/32][S+1][32][H] f32 hidden-state slots (slot = node index), blocked the same way
// One node's edges are contiguous, so a group of lanes reads them with coalesced loads.
#pragma once
#include "mzb_common.cuh"

#define MZB_CHILD_NONE (-1)
#define MZB_CHILD_ILLEGAL (-2)

struct TreeView {
  int G, A, S, P, H;
  double discount;
  uint8_t* nodes;
  size_t rec_bytes, blk_stride;      // blk_stride = (S+1) * 32 * rec_bytes: one block of 32 games
  double* root_prior;
  uint32_t* path;
  double* root_value_sum;
  double* vmin;
  double* vmax;
  float* root_reward;
  int* root_visit;
  int* path_len;
  int* max_depth;
  int* sims_done;
  uint32_t* slot;
  uint32_t* step;
  int8_t* to_play;
  float* hidden;
  const double* log_lut;
  unsigned long long* counters;   // [0] sum of search-path lengths, [1] simulations (bench: mean path length)
  RngKey key;

  __host__ __device__ __forceinline__ size_t rec_index(int g, int n) const {     // in records (or hidden slots)
    return ((size_t)(g >> 5) * (size_t)(S + 1) + (size_t)n) * 32 + (size_t)(g & 31);
  }
  __device__ __forceinline__ uint8_t* rec(int g, int n) const { return nodes + rec_index(g, n) * rec_bytes; }
  __device__ __forceinline__ float* hidden_at(int g, int n) const { return hidden + rec_index(g, n) * (size_t)H; }
  __device__ __forceinline__ static double* value_sum(uint8_t* r) { return (double*)r; }
  __device__ __forceinline__ float* prior(uint8_t* r) const { return (float*)(r + 8 * (size_t)A); }
  __device__ __forceinline__ int* visit(uint8_t* r) const { return (int*)(r + 12 * (size_t)A); }
  __device__ __forceinline__ float* reward(uint8_t* r) const { return (float*)(r + 16 * (size_t)A); }
  __device__ __forceinline__ int* child(uint8_t* r) const { return (int*)(r + 20 * (size_t)A); }
};

struct mzb_tree {
  mzb_tree_config cfg;
  TreeView v;
  int lpg;
  double* d_log_lut;
  size_t bytes;
  // scratch for the modular search driven from C (mzb_search_fc): one entry per game
  int* tmp_parent;
  int* tmp_action;
  float* tmp_value;
  float* tmp_reward;
  float* tmp_priors;
};

// Gamma(alpha, 1) by Marsaglia & Tsang (2000) driven by Philox; alpha < 1 handled by the
// Gamma(alpha+1) * U^(1/alpha) boost.  Device-generated exploration noise only (parity mode injects).
__device__ inline double gamma_sample(RngKey key, uint32_t slot, uint32_t step, uint32_t action, double alpha) {
  const double a = alpha < 1.0 ? alpha + 1.0 : alpha;
  const double d = a - 1.0 / 3.0;
  const double c = 1.0 / sqrt(9.0 * d);
  double g = d;
  for (uint32_t attempt = 0; attempt < 64; ++attempt) {
    const Philox4 r = rng_draw(key, slot, step, MZB_STREAM_NOISE, attempt, action * 2u);
    const double u1 = u01_double(r.x, r.y), u2 = u01_double(r.z, r.w);
    // Box-Muller normal from (u1, u2); u1 == 0 is mapped to the smallest positive value
    const double rad = sqrt(-2.0 * log(u1 > 0.0 ? u1 : 1.1102230246251565e-16));
    const double x = rad * cospi(2.0 * u2);
    const double v0 = 1.0 + c * x;
    if (v0 <= 0.0) continue;
    const double v = v0 * v0 * v0;
    const Philox4 r2 = rng_draw(key, slot, step, MZB_STREAM_NOISE, attempt, action * 2u + 1u);
    const double u = u01_double(r2.x, r2.y);
    if (log(u > 0.0 ? u : 1.1102230246251565e-16) < 0.5 * x * x + d - d * v + d * log(v)) {
      g = d * v;
      if (alpha < 1.0) {
        const double ub = u01_double(r2.z, r2.w);
        g *= pow(ub > 0.0 ? ub : 1.1102230246251565e-16, 1.0 / alpha);
      }
      return g;
    }
  }
  return g;
}
