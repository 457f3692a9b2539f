// Vectorised environments (K9), action selection + recording (K10) and finished-game export.
// One thread owns one game.  Reference: games/{cartpole,tictactoe,connect4,gomoku}.py (Game.step /
// legal_actions / to_play / reset), SelfPlay.play_game + select_action (self_play.py:110-184, 223-246),
// GameHistory (self_play.py:480-512).
#include <math.h>

#include <vector>

#include "mzb_common.cuh"

struct EnvView {
  int kind, G, A, H, W, cells, obs_dim, max_moves;
  double reward_scale;
  // state
  double* cp;            // cartpole [G][4] f64
  int* elapsed;          // cartpole steps in the episode
  int8_t* board;         // board games [G][cells], values -1/0/1
  int8_t* player;        // +1 / -1
  uint32_t* steps;       // per-slot env steps so far, over all episodes (RNG counter c1)
  uint32_t slot0;        // global id of game 0 (rank offset)
  RngKey key;
  // history of the running episode (GameHistory fields, compact)
  int hist_cap;          // max_moves + 1 entries
  int rec_floats;        // observation record: cartpole 4 x f32, boards (cells+1) x i8 packed in floats
  float* h_obs;          // [G][hist_cap][rec_floats]
  int* h_action;         // [G][hist_cap]
  float* h_reward;       // [G][hist_cap]
  int8_t* h_to_play;     // [G][hist_cap]
  uint16_t* h_visits;    // [G][max_moves][A]   root child visit counts (child_visits = count / sum)
  double* h_root_value;  // [G][max_moves]
  int* h_len;            // moves played in the running episode
  uint8_t* finished;     // set by act_step, cleared by harvest
  // counters [8]: 0 games finished, 1 moves of finished games, 2 env steps, 3 export overflow (games dropped)
  unsigned long long* counters;
};

struct ExportView {
  int cap_entries, cap_games;
  float* obs; int* action; float* reward; int8_t* to_play; uint16_t* visits; double* root_value;
  int* game_start; int* game_len; uint32_t* game_slot;
  int* cursor;           // [2]: entries used, games used
  int* plan;             // [G][2]: export entry offset (-1 = not exported) and game index, written by k_harvest_plan
};

struct mzb_env {
  mzb_env_config cfg;
  EnvView v;
  ExportView x;
  size_t bytes;
};

namespace {

constexpr double kGravity = 9.8, kMassPole = 0.1, kTotalMass = 1.1, kLength = 0.5, kPoleMassLength = 0.05;
constexpr double kForceMag = 10.0, kTau = 0.02, kXThreshold = 2.4;
constexpr double kThetaThreshold = 12 * 2 * 3.141592653589793 / 360;
constexpr int kCartpoleLimit = 500;       // TimeLimit of CartPole-v1

__device__ __forceinline__ int8_t* board_of(const EnvView& e, int g) { return e.board + (size_t)g * e.cells; }

__device__ inline bool run_of(const int8_t* b, int H, int W, int player, int len, int dx, int dy) {
  for (int x = 0; x < H; ++x)
    for (int y = 0; y < W; ++y) {
      const int xe = x + dx * (len - 1), ye = y + dy * (len - 1);
      if (xe < 0 || xe >= H || ye < 0 || ye >= W) continue;
      bool all = true;
      for (int k = 0; k < len && all; ++k) all = b[(x + dx * k) * W + (y + dy * k)] == player;
      if (all) return true;
    }
  return false;
}

__device__ inline bool any_empty(const int8_t* b, int n) {
  for (int i = 0; i < n; ++i) if (b[i] == 0) return true;
  return false;
}

__device__ inline void env_reset_game(const EnvView& e, int g) {
  if (e.kind == MZB_ENV_SYNTHETIC_FRAMES) return;
  if (e.kind == MZB_ENV_CARTPOLE) {
    // gym reset: U(-0.05, 0.05)^4, drawn from Philox(slot, steps so far, STREAM_RESET)
    const uint32_t slot = e.slot0 + (uint32_t)g, st = e.steps[g];
    const Philox4 r0 = rng_draw(e.key, slot, st, MZB_STREAM_RESET, 0, 0);
    const Philox4 r1 = rng_draw(e.key, slot, st, MZB_STREAM_RESET, 0, 1);
    double* s = e.cp + (size_t)g * 4;
    s[0] = __dadd_rn(-0.05, __dmul_rn(0.1, u01_double(r0.x, r0.y)));
    s[1] = __dadd_rn(-0.05, __dmul_rn(0.1, u01_double(r0.z, r0.w)));
    s[2] = __dadd_rn(-0.05, __dmul_rn(0.1, u01_double(r1.x, r1.y)));
    s[3] = __dadd_rn(-0.05, __dmul_rn(0.1, u01_double(r1.z, r1.w)));
    e.elapsed[g] = 0;
  } else {
    int8_t* b = board_of(e, g);
    for (int i = 0; i < e.cells; ++i) b[i] = 0;
    e.player[g] = 1;
  }
}

__device__ inline int env_to_play(const EnvView& e, int g) {
  return (e.kind == MZB_ENV_CARTPOLE || e.kind == MZB_ENV_SYNTHETIC_FRAMES) ? 0 : (e.player[g] == 1 ? 0 : 1);
}

__device__ inline bool env_legal(const EnvView& e, int g, int a) {
  switch (e.kind) {
    case MZB_ENV_CARTPOLE: return true;
    case MZB_ENV_SYNTHETIC_FRAMES: return true;
    case MZB_ENV_CONNECT4: return board_of(e, g)[5 * 7 + a] == 0;          // connect4.py:249-254
    default: return board_of(e, g)[a] == 0;                                  // tictactoe.py:270-277, gomoku.py:247-253
  }
}

// Game.step: returns reward (already scaled by the wrapper), sets done
__device__ inline double env_step_game(const EnvView& e, int g, int a, bool& done) {
  if (e.kind == MZB_ENV_SYNTHETIC_FRAMES) { done = false; return 0.0; }   // rewards 0, never done before max_moves
  if (e.kind == MZB_ENV_CARTPOLE) {
    double* s = e.cp + (size_t)g * 4;
    double x = s[0], xd = s[1], th = s[2], thd = s[3];
    const double force = a == 1 ? kForceMag : -kForceMag;
    const double c = cos(th), sn = sin(th);
    // gym CartPoleEnv.step, float64, operation order of the Python source (no FMA contraction)
    const double temp = __ddiv_rn(__dadd_rn(force, __dmul_rn(__dmul_rn(kPoleMassLength, __dmul_rn(thd, thd)), sn)), kTotalMass);
    const double thacc = __ddiv_rn(__dsub_rn(__dmul_rn(kGravity, sn), __dmul_rn(c, temp)),
                                   __dmul_rn(kLength, __dsub_rn(4.0 / 3.0, __ddiv_rn(__dmul_rn(kMassPole, __dmul_rn(c, c)), kTotalMass))));
    const double xacc = __dsub_rn(temp, __ddiv_rn(__dmul_rn(__dmul_rn(kPoleMassLength, thacc), c), kTotalMass));
    x = __dadd_rn(x, __dmul_rn(kTau, xd));
    xd = __dadd_rn(xd, __dmul_rn(kTau, xacc));
    th = __dadd_rn(th, __dmul_rn(kTau, thd));
    thd = __dadd_rn(thd, __dmul_rn(kTau, thacc));
    s[0] = x; s[1] = xd; s[2] = th; s[3] = thd;
    const int el = ++e.elapsed[g];
    done = x < -kXThreshold || x > kXThreshold || th < -kThetaThreshold || th > kThetaThreshold || el >= kCartpoleLimit;
    return 1.0;
  }
  int8_t* b = board_of(e, g);
  const int p = e.player[g];
  double reward = 0.0;
  if (e.kind == MZB_ENV_TICTACTOE) {
    b[a] = (int8_t)p;                                                        // tictactoe.py:256-258
    const bool win = run_of(b, 3, 3, p, 3, 0, 1) || run_of(b, 3, 3, p, 3, 1, 0) || run_of(b, 3, 3, p, 3, 1, 1) ||
                     run_of(b, 3, 3, p, 3, 1, -1);
    done = win || !any_empty(b, 9);
    reward = win ? 1.0 : 0.0;
  } else if (e.kind == MZB_ENV_CONNECT4) {
    for (int r = 0; r < 6; ++r)                                              // lowest empty row; full column = no-op
      if (b[r * 7 + a] == 0) { b[r * 7 + a] = (int8_t)p; break; }
    const bool win = run_of(b, 6, 7, p, 4, 0, 1) || run_of(b, 6, 7, p, 4, 1, 0) || run_of(b, 6, 7, p, 4, 1, 1) ||
                     run_of(b, 6, 7, p, 4, -1, 1);
    bool any = false;
    for (int cidx = 0; cidx < 7; ++cidx) any = any || b[5 * 7 + cidx] == 0;
    done = win || !any;
    reward = win ? 1.0 : 0.0;
  } else {                                                                   // gomoku.py:233-245, 255-284
    b[a] = (int8_t)p;
    bool fin = false;
    for (int colour = -1; colour <= 1 && !fin; colour += 2)
      fin = run_of(b, 11, 11, colour, 5, 1, -1) || run_of(b, 11, 11, colour, 5, 1, 0) ||
            run_of(b, 11, 11, colour, 5, 1, 1) || run_of(b, 11, 11, colour, 5, 0, 1);
    done = fin || !any_empty(b, 121);
    reward = done ? 1.0 : 0.0;
  }
  e.player[g] = (int8_t)-p;
  return reward * e.reward_scale;
}

// observation record of the current state into a history slot (compact native form)
__device__ inline void write_obs_record(const EnvView& e, int g, float* rec) {
  if (e.kind == MZB_ENV_SYNTHETIC_FRAMES) { rec[0] = (float)e.steps[g]; return; }   // frames are regenerable from (slot, step)
  if (e.kind == MZB_ENV_CARTPOLE) {
    const double* s = e.cp + (size_t)g * 4;
    for (int i = 0; i < 4; ++i) rec[i] = (float)s[i];
  } else {
    int8_t* dst = reinterpret_cast<int8_t*>(rec);
    const int8_t* b = board_of(e, g);
    for (int i = 0; i < e.cells; ++i) dst[i] = b[i];
    dst[e.cells] = e.player[g];
  }
}

__global__ void k_env_reset(EnvView e, int all) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= e.G) return;
  if (all) { e.steps[g] = 0; }
  env_reset_game(e, g);
  e.h_len[g] = 0;
  e.finished[g] = 0;
  const size_t h0 = (size_t)g * e.hist_cap;
  write_obs_record(e, g, e.h_obs + h0 * e.rec_floats);
  e.h_action[h0] = 0;
  e.h_reward[h0] = 0.0f;
  e.h_to_play[h0] = (int8_t)env_to_play(e, g);
}

// Game observation / legal_actions / to_play for every game (what play_game hands to MCTS.run :138-150)
__global__ void k_env_observe(EnvView e, float* __restrict__ obs, uint8_t* __restrict__ legal,
                              int8_t* __restrict__ to_play, uint32_t* __restrict__ slot, uint32_t* __restrict__ step) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= e.G) return;
  if (obs) {
    float* o = obs + (size_t)g * e.obs_dim;
    if (e.kind == MZB_ENV_SYNTHETIC_FRAMES) {
      // written by k_synthetic_frames
    } else if (e.kind == MZB_ENV_CARTPOLE) {
      const double* s = e.cp + (size_t)g * 4;
      for (int i = 0; i < 4; ++i) o[i] = (float)s[i];
    } else {
      const int8_t* b = board_of(e, g);
      const float pl = (float)e.player[g];
      for (int i = 0; i < e.cells; ++i) {
        o[i] = b[i] == 1 ? 1.0f : 0.0f;
        o[e.cells + i] = b[i] == -1 ? 1.0f : 0.0f;
        o[2 * e.cells + i] = pl;
      }
    }
  }
  if (legal) for (int a = 0; a < e.A; ++a) legal[(size_t)g * e.A + a] = env_legal(e, g, a) ? 1 : 0;
  if (to_play) to_play[g] = (int8_t)env_to_play(e, g);
  if (slot) slot[g] = e.slot0 + (uint32_t)g;
  if (step) step[g] = e.steps[g];
}

// GameHistory.get_stacked_observations(-1, S) (self_play.py:514-548) for every game, from the running history on the
// device: planes of the current observation, then for k = 1..S the observation k moves ago followed by a plane holding
// the action that was played from it (action_history[index - k + 1], NOT normalised, as in the reference), or
// all-zero planes before the start of the game.  Layout [G][(C (S + 1) + S) * H * W], C planes of H*W per observation.
__global__ void k_env_observe_stacked(EnvView e, int S, int C, float* __restrict__ obs) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= e.G) return;
  const int hw = e.obs_dim / C, len = e.h_len[g];
  float* o = obs + (size_t)g * ((size_t)(C * (S + 1) + S) * hw);
  for (int k = 0; k <= S; ++k) {
    const int idx = len - k;
    float* dst = o + (k == 0 ? 0 : (size_t)C * hw + (size_t)(k - 1) * (C + 1) * hw);
    if (idx < 0) {
      for (int i = 0; i < (C + 1) * hw; ++i) dst[i] = 0.0f;
      continue;
    }
    const float* rec = e.h_obs + ((size_t)g * e.hist_cap + idx) * e.rec_floats;
    if (e.kind == MZB_ENV_CARTPOLE) {
      for (int i = 0; i < hw; ++i) dst[i] = rec[i];
    } else {
      const int8_t* b = reinterpret_cast<const int8_t*>(rec);
      const float pl = (float)b[e.cells];
      for (int i = 0; i < e.cells; ++i) {
        dst[i] = b[i] == 1 ? 1.0f : 0.0f;
        dst[e.cells + i] = b[i] == -1 ? 1.0f : 0.0f;
        dst[2 * e.cells + i] = pl;
      }
    }
    if (k > 0) {
      const float act = (float)e.h_action[(size_t)g * e.hist_cap + idx + 1];
      for (int i = 0; i < hw; ++i) dst[(size_t)C * hw + i] = act;
    }
  }
}

// Synthetic 3x96x96 frames U[0,1) (BASELINE.json: breakout on synthetic frames), a pure function of
// (seed, slot, step, pixel): four pixels per Philox call, coalesced float4 stores.
__global__ void k_synthetic_frames(EnvView e, float* __restrict__ obs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one float4
  const int per = e.obs_dim / 4;
  if (i >= (long long)e.G * per) return;
  const int g = (int)(i / per), q = (int)(i % per);
  const Philox4 r = rng_draw(e.key, e.slot0 + (uint32_t)g, e.steps[g], MZB_STREAM_RESET, 1, (uint32_t)q);
  const float k = 1.0f / 16777216.0f;
  reinterpret_cast<float4*>(obs)[i] = make_float4((r.x >> 8) * k, (r.y >> 8) * k, (r.z >> 8) * k, (r.w >> 8) * k);
}

// SelfPlay.select_action (self_play.py:223-246): children = legal actions in order, counts = visits.
__device__ inline int select_action_dev(const int* visits, const uint8_t* legal, int A, double temperature, double u) {
  int n = 0;
  for (int a = 0; a < A; ++a) n += legal[a] ? 1 : 0;
  if (temperature == 0.0) {                       // numpy.argmax: first maximum
    int best = -1, bv = -1;
    for (int a = 0; a < A; ++a) if (legal[a] && visits[a] > bv) { bv = visits[a]; best = a; }
    return best;
  }
  if (isinf(temperature)) {                       // numpy.random.choice(actions): index floor(u * n)
    int k = (int)(u * (double)n), j = 0;
    for (int a = 0; a < A; ++a) if (legal[a]) { if (j == k) return a; ++j; }
    return -1;
  }
  const double ex = __ddiv_rn(1.0, temperature);
  const bool e1 = ex == 1.0, e2 = ex == 2.0, e4 = ex == 4.0;
  auto powv = [&](int v) {
    const double d = (double)v;
    if (e1) return d;
    if (e2) return __dmul_rn(d, d);
    if (e4) { const double q = __dmul_rn(d, d); return __dmul_rn(q, q); }
    return pow(d, ex);
  };
  double tot = 0.0;                               // Python sum(): left to right
  for (int a = 0; a < A; ++a) if (legal[a]) tot = __dadd_rn(tot, powv(visits[a]));
  double last = 0.0;                              // cdf[-1] of the cumulative sum
  for (int a = 0; a < A; ++a) if (legal[a]) last = __dadd_rn(last, __ddiv_rn(powv(visits[a]), tot));
  double c = 0.0;
  int lastlegal = -1;
  for (int a = 0; a < A; ++a) {
    if (!legal[a]) continue;
    c = __dadd_rn(c, __ddiv_rn(powv(visits[a]), tot));
    lastlegal = a;
    if (__ddiv_rn(c, last) > u) return a;         // searchsorted(cdf / cdf[-1], u, side="right")
  }
  return lastlegal;
}

// One move of play_game for every game (self_play.py:152-182): pick the action from the search result,
// step the environment, append to the running GameHistory.  forced_action != NULL overrides the choice.
__global__ void k_act_step(EnvView e, const int* __restrict__ visits, const double* __restrict__ root_value,
                           const uint8_t* __restrict__ legal, double temperature, int temperature_threshold,
                           const double* __restrict__ uniforms, const int* __restrict__ forced_action,
                           int* __restrict__ out_action, float* __restrict__ out_reward, uint8_t* __restrict__ out_done) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= e.G) return;
  if (e.finished[g]) return;                      // waiting for harvest
  const int m = e.h_len[g];                       // moves so far; len(action_history) = m + 1
  const int A = e.A;
  const int* v = visits ? visits + (size_t)g * A : nullptr;
  const uint8_t* lg = legal + (size_t)g * A;
  int action;
  if (forced_action) {
    action = forced_action[g];
  } else {
    const double T = (temperature_threshold <= 0 || m + 1 < temperature_threshold) ? temperature : 0.0;
    double u;
    if (uniforms) u = uniforms[g];
    else {
      const Philox4 r = rng_draw(e.key, e.slot0 + (uint32_t)g, e.steps[g], MZB_STREAM_ACTION, 0, 0);
      u = u01_double(r.x, r.y);
    }
    action = select_action_dev(v, lg, A, T, u);
  }
  bool done = false;
  const double reward = env_step_game(e, g, action, done);
  e.steps[g] += 1;
  // GameHistory.store_search_statistics + appends (self_play.py:176-182)
  const size_t h0 = (size_t)g * e.hist_cap;
  uint16_t* hv = e.h_visits + ((size_t)g * e.max_moves + m) * A;
  for (int a = 0; a < A; ++a) hv[a] = (uint16_t)((v && lg[a]) ? v[a] : 0);
  e.h_root_value[(size_t)g * e.max_moves + m] = root_value ? root_value[g] : 0.0;
  e.h_action[h0 + m + 1] = action;
  e.h_reward[h0 + m + 1] = (float)reward;
  e.h_to_play[h0 + m + 1] = (int8_t)env_to_play(e, g);
  write_obs_record(e, g, e.h_obs + (h0 + m + 1) * e.rec_floats);
  e.h_len[g] = m + 1;
  const bool fin = done || (m + 1 >= e.max_moves);       // while not done and len(action_history) <= max_moves
  if (fin) e.finished[g] = 1;
  if (out_action) out_action[g] = action;
  if (out_reward) out_reward[g] = (float)reward;
  if (out_done) out_done[g] = fin ? 1 : 0;
  // one atomic per warp (the lanes that returned early are not in the active mask)
  const unsigned am = __activemask();
  if ((threadIdx.x & 31) == (unsigned)(__ffs(am) - 1)) atomicAdd(e.counters + 2, (unsigned long long)__popc(am));
}

// Finished games: one warp per game copies the episode into the export ring (the GameHistory wire format
// handed to ReplayBuffer.save_game, self_play.py:52) and restarts the game (auto-reset keeps the batch full).
// Export placement of the finished games, in GAME ORDER (an ordered scan, not atomics: the ring contents - hence
// everything downstream, e.g. the replay store's sampling - are reproducible from run to run).  Two launches: per-block
// totals of (entries, games) over spans of 8,192 games, then every block offsets its own in-block scan by the totals of
// the blocks before it.  (One block scanning all games took 158 us per move at 606,208 games - 2 % of a cartpole move.)
constexpr int kPlanThreads = 1024, kPlanItems = 8, kPlanSpan = kPlanThreads * kPlanItems;

// inclusive scan of (ve, vg) over the block's threads; s_e / s_g [32] hold the per-warp inclusive totals afterwards
__device__ __forceinline__ void plan_block_scan(int& ve, int& vg, int* s_e, int* s_g) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int ue = __shfl_up_sync(0xFFFFFFFFu, ve, o), ug = __shfl_up_sync(0xFFFFFFFFu, vg, o);
    if (lane >= o) { ve += ue; vg += ug; }
  }
  if (lane == 31) { s_e[warp] = ve; s_g[warp] = vg; }
  __syncthreads();
  if (warp == 0) {
    int we = s_e[lane], wg = s_g[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int ue = __shfl_up_sync(0xFFFFFFFFu, we, o), ug = __shfl_up_sync(0xFFFFFFFFu, wg, o);
      if (lane >= o) { we += ue; wg += ug; }
    }
    s_e[lane] = we; s_g[lane] = wg;
  }
  __syncthreads();
}

// plan[2 G ..]: per-block totals [n_blocks][2], then a snapshot of the ring cursor [2]
__global__ void __launch_bounds__(kPlanThreads) k_harvest_count(EnvView e, ExportView x) {
  __shared__ int s_e[32], s_g[32];
  const int gfirst = blockIdx.x * kPlanSpan + threadIdx.x * kPlanItems;
  int te = 0, tg = 0;
#pragma unroll
  for (int k = 0; k < kPlanItems; ++k) {
    const int g = gfirst + k;
    if (g < e.G && e.finished[g]) { te += e.h_len[g] + 1; tg += 1; }
  }
  plan_block_scan(te, tg, s_e, s_g);
  if (threadIdx.x == 0) {
    int* tot = x.plan + 2 * (size_t)e.G;
    tot[2 * blockIdx.x] = s_e[31];
    tot[2 * blockIdx.x + 1] = s_g[31];
    if (blockIdx.x == 0) { tot[2 * gridDim.x] = x.cursor[0]; tot[2 * gridDim.x + 1] = x.cursor[1]; }
  }
}

__global__ void __launch_bounds__(kPlanThreads) k_harvest_plan(EnvView e, ExportView x) {
  __shared__ int s_e[32], s_g[32], s_base[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int* tot = x.plan + 2 * (size_t)e.G;
  if (warp == 0) {                                          // ring cursor + totals of the blocks before this one
    int be = 0, bg = 0;
    for (int b = lane; b < (int)blockIdx.x; b += 32) { be += tot[2 * b]; bg += tot[2 * b + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { be += __shfl_xor_sync(0xFFFFFFFFu, be, o); bg += __shfl_xor_sync(0xFFFFFFFFu, bg, o); }
    if (lane == 0) { s_base[0] = tot[2 * gridDim.x] + be; s_base[1] = tot[2 * gridDim.x + 1] + bg; }
  }
  __syncthreads();
  const int gfirst = blockIdx.x * kPlanSpan + threadIdx.x * kPlanItems;
  int len[kPlanItems];
  int te = 0, tg = 0;                                       // this thread's totals
#pragma unroll
  for (int k = 0; k < kPlanItems; ++k) {
    const int g = gfirst + k;
    len[k] = (g < e.G && e.finished[g]) ? e.h_len[g] : -1;
    if (len[k] >= 0) { te += len[k] + 1; tg += 1; }
  }
  int ve = te, vg = tg;
  plan_block_scan(ve, vg, s_e, s_g);
  int start = s_base[0] + (warp ? s_e[warp - 1] : 0) + ve - te;
  int gi = s_base[1] + (warp ? s_g[warp - 1] : 0) + vg - tg;
#pragma unroll
  for (int k = 0; k < kPlanItems; ++k) {
    if (len[k] < 0) continue;
    const int g = gfirst + k;
    const bool ok = start + len[k] + 1 <= x.cap_entries && gi < x.cap_games;
    x.plan[2 * g] = ok ? start : -1;
    x.plan[2 * g + 1] = gi;
    if (ok) { x.game_start[gi] = start; x.game_len[gi] = len[k]; x.game_slot[gi] = e.slot0 + (uint32_t)g; }
    else atomicAdd(e.counters + 3, 1ull);
    start += len[k] + 1;
    gi += 1;
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {    // the other blocks read the snapshot, not the cursor
    x.cursor[0] = s_base[0] + s_e[31];
    x.cursor[1] = s_base[1] + s_g[31];
  }
}

__global__ void k_harvest(EnvView e, ExportView x, int do_export) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
  if (warp >= e.G) return;
  const int g = warp;
  if (!e.finished[g]) return;
  const int len = e.h_len[g];
  const int start = do_export ? x.plan[2 * g] : -1;
  const size_t h0 = (size_t)g * e.hist_cap;
  if (start >= 0) {
    const int n_obs = (len + 1) * e.rec_floats;
    for (int i = lane; i < n_obs; i += 32) x.obs[(size_t)start * e.rec_floats + i] = e.h_obs[h0 * e.rec_floats + i];
    for (int i = lane; i <= len; i += 32) {
      x.action[start + i] = e.h_action[h0 + i];
      x.reward[start + i] = e.h_reward[h0 + i];
      x.to_play[start + i] = e.h_to_play[h0 + i];
    }
    for (int i = lane; i < len * e.A; i += 32) x.visits[(size_t)start * e.A + i] = e.h_visits[(size_t)g * e.max_moves * e.A + i];
    for (int i = lane; i < len; i += 32) x.root_value[start + i] = e.h_root_value[(size_t)g * e.max_moves + i];
  }
  __syncwarp();
  if (lane == 0) {
    atomicAdd(e.counters + 0, 1ull);
    atomicAdd(e.counters + 1, (unsigned long long)len);
    env_reset_game(e, g);
    e.h_len[g] = 0;
    e.finished[g] = 0;
    write_obs_record(e, g, e.h_obs + h0 * e.rec_floats);
    e.h_action[h0] = 0;
    e.h_reward[h0] = 0.0f;
    e.h_to_play[h0] = (int8_t)env_to_play(e, g);
  }
}

struct Layout {
  size_t cp, elapsed, board, player, steps, h_obs, h_action, h_reward, h_to_play, h_visits, h_root, h_len, finished,
      counters, x_obs, x_action, x_reward, x_to_play, x_visits, x_root, x_gstart, x_glen, x_gslot, x_cursor, x_plan, total;
};

int describe(const mzb_env_config& c, EnvView& v) {
  v.kind = c.kind; v.G = c.n_games; v.max_moves = c.max_moves;
  switch (c.kind) {
    case MZB_ENV_CARTPOLE: v.A = 2; v.H = 1; v.W = 4; v.cells = 0; v.obs_dim = 4; v.reward_scale = 1; v.rec_floats = 4; break;
    case MZB_ENV_TICTACTOE: v.A = 9; v.H = 3; v.W = 3; v.cells = 9; v.obs_dim = 27; v.reward_scale = 20; v.rec_floats = 3; break;
    case MZB_ENV_CONNECT4: v.A = 7; v.H = 6; v.W = 7; v.cells = 42; v.obs_dim = 126; v.reward_scale = 10; v.rec_floats = 11; break;
    case MZB_ENV_SYNTHETIC_FRAMES: v.A = 4; v.H = 96; v.W = 96; v.cells = 0; v.obs_dim = 3 * 96 * 96; v.reward_scale = 1; v.rec_floats = 1; break;
    case MZB_ENV_GOMOKU: v.A = 121; v.H = 11; v.W = 11; v.cells = 121; v.obs_dim = 363; v.reward_scale = 1; v.rec_floats = 31; break;
    default: mzb_set_error("unknown environment kind %d", c.kind); return MZB_EINVAL;
  }
  v.hist_cap = c.max_moves + 1;
  return MZB_OK;
}

Layout layout(const mzb_env_config& c, const EnvView& v) {
  Layout o;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t at = off; off = mzb_align_up(off + bytes, 256); return at; };
  const size_t G = c.n_games, T = c.max_moves, T1 = T + 1, A = v.A;
  o.cp = take(G * 4 * 8); o.elapsed = take(G * 4);
  o.board = take(G * (size_t)(v.cells ? v.cells : 1)); o.player = take(G);
  o.steps = take(G * 4);
  o.h_obs = take(G * T1 * v.rec_floats * 4); o.h_action = take(G * T1 * 4); o.h_reward = take(G * T1 * 4);
  o.h_to_play = take(G * T1); o.h_visits = take(G * T * A * 2); o.h_root = take(G * T * 8);
  o.h_len = take(G * 4); o.finished = take(G); o.counters = take(8 * 8);
  const size_t E = c.export_entries, XG = c.export_games;
  o.x_obs = take(E * v.rec_floats * 4); o.x_action = take(E * 4); o.x_reward = take(E * 4); o.x_to_play = take(E);
  o.x_visits = take(E * A * 2); o.x_root = take(E * 8);
  o.x_gstart = take(XG * 4); o.x_glen = take(XG * 4); o.x_gslot = take(XG * 4); o.x_cursor = take(2 * 4);
  o.x_plan = take(G * 2 * 4 + ((G + 8191) / 8192 + 1) * 2 * 4);     // + per-block totals and the cursor snapshot of k_harvest_count
  o.total = off;
  return o;
}

int validate(const mzb_env_config* c) {
  MZB_CHECK_ARG(c, "config is NULL");
  MZB_CHECK_ARG(c->n_games > 0 && c->max_moves > 0 && c->max_moves < 65535, "env config out of range");
  MZB_CHECK_ARG(c->export_entries >= 0 && c->export_games >= 0, "negative export capacity");
  return MZB_OK;
}

inline int blocks(int n, int t) { return (n + t - 1) / t; }

}  // namespace

extern "C" {

size_t mzb_env_workspace_bytes(const mzb_env_config* c) {
  if (validate(c) != MZB_OK) return 0;
  EnvView v{};
  if (describe(*c, v) != MZB_OK) return 0;
  return layout(*c, v).total;
}

int mzb_env_create(mzb_env** out, const mzb_env_config* c, void* d_workspace, size_t workspace_bytes, void* stream) {
  MZB_CHECK_ARG(out, "out is NULL");
  *out = nullptr;
  int rc = validate(c);
  if (rc) return rc;
  EnvView v{};
  rc = describe(*c, v);
  if (rc) return rc;
  const Layout o = layout(*c, v);
  MZB_CHECK_ARG(d_workspace && ((uintptr_t)d_workspace & 255) == 0, "workspace NULL or not 256-byte aligned");
  MZB_CHECK_ARG(workspace_bytes >= o.total, "workspace too small: %zu < %zu", workspace_bytes, o.total);
  uint8_t* w = (uint8_t*)d_workspace;
  v.cp = (double*)(w + o.cp); v.elapsed = (int*)(w + o.elapsed); v.board = (int8_t*)(w + o.board);
  v.player = (int8_t*)(w + o.player); v.steps = (uint32_t*)(w + o.steps);
  v.slot0 = c->first_slot; v.key = rng_key(c->seed);
  v.h_obs = (float*)(w + o.h_obs); v.h_action = (int*)(w + o.h_action); v.h_reward = (float*)(w + o.h_reward);
  v.h_to_play = (int8_t*)(w + o.h_to_play); v.h_visits = (uint16_t*)(w + o.h_visits); v.h_root_value = (double*)(w + o.h_root);
  v.h_len = (int*)(w + o.h_len); v.finished = (uint8_t*)(w + o.finished); v.counters = (unsigned long long*)(w + o.counters);
  mzb_env* e = new mzb_env();
  e->cfg = *c; e->v = v; e->bytes = o.total;
  ExportView& x = e->x;
  x.cap_entries = c->export_entries; x.cap_games = c->export_games;
  x.obs = (float*)(w + o.x_obs); x.action = (int*)(w + o.x_action); x.reward = (float*)(w + o.x_reward);
  x.to_play = (int8_t*)(w + o.x_to_play); x.visits = (uint16_t*)(w + o.x_visits); x.root_value = (double*)(w + o.x_root);
  x.game_start = (int*)(w + o.x_gstart); x.game_len = (int*)(w + o.x_glen); x.game_slot = (uint32_t*)(w + o.x_gslot);
  x.cursor = (int*)(w + o.x_cursor);
  x.plan = (int*)(w + o.x_plan);
  cudaStream_t s = (cudaStream_t)stream;
  MZB_CUDA(cudaMemsetAsync(w + o.counters, 0, 64, s));
  MZB_CUDA(cudaMemsetAsync(w + o.x_cursor, 0, 8, s));
  k_env_reset<<<blocks(v.G, 256), 256, 0, s>>>(v, 1);
  MZB_LAUNCH_CHECK();
  *out = e;
  return MZB_OK;
}

int mzb_env_destroy(mzb_env* e) { delete e; return MZB_OK; }

int mzb_env_info(const mzb_env* e, int32_t* n_actions, int32_t* obs_dim, int32_t* rec_floats) {
  MZB_CHECK_ARG(e, "env is NULL");
  if (n_actions) *n_actions = e->v.A;
  if (obs_dim) *obs_dim = e->v.obs_dim;
  if (rec_floats) *rec_floats = e->v.rec_floats;
  return MZB_OK;
}

int mzb_env_reset(mzb_env* e, void* stream) {
  MZB_CHECK_ARG(e, "env is NULL");
  k_env_reset<<<blocks(e->v.G, 256), 256, 0, (cudaStream_t)stream>>>(e->v, 1);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

int mzb_env_observe(mzb_env* e, float* d_obs, uint8_t* d_legal, int8_t* d_to_play, uint32_t* d_slot, uint32_t* d_step,
                    void* stream) {
  MZB_CHECK_ARG(e, "env is NULL");
  k_env_observe<<<blocks(e->v.G, 256), 256, 0, (cudaStream_t)stream>>>(e->v, d_obs, d_legal, d_to_play, d_slot, d_step);
  MZB_LAUNCH_CHECK();
  if (e->v.kind == MZB_ENV_SYNTHETIC_FRAMES && d_obs) {
    const long long n4 = (long long)e->v.G * (e->v.obs_dim / 4);
    k_synthetic_frames<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(e->v, d_obs);
    MZB_LAUNCH_CHECK();
  }
  return MZB_OK;
}

int mzb_env_observe_stacked(mzb_env* e, int32_t stacked_observations, float* d_obs, void* stream) {
  MZB_CHECK_ARG(e && d_obs, "NULL argument");
  MZB_CHECK_ARG(stacked_observations >= 0 && stacked_observations <= e->v.max_moves, "stacked_observations out of range: %d", stacked_observations);
  if (e->v.kind == MZB_ENV_SYNTHETIC_FRAMES) {
    mzb_set_error("stacked observations of regenerated synthetic frames are not kept on the device");
    return MZB_EUNSUPPORTED;
  }
  const int C = e->v.kind == MZB_ENV_CARTPOLE ? 1 : 3;
  k_env_observe_stacked<<<blocks(e->v.G, 128), 128, 0, (cudaStream_t)stream>>>(e->v, stacked_observations, C, d_obs);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

int mzb_env_act_step(mzb_env* e, const int32_t* d_visits, const double* d_root_value, const uint8_t* d_legal,
                     double temperature, int32_t temperature_threshold, const double* d_uniforms,
                     const int32_t* d_forced_action, int32_t* d_action, float* d_reward, uint8_t* d_done, void* stream) {
  MZB_CHECK_ARG(e && d_legal, "NULL argument");
  MZB_CHECK_ARG(d_visits || d_forced_action, "need visit counts or forced actions");
  MZB_CHECK_ARG(temperature >= 0.0, "negative temperature");
  k_act_step<<<blocks(e->v.G, 128), 128, 0, (cudaStream_t)stream>>>(e->v, d_visits, d_root_value, d_legal, temperature,
                                                                    temperature_threshold, d_uniforms, d_forced_action,
                                                                    d_action, d_reward, d_done);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

int mzb_env_harvest(mzb_env* e, int do_export, void* stream) {
  MZB_CHECK_ARG(e, "env is NULL");
  const long long threads = (long long)e->v.G * 32;
  if (do_export) {
    const unsigned nb = (unsigned)((e->v.G + kPlanSpan - 1) / kPlanSpan);
    k_harvest_count<<<nb, kPlanThreads, 0, (cudaStream_t)stream>>>(e->v, e->x);
    MZB_LAUNCH_CHECK();
    k_harvest_plan<<<nb, kPlanThreads, 0, (cudaStream_t)stream>>>(e->v, e->x);
    MZB_LAUNCH_CHECK();
  }
  k_harvest<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(e->v, e->x, do_export);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

int mzb_env_counters_sync(mzb_env* e, uint64_t* h_counters4, void* stream) {
  MZB_CHECK_ARG(e && h_counters4, "NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  MZB_CUDA(cudaMemcpyAsync(h_counters4, e->v.counters, 32, cudaMemcpyDeviceToHost, s));
  MZB_CUDA(cudaStreamSynchronize(s));
  return MZB_OK;
}

int mzb_env_state_ptrs(mzb_env* e, double** d_cartpole_state, int8_t** d_board, int8_t** d_player, int32_t** d_hist_len,
                       uint8_t** d_finished) {
  MZB_CHECK_ARG(e, "env is NULL");
  if (d_cartpole_state) *d_cartpole_state = e->v.cp;
  if (d_board) *d_board = e->v.board;
  if (d_player) *d_player = e->v.player;
  if (d_hist_len) *d_hist_len = e->v.h_len;
  if (d_finished) *d_finished = e->v.finished;
  return MZB_OK;
}

// Drain the export ring to host buffers sized by the caller (capacities from the config); resets the ring.
int mzb_env_export_drain_sync(mzb_env* e, int32_t* h_n_entries, int32_t* h_n_games, float* h_obs, int32_t* h_action,
                              float* h_reward, int8_t* h_to_play, uint16_t* h_visits, double* h_root_value,
                              int32_t* h_game_start, int32_t* h_game_len, uint32_t* h_game_slot, void* stream) {
  MZB_CHECK_ARG(e && h_n_entries && h_n_games, "NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  int cur[2];
  MZB_CUDA(cudaMemcpyAsync(cur, e->x.cursor, 8, cudaMemcpyDeviceToHost, s));
  MZB_CUDA(cudaStreamSynchronize(s));
  // games past capacity were dropped and counted; entries of dropped games may leave a tail gap
  const int ng = cur[1] < e->x.cap_games ? cur[1] : e->x.cap_games;
  const int ne = cur[0] < e->x.cap_entries ? cur[0] : e->x.cap_entries;
  *h_n_entries = ne; *h_n_games = ng;
  const size_t A = e->v.A, R = e->v.rec_floats;
  if (ne > 0) {
    if (h_obs) MZB_CUDA(cudaMemcpyAsync(h_obs, e->x.obs, ne * R * 4, cudaMemcpyDeviceToHost, s));
    if (h_action) MZB_CUDA(cudaMemcpyAsync(h_action, e->x.action, (size_t)ne * 4, cudaMemcpyDeviceToHost, s));
    if (h_reward) MZB_CUDA(cudaMemcpyAsync(h_reward, e->x.reward, (size_t)ne * 4, cudaMemcpyDeviceToHost, s));
    if (h_to_play) MZB_CUDA(cudaMemcpyAsync(h_to_play, e->x.to_play, (size_t)ne, cudaMemcpyDeviceToHost, s));
    if (h_visits) MZB_CUDA(cudaMemcpyAsync(h_visits, e->x.visits, ne * A * 2, cudaMemcpyDeviceToHost, s));
    if (h_root_value) MZB_CUDA(cudaMemcpyAsync(h_root_value, e->x.root_value, (size_t)ne * 8, cudaMemcpyDeviceToHost, s));
  }
  if (ng > 0) {
    if (h_game_start) MZB_CUDA(cudaMemcpyAsync(h_game_start, e->x.game_start, (size_t)ng * 4, cudaMemcpyDeviceToHost, s));
    if (h_game_len) MZB_CUDA(cudaMemcpyAsync(h_game_len, e->x.game_len, (size_t)ng * 4, cudaMemcpyDeviceToHost, s));
    if (h_game_slot) MZB_CUDA(cudaMemcpyAsync(h_game_slot, e->x.game_slot, (size_t)ng * 4, cudaMemcpyDeviceToHost, s));
  }
  MZB_CUDA(cudaMemsetAsync(e->x.cursor, 0, 8, s));
  MZB_CUDA(cudaStreamSynchronize(s));
  return MZB_OK;
}

int mzb_env_export_to_replay(mzb_env* e, mzb_replay* r, int32_t* h_n_games, void* stream) {
  MZB_CHECK_ARG(e && r, "NULL argument");
  // the store copies ring entries with ITS strides: refuse a store laid out for anything but this environment's records
  mzb_replay_config rc_{};
  if (int rc = mzb_replay_get_config(r, &rc_)) return rc;
  const bool board = e->v.kind == MZB_ENV_TICTACTOE || e->v.kind == MZB_ENV_CONNECT4 || e->v.kind == MZB_ENV_GOMOKU;
  MZB_CHECK_ARG(rc_.obs_floats == e->v.rec_floats, "replay store keeps %d floats per observation record, the environment's ring %d "
                "(create the store with this environment as record_env)", rc_.obs_floats, e->v.rec_floats);
  MZB_CHECK_ARG(rc_.n_actions == e->v.A, "replay store has %d actions, the environment %d", rc_.n_actions, e->v.A);
  MZB_CHECK_ARG(rc_.obs_decode == (board ? 1 : 0), "replay store obs_decode = %d does not match environment kind %d",
                rc_.obs_decode, e->v.kind);
  MZB_CHECK_ARG(!board || (rc_.obs_h == e->v.H && rc_.obs_w == e->v.W), "replay store decodes %dx%d boards, the environment plays %dx%d",
                rc_.obs_h, rc_.obs_w, e->v.H, e->v.W);
  MZB_CHECK_ARG(rc_.entry_stride >= e->v.max_moves + 2, "replay store slots hold %d entries, games of this environment up to %d",
                rc_.entry_stride, e->v.max_moves + 2);
  cudaStream_t s = (cudaStream_t)stream;
  int cur[2];
  MZB_CUDA(cudaMemcpyAsync(cur, e->x.cursor, 8, cudaMemcpyDeviceToHost, s));
  MZB_CUDA(cudaStreamSynchronize(s));
  const int ng = cur[1] < e->x.cap_games ? cur[1] : e->x.cap_games;
  if (h_n_games) *h_n_games = ng;
  if (ng > 0) {
    std::vector<int> start((size_t)ng), len((size_t)ng);
    MZB_CUDA(cudaMemcpyAsync(start.data(), e->x.game_start, (size_t)ng * 4, cudaMemcpyDeviceToHost, s));
    MZB_CUDA(cudaMemcpyAsync(len.data(), e->x.game_len, (size_t)ng * 4, cudaMemcpyDeviceToHost, s));
    MZB_CUDA(cudaStreamSynchronize(s));
    for (int g0 = 0; g0 < ng; g0 += MZB_REPLAY_MAX_SAVE) {
      const int n = ng - g0 < MZB_REPLAY_MAX_SAVE ? ng - g0 : MZB_REPLAY_MAX_SAVE;
      const int rc = mzb_replay_save_games(r, n, start.data() + g0, len.data() + g0, e->x.obs, e->x.action, e->x.reward,
                                           e->x.to_play, e->x.root_value, e->x.visits, nullptr, stream);
      if (rc) return rc;
    }
  }
  MZB_CUDA(cudaMemsetAsync(e->x.cursor, 0, 8, s));
  return MZB_OK;
}

}  // extern "C"
