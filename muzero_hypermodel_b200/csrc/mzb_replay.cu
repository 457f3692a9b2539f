// Replay store on the device: the consumer of the self-play export ring.
// Reference: ReplayBuffer (replay_buffer.py) - save_game :33-64, get_batch :69-140, sample_n_games :162-177,
// sample_position :179-192, update_priorities :202-220; targets by K11 (mzb_targets.cu).
//
// Layout (caller-owned workspace).  A game occupies a fixed region of `entry_stride` entries in slot
// game_id % capacity_games, in the export layout of mzb_env / mzb_make_target (entry arrays; a game of n moves has
// n + 1 entries), so eviction is FIFO by game id exactly as the reference's dict (:57-60) and needs no compaction:
//   obs f32 [E][obs_floats] | action i32 [E] | reward f32 [E] | to_play i8 [E] | root_value f64 [E] |
//   visits u16 [E][A] | priority f32 [E]            E = capacity_games * entry_stride
//   game table by slot: start i32 (= slot * entry_stride), len i32, game_priority f32
//   sampling scratch in buffer order (oldest game first): probs f32 [capacity], cdf f64 [capacity]
//
// The float32 / float64 arithmetic of the reference is restated operation by operation (oracle/replay.py): game
// probabilities = priority / numpy.sum(float32) with numpy's PAIRWISE summation order, position probabilities =
// priority / Python's left-to-right float32 sum, numpy.random.choice(p=...) = float64 sequential cumsum, divided by
// its last element, searchsorted(side='right') on one uniform.  Uniforms are injected or drawn from Philox
// (slot = batch element, step = batch counter, streams MZB_STREAM_RGAME / MZB_STREAM_RPOS).
#include <math.h>

#include <vector>

#include "mzb_common.cuh"

extern "C" int mzb_make_target(const float* d_reward, const int8_t* d_to_play, const double* d_root_value,
                               const double* d_reanalysed_root_value, const uint16_t* d_visits, const int32_t* d_action,
                               const int32_t* d_game_start, const int32_t* d_game_len, int32_t n_actions,
                               const int32_t* d_batch_game, const int32_t* d_batch_index, const uint32_t* d_batch_slot,
                               const uint32_t* d_batch_step, int32_t batch, int32_t num_unroll_steps, int32_t td_steps,
                               const double* d_discount_pow, uint64_t seed, double* d_target_value,
                               double* d_target_reward, double* d_target_policy, int32_t* d_actions, void* stream);

namespace {

struct ReplayView {
  float* obs; int* action; float* reward; int8_t* to_play; double* root_value; uint16_t* visits; float* priority;
  double* reanalysed;          // GameHistory.reanalysed_predicted_root_values; = root_value until Reanalyse writes it
  int* g_start; int* g_len; float* g_priority;
  float* probs; double* cdf;
  double* discount_pow;
  int* b_slot; uint32_t* b_step; float* b_w;       // per batch element scratch [max_batch]
  int A, obs_floats, cap, stride, K, td, per, max_batch, decode, oh, ow;
  int S, C;                    // stacked_observations, channels per observation
  double alpha; RngKey key;
  // one observation as network input: `per` floats = C planes of hw; the stacked form adds S x (per + hw)
  __host__ __device__ int hw() const { return decode ? oh * ow : obs_floats / C; }
  __host__ __device__ int per_obs() const { return decode ? 3 * oh * ow : obs_floats; }
  __host__ __device__ int out_floats() const { return per_obs() + S * (per_obs() + hw()); }
};

// Element i of GameHistory.get_stacked_observations(p, S) (self_play.py:514-548) of the game whose entries start at s:
// the observation at p, then for k = 1..S the observation at p - k followed by a plane holding action_history[p - k + 1]
// (not normalised, as in the reference), all zeros before the start of the game.  S = 0: the observation itself.
__device__ __forceinline__ float replay_obs_value(const ReplayView& v, int s, int p, int i) {
  const int hw = v.hw(), per = v.per_obs();
  int k = 0, j = i;
  if (i >= per) { const int r = i - per; k = 1 + r / (per + hw); j = r - (k - 1) * (per + hw); }
  const int idx = p - k;
  if (idx < 0) return 0.0f;
  if (j >= per) return (float)v.action[s + idx + 1];
  const float* src = v.obs + (long long)(s + idx) * v.obs_floats;
  if (!v.decode) return src[j];
  const int8_t* raw = reinterpret_cast<const int8_t*>(src);       // packed board record -> [own, other, to-play] planes
  const int plane = j / hw, c = j - plane * hw;
  return plane == 0 ? (raw[c] == 1 ? 1.0f : 0.0f) : (plane == 1 ? (raw[c] == -1 ? 1.0f : 0.0f) : (float)raw[hw]);
}

// compute_target_value (replay_buffer.py:222-254) on the store's arrays; s = first entry of the game, n = moves
__device__ double target_value(const ReplayView& v, int s, int n, int cur) {
  double value = 0.0;
  const int boot = cur + v.td;
  if (boot < n) {
    const double rv = v.root_value[s + boot];
    const double last = v.to_play[s + boot] == v.to_play[s + cur] ? rv : -rv;
    value = __dmul_rn(last, v.discount_pow[v.td]);
  }
  for (int k = 0; cur + 1 + k <= boot && cur + 1 + k <= n; ++k) {
    const double r = (double)v.reward[s + cur + 1 + k];
    const double signed_r = v.to_play[s + cur] == v.to_play[s + cur + k] ? r : -r;
    value = __dadd_rn(value, __dmul_rn(signed_r, v.discount_pow[k]));
  }
  return value;
}

struct SaveArgs {
  const float* obs; const int* action; const float* reward; const int8_t* to_play; const double* root_value;
  const uint16_t* visits; const float* priorities;
  const int* src_start; const int* len; const int* dst_slot;     // [n_games] device
};

// save_game (:33-64): one block per game copies its n + 1 entries into the slot, then computes the initial priorities
// |root_value - n-step target| ** PER_alpha (float64, stored float32) and the game priority = their maximum.
__global__ void __launch_bounds__(256) k_replay_save(ReplayView v, SaveArgs a) {
  const int g = blockIdx.x, n = a.len[g], src = a.src_start[g], slot = a.dst_slot[g];
  const int dst = slot * v.stride;
  for (int i = threadIdx.x; i <= n; i += blockDim.x) {
    v.action[dst + i] = a.action[src + i];
    v.reward[dst + i] = a.reward[src + i];
    v.to_play[dst + i] = a.to_play[src + i];
    v.root_value[dst + i] = i < n ? a.root_value[src + i] : 0.0;
    v.reanalysed[dst + i] = i < n ? a.root_value[src + i] : 0.0;
  }
  for (long long i = threadIdx.x; i < (long long)(n + 1) * v.obs_floats; i += blockDim.x)
    v.obs[(long long)dst * v.obs_floats + i] = a.obs[(long long)src * v.obs_floats + i];
  for (long long i = threadIdx.x; i < (long long)(n + 1) * v.A; i += blockDim.x)
    v.visits[(long long)dst * v.A + i] = a.visits[(long long)src * v.A + i];
  if (threadIdx.x == 0) { v.g_start[slot] = dst; v.g_len[slot] = n; }
  __syncthreads();
  __shared__ float s_max[256];
  float mx = -CUDART_INF_F;
  if (v.per) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      float p;
      if (a.priorities) {
        p = a.priorities[src + i];
      } else {
        const double d = fabs(__dsub_rn(v.root_value[dst + i], target_value(v, dst, n, i)));
        // numpy's float64 power; the square root is exact, as glibc's pow(x, 0.5) is for all but unobservably rare x
        const double pw = v.alpha == 0.5 ? __dsqrt_rn(d) : (v.alpha == 1.0 ? d : pow(d, v.alpha));
        p = __double2float_rn(pw);
      }
      v.priority[dst + i] = p;
      mx = fmaxf(mx, p);
    }
  }
  s_max[threadIdx.x] = mx;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_max[threadIdx.x] = fmaxf(s_max[threadIdx.x], s_max[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) v.g_priority[slot] = v.per ? s_max[0] : 0.0f;
}

// numpy's float32 pairwise summation (FLOAT_pairwise_sum: blocks of <= 128 with 8 interleaved accumulators, halves
// rounded down to a multiple of 8), restated in oracle/replay.py and checked there against numpy.sum itself.
__device__ float pairwise_block(const float* a, int n) {
  if (n < 8) {
    float r = 0.0f;
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
    return r;
  }
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
  }
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, a[i]);
  return res;
}
__device__ float pairwise_sum(const float* a, int n) {
  if (n <= 128) return pairwise_block(a, n);
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __fadd_rn(pairwise_sum(a, n2), pairwise_sum(a + n2, n - n2));
}

// sample_n_games, PER branch (:163-172): probabilities and their float64 CDF over the buffer in insertion order.
// One thread carries the two order-dependent reductions; the block does the element-wise parts.
__global__ void __launch_bounds__(256) k_replay_game_cdf(ReplayView v, long long first_id, int n_games) {
  __shared__ float s_sum;
  __shared__ double s_last;
  for (int j = threadIdx.x; j < n_games; j += blockDim.x) v.probs[j] = v.g_priority[(int)((first_id + j) % v.cap)];
  __syncthreads();
  if (threadIdx.x == 0) s_sum = pairwise_sum(v.probs, n_games);
  __syncthreads();
  for (int j = threadIdx.x; j < n_games; j += blockDim.x) v.probs[j] = __fdiv_rn(v.probs[j], s_sum);
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = 0.0;
    for (int j = 0; j < n_games; ++j) { acc = __dadd_rn(acc, (double)v.probs[j]); v.cdf[j] = acc; }
    s_last = acc;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n_games; j += blockDim.x) v.cdf[j] = __ddiv_rn(v.cdf[j], s_last);
}

struct SampleArgs {
  const double* u_game; const double* u_pos;
  long long first_id, total_samples; int n_games, B; uint32_t batch_counter;
  long long* game_id; int* pos; float* game_prob; float* pos_prob;
};

// get_batch's sampling (:84-85): thread per batch element.
__global__ void k_replay_sample(ReplayView v, SampleArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  double ug, up;
  if (a.u_game) ug = a.u_game[b];
  else { const Philox4 r = rng_draw(v.key, (uint32_t)b, a.batch_counter, MZB_STREAM_RGAME, 0, 0); ug = u01_double(r.x, r.y); }
  if (a.u_pos) up = a.u_pos[b];
  else { const Philox4 r = rng_draw(v.key, (uint32_t)b, a.batch_counter, MZB_STREAM_RPOS, 0, 0); up = u01_double(r.x, r.y); }
  int j, pos;
  float gp = 0.0f, pp = 0.0f;
  if (v.per) {
    int lo = 0, hi = a.n_games;                         // searchsorted(cdf, u, side='right'): first j with cdf[j] > u
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (v.cdf[mid] > ug) hi = mid; else lo = mid + 1; }
    j = lo < a.n_games ? lo : a.n_games - 1;
    gp = v.probs[j];
  } else {
    j = (int)__dmul_rn(ug, (double)a.n_games);
  }
  const int slot = (int)((a.first_id + j) % v.cap);
  const int s = v.g_start[slot], n = v.g_len[slot];
  if (v.per) {
    const float* p = v.priority + s;
    float tot = 0.0f;
    for (int i = 0; i < n; ++i) tot = __fadd_rn(tot, p[i]);        // Python sum() over float32 scalars
    double last = 0.0;
    for (int i = 0; i < n; ++i) last = __dadd_rn(last, (double)__fdiv_rn(p[i], tot));
    double acc = 0.0;
    pos = n - 1;
    for (int i = 0; i < n; ++i) {
      acc = __dadd_rn(acc, (double)__fdiv_rn(p[i], tot));
      if (__ddiv_rn(acc, last) > up) { pos = i; break; }
    }
    pp = __fdiv_rn(p[pos], tot);
    // 1 / (total_samples * game_prob * position_prob), float32 throughout (:117)
    v.b_w[b] = __fdiv_rn(1.0f, __fmul_rn(__fmul_rn((float)a.total_samples, gp), pp));
  } else {
    pos = (int)__dmul_rn(up, (double)n);
  }
  v.b_slot[b] = slot;
  v.b_step[b] = a.batch_counter;
  a.game_id[b] = a.first_id + j;
  a.pos[b] = pos;
  if (a.game_prob) a.game_prob[b] = gp;
  if (a.pos_prob) a.pos_prob[b] = pp;
}

// The rest of a batch row (:92-118): observation of the sampled position (stacked_observations = 0), gradient scale
// min(num_unroll_steps, len(action_history) - position), importance weight / max weight.
__global__ void __launch_bounds__(256) k_replay_assemble(ReplayView v, int B, const int* pos, float* obs, int* gscale,
                                                         float* weights) {
  __shared__ float s_max[256];
  if (v.per && weights) {
    float mx = -CUDART_INF_F;
    for (int b = threadIdx.x; b < B; b += blockDim.x) mx = fmaxf(mx, v.b_w[b]);
    s_max[threadIdx.x] = mx;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) s_max[threadIdx.x] = fmaxf(s_max[threadIdx.x], s_max[threadIdx.x + o]);
      __syncthreads();
    }
  }
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const int slot = v.b_slot[b], s = v.g_start[slot], n = v.g_len[slot], p = pos[b];
    if (obs) {
      const int nf = v.out_floats();
      for (int i = threadIdx.x; i < nf; i += blockDim.x) obs[(long long)b * nf + i] = replay_obs_value(v, s, p, i);
    }
    if (gscale) {
      const int gs = min(v.K, n + 1 - p);
      for (int i = threadIdx.x; i <= v.K; i += blockDim.x) gscale[b * (v.K + 1) + i] = gs;
    }
    if (v.per && weights && threadIdx.x == 0) weights[b] = __fdiv_rn(v.b_w[b], s_max[0]);
  }
}

// The observations of one stored game as network input [n, obs_out] (Reanalyse, replay_buffer.py:337-349).
__global__ void k_replay_game_obs(ReplayView v, int slot, float* out) {
  const int s = v.g_start[slot], n = v.g_len[slot];
  const int nf = v.out_floats();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * nf) return;
  const int e = (int)(i / nf);
  out[i] = replay_obs_value(v, s, e, (int)(i - (long long)e * nf));
}

__global__ void k_replay_set_reanalysed(ReplayView v, int slot, const float* values) {
  const int s = v.g_start[slot], n = v.g_len[slot];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v.reanalysed[s + i] = (double)values[i];
}

// update_priorities (:202-220).  The reference applies the batch rows in order (a later row overwrites an earlier one
// where their windows overlap), so one thread writes; the per-game maxima are then recomputed in parallel.
__global__ void __launch_bounds__(256) k_replay_update(ReplayView v, int B, const float* prio, const long long* game_id,
                                                       const int* pos, long long first_id, int n_games) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < B; ++i) {
      const long long gid = game_id[i];
      if (gid < first_id || gid >= first_id + n_games) continue;          // evicted since it was sampled (:211)
      const int slot = (int)(gid % v.cap), s = v.g_start[slot], n = v.g_len[slot];
      const int end = min(pos[i] + v.K + 1, n);
      for (int k = pos[i]; k < end; ++k) v.priority[s + k] = prio[(long long)i * (v.K + 1) + (k - pos[i])];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const long long gid = game_id[i];
    if (gid < first_id || gid >= first_id + n_games) continue;
    const int slot = (int)(gid % v.cap), s = v.g_start[slot], n = v.g_len[slot];
    float mx = -CUDART_INF_F;
    for (int k = 0; k < n; ++k) mx = fmaxf(mx, v.priority[s + k]);
    v.g_priority[slot] = mx;                                               // same value from every row of the game
  }
}

}  // namespace

struct mzb_replay {
  mzb_replay_config cfg;
  ReplayView v;
  std::vector<int> h_len;            // by slot
  long long first_id = 0, num_played_games = 0, num_played_steps = 0, total_samples = 0;
  int n_games = 0;
  uint32_t batch_counter = 0;
  bool cdf_valid = false;
  int* d_meta = nullptr;             // [3][max_save] src_start | len | dst_slot
  int max_save = 0;
};

namespace {

struct Layout { size_t obs, action, reward, to_play, root_value, reanalysed, visits, priority, g_start, g_len, g_priority, probs, cdf, dpow, b_slot, b_step, b_w, meta, total; };

Layout replay_layout(const mzb_replay_config& c) {
  Layout L{};
  const size_t E = (size_t)c.capacity_games * c.entry_stride;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off = mzb_align_up(off + bytes, 256); return o; };
  L.obs = take(E * c.obs_floats * sizeof(float));
  L.action = take(E * sizeof(int));
  L.reward = take(E * sizeof(float));
  L.to_play = take(E);
  L.root_value = take(E * sizeof(double));
  L.reanalysed = take(E * sizeof(double));
  L.visits = take(E * c.n_actions * sizeof(uint16_t));
  L.priority = take(E * sizeof(float));
  L.g_start = take((size_t)c.capacity_games * sizeof(int));
  L.g_len = take((size_t)c.capacity_games * sizeof(int));
  L.g_priority = take((size_t)c.capacity_games * sizeof(float));
  L.probs = take((size_t)c.capacity_games * sizeof(float));
  L.cdf = take((size_t)c.capacity_games * sizeof(double));
  L.dpow = take((size_t)(c.td_steps + 1) * sizeof(double));
  L.b_slot = take((size_t)c.max_batch * sizeof(int));
  L.b_step = take((size_t)c.max_batch * sizeof(uint32_t));
  L.b_w = take((size_t)c.max_batch * sizeof(float));
  L.meta = take((size_t)3 * MZB_REPLAY_MAX_SAVE * sizeof(int));
  L.total = off;
  return L;
}

bool config_ok(const mzb_replay_config* c) {
  return c && c->n_actions > 0 && c->n_actions <= 65535 && c->obs_floats > 0 && c->capacity_games > 0 && c->entry_stride > 1 &&
         c->num_unroll_steps >= 0 && c->td_steps > 0 && c->max_batch > 0 && (c->per == 0 || c->per == 1) && c->per_alpha >= 0.0 &&
         (c->obs_decode == 0 || (c->obs_decode == 1 && c->obs_h > 0 && c->obs_w > 0 && c->obs_h * c->obs_w + 1 <= 4 * c->obs_floats)) &&
         c->stacked_observations >= 0 && c->obs_channels >= 0 &&
         (c->stacked_observations == 0 || c->obs_decode == 1 || (c->obs_channels > 0 && c->obs_floats % c->obs_channels == 0));
}

}  // namespace

extern "C" {

size_t mzb_replay_workspace_bytes(const mzb_replay_config* c) { return config_ok(c) ? replay_layout(*c).total : 0; }

int mzb_replay_create(mzb_replay** out, const mzb_replay_config* c, void* d_workspace, size_t workspace_bytes,
                      const double* h_discount_pow, void* stream) {
  MZB_CHECK_ARG(out && d_workspace && h_discount_pow, "NULL argument");
  *out = nullptr;
  MZB_CHECK_ARG(config_ok(c), "replay config out of range");
  const Layout L = replay_layout(*c);
  MZB_CHECK_ARG(workspace_bytes >= L.total, "workspace too small: %zu < %zu", workspace_bytes, L.total);
  MZB_CHECK_ARG((reinterpret_cast<uintptr_t>(d_workspace) & 255) == 0, "workspace must be 256-byte aligned");
  mzb_replay* r = new mzb_replay();
  r->cfg = *c;
  uint8_t* w = (uint8_t*)d_workspace;
  ReplayView& v = r->v;
  v.obs = (float*)(w + L.obs); v.action = (int*)(w + L.action); v.reward = (float*)(w + L.reward);
  v.to_play = (int8_t*)(w + L.to_play); v.root_value = (double*)(w + L.root_value); v.reanalysed = (double*)(w + L.reanalysed); v.visits = (uint16_t*)(w + L.visits);
  v.priority = (float*)(w + L.priority); v.g_start = (int*)(w + L.g_start); v.g_len = (int*)(w + L.g_len);
  v.g_priority = (float*)(w + L.g_priority); v.probs = (float*)(w + L.probs); v.cdf = (double*)(w + L.cdf);
  v.discount_pow = (double*)(w + L.dpow); v.b_slot = (int*)(w + L.b_slot); v.b_step = (uint32_t*)(w + L.b_step);
  v.b_w = (float*)(w + L.b_w);
  v.A = c->n_actions; v.obs_floats = c->obs_floats; v.cap = c->capacity_games; v.stride = c->entry_stride;
  v.decode = c->obs_decode; v.oh = c->obs_h; v.ow = c->obs_w;
  v.K = c->num_unroll_steps; v.td = c->td_steps; v.per = c->per; v.max_batch = c->max_batch; v.alpha = c->per_alpha;
  v.S = c->stacked_observations; v.C = c->obs_channels > 0 ? c->obs_channels : 1;
  v.key = rng_key(c->seed);
  r->d_meta = (int*)(w + L.meta);
  r->max_save = MZB_REPLAY_MAX_SAVE;
  r->h_len.assign((size_t)c->capacity_games, 0);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemcpyAsync(v.discount_pow, h_discount_pow, sizeof(double) * (c->td_steps + 1), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { delete r; mzb_set_error("replay create: %s", cudaGetErrorString(e)); return MZB_ECUDA; }
  // the recursive pairwise sum needs a little more stack than the default for very large buffers
  *out = r;
  return MZB_OK;
}

int mzb_replay_destroy(mzb_replay* r) {
  delete r;
  return MZB_OK;
}

int mzb_replay_save_games(mzb_replay* r, int32_t n_games, const int32_t* h_src_start, const int32_t* h_len,
                          const float* d_obs, const int32_t* d_action, const float* d_reward, const int8_t* d_to_play,
                          const double* d_root_value, const uint16_t* d_visits, const float* d_priorities, void* stream) {
  MZB_CHECK_ARG(r && h_src_start && h_len && d_obs && d_action && d_reward && d_to_play && d_root_value && d_visits, "NULL argument");
  MZB_CHECK_ARG(n_games > 0 && n_games <= r->max_save, "1..%d games per call", r->max_save);
  std::vector<int> meta((size_t)3 * n_games);
  for (int g = 0; g < n_games; ++g) {
    MZB_CHECK_ARG(h_len[g] >= 1 && h_len[g] + 1 <= r->cfg.entry_stride, "game %d: %d moves do not fit entry_stride %d", g,
                  h_len[g], r->cfg.entry_stride);
    MZB_CHECK_ARG(h_src_start[g] >= 0, "negative source offset");
  }
  cudaStream_t s = (cudaStream_t)stream;
  for (int g = 0; g < n_games; ++g) {
    const long long id = r->num_played_games;
    const int slot = (int)(id % r->cfg.capacity_games);
    if (r->n_games == r->cfg.capacity_games) {          // del self.buffer[del_id] (:57-60): the slot's previous game
      r->total_samples -= r->h_len[slot];
      r->first_id += 1;
      r->n_games -= 1;
    }
    r->h_len[slot] = h_len[g];
    r->n_games += 1;
    r->num_played_games += 1;
    r->num_played_steps += h_len[g];
    r->total_samples += h_len[g];
    meta[g] = h_src_start[g]; meta[n_games + g] = h_len[g]; meta[2 * n_games + g] = slot;
  }
  // more games than slots in one call: the early ones are evicted by the later ones of the same call (the counters
  // above already say so) - only the last `capacity` games are materialised, so no two blocks write one slot
  const int first = n_games > r->cfg.capacity_games ? n_games - r->cfg.capacity_games : 0, n_copy = n_games - first;
  std::vector<int> packed((size_t)3 * n_copy);
  for (int g = 0; g < n_copy; ++g)
    for (int k = 0; k < 3; ++k) packed[(size_t)k * n_copy + g] = meta[(size_t)k * n_games + first + g];
  MZB_CUDA(cudaMemcpyAsync(r->d_meta, packed.data(), sizeof(int) * packed.size(), cudaMemcpyHostToDevice, s));
  MZB_CUDA(cudaStreamSynchronize(s));                     // `packed` is a stack-lifetime staging buffer
  SaveArgs a{d_obs, d_action, d_reward, d_to_play, d_root_value, d_visits, d_priorities, r->d_meta, r->d_meta + n_copy,
             r->d_meta + 2 * n_copy};
  k_replay_save<<<n_copy, 256, 0, s>>>(r->v, a);
  MZB_LAUNCH_CHECK();
  r->cdf_valid = false;
  return MZB_OK;
}

int mzb_replay_get_batch(mzb_replay* r, int32_t batch, const double* d_u_game, const double* d_u_pos, int64_t* d_game_id,
                         int32_t* d_pos, float* d_game_prob, float* d_pos_prob, float* d_obs, int32_t* d_actions,
                         double* d_values, double* d_rewards, double* d_policies, float* d_weights,
                         int32_t* d_gradient_scale, void* stream) {
  MZB_CHECK_ARG(r && d_game_id && d_pos, "NULL argument");
  MZB_CHECK_ARG(batch > 0 && batch <= r->cfg.max_batch, "batch %d outside 1..%d", batch, r->cfg.max_batch);
  if (r->n_games == 0) { mzb_set_error("replay buffer is empty"); return MZB_ESTATE; }
  cudaStream_t s = (cudaStream_t)stream;
  if (r->cfg.per && !r->cdf_valid) {
    static bool stack_set = false;
    if (!stack_set) { cudaDeviceSetLimit(cudaLimitStackSize, 8192); stack_set = true; }
    k_replay_game_cdf<<<1, 256, 0, s>>>(r->v, r->first_id, r->n_games);
    MZB_LAUNCH_CHECK();
    r->cdf_valid = true;
  }
  SampleArgs a{d_u_game, d_u_pos, r->first_id, r->total_samples, r->n_games, batch, r->batch_counter,
               (long long*)d_game_id, d_pos, d_game_prob, d_pos_prob};
  k_replay_sample<<<(batch + 127) / 128, 128, 0, s>>>(r->v, a);
  MZB_LAUNCH_CHECK();
  if (d_actions || d_values || d_rewards || d_policies) {
    MZB_CHECK_ARG(d_actions && d_values && d_rewards && d_policies, "targets: all four outputs or none");
    const int rc = mzb_make_target(r->v.reward, r->v.to_play, r->v.root_value, r->v.reanalysed, r->v.visits, r->v.action, r->v.g_start,
                                   r->v.g_len, r->v.A, r->v.b_slot, d_pos, nullptr, r->v.b_step, batch, r->v.K, r->v.td,
                                   r->v.discount_pow, r->cfg.seed, d_values, d_rewards, d_policies, d_actions, stream);
    if (rc) return rc;
  }
  if (d_obs || d_gradient_scale || d_weights) {
    k_replay_assemble<<<batch < 1184 ? batch : 1184, 256, 0, s>>>(r->v, batch, d_pos, d_obs, d_gradient_scale, d_weights);
    MZB_LAUNCH_CHECK();
  }
  r->batch_counter += 1;
  return MZB_OK;
}

int mzb_replay_update_priorities(mzb_replay* r, int32_t batch, const float* d_priorities, const int64_t* d_game_id,
                                 const int32_t* d_pos, void* stream) {
  MZB_CHECK_ARG(r && d_priorities && d_game_id && d_pos && batch > 0, "NULL argument");
  if (!r->cfg.per) return MZB_OK;
  k_replay_update<<<1, 256, 0, (cudaStream_t)stream>>>(r->v, batch, d_priorities, (const long long*)d_game_id, d_pos,
                                                     r->first_id, r->n_games);
  MZB_LAUNCH_CHECK();
  r->cdf_valid = false;
  return MZB_OK;
}

int mzb_replay_info(const mzb_replay* r, int64_t* out5) {
  MZB_CHECK_ARG(r && out5, "NULL argument");
  out5[0] = r->total_samples; out5[1] = r->num_played_games; out5[2] = r->num_played_steps; out5[3] = r->n_games;
  out5[4] = r->first_id;
  return MZB_OK;
}

int mzb_replay_get_config(const mzb_replay* r, mzb_replay_config* out) {
  MZB_CHECK_ARG(r && out, "NULL argument");
  *out = r->cfg;
  return MZB_OK;
}

int mzb_replay_set_batch_counter(mzb_replay* r, uint32_t counter) {
  MZB_CHECK_ARG(r, "NULL argument");
  r->batch_counter = counter;
  return MZB_OK;
}

int mzb_replay_game_priorities_sync(mzb_replay* r, int64_t game_id, float* h_priorities, float* h_game_priority,
                                    int32_t* h_len, void* stream) {
  MZB_CHECK_ARG(r && h_len, "NULL argument");
  MZB_CHECK_ARG(game_id >= r->first_id && game_id < r->first_id + r->n_games, "game %lld is not in the buffer", (long long)game_id);
  const int slot = (int)(game_id % r->cfg.capacity_games), n = r->h_len[slot];
  cudaStream_t s = (cudaStream_t)stream;
  *h_len = n;
  if (h_priorities) MZB_CUDA(cudaMemcpyAsync(h_priorities, r->v.priority + (size_t)slot * r->cfg.entry_stride, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
  if (h_game_priority) MZB_CUDA(cudaMemcpyAsync(h_game_priority, r->v.g_priority + slot, sizeof(float), cudaMemcpyDeviceToHost, s));
  MZB_CUDA(cudaStreamSynchronize(s));
  return MZB_OK;
}

int mzb_replay_game_observations(mzb_replay* r, int64_t game_id, float* d_obs, int32_t* h_len, void* stream) {
  MZB_CHECK_ARG(r && h_len, "NULL argument");
  MZB_CHECK_ARG(game_id >= r->first_id && game_id < r->first_id + r->n_games, "game %lld is not in the buffer", (long long)game_id);
  const int slot = (int)(game_id % r->cfg.capacity_games), n = r->h_len[slot];
  *h_len = n;
  if (d_obs) {
    const long long per = r->v.out_floats();
    k_replay_game_obs<<<(unsigned)((n * per + 255) / 256), 256, 0, (cudaStream_t)stream>>>(r->v, slot, d_obs);
    MZB_LAUNCH_CHECK();
  }
  return MZB_OK;
}

int mzb_replay_set_reanalysed(mzb_replay* r, int64_t game_id, const float* d_values, void* stream) {
  MZB_CHECK_ARG(r && d_values, "NULL argument");
  if (game_id < r->first_id || game_id >= r->first_id + r->n_games) return MZB_OK;     // evicted meanwhile (:194-196)
  const int slot = (int)(game_id % r->cfg.capacity_games), n = r->h_len[slot];
  k_replay_set_reanalysed<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(r->v, slot, d_values);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

int mzb_replay_export_game_sync(mzb_replay* r, int64_t game_id, int32_t* h_len, float* h_obs, int32_t* h_action,
                                float* h_reward, int8_t* h_to_play, double* h_root_value, uint16_t* h_visits, void* stream) {
  MZB_CHECK_ARG(r && h_len, "NULL argument");
  MZB_CHECK_ARG(game_id >= r->first_id && game_id < r->first_id + r->n_games, "game %lld is not in the buffer", (long long)game_id);
  const int slot = (int)(game_id % r->cfg.capacity_games), n = r->h_len[slot];
  const size_t e0 = (size_t)slot * r->cfg.entry_stride;
  cudaStream_t s = (cudaStream_t)stream;
  *h_len = n;
  const ReplayView& v = r->v;
  if (h_obs) MZB_CUDA(cudaMemcpyAsync(h_obs, v.obs + e0 * v.obs_floats, sizeof(float) * (n + 1) * v.obs_floats, cudaMemcpyDeviceToHost, s));
  if (h_action) MZB_CUDA(cudaMemcpyAsync(h_action, v.action + e0, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost, s));
  if (h_reward) MZB_CUDA(cudaMemcpyAsync(h_reward, v.reward + e0, sizeof(float) * (n + 1), cudaMemcpyDeviceToHost, s));
  if (h_to_play) MZB_CUDA(cudaMemcpyAsync(h_to_play, v.to_play + e0, (size_t)(n + 1), cudaMemcpyDeviceToHost, s));
  if (h_root_value) MZB_CUDA(cudaMemcpyAsync(h_root_value, v.root_value + e0, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  if (h_visits) MZB_CUDA(cudaMemcpyAsync(h_visits, v.visits + e0 * v.A, sizeof(uint16_t) * (size_t)n * v.A, cudaMemcpyDeviceToHost, s));
  MZB_CUDA(cudaStreamSynchronize(s));
  return MZB_OK;
}

}  // extern "C"
