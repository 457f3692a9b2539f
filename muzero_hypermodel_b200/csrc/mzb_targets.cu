// K11: n-step value / reward / policy / action targets for sampled positions, on device.
// Reference: ReplayBuffer.compute_target_value (replay_buffer.py:222-254), make_target (:256-295).
// Game records use the export-ring layout of mzb_env (entry arrays + per-game start/length).
#include "mzb_common.cuh"

namespace {

struct TargetArgs {
  const float* reward; const int8_t* to_play; const double* root_value; const double* reanalysed;
  const uint16_t* visits; const int* action; const int* game_start; const int* game_len;
  const int* b_game; const int* b_index; const uint32_t* b_slot; const uint32_t* b_step;
  const double* discount_pow;      // [td_steps + 1] discount ** i as the host's pow() rounds it
  int B, K, td, A; RngKey key;
  double* t_value; double* t_reward; double* t_policy; int* t_action;
};

__global__ void k_make_target(TargetArgs p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.B * (p.K + 1)) return;
  const int b = (int)(i / (p.K + 1)), row = (int)(i % (p.K + 1));
  const int gm = p.b_game[b];
  const int s = p.game_start[gm], n = p.game_len[gm];       // n = len(root_values); histories hold n + 1 entries
  const int cur = p.b_index[b] + row;
  // compute_target_value(game_history, cur)
  double value = 0.0;
  const int boot = cur + p.td;
  if (boot < n) {
    const double rv = p.reanalysed ? p.reanalysed[s + boot] : p.root_value[s + boot];
    const double last = p.to_play[s + boot] == p.to_play[s + cur] ? rv : -rv;
    value = __dmul_rn(last, p.discount_pow[p.td]);
  }
  // reward_history[cur + 1 : boot + 1]; sign decided by to_play_history[cur + k] (:246-250)
  for (int k = 0; cur + 1 + k <= boot && cur + 1 + k <= n; ++k) {
    const double r = (double)p.reward[s + cur + 1 + k];
    const double signed_r = p.to_play[s + cur] == p.to_play[s + cur + k] ? r : -r;
    value = __dadd_rn(value, __dmul_rn(signed_r, p.discount_pow[k]));
  }
  double* pol = p.t_policy + i * p.A;
  if (cur < n) {
    p.t_value[i] = value;
    p.t_reward[i] = (double)p.reward[s + cur];
    const uint16_t* v = p.visits + (size_t)(s + cur) * p.A;
    int tot = 0;
    for (int a = 0; a < p.A; ++a) tot += v[a];
    for (int a = 0; a < p.A; ++a) pol[a] = v[a] ? __ddiv_rn((double)v[a], (double)tot) : 0.0;
    p.t_action[i] = p.action[s + cur];
  } else {
    const double uni = __ddiv_rn(1.0, (double)p.A);
    for (int a = 0; a < p.A; ++a) pol[a] = uni;
    p.t_value[i] = 0.0;
    if (cur == n) {
      p.t_reward[i] = (double)p.reward[s + cur];
      p.t_action[i] = p.action[s + cur];
    } else {                                             // absorbing state past the end: random action (:291)
      p.t_reward[i] = 0.0;
      // the k-th past-the-end row of this position draws pad index k
      const int k = cur - n - 1;
      const uint32_t slot = p.b_slot ? p.b_slot[b] : (uint32_t)b, step = p.b_step ? p.b_step[b] : 0u;
      p.t_action[i] = (int)__umulhi(rng_draw(p.key, slot, step, MZB_STREAM_PAD, 0, (uint32_t)k).x, (uint32_t)p.A);
    }
  }
}

}  // namespace

extern "C" int mzb_make_target(const float* d_reward, const int8_t* d_to_play, const double* d_root_value,
                               const double* d_reanalysed_root_value, const uint16_t* d_visits, const int32_t* d_action,
                               const int32_t* d_game_start, const int32_t* d_game_len, int32_t n_actions,
                               const int32_t* d_batch_game, const int32_t* d_batch_index, const uint32_t* d_batch_slot,
                               const uint32_t* d_batch_step, int32_t batch, int32_t num_unroll_steps, int32_t td_steps,
                               const double* d_discount_pow, uint64_t seed, double* d_target_value,
                               double* d_target_reward, double* d_target_policy, int32_t* d_actions, void* stream) {
  MZB_CHECK_ARG(d_reward && d_to_play && d_root_value && d_visits && d_action && d_game_start && d_game_len, "NULL game array");
  MZB_CHECK_ARG(d_batch_game && d_batch_index && d_discount_pow, "NULL batch array");
  MZB_CHECK_ARG(d_target_value && d_target_reward && d_target_policy && d_actions, "NULL output");
  MZB_CHECK_ARG(batch > 0 && num_unroll_steps >= 0 && td_steps > 0 && n_actions > 0, "bad size");
  TargetArgs p{d_reward, d_to_play, d_root_value, d_reanalysed_root_value, d_visits, d_action, d_game_start, d_game_len,
               d_batch_game, d_batch_index, d_batch_slot, d_batch_step, d_discount_pow, batch, num_unroll_steps, td_steps,
               n_actions, rng_key(seed), d_target_value, d_target_reward, d_target_policy, d_actions};
  const long long n = (long long)batch * (num_unroll_steps + 1);
  k_make_target<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}
