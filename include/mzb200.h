/* mzb200.h - C ABI of libmzb200.so: the B200 (sm_100a) self-play hot path of muzero-hypermodel.
 *
 * The reference has no FFI: its hot path is a set of duck-typed Python call sites
 * (SURVEY.md §8b).  Each entry point below names the reference call site it replaces
 * (file:line under /root/reference).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - every function returns int: 0 = MZB_OK, <0 = MZB_E*; message via mzb_last_error() (thread-local).
 *  - no C++ exceptions, no torch types; plain pointers and sizes.
 *  - pointers named d_* are DEVICE pointers owned by the caller (e.g. torch tensors' data_ptr());
 *    h_* are HOST pointers.  Handles own no data buffers except where stated.
 *  - every launch takes the cudaStream_t (as void*) to run on and is asynchronous; no hidden syncs
 *    unless the function name ends in _sync or it takes h_* output pointers.
 *  - handles are not thread-safe (one owner thread), one handle set per GPU.
 */
#ifndef MZB200_H
#define MZB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MZB_VERSION 100           /* 0.1.0 */

#define MZB_OK 0
#define MZB_EINVAL (-1)           /* bad argument (shape, range, NULL)                      */
#define MZB_ECUDA (-2)            /* CUDA runtime error, text in mzb_last_error()           */
#define MZB_EUNSUPPORTED (-3)     /* e.g. more than two players (self_play.py:431)          */
#define MZB_ESTATE (-4)           /* call out of order (e.g. expand before select)          */

int mzb_version(void);
const char* mzb_last_error(void);
/* Number of this library's kernels launched since load / last reset (bench.py "gpu_launches"). */
uint64_t mzb_launch_count(void);
void mzb_reset_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Counter-based RNG (Philox4x32-10), see oracle/rng.py.  Replaces numpy's global generator
 * (self_play.py:22-23) with a pure function of (seed, slot, step, stream, sim, index).
 * ------------------------------------------------------------------------------------------- */
#define MZB_STREAM_TIE 0
#define MZB_STREAM_NOISE 1
#define MZB_STREAM_ACTION 2
#define MZB_STREAM_RESET 3
#define MZB_STREAM_PAD 4
#define MZB_STREAM_RGAME 5        /* replay: game draw of batch element (slot) in batch (step)     */
#define MZB_STREAM_RPOS 6         /* replay: position draw                                         */
/* Host-side evaluation (tests): out[4]. */
void mzb_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out);

/* ---------------------------------------------------------------------------------------------
 * Tree store: structure-of-arrays MCTS trees for G independent games in HBM.
 * Replaces the Python Node graph + MinMaxStats (self_play.py:434-477, 551-568).
 * Node k (k>=1) is the node created by simulation k; node 0 is the root.
 * ------------------------------------------------------------------------------------------- */
typedef struct mzb_tree mzb_tree;

typedef struct {
  int32_t n_games;          /* G                                                             */
  int32_t n_actions;        /* A = len(config.action_space), 1..65535                        */
  int32_t num_simulations;  /* config.num_simulations, 1..65534                              */
  int32_t n_players;        /* len(config.players): 1 or 2                                   */
  double discount;          /* config.discount                                               */
  double pb_c_base;         /* config.pb_c_base                                              */
  double pb_c_init;         /* config.pb_c_init                                              */
  int32_t hidden_floats;    /* fp32 elements of one hidden-state slot (0: no slots)          */
  int32_t reserved;
  uint64_t seed;            /* Philox key                                                    */
} mzb_tree_config;

/* Bytes of device workspace a tree of this config needs. */
size_t mzb_tree_workspace_bytes(const mzb_tree_config* cfg);
/* d_workspace: caller-owned device memory (>= workspace_bytes, 256-byte aligned), kept alive until destroy.
 * h_log_lut: optional host array [num_simulations+1] with log((N+base+1)/base)+init per parent visit
 * count N as the caller's math.log computes it (self_play.py:385-390); NULL = computed with libm log(). */
int mzb_tree_create(mzb_tree** out, const mzb_tree_config* cfg, void* d_workspace, size_t workspace_bytes,
                    const double* h_log_lut);
int mzb_tree_destroy(mzb_tree* t);
/* Device pointer to the hidden-state slots, fp32, blocked by 32 games and node-major inside a block:
 * [ceil(G/32)][num_simulations+1][32][hidden_floats]; game g, slot n lives at
 * ((g >> 5) * (num_simulations+1) * 32 + n * 32 + (g & 31)) * hidden_floats. */
float* mzb_tree_hidden_ptr(mzb_tree* t);

/* Root expansion (+ exploration noise).  Replaces Node.expand at the root and
 * Node.add_exploration_noise (self_play.py:303-314, 452-477) and resets MinMaxStats (:317).
 *  d_reward   [G] f32  support_to_scalar(reward logits) of initial_inference
 *  d_policy   [G,A] f32 policy logits (policy_is_logits=1: softmax over LEGAL actions computed here)
 *             or priors over the legal actions (0: taken as given, float32 widened like .tolist())
 *  d_legal    [G,A] u8 or NULL (all legal).  The caller asserts non-empty legal sets (self_play.py:297-302).
 *  d_to_play  [G] i8 or NULL (0)
 *  d_noise    [G,A] f64 Dirichlet sample laid out by ACTION (entries of illegal actions ignored) or NULL;
 *             NULL with frac>0 -> generated on device: Gamma(alpha) by Marsaglia-Tsang from Philox.
 *  frac       root_exploration_fraction, 0 disables noise (add_exploration_noise=False)
 *  d_slot,d_step [G] u32 RNG counters c0,c1 per game or NULL (slot=g, step=0)                      */
int mzb_tree_root_init(mzb_tree* t, const float* d_reward, const float* d_policy, int policy_is_logits,
                       const uint8_t* d_legal, const int8_t* d_to_play, const double* d_noise, double alpha,
                       double frac, const uint32_t* d_slot, const uint32_t* d_step, void* stream);

/* One select walk per game: pUCT argmax from the root to an unexpanded edge.
 * Replaces the `while node.expanded()` loop, select_child and ucb_score (self_play.py:326-335, 364-405).
 * Outputs (any may be NULL): d_parent_slot [G] i32 node whose hidden state feeds recurrent_inference,
 * d_action [G] i32 last action, d_depth [G] i32 search-path length (= current_tree_depth). */
int mzb_tree_select(mzb_tree* t, int32_t* d_parent_slot, int32_t* d_action, int32_t* d_depth, void* stream);

/* Expand the selected leaf with the network outputs and back the value up the search path.
 * Replaces node.expand + backpropagate (self_play.py:346-354, 407-431).
 *  d_value, d_reward [G] f32 (support_to_scalar outputs), d_policy [G,A] f32 logits or priors. */
int mzb_tree_expand_backup(mzb_tree* t, const float* d_value, const float* d_reward, const float* d_policy,
                           int policy_is_logits, void* stream);

/* Root statistics after (or during) a search; any output may be NULL.
 *  d_visits [G,A] i32 (0 for illegal), d_root_value [G] f64 = value_sum/visit_count (Node.value),
 *  d_max_depth [G] i32, d_child_value_sum [G,A] f64, d_child_reward [G,A] f32, d_child_prior [G,A] f64,
 *  d_minmax [G,2] f64. Feeds select_action / store_search_statistics (self_play.py:223-246, 497-512). */
int mzb_tree_root_stats(mzb_tree* t, int32_t* d_visits, double* d_root_value, int32_t* d_max_depth,
                        double* d_child_value_sum, float* d_child_reward, double* d_child_prior,
                        double* d_minmax, void* stream);

/* h_counters2 = {sum of search-path lengths (nodes below the root), simulations} since creation / last reset. */
int mzb_tree_counters_sync(mzb_tree* t, uint64_t* h_counters2, int reset, void* stream);

/* Test hook: d_fast[i] = a[i] / b[i] through the whole-search kernel's hoisted-reciprocal float64 division
 * (csrc/mzb_common.cuh ddiv_rcp: the compiler's own div.rn.f64 sequence with the refinement of 1/b computed once per
 * divisor), d_ref[i] = the compiler's division.  Python's float division (self_play.py:393-404, 563-568) is the
 * correctly rounded quotient; the two must agree bit for bit. */
int mzb_debug_ddiv_rcp(const double* d_a, const double* d_b, int64_t n, double* d_fast, double* d_ref, void* stream);

/* Copy one game's complete tree to the host (synchronises `stream`): used to materialise the
 * reference's Node graph for callers that walk it (diagnose_model.py:161-252).
 * Arrays are [num_simulations+1][A] unless noted; h_root_prior [A] f64; h_scalars[4] =
 * {root_visit, root_value_sum, root_reward, nodes_used}. */
int mzb_tree_export_game_sync(mzb_tree* t, int32_t game, double* h_value_sum, float* h_prior, int32_t* h_visit,
                              float* h_reward, int32_t* h_child, double* h_root_prior, double* h_scalars,
                              void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fully-connected MuZero network (models.py:80-195): cartpole, tictactoe-FC.
 * ------------------------------------------------------------------------------------------- */
typedef struct mzb_fc_model mzb_fc_model;

typedef struct {
  int32_t obs_dim;          /* C*H*W*(stacked+1) + stacked*H*W  (models.py:100-107)              */
  int32_t encoding_size;    /* config.encoding_size                                              */
  int32_t n_actions;        /* len(config.action_space)                                          */
  int32_t support_size;     /* config.support_size; logits are 2*support_size+1 wide             */
  int32_t n_rep; int32_t rep[3];   /* config.fc_representation_layers (hidden widths, <= 3)     */
  int32_t n_dyn; int32_t dyn[3];   /* config.fc_dynamics_layers                                  */
  int32_t n_rew; int32_t rew[3];   /* config.fc_reward_layers                                    */
  int32_t n_val; int32_t val[3];   /* config.fc_value_layers                                     */
  int32_t n_pol; int32_t pol[3];   /* config.fc_policy_layers                                    */
} mzb_fc_config;

int mzb_fc_create(mzb_fc_model** out, const mzb_fc_config* cfg);
int mzb_fc_destroy(mzb_fc_model* m);
int mzb_fc_num_tensors(const mzb_fc_model* m);
/* Replaces AbstractNetwork.set_weights (models.py:72-73).  h_tensors: HOST float32 pointers in
 * state_dict() order (representation, dynamics_encoded_state, dynamics_reward, prediction_policy,
 * prediction_value; per Linear: weight [out][in] row-major, then bias [out]).  The weights are copied
 * into a library-owned packed device blob; the call synchronises `stream`. */
int mzb_fc_set_weights(mzb_fc_model* m, const float* const* h_tensors, int n_tensors, void* stream);

/* initial_inference (models.py:172-190) for B rows, fused with support_to_scalar (self_play.py:293-296)
 * and the root softmax of Node.expand (:459-461).  Outputs may be NULL.
 *  d_obs [B, obs_dim] f32; d_legal [B,A] u8 or NULL (softmax over all actions)
 *  d_state_out: row r is written at d_state_out + r*out_row_stride + out_offset (floats), encoding_size wide
 *  d_value_logits/d_reward_logits [B, 2S+1], d_policy_logits [B,A]  (the reference's return values)
 *  d_value/d_reward [B] scalars, d_priors [B,A] (0 for illegal) */
int mzb_fc_initial(mzb_fc_model* m, int64_t B, const float* d_obs, const uint8_t* d_legal, float* d_state_out,
                   int64_t out_row_stride, int64_t out_offset, float* d_value_logits, float* d_reward_logits,
                   float* d_policy_logits, float* d_value, float* d_reward, float* d_priors, void* stream);

/* recurrent_inference (models.py:192-195) for B rows.  Row r reads its state at
 * d_state_in + r*in_row_stride + (d_in_slot ? d_in_slot[r] : 0)*slot_stride, so the hidden-state slots of a
 * tree store can be consumed in place (d_in_slot = parent slots from mzb_tree_select). */
int mzb_fc_recurrent(mzb_fc_model* m, int64_t B, const float* d_state_in, int64_t in_row_stride,
                     const int32_t* d_in_slot, int64_t slot_stride, const int32_t* d_action, float* d_state_out,
                     int64_t out_row_stride, int64_t out_offset, float* d_value_logits, float* d_reward_logits,
                     float* d_policy_logits, float* d_value, float* d_reward, float* d_priors, void* stream);

/* uint8 image observations -> the float32 frames the reference's Atari wrapper builds on the host
 * (games/breakout.py:141-159: `numpy.asarray(frame, dtype="float32") / 255.0`): out[i] = (float)in[i] / 255.0f, one
 * correctly rounded float32 division per element, so a caller that ships the emulator's uint8 frames (4x fewer bytes
 * over PCIe) feeds the network the same values bit for bit. */
int mzb_u8_to_unit_float(const uint8_t* d_in, int64_t n, float* d_out, void* stream);

/* models.support_to_scalar (models.py:641-662): d_logits [B, 2S+1] -> d_out [B]. */
int mzb_support_to_scalar(const float* d_logits, int64_t B, int support_size, float* d_out, void* stream);
/* models.scalar_to_support (models.py:665-685): d_x [n] -> d_out [n, 2S+1]. */
int mzb_scalar_to_support(const float* d_x, int64_t n, int support_size, float* d_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Batched MCTS.run for fully-connected networks (self_play.py:261-362): G searches in one call.
 *  d_obs [G, obs_dim] f32 (the stacked observation MCTS.run receives), d_legal [G,A] u8 | NULL,
 *  d_to_play [G] i8 | NULL; noise/alpha/frac/slot/step as in mzb_tree_root_init (frac=0 <=>
 *  add_exploration_noise=False).  num_simulations <= the tree's capacity.
 *  allow_fused=1: shapes with a compiled whole-search kernel (cartpole, tictactoe-FC; see
 *  mzb_search_fc_is_fused) run as ONE launch, one thread per game; otherwise - and always with
 *  allow_fused=0 - the modular kernels run: fc_initial, root_init, then select / fc_recurrent /
 *  expand_backup per simulation.  Both paths leave the identical tree in `t` (bit-for-bit), so
 *  mzb_tree_root_stats / mzb_tree_export_game_sync work after either.
 *  Outputs (NULL to skip): d_visits [G,A] i32, d_root_value [G] f64 (root.value()),
 *  d_root_predicted_value [G] f32 (extra_info["root_predicted_value"]), d_max_depth [G] i32.      */
int mzb_search_fc(mzb_tree* t, mzb_fc_model* m, const float* d_obs, const uint8_t* d_legal, const int8_t* d_to_play,
                  const double* d_noise, double alpha, double frac, const uint32_t* d_slot, const uint32_t* d_step,
                  int32_t num_simulations, int allow_fused, int32_t* d_visits, double* d_root_value,
                  float* d_root_predicted_value, int32_t* d_max_depth, void* stream);
int mzb_search_fc_is_fused(const mzb_fc_model* m);

/* ---------------------------------------------------------------------------------------------
 * Vectorised environments + the play_game bookkeeping around each search.
 * Replaces Game.step/legal_actions/to_play/reset (games/*.py), SelfPlay.select_action
 * (self_play.py:223-246), GameHistory appends (:176-182) and the hand-over of finished games
 * (:52) for G games per GPU.  Game g has the global id first_slot + g (RNG counter c0).
 * ------------------------------------------------------------------------------------------- */
#define MZB_ENV_CARTPOLE 0   /* games/cartpole.py (gym CartPole-v1 physics)  */
#define MZB_ENV_TICTACTOE 1  /* games/tictactoe.py                           */
#define MZB_ENV_CONNECT4 2   /* games/connect4.py                            */
#define MZB_ENV_GOMOKU 3     /* games/gomoku.py                              */
#define MZB_ENV_SYNTHETIC_FRAMES 4 /* breakout stand-in (BASELINE.json): 3x96x96 frames U[0,1) from Philox, A = 4,
                                      reward 0, never done before max_moves (the ALE emulator is out of scope) */

typedef struct mzb_env mzb_env;
typedef struct {
  int32_t kind;             /* MZB_ENV_*                                                        */
  int32_t n_games;          /* G                                                                */
  int32_t max_moves;        /* config.max_moves                                                 */
  int32_t export_entries;   /* capacity of the finished-game ring, in history entries (moves+1) */
  int32_t export_games;     /* capacity of the ring's game index                                */
  uint32_t first_slot;      /* global id of game 0 (= rank * G when games are sharded)          */
  uint64_t seed;            /* Philox key (config.seed)                                         */
} mzb_env_config;

size_t mzb_env_workspace_bytes(const mzb_env_config* cfg);
/* Creates the environments in caller-owned device memory and resets every game (Game.reset). */
int mzb_env_create(mzb_env** out, const mzb_env_config* cfg, void* d_workspace, size_t workspace_bytes, void* stream);
int mzb_env_destroy(mzb_env* e);
int mzb_env_info(const mzb_env* e, int32_t* n_actions, int32_t* obs_dim, int32_t* rec_floats);
int mzb_env_reset(mzb_env* e, void* stream);
/* What play_game passes to MCTS.run (self_play.py:138-150): d_obs [G, obs_dim] f32 (C*H*W planes,
 * to-play plane +1/-1), d_legal [G,A] u8, d_to_play [G] i8, plus the RNG counters of this move:
 * d_slot [G] u32 (global game id), d_step [G] u32 (env steps the slot has taken).  NULL = skip. */
int mzb_env_observe(mzb_env* e, float* d_obs, uint8_t* d_legal, int8_t* d_to_play, uint32_t* d_slot, uint32_t* d_step,
                    void* stream);
/* GameHistory.get_stacked_observations(-1, stacked_observations) (self_play.py:514-548) of every running game, built
 * from the history the environment keeps on the device: d_obs [G][(C (S + 1) + S) * H * W] float32 - the current
 * observation, then for k = 1..S the observation k moves ago and a plane filled with the action played from it (zeros
 * before the start of the game).  S = 0 is mzb_env_observe's observation.  Not available for synthetic frames. */
int mzb_env_observe_stacked(mzb_env* env, int32_t stacked_observations, float* d_obs, void* stream);

/* One move for every running game: select_action from the root visit counts at `temperature`
 * (0 past temperature_threshold, <=0 disables the threshold), Game.step, GameHistory appends.
 *  d_uniforms [G] f64 injects the random draw (NULL = Philox(slot, step, STREAM_ACTION));
 *  d_forced_action [G] i32 overrides the choice (parity tests / opponents).
 *  Outputs (NULL = skip): d_action [G] i32, d_reward [G] f32, d_done [G] u8 (episode finished). */
int mzb_env_act_step(mzb_env* e, const int32_t* d_visits, const double* d_root_value, const uint8_t* d_legal,
                     double temperature, int32_t temperature_threshold, const double* d_uniforms,
                     const int32_t* d_forced_action, int32_t* d_action, float* d_reward, uint8_t* d_done, void* stream);
/* Finished games are appended to the export ring (do_export=1) and restarted (auto-reset). */
int mzb_env_harvest(mzb_env* e, int do_export, void* stream);
/* h_counters4: {games finished, moves of finished games, env steps, games dropped by a full ring}. */
int mzb_env_counters_sync(mzb_env* e, uint64_t* h_counters4, void* stream);
int mzb_env_state_ptrs(mzb_env* e, double** d_cartpole_state, int8_t** d_board, int8_t** d_player, int32_t** d_hist_len,
                       uint8_t** d_finished);
/* Copy the ring to host arrays sized for the configured capacities and empty it.  Entry arrays:
 * h_obs [E, rec_floats] f32 (cartpole: 4 floats; boards: cells+1 int8 packed), h_action/h_reward/h_to_play [E],
 * h_visits [E, A] u16 (root child visit counts; child_visits = count / sum), h_root_value [E] f64;
 * game g occupies entries [h_game_start[g], h_game_start[g] + h_game_len[g]] (len moves, len+1 entries). */
int mzb_env_export_drain_sync(mzb_env* e, int32_t* h_n_entries, int32_t* h_n_games, float* h_obs, int32_t* h_action,
                              float* h_reward, int8_t* h_to_play, uint16_t* h_visits, double* h_root_value,
                              int32_t* h_game_start, int32_t* h_game_len, uint32_t* h_game_slot, void* stream);

/* ---------------------------------------------------------------------------------------------
 * n-step targets (ReplayBuffer.make_target / compute_target_value, replay_buffer.py:222-295) for a
 * batch of (game, position) pairs.  Games are given in the export layout of mzb_env: entry arrays
 * d_reward/d_to_play/d_action [E], d_root_value [E] f64, d_visits [E,A] u16, game g occupying
 * entries d_game_start[g] .. d_game_start[g] + d_game_len[g] (len = number of moves).
 * d_reanalysed_root_value [E] f64 or NULL (GameHistory.reanalysed_predicted_root_values).
 * d_discount_pow [td_steps+1] f64 = discount ** i as the caller's pow rounds it.
 * d_batch_slot/d_batch_step [B] u32 (or NULL): counters of the random padding action (:291).
 * Outputs, one row per unroll step: d_target_value, d_target_reward [B, K+1] f64,
 * d_target_policy [B, K+1, A] f64, d_actions [B, K+1] i32. */
int mzb_make_target(const float* d_reward, const int8_t* d_to_play, const double* d_root_value,
                    const double* d_reanalysed_root_value, const uint16_t* d_visits, const int32_t* d_action,
                    const int32_t* d_game_start, const int32_t* d_game_len, int32_t n_actions,
                    const int32_t* d_batch_game, const int32_t* d_batch_index, const uint32_t* d_batch_slot,
                    const uint32_t* d_batch_step, int32_t batch, int32_t num_unroll_steps, int32_t td_steps,
                    const double* d_discount_pow, uint64_t seed, double* d_target_value, double* d_target_reward,
                    double* d_target_policy, int32_t* d_actions, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Residual MuZero network (models.py:206-619): tictactoe (default), connect4, gomoku, breakout.
 * ------------------------------------------------------------------------------------------- */
typedef struct mzb_resnet_model mzb_resnet_model;

typedef struct {
  int32_t obs_channels;              /* observation_shape[0]*(stacked+1)+stacked                        */
  int32_t height, width;             /* observation_shape[1:], latent = same, or ceil(/16) when downsampling */
  int32_t n_actions;
  int32_t blocks, channels;          /* config.blocks, config.channels                                  */
  int32_t reduced_channels_reward, reduced_channels_value, reduced_channels_policy;
  int32_t n_fc_reward; int32_t fc_reward[3];     /* config.resnet_fc_reward_layers                       */
  int32_t n_fc_value; int32_t fc_value[3];       /* config.resnet_fc_value_layers                        */
  int32_t n_fc_policy; int32_t fc_policy[3];     /* config.resnet_fc_policy_layers                       */
  int32_t support_size;
  int32_t downsample;                /* 0: False, 1: "resnet" (models.py:233-275); "CNN" unsupported     */
  int32_t precision;                 /* 0: fp32 CUDA-core path, 1: bf16 tcgen05 path (fp32 accumulate)   */
} mzb_resnet_config;

int mzb_resnet_create(mzb_resnet_model** out, const mzb_resnet_config* cfg);
/* Test hook: 0 makes the bf16 path run the CUDA-core direct convolution instead of the tcgen05 kernel. */
void mzb_conv_tc_enable(int on);
/* Which 64 -> 64 channel layers with resident weights run as CTA pairs (tcgen05.mma cta_group::2, each CTA holding half of
 * the weight rows) on batches that give every CTA of the grid at least two tiles: 0 none, 1 plain layers (no residual /
 * action plane / head projection; the default), 2 every eligible layer, -1 back to the MZB_TC_PAIR environment value.
 * The forms are bit-identical. */
void mzb_conv_tc_pair_enable(int mode);
/* Test hook: 0 makes narrow networks (16 channels, <= 48 latent positions: Breakout) run recurrent inference layer by
 * layer instead of the one-kernel warp-per-image path (csrc/mzb_tower16.cu, models.py:363-404, 447-456, 551-595). */
void mzb_tower16_enable(int on);
/* Test hook: 0 makes the bf16 DownSample stem (models.py:226-275) run its residual blocks layer by layer on the tcgen05
 * convolution instead of one launch per resolution with the image resident in shared memory (csrc/mzb_stem16.cu). */
void mzb_stem16_enable(int on);
int mzb_resnet_destroy(mzb_resnet_model* m);
int mzb_resnet_num_tensors(const mzb_resnet_model* m);
int mzb_resnet_latent_dims(const mzb_resnet_model* m, int32_t* C, int32_t* H, int32_t* W);
/* set_weights (models.py:72-73): HOST fp32 tensors in state_dict() order with the integer
 * `num_batches_tracked` entries skipped; h_numel gives each tensor's element count (validated).
 * Eval-mode batch-norm is folded to per-channel scale/shift here.  Synchronises the device. */
int mzb_resnet_set_weights(mzb_resnet_model* m, const float* const* h_tensors, const int64_t* h_numel, int n_tensors,
                           void* stream);
size_t mzb_resnet_workspace_bytes(const mzb_resnet_model* m, int64_t max_batch);
/* Must be called once on a freshly allocated workspace (zeroes it: the padded activation layout of the
 * tensor-core path relies on zero pad rows that no kernel ever writes). */
int mzb_resnet_workspace_init(const mzb_resnet_model* m, void* d_workspace, size_t workspace_bytes, void* stream);
/* State layouts: 0 = NCHW fp32 (the reference's tensors), 1 = NHWC fp32, 2 = NHWC bf16 (internal pools;
 * must match the model precision).  Row r of a state lives at base + r*row_stride (+ slot*slot_stride on
 * input, + offset on output), in elements of the layout's type.
 * initial_inference (models.py:597-614): d_obs [B, obs_channels, H, W] fp32; outputs as mzb_fc_initial. */
int mzb_resnet_initial(mzb_resnet_model* m, int64_t B, const float* d_obs, const uint8_t* d_legal, void* d_workspace,
                       size_t workspace_bytes, void* d_state_out, int state_layout, int64_t out_row_stride,
                       int64_t out_offset, float* d_value_logits, float* d_reward_logits, float* d_policy_logits,
                       float* d_value, float* d_reward, float* d_priors, void* stream);
/* recurrent_inference (models.py:616-619). */
int mzb_resnet_recurrent(mzb_resnet_model* m, int64_t B, const void* d_state_in, int in_layout, int64_t in_row_stride,
                         const int32_t* d_in_slot, int64_t slot_stride, const int32_t* d_action, void* d_workspace,
                         size_t workspace_bytes, void* d_state_out, int out_layout, int64_t out_row_stride,
                         int64_t out_offset, float* d_value_logits, float* d_reward_logits, float* d_policy_logits,
                         float* d_value, float* d_reward, float* d_priors, void* stream);

/* Measurement hook for the roofline line of bench.py: launches the first residual-block convolution of the
 * representation tower (C -> C, latent resolution, batch B; conv3x3 + BatchNorm2d + relu, models.py:206-229)
 * `iters` times on the workspace's activation buffers, so the dominant kernel can be timed alone. */
int mzb_resnet_conv_probe(mzb_resnet_model* m, int64_t B, void* d_workspace, size_t workspace_bytes, int32_t iters,
                          void* stream);

/* Batched MCTS.run for residual networks (self_play.py:261-362): as mzb_search_fc, with the tree kernels and
 * the resnet layer program launched per simulation.  d_hidden_pool: caller-owned hidden-state slots
 * [G][tree capacity + 1][H*W*C] dense NHWC, bf16 (precision 1) or fp32 (precision 0);
 * d_workspace as for mzb_resnet_recurrent with batch G.
 * From the second call with an identical argument set (same handles and buffers - the self-play loop) the launch
 * sequence of the whole search is captured once into a CUDA graph and replayed; everything that varies from move to
 * move is read through the device pointers.  MZB_NO_GRAPH=1 in the environment disables this. */
int mzb_search_resnet(mzb_tree* t, mzb_resnet_model* m, const float* d_obs, const uint8_t* d_legal,
                      const int8_t* d_to_play, const double* d_noise, double alpha, double frac, const uint32_t* d_slot,
                      const uint32_t* d_step, int32_t num_simulations, void* d_hidden_pool, void* d_workspace,
                      size_t workspace_bytes, int32_t* d_visits, double* d_root_value, float* d_root_predicted_value,
                      int32_t* d_max_depth, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Replay store on the device: replaces ReplayBuffer (replay_buffer.py:11-220), the consumer
 * of the self-play output and the producer of the trainer's batches.
 * Games live in fixed slots (game id % capacity_games, FIFO eviction like the reference's dict :57-60) in the export
 * layout of mzb_env / mzb_make_target; a game of n moves has n + 1 entries and must fit entry_stride.
 * All float32 / float64 arithmetic of the reference (initial priorities :37-50, sampling probabilities :162-192,
 * importance weights :117-121) is restated operation by operation; random draws are one float64 uniform per batch
 * element and draw, injected (d_u_*) or Philox(seed; slot = element, step = batch counter; MZB_STREAM_RGAME/RPOS).
 * ------------------------------------------------------------------------------------------- */
typedef struct mzb_replay mzb_replay;
#define MZB_REPLAY_MAX_SAVE 65536   /* games per mzb_replay_save_games call */
typedef struct {
  int32_t n_actions;         /* len(config.action_space)                                      */
  int32_t obs_floats;        /* floats per stored observation record (stacked_observations = 0) */
  int32_t obs_decode;        /* 0: the record IS the observation; 1: packed board record of mzb_env
                              * (obs_h * obs_w int8 cells + int8 player, padded to obs_floats floats), decoded by
                              * get_batch into the 3 planes [own, other, to-play] of the board games        */
  int32_t obs_h, obs_w;      /* board size for obs_decode = 1                                 */
  int32_t capacity_games;    /* config.replay_buffer_size                                     */
  int32_t entry_stride;      /* entries per slot >= max_moves + 1                             */
  int32_t num_unroll_steps;  /* config.num_unroll_steps                                       */
  int32_t td_steps;          /* config.td_steps                                               */
  int32_t per;               /* config.PER                                                    */
  int32_t max_batch;         /* largest batch get_batch / update_priorities will see          */
  int32_t stacked_observations; /* config.stacked_observations: get_batch / game_observations return
                              * GameHistory.get_stacked_observations(position, S) (self_play.py:514-548)          */
  int32_t obs_channels;      /* C of the observation (planes of H*W); used for the action planes when S > 0  */
  double per_alpha;          /* config.PER_alpha                                              */
  uint64_t seed;             /* Philox key                                                    */
} mzb_replay_config;

size_t mzb_replay_workspace_bytes(const mzb_replay_config* cfg);
/* h_discount_pow [td_steps + 1] = discount ** i as the caller's pow rounds it (replay_buffer.py:240, 253).
 * d_workspace: caller-owned, 256-byte aligned, >= mzb_replay_workspace_bytes. */
int mzb_replay_create(mzb_replay** out, const mzb_replay_config* cfg, void* d_workspace, size_t workspace_bytes,
                      const double* h_discount_pow, void* stream);
int mzb_replay_destroy(mzb_replay* r);
/* save_game (replay_buffer.py:33-64) for n_games games given as entry arrays on the device (export layout: game g
 * occupies source entries h_src_start[g] .. h_src_start[g] + h_len[g]); appended in order, the oldest games evicted
 * beyond capacity.  PER: initial priorities |root_value - target_value| ** PER_alpha (float64 -> float32) and the
 * game priority, or d_priorities [entries] f32 if the histories already carry them (:35-37).  Synchronises the
 * stream once (staging of the per-game table). */
int mzb_replay_save_games(mzb_replay* r, int32_t n_games, const int32_t* h_src_start, const int32_t* h_len,
                          const float* d_obs, const int32_t* d_action, const float* d_reward, const int8_t* d_to_play,
                          const double* d_root_value, const uint16_t* d_visits, const float* d_priorities, void* stream);
/* get_batch (replay_buffer.py:69-140): sample_n_games + sample_position per element, then make_target, the
 * observation of the position, gradient scale and (PER) importance weights normalised by their maximum.
 * d_u_game / d_u_pos [batch] f64 inject the draws (NULL = Philox).  Outputs: d_game_id [B] i64 + d_pos [B] i32
 * (index_batch), optional d_game_prob / d_pos_prob [B] f32, d_obs [B, obs_floats] f32, d_actions [B, K+1] i32,
 * d_values / d_rewards [B, K+1] f64, d_policies [B, K+1, A] f64, d_weights [B] f32 (PER), d_gradient_scale
 * [B, K+1] i32.  Output groups may be NULL.  Increments the batch counter. */
int mzb_replay_get_batch(mzb_replay* r, int32_t batch, const double* d_u_game, const double* d_u_pos, int64_t* d_game_id,
                         int32_t* d_pos, float* d_game_prob, float* d_pos_prob, float* d_obs, int32_t* d_actions,
                         double* d_values, double* d_rewards, double* d_policies, float* d_weights,
                         int32_t* d_gradient_scale, void* stream);
/* update_priorities (replay_buffer.py:202-220): d_priorities [batch, K+1] f32 written to positions
 * [pos, min(pos + K + 1, len)) of each still-buffered game, rows applied in order; game priority = max. */
int mzb_replay_update_priorities(mzb_replay* r, int32_t batch, const float* d_priorities, const int64_t* d_game_id,
                                 const int32_t* d_pos, void* stream);
/* The device-to-device hop that replaces `replay_buffer.save_game.remote(game_history)` (self_play.py:52): every
 * finished game in the export ring of `env` is appended to the store and the ring is emptied; only the per-game
 * (start, length) table crosses to the host.  The store must have been created with the environment's record
 * format (obs_floats = mzb_env_info's rec_floats, obs_decode = 1 with the board size for the board games, the same
 * n_actions, entry_stride >= max_moves + 2): anything else returns MZB_EINVAL and leaves ring and store untouched. */
int mzb_env_export_to_replay(mzb_env* env, mzb_replay* r, int32_t* h_n_games, void* stream);
/* The configuration the store was created with (layout checks of callers that copy into it). */
int mzb_replay_get_config(const mzb_replay* r, mzb_replay_config* out);
/* out5 = total_samples, num_played_games, num_played_steps, games in the buffer, id of the oldest game. */
int mzb_replay_info(const mzb_replay* r, int64_t* out5);
int mzb_replay_set_batch_counter(mzb_replay* r, uint32_t counter);
/* Reanalyse (replay_buffer.py:298-361).  mzb_replay_game_observations writes the len observations of a buffered game
 * as network input d_obs [len, observation floats] (board records decoded; d_obs may be NULL to query len);
 * mzb_replay_set_reanalysed stores fresh root-value predictions d_values [len] f32 as the game's
 * `reanalysed_predicted_root_values`: get_batch's n-step targets then bootstrap from them (:229-233, accumulated in
 * float64).  A game evicted in the meantime is ignored (update_game_history :194-200). */
int mzb_replay_game_observations(mzb_replay* r, int64_t game_id, float* d_obs, int32_t* h_len, void* stream);
int mzb_replay_set_reanalysed(mzb_replay* r, int64_t game_id, const float* d_values, void* stream);
/* get_buffer (replay_buffer.py:66-67, persisted as replay_buffer.pkl by muzero.py:315-323): one buffered game back
 * to the host as entry arrays (len + 1 entries; root values and visit counts for the len moves). */
int mzb_replay_export_game_sync(mzb_replay* r, int64_t game_id, int32_t* h_len, float* h_obs, int32_t* h_action,
                                float* h_reward, int8_t* h_to_play, double* h_root_value, uint16_t* h_visits, void* stream);
/* Inspection (tests, checkpoints): priorities [len] and game priority of a buffered game. */
int mzb_replay_game_priorities_sync(mzb_replay* r, int64_t game_id, float* h_priorities, float* h_game_priority,
                                    int32_t* h_len, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training step of the fully-connected network as ONE kernel: unrolled forward, categorical
 * cross-entropies, backward through time, batch reduction into the flat gradient bucket.  Replaces the autograd graph
 * of Trainer.update_weights (trainer.py:124-255; loss_function :267-284) for MuZeroFullyConnectedNetwork
 * (models.py:80-195): gradient hooks (0.5 into the dynamics function, 1 / gradient_scale on the losses of the unrolled
 * steps), value loss weight, PER importance weights, priorities |support_to_scalar(value) - target| ** PER_alpha.
 * The layer table indexes the trainer's flat parameter bucket (model.parameters() order: weight [out][in], bias [out]);
 * networks in the order representation, dynamics, reward, policy, value; layer l of a network with n_layers layers is
 * followed by ELU unless it is the last (models.py:626-638).  Deterministic: no atomics on the gradients. */
typedef struct {
  int32_t obs_dim, encoding_size, n_actions, support_size;
  int32_t n_layers[5];
  int32_t in[5][4], out[5][4];
  int64_t w_off[5][4], b_off[5][4];
} mzb_fc_train_desc;
/* Bytes of device scratch mzb_fc_train_grad needs (zero-initialised ONCE by the caller), or -1 for a bad table. */
int64_t mzb_fc_train_workspace_bytes(const mzb_fc_train_desc* desc, int32_t batch, int32_t unroll_plus_1);
/* 1 when weights + per-sample activations of the unrolled steps fit the kernel's shared memory, else 0. */
int mzb_fc_train_fits(const mzb_fc_train_desc* desc, int32_t batch, int32_t unroll_plus_1);
/* d_obs [B][obs_dim] f32, d_action [B][K+1] i64 (column 0 unused), target supports [B][K+1][2S+1], target policy
 * [B][K+1][A], target value scalars [B][K+1], PER weights [B] or NULL, gradient scale [B][K+1].
 * Outputs: d_grad [n_params] (overwritten), d_losses [3][B] = per-sample value / reward / policy loss sums,
 * d_priorities [B][K+1], d_loss [1] = the batch objective. */
int mzb_fc_train_grad(const mzb_fc_train_desc* desc, const float* d_params, int64_t n_params, int32_t batch, int32_t unroll_plus_1,
                      const float* d_obs, const int64_t* d_action, const float* d_target_value_support,
                      const float* d_target_reward_support, const float* d_target_policy, const float* d_target_value_scalar,
                      const float* d_weight, const float* d_gradient_scale, double value_loss_weight, double per_alpha,
                      float* d_grad, float* d_losses, float* d_priorities, float* d_loss, void* d_workspace,
                      int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimiser step on one flat float32 bucket: the trainer keeps all parameters and all
 * gradients as views into two flat buffers, all-reduces the gradient bucket (NCCL) and applies ONE of these launches.
 * Replaces torch.optim.Adam / SGD as configured in trainer.py:35-52 (L2 weight decay added to the gradient; Adam
 * without amsgrad; SGD with momentum, no dampening / nesterov).  step counts from 1; grad_scale multiplies the
 * gradient first (1 / world size for the data-parallel mean). */
int mzb_adam_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n, double lr,
                  double beta1, double beta2, double eps, double weight_decay, int64_t step, double grad_scale, void* stream);
int mzb_sgd_step(float* d_param, const float* d_grad, float* d_momentum_buffer, int64_t n, double lr, double momentum,
                 double weight_decay, int64_t step, double grad_scale, void* stream);

/* Gradient all-reduce fused into the optimiser step over NVLink peer memory (single node, <= 8 ranks; trainer.py:35-53
 * + the data-parallel mean).  Every rank keeps two flat gradient buckets (step parity) and a flag array of `world`
 * uint32 in mzb_p2p_alloc'd memory, exports them (mzb_p2p_export -> 64-byte IPC handle, exchanged by the caller) and
 * maps its peers' (mzb_p2p_import).  One launch per rank and step: barrier on the flags (release/acquire at system
 * scope, bounded spin), then param[i] is updated from sum_r bucket_r[i] / world, summed in rank order on every rank so
 * the replicas stay bit-identical.  h_peer_grads / h_peer_flags: HOST arrays of `world` device pointers valid on this
 * rank (own buffers at index `rank`); `seq` = the step sequence number (starts at 1, the same on all ranks; selects
 * nothing by itself - the caller passes the buckets of parity seq & 1). */
int mzb_p2p_alloc(void** d_ptr, size_t bytes);
int mzb_p2p_free(void* d_ptr);
int mzb_p2p_export(void* d_ptr, uint8_t* handle64);
int mzb_p2p_import(const uint8_t* handle64, void** d_ptr);
int mzb_p2p_close(void* d_ptr);
int mzb_adam_step_allreduce(float* d_param, const float* const* h_peer_grads, uint32_t* const* h_peer_flags, int32_t rank,
                            int32_t world, uint32_t seq, float* d_exp_avg, float* d_exp_avg_sq, int64_t n, double lr,
                            double beta1, double beta2, double eps, double weight_decay, int64_t step, void* stream);
int mzb_sgd_step_allreduce(float* d_param, const float* const* h_peer_grads, uint32_t* const* h_peer_flags, int32_t rank,
                           int32_t world, uint32_t seq, float* d_momentum_buffer, int64_t n, double lr, double momentum,
                           double weight_decay, int64_t step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MZB200_H */
