// Batched MCTS.run for residual networks (self_play.py:261-362): tree kernels + the resnet layer program per
// simulation.  Hidden states live in a caller-owned pool [G][S+1][H*W*C] in dense NHWC (bf16 on the tensor-core
// path, fp32 on the exact path); slot = node index, so recurrent_inference reads the parent's slot in place.
#include "mzb_resnet_model.h"
#include "mzb_tree.cuh"

extern "C" int mzb_search_resnet(mzb_tree* t, mzb_resnet_model* m, const float* d_obs, const uint8_t* d_legal,
                                 const int8_t* d_to_play, const double* d_noise, double alpha, double frac,
                                 const uint32_t* d_slot, const uint32_t* d_step, int32_t num_simulations,
                                 void* d_hidden_pool, void* d_workspace, size_t workspace_bytes, int32_t* d_visits,
                                 double* d_root_value, float* d_root_predicted_value, int32_t* d_max_depth, void* stream) {
  MZB_CHECK_ARG(t && m && d_obs && d_hidden_pool && d_workspace, "NULL argument");
  MZB_CHECK_ARG(t->v.A == m->A, "tree has %d actions, network %d", t->v.A, m->A);
  MZB_CHECK_ARG(num_simulations > 0 && num_simulations <= t->v.S, "num_simulations %d outside 1..%d", num_simulations, t->v.S);
  MZB_CHECK_ARG(frac >= 0.0 && frac <= 1.0, "exploration fraction out of [0,1]: %f", frac);
  const int G = t->v.G;
  const int layout = m->precision == 1 ? 2 : 1;
  const int64_t state = (int64_t)m->Hl * m->Wl * m->C;
  const int64_t row_stride = (int64_t)(t->v.S + 1) * state;
  int rc = mzb_resnet_initial(m, G, d_obs, d_legal, d_workspace, workspace_bytes, d_hidden_pool, layout, row_stride, 0,
                              nullptr, nullptr, nullptr, d_root_predicted_value, t->tmp_reward, t->tmp_priors, stream);
  if (rc) return rc;
  rc = mzb_tree_root_init(t, t->tmp_reward, t->tmp_priors, 0, d_legal, d_to_play, d_noise, alpha, frac, d_slot, d_step, stream);
  if (rc) return rc;
  for (int sim = 0; sim < num_simulations; ++sim) {
    rc = mzb_tree_select(t, t->tmp_parent, t->tmp_action, nullptr, stream);
    if (rc) return rc;
    rc = mzb_resnet_recurrent(m, G, d_hidden_pool, layout, row_stride, t->tmp_parent, state, t->tmp_action, d_workspace,
                              workspace_bytes, d_hidden_pool, layout, row_stride, (int64_t)(sim + 1) * state, nullptr,
                              nullptr, nullptr, t->tmp_value, t->tmp_reward, t->tmp_priors, stream);
    if (rc) return rc;
    rc = mzb_tree_expand_backup(t, t->tmp_value, t->tmp_reward, t->tmp_priors, 0, stream);
    if (rc) return rc;
  }
  if (d_visits || d_root_value || d_max_depth)
    return mzb_tree_root_stats(t, d_visits, d_root_value, d_max_depth, nullptr, nullptr, nullptr, nullptr, stream);
  return MZB_OK;
}
