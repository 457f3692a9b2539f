// Batched MCTS.run for residual networks (self_play.py:261-362): tree kernels + the resnet layer program per
// simulation.  Hidden states live in a caller-owned pool [G][S+1][H*W*C] in dense NHWC (bf16 on the tensor-core
// path, fp32 on the exact path); slot = node index, so recurrent_inference reads the parent's slot in place.
#include <stdlib.h>
#include <string.h>

#include "mzb_resnet_model.h"
#include "mzb_tree.cuh"

namespace {

int search_resnet_launches(mzb_tree* t, mzb_resnet_model* m, const float* d_obs, const uint8_t* d_legal,
                           const int8_t* d_to_play, const double* d_noise, double alpha, double frac,
                           const uint32_t* d_slot, const uint32_t* d_step, int32_t num_simulations, void* d_hidden_pool,
                           void* d_workspace, size_t workspace_bytes, int32_t* d_visits, double* d_root_value,
                           float* d_root_predicted_value, int32_t* d_max_depth, cudaStream_t stream);

// ---- CUDA-graph replay.  One search is ~17 launches per simulation (3,400 for connect4); self-play calls it once per
// move with the SAME buffers, so the launch sequence of a search is captured once and replayed: the second call with
// an identical argument set captures (on an internal stream - torch's default stream cannot be captured), later calls
// launch the instantiated graph into the caller's stream.  Every value that changes from move to move (observations,
// legal masks, RNG counters, the tree) lives in device memory behind those pointers.  MZB_NO_GRAPH=1 disables.
struct GraphKey {
  const void* p[20];
  double alpha, frac;
  long long sims, ws;
};
struct GraphSlot {
  GraphKey key;
  cudaGraph_t graph = nullptr;
  // two instances of the same graph, launched alternately: an executable graph runs once at a time, so re-launching
  // the ONE instance makes the driver wait for the move in flight before it can stage the next one (the GPU then
  // idles for the staging latency of a 160 - 3,400-node graph every move); with two, move n+1 is staged while n runs
  cudaGraphExec_t exec = nullptr, exec2 = nullptr;
  unsigned long long n_replays = 0;
  long long launches = 0;
  int seen = 0;
};
GraphSlot g_slot[4];
cudaStream_t g_capture_stream = nullptr;

void slot_reset(GraphSlot& s) {
  if (s.exec) cudaGraphExecDestroy(s.exec);
  if (s.exec2) cudaGraphExecDestroy(s.exec2);
  if (s.graph) cudaGraphDestroy(s.graph);
  s = GraphSlot{};
}

bool graphs_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MZB_NO_GRAPH"); v = (e && atoi(e) != 0) ? 0 : 1; }
  return v == 1;
}

}  // namespace

// called by mzb_tree_destroy / mzb_resnet_destroy: drop graphs that reference the handle
void mzb_search_graph_forget(const void* handle) {
  for (auto& s : g_slot)
    if (s.seen && (s.key.p[0] == handle || s.key.p[1] == handle)) slot_reset(s);
}

extern "C" int mzb_search_resnet(mzb_tree* t, mzb_resnet_model* m, const float* d_obs, const uint8_t* d_legal,
                                 const int8_t* d_to_play, const double* d_noise, double alpha, double frac,
                                 const uint32_t* d_slot, const uint32_t* d_step, int32_t num_simulations,
                                 void* d_hidden_pool, void* d_workspace, size_t workspace_bytes, int32_t* d_visits,
                                 double* d_root_value, float* d_root_predicted_value, int32_t* d_max_depth, void* stream) {
  MZB_CHECK_ARG(t && m && d_obs && d_hidden_pool && d_workspace, "NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (!graphs_enabled())
    return search_resnet_launches(t, m, d_obs, d_legal, d_to_play, d_noise, alpha, frac, d_slot, d_step, num_simulations,
                                  d_hidden_pool, d_workspace, workspace_bytes, d_visits, d_root_value,
                                  d_root_predicted_value, d_max_depth, s);
  GraphKey key;
  memset(&key, 0, sizeof(key));
  const void* ptrs[] = {t, m, d_obs, d_legal, d_to_play, d_noise, d_slot, d_step, d_hidden_pool, d_workspace, d_visits,
                        d_root_value, d_root_predicted_value, d_max_depth, t->v.nodes, t->tmp_parent, m->rep_conv.w};
  for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); ++i) key.p[i] = ptrs[i];
  key.alpha = alpha; key.frac = frac; key.sims = num_simulations; key.ws = (long long)workspace_bytes;
  GraphSlot* slot = nullptr;
  for (auto& c : g_slot)
    if (c.seen && memcmp(&c.key, &key, sizeof(key)) == 0) slot = &c;
  if (!slot) {                                      // first call with these arguments: plain launches, remember the key
    static int next = 0;
    slot = &g_slot[next++ % 4];
    slot_reset(*slot);
    slot->key = key;
    slot->seen = 1;
    return search_resnet_launches(t, m, d_obs, d_legal, d_to_play, d_noise, alpha, frac, d_slot, d_step, num_simulations,
                                  d_hidden_pool, d_workspace, workspace_bytes, d_visits, d_root_value,
                                  d_root_predicted_value, d_max_depth, s);
  }
  if (!slot->exec) {                                // second call: capture
    if (!g_capture_stream) MZB_CUDA(cudaStreamCreateWithFlags(&g_capture_stream, cudaStreamNonBlocking));
    const uint64_t before = mzb_launch_count();
    MZB_CUDA(cudaStreamBeginCapture(g_capture_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = search_resnet_launches(t, m, d_obs, d_legal, d_to_play, d_noise, alpha, frac, d_slot, d_step,
                                          num_simulations, d_hidden_pool, d_workspace, workspace_bytes, d_visits,
                                          d_root_value, d_root_predicted_value, d_max_depth, g_capture_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(g_capture_stream, &graph);
    slot->launches = (long long)(mzb_launch_count() - before);
    mzb_count_launch(-(int)slot->launches);         // nothing ran yet
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    MZB_CUDA(ce);
    slot->graph = graph;
    MZB_CUDA(cudaGraphInstantiate(&slot->exec, graph, 0));
    MZB_CUDA(cudaGraphInstantiate(&slot->exec2, graph, 0));
  }
  MZB_CUDA(cudaGraphLaunch((slot->n_replays++ & 1) ? slot->exec2 : slot->exec, s));
  mzb_count_launch((int)slot->launches);
  return MZB_OK;
}

namespace {

int search_resnet_launches(mzb_tree* t, mzb_resnet_model* m, const float* d_obs, const uint8_t* d_legal,
                           const int8_t* d_to_play, const double* d_noise, double alpha, double frac,
                           const uint32_t* d_slot, const uint32_t* d_step, int32_t num_simulations, void* d_hidden_pool,
                           void* d_workspace, size_t workspace_bytes, int32_t* d_visits, double* d_root_value,
                           float* d_root_predicted_value, int32_t* d_max_depth, cudaStream_t stream) {
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

}  // namespace
