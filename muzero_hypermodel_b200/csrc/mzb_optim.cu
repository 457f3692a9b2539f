// Optimiser step on ONE flat float32 parameter bucket: the trainer keeps every parameter and
// every gradient as a view into two flat buffers, so the data-parallel step is one NCCL all-reduce of the gradient
// bucket followed by one launch here - instead of torch.optim's per-tensor (or per-chunk foreach) launches.
// Reference: torch.optim.Adam / torch.optim.SGD as configured by trainer.py:35-52 (L2 weight decay added to the
// gradient, Adam without amsgrad, SGD with momentum and no dampening / nesterov); same operation order.
#include <string.h>

#include "mzb_common.cuh"

namespace {

__global__ void k_adam_flat(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float grad_scale, float weight_decay, float beta1, float beta2, float step_size,
                            float bias2_sqrt, float eps) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float w = p[i];
  const float grad = fmaf(weight_decay, w, g[i] * grad_scale);                // grad.add(param, alpha=weight_decay)
  const float m1 = fmaf(grad - m[i], 1.0f - beta1, m[i]);                     // exp_avg.lerp_(grad, 1 - beta1)
  const float v1 = fmaf((1.0f - beta2) * grad, grad, v[i] * beta2);           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  m[i] = m1;
  v[i] = v1;
  const float denom = sqrtf(v1) / bias2_sqrt + eps;
  p[i] = w - step_size * (m1 / denom);                                        // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void k_sgd_flat(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long n,
                           float grad_scale, float weight_decay, float momentum, float lr, int first) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float w = p[i];
  float grad = fmaf(weight_decay, w, g[i] * grad_scale);
  if (momentum != 0.0f) {
    const float b = first ? grad : fmaf(momentum, buf[i], grad);             // buf.mul_(momentum).add_(grad)
    buf[i] = b;
    grad = b;
  }
  p[i] = w - lr * grad;
}

// ---- gradient all-reduce FUSED into the optimiser step, over NVLink peer memory (the trainer's
// gradient all-reduce is the only collective of the system).  Every rank's flat gradient bucket lives in an
// IPC-exported allocation that all ranks of the node map; one launch per rank
//   1. signals "my bucket of step `seq` is complete" into every peer's flag array (system-scope release),
//   2. waits until all peers signalled `seq` (bounded spin: a missing peer traps instead of hanging the GPU),
//   3. streams through the parameters: g = sum over ranks IN RANK ORDER of bucket_r[i] (peer loads over NVLink /
//      NVSwitch, .cv: never served from a stale line) * 1/world, then the Adam / SGD update of k_adam_flat / k_sgd_flat.
// All ranks sum in the same order, so the replicas stay bit-identical without a broadcast; the transfer of element i
// overlaps the update of element i-1 in the same kernel, and a 1.5 K - 5.5 M parameter model pays ONE launch latency
// instead of an NCCL all-reduce + an optimiser launch.  Buckets are double buffered by step parity: a rank overwrites
// bucket (seq & 1) only after its step seq-1 kernel, which waited for every peer's `seq-1` signal, i.e. for every
// peer to have finished the step seq-2 kernel that read it.
constexpr int kMaxPeers = 8;
struct PeerArgs {
  const float* grad[kMaxPeers];      // this step's bucket of every rank (own included), peer-mapped
  uint32_t* flags[kMaxPeers];        // flags[r][q]: rank q's latest completed step, stored in rank r's memory
  int rank, world;
  uint32_t seq;
};

__device__ __forceinline__ void peer_barrier(const PeerArgs& a) {
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.flags[threadIdx.x] + a.rank), "r"(a.seq) : "memory");
  }
  if (threadIdx.x < a.world) {
    const uint32_t* f = a.flags[a.rank] + threadIdx.x;
    uint32_t v = 0, spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (++spins > (1u << 28)) __trap();                 // a peer never arrived: fail the launch, do not hang
    } while ((int32_t)(v - a.seq) < 0);
  }
  __syncthreads();
}

__device__ __forceinline__ float peer_sum(const PeerArgs& a, long long i) {
  float g = 0.0f;
  for (int r = 0; r < a.world; ++r) g += __ldcv(a.grad[r] + i);
  return g;
}

__global__ void __launch_bounds__(256) k_adam_allreduce(float* __restrict__ p, PeerArgs a, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, float grad_scale, float weight_decay,
                                                        float beta1, float beta2, float step_size, float bias2_sqrt, float eps) {
  peer_barrier(a);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float w = p[i];
    const float grad = fmaf(weight_decay, w, peer_sum(a, i) * grad_scale);
    const float m1 = fmaf(grad - m[i], 1.0f - beta1, m[i]);
    const float v1 = fmaf((1.0f - beta2) * grad, grad, v[i] * beta2);
    m[i] = m1;
    v[i] = v1;
    const float denom = sqrtf(v1) / bias2_sqrt + eps;
    p[i] = w - step_size * (m1 / denom);
  }
}

__global__ void __launch_bounds__(256) k_sgd_allreduce(float* __restrict__ p, PeerArgs a, float* __restrict__ buf, long long n,
                                                       float grad_scale, float weight_decay, float momentum, float lr, int first) {
  peer_barrier(a);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float w = p[i];
    float grad = fmaf(weight_decay, w, peer_sum(a, i) * grad_scale);
    if (momentum != 0.0f) {
      const float b = first ? grad : fmaf(momentum, buf[i], grad);
      buf[i] = b;
      grad = b;
    }
    p[i] = w - lr * grad;
  }
}

int fill_peers(PeerArgs& a, const float* const* h_peer_grads, uint32_t* const* h_peer_flags, int rank, int world, uint32_t seq) {
  MZB_CHECK_ARG(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "rank %d / world %d outside 1..%d", rank, world, kMaxPeers);
  for (int r = 0; r < world; ++r) {
    MZB_CHECK_ARG(h_peer_grads[r] && h_peer_flags[r], "peer %d: NULL bucket or flag pointer", r);
    a.grad[r] = h_peer_grads[r];
    a.flags[r] = h_peer_flags[r];
  }
  a.rank = rank; a.world = world; a.seq = seq;
  return MZB_OK;
}

}  // namespace

extern "C" {

// ---- peer-mapped allocations: cudaMalloc'd (IPC-exportable, unlike a caching allocator's sub-blocks), zero-filled
int mzb_p2p_alloc(void** d_ptr, size_t bytes) {
  MZB_CHECK_ARG(d_ptr && bytes > 0, "bad argument");
  MZB_CUDA(cudaMalloc(d_ptr, bytes));
  MZB_CUDA(cudaMemset(*d_ptr, 0, bytes));
  return MZB_OK;
}
int mzb_p2p_free(void* d_ptr) {
  if (d_ptr) MZB_CUDA(cudaFree(d_ptr));
  return MZB_OK;
}
int mzb_p2p_export(void* d_ptr, uint8_t* handle64) {
  MZB_CHECK_ARG(d_ptr && handle64, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  MZB_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
  memcpy(handle64, &h, 64);
  return MZB_OK;
}
int mzb_p2p_import(const uint8_t* handle64, void** d_ptr) {
  MZB_CHECK_ARG(d_ptr && handle64, "NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  MZB_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return MZB_OK;
}
int mzb_p2p_close(void* d_ptr) {
  if (d_ptr) MZB_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return MZB_OK;
}

int mzb_adam_step_allreduce(float* d_param, const float* const* h_peer_grads, uint32_t* const* h_peer_flags, int32_t rank,
                            int32_t world, uint32_t seq, float* d_exp_avg, float* d_exp_avg_sq, int64_t n, double lr,
                            double beta1, double beta2, double eps, double weight_decay, int64_t step, void* stream) {
  MZB_CHECK_ARG(d_param && h_peer_grads && h_peer_flags && d_exp_avg && d_exp_avg_sq && n > 0 && step >= 1, "bad argument");
  PeerArgs a{};
  if (int rc = fill_peers(a, h_peer_grads, h_peer_flags, rank, world, seq)) return rc;
  const double bias1 = 1.0 - pow(beta1, (double)step), bias2 = 1.0 - pow(beta2, (double)step);
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;             // every block takes part in the barrier: keep them co-resident
  k_adam_allreduce<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_param, a, d_exp_avg, d_exp_avg_sq, n, (float)(1.0 / world),
                                                                        (float)weight_decay, (float)beta1, (float)beta2,
                                                                        (float)(lr / bias1), (float)sqrt(bias2), (float)eps);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

int mzb_sgd_step_allreduce(float* d_param, const float* const* h_peer_grads, uint32_t* const* h_peer_flags, int32_t rank,
                           int32_t world, uint32_t seq, float* d_momentum_buffer, int64_t n, double lr, double momentum,
                           double weight_decay, int64_t step, void* stream) {
  MZB_CHECK_ARG(d_param && h_peer_grads && h_peer_flags && n > 0 && step >= 1 && (momentum == 0.0 || d_momentum_buffer), "bad argument");
  PeerArgs a{};
  if (int rc = fill_peers(a, h_peer_grads, h_peer_flags, rank, world, seq)) return rc;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_sgd_allreduce<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_param, a, d_momentum_buffer, n, (float)(1.0 / world),
                                                                       (float)weight_decay, (float)momentum, (float)lr, step == 1);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

int mzb_adam_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n, double lr,
                  double beta1, double beta2, double eps, double weight_decay, int64_t step, double grad_scale, void* stream) {
  MZB_CHECK_ARG(d_param && d_grad && d_exp_avg && d_exp_avg_sq && n > 0 && step >= 1, "bad argument");
  const double bias1 = 1.0 - pow(beta1, (double)step), bias2 = 1.0 - pow(beta2, (double)step);
  k_adam_flat<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_param, d_grad, d_exp_avg, d_exp_avg_sq, n, (float)grad_scale, (float)weight_decay, (float)beta1, (float)beta2,
      (float)(lr / bias1), (float)sqrt(bias2), (float)eps);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

int mzb_sgd_step(float* d_param, const float* d_grad, float* d_momentum_buffer, int64_t n, double lr, double momentum,
                 double weight_decay, int64_t step, double grad_scale, void* stream) {
  MZB_CHECK_ARG(d_param && d_grad && n > 0 && step >= 1 && (momentum == 0.0 || d_momentum_buffer), "bad argument");
  k_sgd_flat<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_param, d_grad, d_momentum_buffer, n, (float)grad_scale, (float)weight_decay, (float)momentum, (float)lr, step == 1);
  MZB_LAUNCH_CHECK();
  return MZB_OK;
}

}  // extern "C"
