// Optimiser step on ONE flat float32 parameter bucket: the trainer keeps every parameter and
// every gradient as a view into two flat buffers, so the data-parallel step is one NCCL all-reduce of the gradient
// bucket followed by one launch here - instead of torch.optim's per-tensor (or per-chunk foreach) launches.
// Reference: torch.optim.Adam / torch.optim.SGD as configured by trainer.py:35-52 (L2 weight decay added to the
// gradient, Adam without amsgrad, SGD with momentum and no dampening / nesterov); same operation order.
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

}  // namespace

extern "C" {

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
