// optim.cu -- optimizer step on flat fp32 buffers (torch.optim.Adam / SGD as configured at train.py:402-405).
//
// The engine keeps all parameters, gradients and optimizer state of the model in a few contiguous fp32
// buffers, so one launch updates everything: 16 B/param read (p, g, m, v) + 12 B/param written -> HBM-bound.
// Semantics follow torch.optim.Adam (no amsgrad, L2 weight decay added to the gradient) and torch.optim.SGD
// (momentum with dampening 0, optional nesterov), with an optional gradient pre-scale (1/world for the
// data-parallel average, 1/loss_scale for AMP).
#include "b2u_internal.h"

namespace b2u {

__global__ void adam_step_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                 float4* __restrict__ v, long long n4, float lr, float b1, float b2, float eps, float wd,
                                 float bc1, float bc2_sqrt, float gscale) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
  float* P = reinterpret_cast<float*>(&pp); float* G = reinterpret_cast<float*>(&gg);
  float* M = reinterpret_cast<float*>(&mm); float* V = reinterpret_cast<float*>(&vv);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float grad = G[k] * gscale;
    if (wd != 0.f) grad += wd * P[k];
    M[k] = M[k] + (1.f - b1) * (grad - M[k]);            // lerp form used by torch
    V[k] = b2 * V[k] + (1.f - b2) * grad * grad;
    const float denom = sqrtf(V[k]) / bc2_sqrt + eps;
    P[k] = P[k] - (lr / bc1) * (M[k] / denom);
  }
  p[i] = pp; m[i] = mm; v[i] = vv;
}

__global__ void sgd_step_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ buf,
                                long long n4, float lr, float mom, float wd, int nesterov, int first, float gscale) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 pp = p[i], gg = g[i], bb = buf ? buf[i] : make_float4(0, 0, 0, 0);
  float* P = reinterpret_cast<float*>(&pp); float* G = reinterpret_cast<float*>(&gg); float* B = reinterpret_cast<float*>(&bb);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float grad = G[k] * gscale;
    if (wd != 0.f) grad += wd * P[k];
    if (mom != 0.f) {
      B[k] = first ? grad : mom * B[k] + grad;
      grad = nesterov ? grad + mom * B[k] : B[k];
    }
    P[k] -= lr * grad;
  }
  p[i] = pp;
  if (buf) buf[i] = bb;
}

}  // namespace b2u

extern "C" {
using namespace b2u;

// n must be a multiple of 4 and the buffers 16-byte aligned (the engine pads its flat buffers).
int b2u_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
  if (n <= 0 || n % 4 != 0 || step < 1) return set_error(B2U_ERR_ARG, "adam_step: n must be a positive multiple of 4, step >= 1");
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
  const long long n4 = n / 4;
  adam_step_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float4*>(param), reinterpret_cast<const float4*>(grad), reinterpret_cast<float4*>(exp_avg),
      reinterpret_cast<float4*>(exp_avg_sq), n4, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), grad_scale);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B2U_ERR_CUDA, "adam_step launch: %s", cudaGetErrorString(e));
  note_launch();
  return 0;
}

int b2u_sgd_step(float* param, const float* grad, float* momentum_buf, long long n, float lr, float momentum,
                 float weight_decay, int nesterov, int first_step, float grad_scale, void* stream) {
  if (n <= 0 || n % 4 != 0) return set_error(B2U_ERR_ARG, "sgd_step: n must be a positive multiple of 4");
  if (momentum != 0.f && !momentum_buf) return set_error(B2U_ERR_ARG, "sgd_step: momentum needs a buffer");
  const long long n4 = n / 4;
  sgd_step_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float4*>(param), reinterpret_cast<const float4*>(grad), reinterpret_cast<float4*>(momentum_buf), n4,
      lr, momentum, weight_decay, nesterov, first_step, grad_scale);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B2U_ERR_CUDA, "sgd_step launch: %s", cudaGetErrorString(e));
  note_launch();
  return 0;
}

}  // extern "C"
