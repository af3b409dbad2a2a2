// ppo_loss.cu - K6: PPO objective of one minibatch, forward and backward in one pass.
//
// What RLlib's PPOTorchLearner computes for the reference's algorithms (hyper-parameters forwarded at
// src/algorithms/ippo.py:145-160; hysteretic weighting of negative advantages: learners/hysteretic_learner.py:39-42),
// per agent-sample i with action mean mu_i[S], state-independent log_std[S] (floored, rlmodules/base.py:473-478):
//   logp_i  = sum_s -0.5 z^2 - log_std_s - 0.5 log(2 pi),  z = (a - mu) / std
//   ratio_i = exp(logp_i - logp_old_i),  adv'_i = adv_i (* beta where negative)
//   L = -mean(min(ratio adv', clip(ratio, 1-c, 1+c) adv')) + vf_coeff mean(min((v - target)^2, vf_clip)) - ent_coeff H
// One thread per agent-sample writes dL/dmu and dL/dv; the sums (surrogate, value loss, dL/dlog_std) are reduced per
// warp and added to a small float64 accumulator. Replaces ~25 elementwise PyTorch kernels and their autograd graph.
#include <cuda_runtime.h>

#include "lib_common.h"

namespace marlsc {
namespace {

constexpr int kMaxS = 512;

__global__ void __launch_bounds__(256)
ppo_loss_kernel(const float* __restrict__ mean, const float* __restrict__ actions, const float* __restrict__ log_std, float floor_,
                const float* __restrict__ logp_old, const float* __restrict__ adv, const float* __restrict__ value,
                const float* __restrict__ targets, long long N, int S, float clip, float vf_clip, float vf_coeff, float beta,
                float* __restrict__ grad_mean, float* __restrict__ grad_value, double* __restrict__ sums) {
  extern __shared__ float sh[];                       // [S] floored log_std, [S] 1/std
  float* s_ls = sh;
  float* s_inv = sh + S;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const float ls = fmaxf(log_std[s], floor_);
    s_ls[s] = ls;
    s_inv[s] = expf(-ls);
  }
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < N;
  const float invN = 1.0f / (float)N;
  float surr = 0.f, vf = 0.f, coef = 0.f;             // coef = dL/dlogp_i
  if (live) {
    const float* mu = mean + i * S;
    const float* a = actions + i * S;
    float logp = 0.f;
    for (int s = 0; s < S; ++s) {
      const float z = (a[s] - mu[s]) * s_inv[s];
      logp += -0.5f * z * z - s_ls[s] - 0.9189385332046727f;
    }
    const float ratio = expf(logp - logp_old[i]);
    float ad = adv[i];
    if (beta >= 0.f && ad < 0.f) ad *= beta;
    const float s1 = ratio * ad, s2 = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip) * ad;
    surr = fminf(s1, s2);
    // d min(s1, s2) / d ratio: through s1 when it is the smaller one (inside the clip range both are the same function)
    const bool inside = ratio >= 1.f - clip && ratio <= 1.f + clip;
    const float dsurr = (inside || s1 < s2) ? ad : 0.f;
    coef = -invN * dsurr * ratio;
    const float d = value[i] - targets[i];
    const float sq = d * d;
    vf = fminf(sq, vf_clip);
    grad_value[i] = sq <= vf_clip ? vf_coeff * invN * 2.f * d : 0.f;
    float* gm = grad_mean + i * S;
    for (int s = 0; s < S; ++s) gm[s] = coef * (a[s] - mu[s]) * s_inv[s] * s_inv[s];   // dlogp/dmu = z / std
  }
  // warp sums -> float64 accumulators: [0] surrogate, [1] value loss, [2 + s] dL/dlog_std_s (without the entropy term)
  double a0 = surr, a1 = vf;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&sums[0], a0);
    atomicAdd(&sums[1], a1);
  }
  for (int s = 0; s < S; ++s) {
    float g = 0.f;
    if (live) {
      const float z = (actions[i * S + s] - mean[i * S + s]) * s_inv[s];
      g = coef * (z * z - 1.f);                       // dlogp/dlog_std = z^2 - 1
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
    if ((threadIdx.x & 31) == 0 && g != 0.f) atomicAdd(&sums[2 + s], (double)g);
  }
}

}  // namespace
}  // namespace marlsc

using namespace marlsc;

extern "C" {

int marlsc_ppo_loss(const float* mean, const float* actions, const float* log_std, float logstd_floor, const float* logp_old,
                    const float* adv, const float* value, const float* targets, int64_t n_samples, int32_t action_dim,
                    float clip_param, float vf_clip_param, float vf_loss_coeff, float hysteretic_beta, float* grad_mean,
                    float* grad_value, double* sums, void* stream) {
  if (!mean || !actions || !log_std || !logp_old || !adv || !value || !targets || !grad_mean || !grad_value || !sums)
    return set_error(MARLSC_EINVAL, "null argument");
  if (n_samples < 1 || action_dim < 1 || action_dim > kMaxS) return set_error(MARLSC_EINVAL, "n_samples must be positive and action_dim in [1, 512]");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MARLSC_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * (2 + (size_t)action_dim), s));
  const unsigned grid = (unsigned)((n_samples + 255) / 256);
  ppo_loss_kernel<<<grid, 256, 2 * action_dim * sizeof(float), s>>>(mean, actions, log_std, logstd_floor, logp_old, adv, value, targets,
                                                                    n_samples, action_dim, clip_param, vf_clip_param, vf_loss_coeff,
                                                                    hysteretic_beta, grad_mean, grad_value, sums);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MARLSC_CUDA(cudaGetLastError());
  return MARLSC_OK;
}

}  // extern "C"
