// ppo_loss.cu - K6: PPO objective of one minibatch, forward and backward in one pass.
//
// What RLlib's PPOTorchLearner computes for the reference's algorithms (hyper-parameters forwarded at
// src/algorithms/ippo.py:145-160 / mappo.py:142-157; hysteretic weighting of negative advantages:
// learners/hysteretic_learner.py:39-42), per agent-sample i of policy p with action mean mu_i[S] and that policy's
// state-independent log_std_p[S] (floored, rlmodules/base.py:473-478):
//   logp_i  = sum_s -0.5 z^2 - log_std_s - 0.5 log(2 pi),  z = (a - mu) / std
//   ratio_i = exp(logp_i - logp_old_i),  adv'_i = adv_i (* beta where negative)
//   L_p = -mean_p(min(ratio adv', clip(ratio, 1-c, 1+c) adv')) + vf_coeff mean_p(min((v - target)^2, vf_clip))
//         + kl_coeff mean_p(KL(old_i || new_i))                                   (use_kl_loss, ippo.py:146)
// and L = sum_p L_p: with parameter sharing (ippo.py:106-110) there is one policy and the mean runs over every
// agent-sample; without it (ippo.py:111-115) sample i belongs to policy i % n_policies (the warehouse index is the
// fastest axis of the flattened [batch, warehouse] minibatch) and every policy averages over its own samples, as RLlib's
// per-module losses do. The entropy bonus only depends on log_std and is left to the caller.
// One thread per agent-sample writes dL/dmu and dL/dv; the sums (surrogate, value loss, KL, dL/dlog_std) are reduced per
// warp and added to a small float64 accumulator. Replaces ~30 elementwise PyTorch kernels and their autograd graph.
#include <cuda_runtime.h>

#include "lib_common.h"

namespace marlsc {
namespace {

constexpr int kMaxS = 512;

// sums layout: [n_policies][3 + S] = surrogate, value loss, KL, dL/dlog_std_s
__global__ void __launch_bounds__(256)
ppo_loss_kernel(const float* __restrict__ mean, const float* __restrict__ actions, const float* __restrict__ log_std, int P,
                float floor_, const float* __restrict__ logp_old, const float* __restrict__ adv, const float* __restrict__ value,
                const float* __restrict__ targets, const float* __restrict__ mean_old, const float* __restrict__ log_std_old,
                float kl_coeff, long long N, int S, float clip, float vf_clip, float vf_coeff, float beta,
                float* __restrict__ grad_mean, float* __restrict__ grad_value, double* __restrict__ sums) {
  extern __shared__ float sh[];                       // [P,S] floored log_std, [P,S] 1/std, [P,S] old variance (KL only)
  float* s_ls = sh;
  float* s_inv = sh + P * S;
  float* s_var_old = sh + 2 * P * S;
  const bool use_kl = mean_old != nullptr;
  for (int k = threadIdx.x; k < P * S; k += blockDim.x) {
    const float ls = fmaxf(log_std[k], floor_);
    s_ls[k] = ls;
    s_inv[k] = expf(-ls);
    if (use_kl) s_var_old[k] = expf(2.f * log_std_old[k]);
  }
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < N;
  const int p = P > 1 ? (int)(i % P) : 0;
  const float invN = (float)P / (float)N;             // every policy averages over its own N / P samples
  const float* ls_p = s_ls + p * S;
  const float* inv_p = s_inv + p * S;
  float surr = 0.f, vf = 0.f, kl = 0.f, coef = 0.f;   // coef = dL/dlogp_i
  if (live) {
    const float* mu = mean + i * S;
    const float* a = actions + i * S;
    float logp = 0.f;
    for (int s = 0; s < S; ++s) {
      const float z = (a[s] - mu[s]) * inv_p[s];
      logp += -0.5f * z * z - ls_p[s] - 0.9189385332046727f;
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
    const float* mo = use_kl ? mean_old + i * S : nullptr;
    for (int s = 0; s < S; ++s) {
      float g = coef * (a[s] - mu[s]) * inv_p[s] * inv_p[s];   // dlogp/dmu = z / std
      if (use_kl) {
        // KL(N(mo, so) || N(mu, sn)) = log(sn / so) + (so^2 + (mo - mu)^2) / (2 sn^2) - 1/2
        const float dm = mu[s] - mo[s], iv = inv_p[s] * inv_p[s];
        kl += ls_p[s] - 0.5f * logf(s_var_old[p * S + s]) + 0.5f * (s_var_old[p * S + s] + dm * dm) * iv - 0.5f;
        g += kl_coeff * invN * dm * iv;
      }
      gm[s] = g;
    }
  }
  // warp sums -> float64 accumulators. With several policies the lanes of a warp belong to different policies: the
  // reduction runs once per policy present in the warp (P <= a handful).
  for (int q = 0; q < P; ++q) {
    const bool mine = live && p == q;
    double a0 = mine ? surr : 0.f, a1 = mine ? vf : 0.f, a2 = mine ? kl : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    double* sq_ = sums + (size_t)q * (3 + S);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&sq_[0], a0);
      atomicAdd(&sq_[1], a1);
      if (use_kl) atomicAdd(&sq_[2], a2);
    }
    for (int s = 0; s < S; ++s) {
      float g = 0.f;
      if (mine) {
        const float dm = actions[i * S + s] - mean[i * S + s];
        const float z = dm * inv_p[s];
        g = coef * (z * z - 1.f);                     // dlogp/dlog_std = z^2 - 1
        if (use_kl) {
          const float dk = mean[i * S + s] - mean_old[i * S + s];
          g += kl_coeff * invN * (1.f - (s_var_old[q * S + s] + dk * dk) * inv_p[s] * inv_p[s]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
      if ((threadIdx.x & 31) == 0 && g != 0.f) atomicAdd(&sq_[3 + s], (double)g);
    }
  }
}

}  // namespace
}  // namespace marlsc

using namespace marlsc;

extern "C" {

int marlsc_ppo_loss(const float* mean, const float* actions, const float* log_std, int32_t n_policies, float logstd_floor,
                    const float* logp_old, const float* adv, const float* value, const float* targets, const float* mean_old,
                    const float* log_std_old, float kl_coeff, int64_t n_samples, int32_t action_dim, float clip_param,
                    float vf_clip_param, float vf_loss_coeff, float hysteretic_beta, float* grad_mean, float* grad_value,
                    double* sums, void* stream) {
  if (!mean || !actions || !log_std || !logp_old || !adv || !value || !targets || !grad_mean || !grad_value || !sums)
    return set_error(MARLSC_EINVAL, "null argument");
  if (n_samples < 1 || action_dim < 1 || action_dim > kMaxS) return set_error(MARLSC_EINVAL, "n_samples must be positive and action_dim in [1, 512]");
  if (n_policies < 1 || n_policies > 64 || n_samples % n_policies != 0)
    return set_error(MARLSC_EINVAL, "n_policies must be in [1, 64] and divide n_samples (the policy index is the fastest axis)");
  if ((mean_old == nullptr) != (log_std_old == nullptr)) return set_error(MARLSC_EINVAL, "mean_old and log_std_old go together");
  const size_t smem = 3 * (size_t)n_policies * action_dim * sizeof(float);
  if (smem > 48 * 1024) return set_error(MARLSC_EINVAL, "n_policies * action_dim too large");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MARLSC_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)n_policies * (3 + (size_t)action_dim), s));
  const unsigned grid = (unsigned)((n_samples + 255) / 256);
  ppo_loss_kernel<<<grid, 256, smem, s>>>(mean, actions, log_std, n_policies, logstd_floor, logp_old, adv, value, targets, mean_old,
                                          log_std_old, kl_coeff, n_samples, action_dim, clip_param, vf_clip_param, vf_loss_coeff,
                                          hysteretic_beta, grad_mean, grad_value, sums);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MARLSC_CUDA(cudaGetLastError());
  return MARLSC_OK;
}

}  // extern "C"
