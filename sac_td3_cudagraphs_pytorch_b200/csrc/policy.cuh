// policy.cuh — the action heads: SAC tanh-Gaussian (sample, log-prob and their backward) and the
// TD3 deterministic tanh policy. Restates agents/nets.py:143-147 (Actor.forward), :206-234
// (TanhGaussActor.bound_log_std / forward / get_action) and torch.distributions.Normal
// (rsample = loc + eps*scale; log_prob = -(x-loc)^2/(2 var) - log(scale) - log(sqrt(2 pi))).
// Rounding follows the reference's operation order (no FMA contraction across its op boundaries).
#pragma once
#include "common.cuh"

namespace b2rl {

constexpr float LOG_STD_MIN = -5.0f, LOG_STD_MAX = 2.0f;  // agents/nets.py:13
constexpr float LOG_SQRT_2PI = 0.9189385332046727f;

struct GaussSample {  // everything the backward pass needs for one (row, action-dim)
  float action, logp;  // a = tanh(x)*scale + bias ; log N(x) - log(scale*(1-y^2)+1e-6)
  float y, sigma, th;  // tanh(x), exp(log_std), tanh(raw log_std)
};

__device__ __forceinline__ GaussSample gauss_sample(float mu, float ls_raw, float eps, float scale, float bias) {
  GaussSample s;
  s.th = tanhf(ls_raw);
  const float ls = __fadd_rn(LOG_STD_MIN, __fmul_rn(0.5f * (LOG_STD_MAX - LOG_STD_MIN), __fadd_rn(s.th, 1.0f)));
  s.sigma = expf(ls);
  const float x = __fadd_rn(mu, __fmul_rn(eps, s.sigma));
  s.y = tanhf(x);
  s.action = __fadd_rn(__fmul_rn(s.y, scale), bias);
  const float d = __fsub_rn(x, mu);
  const float var = __fmul_rn(s.sigma, s.sigma);
  const float logn = __fsub_rn(__fsub_rn(-__fdiv_rn(__fmul_rn(d, d), __fmul_rn(2.0f, var)), logf(s.sigma)), LOG_SQRT_2PI);
  const float corr = logf(__fadd_rn(__fmul_rn(scale, __fsub_rn(1.0f, __fmul_rn(s.y, s.y))), 1e-6f));
  s.logp = __fsub_rn(logn, corr);
  return s;
}

// Backward of  L = sum_a [ ga*action + c_pi*logp ]  w.r.t. (mu, ls_raw).
// With x = mu + eps*sigma the Normal term is -eps^2/2 - log(sigma) - const, so dlogN/dmu = 0 and
// dlogN/dsigma = -1/sigma (autograd reaches the same values as a two-path sum whose residual is
// fp32 noise); the tanh correction and the action path go through y = tanh(x).
__device__ __forceinline__ void gauss_backward(const GaussSample& s, float eps, float scale, float ga, float c_pi,
                                               float& g_mu, float& g_ls_raw) {
  const float one_m_y2 = 1.0f - s.y * s.y;
  const float dcorr_dy = (-2.0f * s.y * scale) / (scale * one_m_y2 + 1e-6f);
  const float gy = ga * scale - c_pi * dcorr_dy;
  const float gx = gy * one_m_y2;
  g_mu = gx;
  const float g_sigma = gx * eps - c_pi / s.sigma;
  const float g_ls = g_sigma * s.sigma;
  g_ls_raw = g_ls * (0.5f * (LOG_STD_MAX - LOG_STD_MIN)) * (1.0f - s.th * s.th);
}

__device__ __forceinline__ float td3_action(float u, float scale, float bias, float& t) {
  t = tanhf(u);
  return __fadd_rn(__fmul_rn(t, scale), bias);
}

}  // namespace b2rl
