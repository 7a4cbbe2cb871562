// adam_math.cuh — the per-element optimizer arithmetic, shared by adam.cu (stand-alone launch) and wgrad.cu (Adam
// fused into the weight-gradient kernel) so that both paths are bitwise identical. torch/optim/adam.py `capturable`
// branch (:478-527), the one the reference runs on the GPU (agents/agent.py:118):
//   m = lerp(m, g, 1-b1);  v = v*b2 + (1-b2) g g;
//   ss = -(lr / (1-b1^t));  denom = sqrt(v) / (sqrt(1-b2^t) * ss) + eps/ss;  p += m/denom
// Polyak (torch.lerp, weight < 0.5):  targ = targ + polyak * (p_new - targ)   (agents/agent.py:328-331).
// Every contraction is written out (fmaf / __fmul_rn) so that the compiler cannot choose differently per call site.
#pragma once
#include "common.cuh"

namespace b2rl {

struct AdamScalars {
  float ssn, bc2s, gscale;
  float c1, c2;  // sqrt(1-b2^t) * ss  and  eps / ss: the per-step constants of denom (set by adam_finish)
};
__device__ __forceinline__ void adam_finish(AdamScalars& k, float eps) {
  k.c1 = __fmul_rn(k.bc2s, k.ssn);
  k.c2 = __fdiv_rn(eps, k.ssn);
}
static __device__ __noinline__ float pow_once(float b, float t) { return powf(b, t); }  // (one copy of powf's code)
__device__ __forceinline__ AdamScalars adam_scalars(float lr, float beta1, float beta2, float t) {
  const float bc1 = 1.0f - pow_once(beta1, t), bc2 = 1.0f - pow_once(beta2, t);
  AdamScalars k;
  k.ssn = -(lr / bc1);
  k.bc2s = sqrtf(bc2);
  k.gscale = 1.0f;
  k.c1 = k.c2 = 0.f;
  return k;
}
// g is the (already scaled) gradient
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamScalars& k, float beta2,
                                          float omb1, float omb2, float eps) {
  m = fmaf(omb1, g - m, m);
  v = fmaf(__fmul_rn(omb2, g), g, __fmul_rn(v, beta2));
  const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), k.c1), k.c2);  // (call adam_finish(k, eps) once per step)
  p = __fadd_rn(p, __fdiv_rn(m, denom));
}
__device__ __forceinline__ float polyak_elem(float tg, float p, float polyak) { return fmaf(polyak, p - tg, tg); }

}  // namespace b2rl
