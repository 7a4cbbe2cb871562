// adam.cu — one launch for "Adam step on some spans + Polyak averaging on some spans" of the flat
// parameter arena. Replaces torch's _multi_tensor_adam (~16 foreach launches per optimizer.step(),
// agents/agent.py:236,286) and tensordict's lerp_ (agents/agent.py:328-331).
//
// Arithmetic: torch/optim/adam.py `capturable` branch (:478-527) — the one the reference runs on the
// GPU (agents/agent.py:118):
//   m = m + (1-b1)(g-m);  v = v*b2 + (1-b2) g g;
//   ss = -(lr / (1-b1^t));  denom = sqrt(v) / (sqrt(1-b2^t) * ss) + eps/ss;  p += m/denom
// Polyak (torch.lerp, weight < 0.5):  targ = targ + polyak * (p_new - targ).
// HBM-bound: 28 B/param for Adam, +8 B/param when fused with Polyak (the new p is still in registers),
// 12 B/param for Polyak alone.
#include "adam_math.cuh"

namespace b2rl {

using SegScalars = AdamScalars;

__global__ void __launch_bounds__(256) adam_polyak_kernel(const __grid_constant__ b2rl_adam_args_t A) {
  pdl_enter();
  const int agent = blockIdx.y;
  float* P = A.arena + (size_t)agent * A.arena_agent_stride;
  float* T = P + A.region_stride;
  float* M1 = P + 2 * A.region_stride;
  float* V = P + 3 * A.region_stride;
  float* G = P + 4 * A.region_stride;
  const uint64_t* ctr = A.counters + (size_t)agent * 8;

  __shared__ SegScalars sc[B2RL_MAX_SEG];
  if (threadIdx.x < A.n_seg) {
    const b2rl_seg_t& s = A.seg[threadIdx.x];
    SegScalars v = {0.f, 1.f, 1.f, 1.f, 0.f};
    if (s.do_adam) {
      const float t = (float)ctr[s.counter];  // step count AFTER this step's bump (>= 1)
      v = adam_scalars(s.lr, A.beta1, A.beta2, t);
      adam_finish(v, A.eps);
      v.gscale = s.grad_scale;
      if (s.clip) {  // clip_grad_norm_: g *= min(1, max_norm / (norm + 1e-6))
        const float norm = sqrtf(A.grad_sumsq[agent]) * s.grad_scale;
        v.gscale *= fminf(1.0f, A.clip_norm / (norm + 1e-6f));
      }
    }
    sc[threadIdx.x] = v;
  }
  __syncthreads();

  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int si = 0; si < A.n_seg; ++si) {
    const b2rl_seg_t& s = A.seg[si];
    const SegScalars k = sc[si];
    const float omb1 = 1.0f - A.beta1, omb2 = 1.0f - A.beta2;
    for (int64_t i = s.begin + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < s.end; i += stride) {
      float4 p = *reinterpret_cast<const float4*>(P + i);
      if (s.do_adam) {
        float4 g = *reinterpret_cast<const float4*>(G + i);
        float4 m = *reinterpret_cast<const float4*>(M1 + i);
        float4 v = *reinterpret_cast<const float4*>(V + i);
        float* pp = &p.x; float* gp = &g.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float gg = __fmul_rn(gp[c], k.gscale);
          gp[c] = gg;
          adam_elem(pp[c], gg, mp[c], vp[c], k, A.beta2, omb1, omb2, A.eps);
        }
        // clip_grad_norm_ scales .grad in place (agent.py:284-285): keep region 4 what torch would show
        if (s.clip) *reinterpret_cast<float4*>(G + i) = g;
        *reinterpret_cast<float4*>(P + i) = p;
        *reinterpret_cast<float4*>(M1 + i) = m;
        *reinterpret_cast<float4*>(V + i) = v;
      }
      if (s.do_polyak) {
        float4 tg = *reinterpret_cast<const float4*>(T + i);
        tg.x = polyak_elem(tg.x, p.x, A.polyak);
        tg.y = polyak_elem(tg.y, p.y, A.polyak);
        tg.z = polyak_elem(tg.z, p.z, A.polyak);
        tg.w = polyak_elem(tg.w, p.w, A.polyak);
        *reinterpret_cast<float4*>(T + i) = tg;
      }
    }
  }
}

// sum of squares of a gradient span: stage 1 (per-CTA partials) and stage 2 (fixed-order total)
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* arena, int64_t region_stride,
                                                            int64_t agent_stride, int64_t begin, int64_t end,
                                                            float* scratch) {
  pdl_enter();
  const int agent = blockIdx.y;
  const float* G = arena + (size_t)agent * agent_stride + 4 * region_stride;
  float s = 0.f;
  for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = G[i];
    s = fmaf(g, g, s);
  }
  __shared__ float ws[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tsum = 0.f;
    for (int w = 0; w < 8; ++w) tsum += ws[w];
    scratch[(size_t)agent * gridDim.x + blockIdx.x] = tsum;
  }
}
__global__ void sumsq_final_kernel(const float* scratch, int n_part, float* sumsq) {
  pdl_enter();
  const int agent = blockIdx.x;
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < n_part; ++i) s += scratch[(size_t)agent * n_part + i];
    sumsq[agent] = s;
  }
}

__global__ void bump_kernel(uint64_t* counters, int which, int n_agents) {
  pdl_enter();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < n_agents) counters[(size_t)a * 8 + which] += 1ULL;
}

constexpr int SUMSQ_PARTS = 64;

cudaError_t init_adam() {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, adam_polyak_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, sumsq_partial_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, sumsq_final_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, bump_kernel);
  return e;
}

cudaError_t launch_adam(const b2rl_adam_args_t& a, cudaStream_t st) {
  int64_t total = 0;
  for (int i = 0; i < a.n_seg; ++i) total += a.seg[i].end - a.seg[i].begin;
  int ctas = (int)((total / 4 + 255) / 256);
  if (ctas < 1) ctas = 1;
  if (ctas > 148 * 4) ctas = 148 * 4;  // grid-stride beyond four CTAs per SM
  return launch_k(adam_polyak_kernel, dim3(ctas, a.n_agents), dim3(256), 1, 0, st, a);
}

cudaError_t launch_sumsq(const float* arena, int64_t region_stride, int64_t agent_stride, int64_t begin, int64_t end,
                         int n_agents, float* sumsq, float* scratch, cudaStream_t st) {
  cudaError_t e = launch_k(sumsq_partial_kernel, dim3(SUMSQ_PARTS, n_agents), dim3(256), 1, 0, st, arena, region_stride, agent_stride,
                           begin, end, scratch);
  if (e != cudaSuccess) return e;
  return launch_k(sumsq_final_kernel, dim3(n_agents), dim3(32), 1, 0, st, (const float*)scratch, (int)SUMSQ_PARTS, sumsq);
}

cudaError_t launch_bump(uint64_t* counters, int which, int n_agents, cudaStream_t st) {
  return launch_k(bump_kernel, dim3((n_agents + 127) / 128), dim3(128), 1, 0, st, counters, which, n_agents);
}

}  // namespace b2rl
