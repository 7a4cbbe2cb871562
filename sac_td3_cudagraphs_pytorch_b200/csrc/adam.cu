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
  float* LO = A.lo ? A.lo + (size_t)agent * A.lo_agent_stride : nullptr;  // 3xTF32 mirror: lo parts of what is written here

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

  const float omb1 = 1.0f - A.beta1, omb2 = 1.0f - A.beta2;
  const int main_ctas = (int)gridDim.x - 64 * A.n_shadow;
  if ((int)blockIdx.x >= main_ctas) {
    // ---- shadow pairs: one 32 x 32 tile of a net's w2t per CTA; the step is computed once, at the w2t position, and
    // the results go to w2t (as they are read: rows of 128 bytes) and, transposed through shared memory, to w2n.
    __shared__ float tr[32][33];
    const int tile = (int)blockIdx.x - main_ctas, pr = tile >> 6, tin = tile & 63;
    const int64_t src = A.shadow_src[pr], dst = A.shadow_dst[pr];
    int si = -1;
    for (int q = 0; q < A.n_seg; ++q)
      if (src >= A.seg[q].begin && src < A.seg[q].end) si = q;
    if (si < 0) return;  // (this launch does not touch that net)
    const b2rl_seg_t& s = A.seg[si];
    const SegScalars k = sc[si];
    const int k0 = (tin >> 3) * 32, j0 = (tin & 7) * 32, r = threadIdx.x >> 3, c4 = (threadIdx.x & 7) * 4;
    const int64_t i = src + (int64_t)(k0 + r) * HID + j0 + c4;        // w2t[k0 + r][j0 + c4 ..]
    const int64_t it = dst + (int64_t)(j0 + r) * HID + k0 + c4;       // w2n[j0 + r][k0 + c4 ..]
    auto put_t = [&](float* base, const float4& x) {  // x (held for w2t[k0 + r][j0 + c4..]) -> base[w2n tile], transposed
      __syncthreads();
      tr[r][c4] = x.x, tr[r][c4 + 1] = x.y, tr[r][c4 + 2] = x.z, tr[r][c4 + 3] = x.w;
      __syncthreads();
      *reinterpret_cast<float4*>(base + it) = make_float4(tr[c4][r], tr[c4 + 1][r], tr[c4 + 2][r], tr[c4 + 3][r]);
    };
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
      if (s.clip) *reinterpret_cast<float4*>(G + i) = g;
      *reinterpret_cast<float4*>(P + i) = p;
      *reinterpret_cast<float4*>(M1 + i) = m;
      *reinterpret_cast<float4*>(V + i) = v;
      put_t(P, p);
      put_t(M1, m);
      put_t(V, v);
      if (LO) {
        const float4 l = tf32_lo4(p);
        *reinterpret_cast<float4*>(LO + i) = l;
        put_t(LO, l);
      }
    }
    if (s.do_polyak) {
      float4 tg = *reinterpret_cast<const float4*>(T + i);
      tg.x = polyak_elem(tg.x, p.x, A.polyak);
      tg.y = polyak_elem(tg.y, p.y, A.polyak);
      tg.z = polyak_elem(tg.z, p.z, A.polyak);
      tg.w = polyak_elem(tg.w, p.w, A.polyak);
      *reinterpret_cast<float4*>(T + i) = tg;
      put_t(T, tg);
      if (LO) {
        const float4 l = tf32_lo4(tg);
        *reinterpret_cast<float4*>(LO + A.region_stride + i) = l;
        put_t(LO + A.region_stride, l);
      }
    }
    return;
  }
  const int64_t stride = (int64_t)main_ctas * blockDim.x * 4;
  for (int si = 0; si < A.n_seg; ++si) {
    const b2rl_seg_t& s = A.seg[si];
    const SegScalars k = sc[si];
    for (int64_t i = s.begin + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < s.end; i += stride) {
      bool paired = false;  // w2t / w2n of a shadow pair: stepped by the tile CTAs above
      for (int q = 0; q < A.n_shadow; ++q)
        paired |= (i >= A.shadow_src[q] && i < A.shadow_src[q] + HID * HID) || (i >= A.shadow_dst[q] && i < A.shadow_dst[q] + HID * HID);
      if (paired) continue;
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
        if (LO) *reinterpret_cast<float4*>(LO + i) = tf32_lo4(p);
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
        if (LO) *reinterpret_cast<float4*>(LO + A.region_stride + i) = tf32_lo4(tg);
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

cudaError_t launch_adam(const b2rl_adam_args_t& a_in, cudaStream_t st) {
  b2rl_adam_args_t a = a_in;  // keep only the shadow pairs this launch touches (w2t inside one of its segments)
  a.n_shadow = 0;
  int64_t total = 0;
  for (int i = 0; i < a.n_seg; ++i) total += a.seg[i].end - a.seg[i].begin;
  for (int q = 0; q < a_in.n_shadow; ++q)
    for (int i = 0; i < a.n_seg; ++i)
      if (a_in.shadow_src[q] >= a.seg[i].begin && a_in.shadow_src[q] < a.seg[i].end) {
        a.shadow_src[a.n_shadow] = a_in.shadow_src[q];
        a.shadow_dst[a.n_shadow++] = a_in.shadow_dst[q];
        total -= 2 * (int64_t)HID * HID;  // (both layouts are stepped by the pair's 64 tile CTAs)
        break;
      }
  int ctas = (int)((total / 4 + 255) / 256);
  if (ctas < 1) ctas = 1;
  if (ctas > 148 * 4) ctas = 148 * 4;  // grid-stride beyond four CTAs per SM
  // + 64 tile CTAs per shadow pair (blockIdx.x >= ctas)
  return launch_k(adam_polyak_kernel, dim3(ctas + 64 * a.n_shadow, a.n_agents), dim3(256), 1, 0, st, a);
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
