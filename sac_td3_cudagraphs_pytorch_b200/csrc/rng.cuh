// rng.cuh — counter-based RNG for the device-side draws (replay indices, SAC/TD3 noise).
// Philox-4x32-10 (Salmon et al., SC'11); the numpy twin used by the tests is
// oracle/sac_td3_oracle.py::philox4x32_10. Counter = (row, block, step, agent<<2 | stream),
// key = seed, so draws are independent of grid shape and of how agents are sharded over GPUs.
#pragma once
#include <stdint.h>

namespace b2rl {

enum : uint32_t { STREAM_INDEX = 0, STREAM_CRITIC_EPS = 1, STREAM_ACTOR_EPS = 2, STREAM_ALPHA_EPS = 3 };

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ uint4 philox_block(uint64_t seed, uint32_t row, uint32_t blk, uint64_t step,
                                              uint32_t agent, uint32_t stream) {
  return philox4x32_10(make_uint4(row, blk, (uint32_t)step, (agent << 2) | stream),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

// uniform index in [0, n): multiply-shift (n < 2^32)
__device__ __forceinline__ int64_t philox_index(uint64_t seed, uint32_t row, uint64_t step, uint32_t agent,
                                                uint64_t n) {
  const uint4 r = philox_block(seed, row, 0u, step, agent, STREAM_INDEX);
  return (int64_t)(((uint64_t)r.x * n) >> 32);
}

// four N(0,1) from one Philox block (Box-Muller, two pairs)
__device__ __forceinline__ float4 box_muller4(uint4 r) {
  const float k = 2.3283064365386963e-10f;  // 2^-32
  const float u1a = ((float)r.x + 1.0f) * k, u2a = (float)r.y * k;
  const float u1b = ((float)r.z + 1.0f) * k, u2b = (float)r.w * k;
  // hardware log2 / sin / cos (MUFU): the draws feed exploration noise, whose distribution a few ulp do not
  // change, and the accurate libm paths were a third of the fused kernels' prologue (and of their code size)
  const float ra = sqrtf(-2.0f * __logf(fminf(u1a, 1.0f))), rb = sqrtf(-2.0f * __logf(fminf(u1b, 1.0f)));
  float sa, ca, sb, cb;
  __sincosf(6.283185307179586f * u2a, &sa, &ca);
  __sincosf(6.283185307179586f * u2b, &sb, &cb);
  return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}

// noise element (row, a): from `eps` if given, else Philox(step, stream)
__device__ __forceinline__ float noise_at(const float* __restrict__ eps, int64_t elem, uint64_t seed,
                                          uint32_t row, int a, uint64_t step, uint32_t agent, uint32_t stream) {
  if (eps) return eps[elem];
  const float4 z = box_muller4(philox_block(seed, row, (uint32_t)(a >> 2), step, agent, stream));
  const int i = a & 3;
  return i == 0 ? z.x : (i == 1 ? z.y : (i == 2 ? z.z : z.w));
}

// The noise of one 8-row tile, dst[r][a] = noise_at(row b0 + r, action dim a): ONE Philox block (four normals) per lane
// and step instead of one per element — the same values (noise_at takes component a & 3 of block a >> 2), a quarter of
// the Philox / Box-Muller work (Humanoid, 17 action dims: 5 rounds of a lone warp in the prologue become 2).
template <int ROWS, int LD>
__device__ __forceinline__ void tile_noise(float (*dst)[LD], const float* __restrict__ eps, float* __restrict__ eps_out,
                                           uint64_t seed, int64_t row0_elem, int b0, int nvalid, int AD, uint64_t step,
                                           uint32_t agent, uint32_t stream, bool need, int lane) {
  const int nb = (AD + 3) >> 2;
  for (int i = lane; i < ROWS * nb; i += 32) {
    const int r = i / nb, q = i - r * nb;
    const bool on = need && r < nvalid;
    const int64_t e = (row0_elem + r) * AD + 4 * q;
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (on) {
      if (eps) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (4 * q + u < AD) z[u] = eps[e + u];
      } else {
        const float4 v = box_muller4(philox_block(seed, (uint32_t)(b0 + r), (uint32_t)q, step, agent, stream));
        z[0] = v.x, z[1] = v.y, z[2] = v.z, z[3] = v.w;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (4 * q + u < AD) {
        dst[r][4 * q + u] = z[u];
        if (on && eps_out) eps_out[e + u] = z[u];
      }
  }
}


}  // namespace b2rl
