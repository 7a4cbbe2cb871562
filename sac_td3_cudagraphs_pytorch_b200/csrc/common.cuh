// common.cuh — shared device helpers for libb2rl (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b2rl.h"

namespace b2rl {

constexpr int HID = B2RL_HID;    // 256
constexpr int ROWS = B2RL_ROWS;  // batch rows per CTA
constexpr int NT = 512;          // threads per CTA in the fused kernels: 16 warps feed the FFMA pipes
constexpr int NW = NT / 32;
constexpr int ET = HID;          // "epilogue threads": thread t < ET <-> hidden unit t (warps 0..7)
constexpr int EW = ET / 32;
constexpr int KSPLIT = 8;        // GEMM k-slices (x 2 column halves = 16 warps)
constexpr int MAX_OUT = B2RL_MAX_OUT;
constexpr float LN_EPS = 1e-5f;  // torch.nn.LayerNorm default (agents/nets.py:70)

static_assert(ROWS == 4, "row tiles are stored as float4 (one float per batch row)");

// workspace layout per agent (floats); `slot` = 0/1 for the twin critics, 0 for the actor
//   H1, H2, DZ1, DZ2 : [2][B][256]   DZ3 : [2][B][MAX_OUT]   PART : [2][B/ROWS][PART_LEN]
constexpr int PART_VEC = 6;                               // db1 dg1 dbe1 db2 dg2 dbe2
constexpr int PART_LEN = PART_VEC * HID + MAX_OUT + 8;    // + db3[MAX_OUT] + {loss terms}
constexpr int PART_DB3 = PART_VEC * HID;
constexpr int PART_SCAL = PART_VEC * HID + MAX_OUT;       // [0] sum sq err / actor loss, [1] logpi sum, ...

struct Workspace {
  float *h1, *h2, *dz1, *dz2, *dz3, *part;
};
__host__ __device__ inline int64_t ws_floats(int B) {
  return (int64_t)2 * B * HID * 4 + (int64_t)2 * B * MAX_OUT + (int64_t)2 * (B / ROWS) * PART_LEN;
}
__host__ __device__ inline Workspace ws_carve(float* base, int B, int slot) {
  Workspace w;
  const int64_t bh = (int64_t)B * HID;
  w.h1 = base + (0 * 2 + slot) * bh;
  w.h2 = base + (1 * 2 + slot) * bh;
  w.dz1 = base + (2 * 2 + slot) * bh;
  w.dz2 = base + (3 * 2 + slot) * bh;
  float* p = base + 8 * bh;
  w.dz3 = p + (int64_t)slot * B * MAX_OUT;
  p += (int64_t)2 * B * MAX_OUT;
  w.part = p + (int64_t)slot * (B / ROWS) * PART_LEN;
  return w;
}

// ---- optional in-kernel phase timing (build with B2RL_EXTRA_NVCC_FLAGS=-DB2RL_TIMING; tools/phase_timing.py)
#ifdef B2RL_TIMING
static __device__ long long g_b2rl_timing[64];  // one copy per translation unit
#define B2RL_TICK(slot)                                                                    \
  do {                                                                                     \
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_b2rl_timing[slot] = clock64(); \
  } while (0)
#else
#define B2RL_TICK(slot) do {} while (0)
#endif

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// streaming 128-bit load/store that do not allocate in L1 (replay rows are touched once)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// barrier over the ET epilogue threads only (named barrier 1; barrier 0 is __syncthreads)
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory"); }

// Sum of v[0..3] over the ET epilogue threads, result broadcast to each of them; fixed order =>
// deterministic. `buf` is EW float4 of shared memory; callers alternate between two buffers so that
// one barrier per call is enough. Must be called by exactly the threads t < ET.
__device__ __forceinline__ float4 block_sum4(float4 v, float4* buf) {
  v.x = warp_sum(v.x);
  v.y = warp_sum(v.y);
  v.z = warp_sum(v.z);
  v.w = warp_sum(v.w);
  if ((threadIdx.x & 31) == 0) buf[threadIdx.x >> 5] = v;
  epi_sync();
  float4 s = buf[0];
#pragma unroll
  for (int w = 1; w < EW; ++w) {
    const float4 t = buf[w];
    s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
  }
  return s;
}

__device__ __forceinline__ float f4get(const float4& v, int i) {
  return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

}  // namespace b2rl
