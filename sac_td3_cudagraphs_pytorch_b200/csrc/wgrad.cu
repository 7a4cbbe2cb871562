// wgrad.cu — weight gradients: the contraction over the batch that the row-tiled fused kernels
// cannot do locally. For every network with trainable parameters in this step:
//   d w1t[k][j] = sum_b X0[b][k] * dZ1[b][j]      (X0 = the rows' [obs|act] or [obs] prefix)
//   d w2t[k][j] = sum_b H1[b][k] * dZ2[b][j]      (+ the transposed copy into the w2n shadow)
//   d w3 [o][k] = sum_b dZ3[b][o] * H2[b][k]
// plus the column sums (bias / LayerNorm affine gradients, per-CTA partials -> totals), the scalar
// losses, and the optimizer's step counter bump. One launch; every output element is produced by
// exactly one CTA with a fixed summation order (no atomics): bitwise reproducible.
// Replaces the weight/bias/LayerNorm-parameter part of `loss.backward()` (agents/agent.py:235,283).
#include "common.cuh"

namespace b2rl {

constexpr int TM = 32, TN = 32;  // output tile; N (the 256-wide side) is always a multiple of TN
constexpr int WT = 256, WW = WT / 32;  // threads / warps per CTA of this kernel

struct GemmJob {
  const float* A;  // [B][lda], M columns used
  const float* Bm; // [B][ldb], 256 columns
  float* C;        // [M][256]
  float* Ct;       // [256][M] transposed copy or nullptr
  int lda, ldb, M;
};

constexpr int CHUNK = 32;  // batch rows staged per warp at a time
struct WgradSmem {
  union {
    float stage[WW][2][CHUNK][TM];  // per warp: A rows then B rows of the current chunk (8 KB per warp)
    float red[WW][TM][TN + 1];      // after the k loop: per-warp partial tiles
  };
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;  // src-size 0 => the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}

// C[m0+i][n0+j] = sum_b A[b][m0+i] * Bm[b][n0+j]; warp w sums its slice of b, lanes hold a 4x8 sub-tile.
// Each warp first fires ALL global->shared copies of its batch slice (cp.async: no register dependency, so
// the whole slice is one L2 round trip instead of one per row - the first version, with plain loads, was
// bound by exactly that: 32 dependent round trips, 10 us), then multiplies out of shared memory.
__device__ void gemm_tile(const GemmJob& J, int B, int m0, int n0, WgradSmem& S) {
  const int t = threadIdx.x, w = t >> 5, l = t & 31;
  const int mi = (l >> 2) * 4, ni = (l & 3) * 8;
  const int bs = (B + WW - 1) / WW;
  const int bA = min(B, w * bs), bB = min(B, bA + bs);
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float(*As)[TM] = S.stage[w][0];
  float(*Bs)[TM] = S.stage[w][1];
  for (int b0 = bA; b0 < bB; b0 += CHUNK) {
    const int nb = min(CHUNK, bB - b0);
    for (int idx = l; idx < nb * 8; idx += 32) {
      const int row = idx >> 3, seg = (idx & 7) * 4;
      cp_async16(&As[row][seg], J.A + (size_t)(b0 + row) * J.lda + m0 + seg, m0 + seg + 3 < J.lda);
      cp_async16(&Bs[row][seg], J.Bm + (size_t)(b0 + row) * J.ldb + n0 + seg, true);
    }
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
#pragma unroll 4
    for (int row = 0; row < nb; ++row) {
      const float4 a = *reinterpret_cast<const float4*>(&As[row][mi]);
      const float4 x0 = *reinterpret_cast<const float4*>(&Bs[row][ni]);
      const float4 x1 = *reinterpret_cast<const float4*>(&Bs[row][ni + 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], xv[j], acc[i][j]);
    }
    __syncwarp();
  }
  __syncthreads();  // staging area is dead: reuse it for the partial tiles
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) S.red[w][mi + i][ni + j] = acc[i][j];
  __syncthreads();
  // 1024 outputs / 256 threads: thread -> (row i = t/8, 4 consecutive columns)
  const int i = t >> 3, j0 = (t & 7) * 4;
  float s[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v = S.red[0][i][j0 + j];
#pragma unroll
    for (int ww = 1; ww < WW; ++ww) v += S.red[ww][i][j0 + j];
    s[j] = v;
  }
  if (m0 + i < J.M) {
    *reinterpret_cast<float4*>(J.C + (size_t)(m0 + i) * HID + n0 + j0) = make_float4(s[0], s[1], s[2], s[3]);
  }
  if (J.Ct) {  // transposed copy: stage the reduced tile in red[0] and read it column-wise
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) S.red[0][i][j0 + j] = s[j];
    __syncthreads();
    const int jj = t >> 3, i0 = (t & 7) * 4;  // output row n0+jj, columns m0+i0..+3
    if (m0 + i0 + 3 < J.M) {
      *reinterpret_cast<float4*>(J.Ct + (size_t)(n0 + jj) * J.M + m0 + i0) =
          make_float4(S.red[0][i0][jj], S.red[0][i0 + 1][jj], S.red[0][i0 + 2][jj], S.red[0][i0 + 3][jj]);
    }
  }
}

struct WgradArgs {
  b2rl_update_args_t u;
  int actor_step;  // 0: the two critics are trained (slots 0,1); 1: the actor (slot 0)
  int bump_counter;
};

__device__ __forceinline__ const b2rl_net_t& net_of(const WgradArgs& W, int n) {
  return W.actor_step ? W.u.actor : W.u.critic[n];
}

__host__ __device__ inline int tiles_of_net(const b2rl_net_t& n) {
  const int nt = HID / TN;
  return ((n.in_dim + TM - 1) / TM) * nt + (HID / TM) * nt + ((n.out_dim + TM - 1) / TM) * nt;
}
constexpr int VEC_CTAS = PART_VEC + 1;  // 6 column vectors + {db3, scalars}

__global__ void __launch_bounds__(WT) wgrad_kernel(const __grid_constant__ WgradArgs W) {
  extern __shared__ __align__(16) unsigned char wsm_raw[];
  WgradSmem& S = *reinterpret_cast<WgradSmem*>(wsm_raw);
  const b2rl_update_args_t& A = W.u;
  const int agent = blockIdx.y, t = threadIdx.x;
  const int B = A.batch, nblk = row_blocks(B);
  const int n_nets = W.actor_step ? 1 : 2;
  float* arena = A.arena + (size_t)agent * A.arena_agent_stride;
  float* G = arena + 4 * A.region_stride;  // region 4: gradients
  const float* rows = A.rows + (size_t)agent * A.rows_agent_stride;
  float* wsb = A.workspace + (size_t)agent * A.workspace_agent_stride;
  const int nt = HID / TN;

  int id = blockIdx.x;
  for (int n = 0; n < n_nets; ++n) {
    const b2rl_net_t& net = net_of(W, n);
    const Workspace ws = ws_carve(wsb, B, n);
    const int t1 = ((net.in_dim + TM - 1) / TM) * nt, t2 = (HID / TM) * nt, t3 = ((net.out_dim + TM - 1) / TM) * nt;
    if (id < t1 + t2 + t3) {
      GemmJob J;
      if (id < t1) {
        J = {rows, ws.dz1, G + net.w1t, nullptr, A.fmt.row_stride, HID, net.in_dim};
      } else if (id < t1 + t2) {
        id -= t1;
        J = {ws.h1, ws.dz2, G + net.w2t, net.w2n >= 0 ? G + net.w2n : nullptr, HID, HID, HID};
      } else {
        id -= t1 + t2;
        J = {ws.dz3, ws.h2, G + net.w3, nullptr, MAX_OUT, HID, net.out_dim};
      }
      gemm_tile(J, B, (id / nt) * TM, (id % nt) * TN, S);
      return;
    }
    id -= t1 + t2 + t3;
    if (id < VEC_CTAS) {
      if (id < PART_VEC) {  // one 256-wide column vector: sum the per-row-block partials
        if (!net.layer_norm && (id % 3) != 0) return;
        const int64_t off = id == 0 ? net.b1 : id == 1 ? net.g1 : id == 2 ? net.be1 : id == 3 ? net.b2 : id == 4 ? net.g2 : net.be2;
        const float* p = ws.part + (size_t)id * HID + t;
        float s = 0.f;
#pragma unroll 16
        for (int i = 0; i < nblk; ++i) s += p[(size_t)i * PART_LEN];  // loads are independent: 16 in flight
        G[off + t] = s;
      } else {  // head bias gradient and the scalar outputs
        if (t < net.out_dim) {
          float s = 0.f;
#pragma unroll 16
          for (int i = 0; i < nblk; ++i) s += ws.part[(size_t)i * PART_LEN + PART_DB3 + t];
          G[net.b3 + t] = s;
        }
        if ((t >> 5) == 2 && n == n_nets - 1) {  // warp 2: the step's scalar outputs (lane-strided, fixed tree)
          const int lane = t & 31;
          float* out = A.out + (size_t)agent * 8;
          float tot = 0.f, lp = 0.f;
          for (int m = 0; m < n_nets; ++m) {  // sum over critics of the per-critic mean (agent.py:233)
            const float* pm = ws_carve(wsb, B, m).part;
            float s0 = 0.f, s1 = 0.f;
            for (int i = lane; i < nblk; i += 32) {
              s0 += pm[(size_t)i * PART_LEN + PART_SCAL];
              s1 += pm[(size_t)i * PART_LEN + PART_SCAL + 1];
            }
            tot += warp_sum(s0) / (float)B;
            lp += warp_sum(s1);
          }
          if (lane == 0) {
            if (W.actor_step) {
              out[B2RL_OUT_ACTOR_LOSS] = tot;
              out[B2RL_OUT_LOGPI_MEAN] = lp / (float)B;
              if (!A.hp.td3) out[B2RL_OUT_ALPHA] = expf(A.log_alpha[(size_t)agent * 5]);
            } else {
              out[B2RL_OUT_QF_LOSS] = tot;
            }
          }
        }
      }
      return;
    }
    id -= VEC_CTAS;
  }
  // the last CTA of the grid: bump the optimizer's step counter (the kernels that consumed the old
  // value for their noise streams ran earlier in the stream; Adam, later, reads the new one)
  if (id == 0 && t == 0 && W.bump_counter >= 0) A.counters[(size_t)agent * 8 + W.bump_counter] += 1ULL;
}

cudaError_t init_wgrad() {
  return cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WgradSmem));
}

cudaError_t launch_wgrad(const b2rl_update_args_t& a, int actor_step, int bump_counter, cudaStream_t st) {
  WgradArgs W;
  W.u = a;
  W.actor_step = actor_step;
  W.bump_counter = bump_counter;
  int ctas = 1;
  const int n_nets = actor_step ? 1 : 2;
  for (int n = 0; n < n_nets; ++n) ctas += tiles_of_net(actor_step ? a.actor : a.critic[n]) + VEC_CTAS;
  dim3 grid(ctas, a.n_agents);
  wgrad_kernel<<<grid, WT, sizeof(WgradSmem), st>>>(W);
  return cudaGetLastError();
}

}  // namespace b2rl
