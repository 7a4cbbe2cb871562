// wgrad.cu — weight gradients: the contraction over the batch that the row-tiled fused kernels
// cannot do locally. For every network with trainable parameters in this step:
//   d w1t[k][j] = sum_b X0[b][k] * dZ1[b][j]      (X0 = the rows' [obs|act] or [obs] prefix)
//   d w2t[k][j] = sum_b H1[b][k] * dZ2[b][j]      (+ the transposed copy into the w2n shadow)
//   d w3 [o][k] = sum_b dZ3[b][o] * H2[b][k]
// plus the column sums (bias / LayerNorm affine gradients, per-CTA partials -> totals), the scalar
// losses, and the optimizer's step counter bump. One launch; every output element is produced by
// exactly one CTA with a fixed summation order (no atomics): bitwise reproducible.
// Replaces the weight/bias/LayerNorm-parameter part of `loss.backward()` (agents/agent.py:235,283).
//
// Fused optimizer (b2rl_*_update_opt): the thread that has just produced a gradient element also applies Adam
// to that parameter (and the Polyak average to its target) — `optimizer.step()` (agents/agent.py:236,286) and
// `update_targ_nets` (:320-331) without a launch of their own and without re-reading the gradients. Same
// arithmetic as adam.cu (adam_math.cuh), hence bitwise-equal results. The step counter the bias corrections need
// is bumped by the LAST CTA TO FINISH (ticket), i.e. after every CTA has read it.
#include "adam_math.cuh"

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

struct OptCtx {  // the fused optimizer, per CTA (shared memory)
  int adam, polyak;
  AdamScalars k;
  float beta2, omb1, omb2, eps, pk;
  int64_t region;  // floats between arena regions; gradients are region 4: P = G - 4*region, T = G - 3*region, ...
};
// Adam (+ Polyak) on the 4 consecutive parameters (16-byte aligned) / the one parameter whose gradients were just
// written at gptr. __noinline__: ONE copy of the arithmetic for all call sites — this kernel runs for a few
// microseconds, of which instruction fetch is a visible part (the first fused version, with this inlined at four
// sites, doubled the kernel's SASS and cost 2 us). The loads are issued together (one L2 round trip).
static __device__ __noinline__ void opt_apply4(const OptCtx* ocp, float* gptr) {
  const OptCtx& oc = *ocp;
  float* P = gptr - 4 * oc.region;
  float4 p = *reinterpret_cast<const float4*>(P);
  float4 tg = make_float4(0.f, 0.f, 0.f, 0.f);
  if (oc.polyak) tg = *reinterpret_cast<const float4*>(P + oc.region);
  if (oc.adam) {
    const float4 g = *reinterpret_cast<const float4*>(gptr);
    float4 m = *reinterpret_cast<const float4*>(P + 2 * oc.region);
    float4 v = *reinterpret_cast<const float4*>(P + 3 * oc.region);
    adam_elem(p.x, g.x, m.x, v.x, oc.k, oc.beta2, oc.omb1, oc.omb2, oc.eps);
    adam_elem(p.y, g.y, m.y, v.y, oc.k, oc.beta2, oc.omb1, oc.omb2, oc.eps);
    adam_elem(p.z, g.z, m.z, v.z, oc.k, oc.beta2, oc.omb1, oc.omb2, oc.eps);
    adam_elem(p.w, g.w, m.w, v.w, oc.k, oc.beta2, oc.omb1, oc.omb2, oc.eps);
    *reinterpret_cast<float4*>(P) = p;
    *reinterpret_cast<float4*>(P + 2 * oc.region) = m;
    *reinterpret_cast<float4*>(P + 3 * oc.region) = v;
  }
  if (oc.polyak) {
    tg.x = polyak_elem(tg.x, p.x, oc.pk); tg.y = polyak_elem(tg.y, p.y, oc.pk);
    tg.z = polyak_elem(tg.z, p.z, oc.pk); tg.w = polyak_elem(tg.w, p.w, oc.pk);
    *reinterpret_cast<float4*>(P + oc.region) = tg;
  }
}
static __device__ __noinline__ void opt_apply1(const OptCtx* ocp, float* gptr) {
  const OptCtx& oc = *ocp;
  float* P = gptr - 4 * oc.region;
  float p = *P;
  float tg = oc.polyak ? P[oc.region] : 0.f;
  if (oc.adam) {
    float m = P[2 * oc.region], v = P[3 * oc.region];
    adam_elem(p, *gptr, m, v, oc.k, oc.beta2, oc.omb1, oc.omb2, oc.eps);
    *P = p; P[2 * oc.region] = m; P[3 * oc.region] = v;
  }
  if (oc.polyak) P[oc.region] = polyak_elem(tg, p, oc.pk);
}

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
template <typename Setup>
__device__ void gemm_tile(const GemmJob& J, int B, int m0, int n0, WgradSmem& S, const OptCtx* oc, Setup setup) {
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
  setup();          // (thread 0: optimizer scalars + ticket; its counter load was issued at kernel start)
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
    float* gp = J.C + (size_t)(m0 + i) * HID + n0 + j0;
    const float4 g4 = make_float4(s[0], s[1], s[2], s[3]);
    *reinterpret_cast<float4*>(gp) = g4;
    if (oc) opt_apply4(oc, gp);
  }
  if (J.Ct) {  // transposed copy: stage the reduced tile in red[0] and read it column-wise
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) S.red[0][i][j0 + j] = s[j];
    __syncthreads();
    const int jj = t >> 3, i0 = (t & 7) * 4;  // output row n0+jj, columns m0+i0..+3
    if (m0 + i0 + 3 < J.M) {
      float* gp = J.Ct + (size_t)(n0 + jj) * J.M + m0 + i0;
      const float4 g4 = make_float4(S.red[0][i0][jj], S.red[0][i0 + 1][jj], S.red[0][i0 + 2][jj], S.red[0][i0 + 3][jj]);
      *reinterpret_cast<float4*>(gp) = g4;
      if (oc) opt_apply4(oc, gp);  // the shadow is a parameter in its own right: same gradients, same arithmetic
    }
  }
}

constexpr int XTRA_CTAS = 16;  // trailing CTAs for the optimizer's extra (Polyak-only) segments

struct WgradArgs {
  b2rl_update_args_t u;
  int actor_step;  // 0: the two critics are trained (slots 0,1); 1: the actor (slot 0)
  int bump_counter;
  int has_opt;     // opt.seg[0]: Adam (+ Polyak) on the trained nets; opt.seg[1..]: Polyak-only spans
  int work_ctas;   // CTAs before the trailing extra-segment CTAs
  int skip_vec;    // the vector / scalar reductions are done by the caller (wide path)
  b2rl_adam_args_t opt;
};

__device__ __forceinline__ const b2rl_net_t& net_of(const WgradArgs& W, int n) {
  return W.actor_step ? W.u.actor : W.u.critic[n];
}

__host__ __device__ inline int tiles_of_net(const b2rl_net_t& n) {
  const int nt = HID / TN;
  return ((n.in_dim + TM - 1) / TM) * nt + (HID / TM) * nt + ((n.out_dim + TM - 1) / TM) * nt;
}
constexpr int VEC_CTAS = PART_VEC + 1;  // 6 column vectors + {db3, scalars}

// Every CTA runs this once (thread 0), after its bulk loads are in flight: the optimizer scalars from the step
// counter (read at kernel start), then the ticket — the last CTA to TAKE ITS TICKET has, like all the others,
// already read the counter, so it bumps it (nothing in this kernel reads it again; the kernels that consumed the
// old value for their noise streams ran earlier in the stream) and re-arms the ticket. The atomic is only ISSUED
// here; its result is looked at in ticket_finish, the CTA's last action, so that neither its L2 round trip nor the
// serialisation of ~300 same-address atomics sits on the CTA's critical path.
static __device__ __noinline__ unsigned long long setup_and_ticket(const WgradArgs& W, int agent, uint64_t cnt, OptCtx& oc_s) {
  uint64_t* ctr = W.u.counters + (size_t)agent * 8;
  if (W.has_opt) {
    const b2rl_seg_t& s0 = W.opt.seg[0];
    OptCtx o;
    o.adam = s0.do_adam; o.polyak = s0.do_polyak;
    o.k = adam_scalars(s0.lr, W.opt.beta1, W.opt.beta2, (float)(cnt + 1ULL));  // step count AFTER this step's bump
    adam_finish(o.k, W.opt.eps);
    o.beta2 = W.opt.beta2; o.omb1 = 1.0f - W.opt.beta1; o.omb2 = 1.0f - W.opt.beta2; o.eps = W.opt.eps;
    o.pk = W.opt.polyak; o.region = W.u.region_stride;
    oc_s = o;
  }
  if (W.bump_counter < 0) return 0ULL;
  __threadfence();
  return atomicAdd((unsigned long long*)&ctr[B2RL_CTR_TICKET], 1ULL + (cnt & 0ULL));
}
template <bool OPT>
__device__ __forceinline__ void ticket_finish(const WgradArgs& W, int agent, uint64_t cnt, unsigned long long ticket) {
  if (!OPT) return;
  if (threadIdx.x != 0 || W.bump_counter < 0 || ticket != (unsigned long long)gridDim.x - 1) return;
  uint64_t* ctr = W.u.counters + (size_t)agent * 8;
  ctr[W.bump_counter] = cnt + 1ULL;
  ctr[B2RL_CTR_TICKET] = 0;
}

// OPT: with the fused optimizer (two instantiations: the plain one carries none of the optimizer's code)
template <bool OPT>
__global__ void __launch_bounds__(WT) wgrad_kernel(const __grid_constant__ WgradArgs W) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char wsm_raw[];
  WgradSmem& S = *reinterpret_cast<WgradSmem*>(wsm_raw);
  __shared__ OptCtx oc_s;
  const b2rl_update_args_t& A = W.u;
  const int agent = blockIdx.y, t = threadIdx.x;
  const int B = A.batch, nblk = row_blocks(B);
  const int n_nets = W.actor_step ? 1 : 2;
  float* arena = A.arena + (size_t)agent * A.arena_agent_stride;
  float* G = arena + 4 * A.region_stride;  // region 4: gradients
  const float* rows = A.rows + (size_t)agent * A.rows_agent_stride;
  float* wsb = A.workspace + (size_t)agent * A.workspace_agent_stride;
  const int nt = HID / TN;
  const OptCtx* oc = OPT ? &oc_s : nullptr;
  // the step counter this kernel bumps is also the one Adam's bias corrections are taken from
  const int cidx = W.bump_counter >= 0 ? W.bump_counter : (W.has_opt ? W.opt.seg[0].counter : 0);
  const uint64_t cnt = (OPT && t == 0) ? A.counters[(size_t)agent * 8 + cidx] : 0ULL;
  unsigned long long ticket = 0ULL;
  // (plain instantiation: nothing in this kernel reads the counter, so one extra CTA bumps it — no ticket)
  auto setup = [&]() { if (OPT && t == 0) ticket = setup_and_ticket(W, agent, cnt, oc_s); };

  int id = blockIdx.x;
  if (OPT && id >= W.work_ctas) {  // trailing CTAs: the optimizer's extra Polyak-only spans (e.g. TD3's actor target in an
    const int64_t stride = (int64_t)XTRA_CTAS * WT * 4;  // iteration without actor update)
    float* P = arena;
    float* T = arena + A.region_stride;
    for (int si = 1; si < W.opt.n_seg; ++si) {
      const b2rl_seg_t& sg = W.opt.seg[si];
      for (int64_t i = sg.begin + ((int64_t)(id - W.work_ctas) * WT + t) * 4; i < sg.end; i += stride) {
        const float4 p = *reinterpret_cast<const float4*>(P + i);
        float4 tg = *reinterpret_cast<const float4*>(T + i);
        tg.x = polyak_elem(tg.x, p.x, W.opt.polyak); tg.y = polyak_elem(tg.y, p.y, W.opt.polyak);
        tg.z = polyak_elem(tg.z, p.z, W.opt.polyak); tg.w = polyak_elem(tg.w, p.w, W.opt.polyak);
        *reinterpret_cast<float4*>(T + i) = tg;
      }
    }
    setup();
    ticket_finish<OPT>(W, agent, cnt, ticket);
    return;
  }
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
      gemm_tile(J, B, (id / nt) * TM, (id % nt) * TN, S, oc, setup);
      ticket_finish<OPT>(W, agent, cnt, ticket);
      return;
    }
    id -= t1 + t2 + t3;
    if (id < VEC_CTAS) {
      if (W.skip_vec) return;
      setup();
      __syncthreads();
      if (id < PART_VEC) {  // one 256-wide column vector: sum the per-row-block partials
        if (net.layer_norm || (id % 3) == 0) {
          const int64_t off = id == 0 ? net.b1 : id == 1 ? net.g1 : id == 2 ? net.be1 : id == 3 ? net.b2 : id == 4 ? net.g2 : net.be2;
          const float* p = ws.part + (size_t)id * HID + t;
          float s = 0.f;
#pragma unroll 16
          for (int i = 0; i < nblk; ++i) s += p[(size_t)i * PART_LEN];  // loads are independent: 16 in flight
          G[off + t] = s;
          if (oc) opt_apply1(oc, G + off + t);
        }
      } else {  // head bias gradient and the scalar outputs
        if (t < net.out_dim) {
          float s = 0.f;
#pragma unroll 16
          for (int i = 0; i < nblk; ++i) s += ws.part[(size_t)i * PART_LEN + PART_DB3 + t];
          G[net.b3 + t] = s;
          if (oc) opt_apply1(oc, G + net.b3 + t);
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
      ticket_finish<OPT>(W, agent, cnt, ticket);
      return;
    }
    id -= VEC_CTAS;
  }
  if (!OPT) {  // the CTA after the last working one: bump the optimizer's step counter (the kernels that consumed the
    // old value for their noise streams ran earlier in the stream; Adam, later, reads the new one)
    if (id == 0 && t == 0 && W.bump_counter >= 0) A.counters[(size_t)agent * 8 + W.bump_counter] += 1ULL;
    if (id == 0 && !W.actor_step && A.new_rows) {  // the replay write folded into the critic step: every CTA of
      uint64_t* ctr = A.counters + (size_t)agent * 8;  // critic_fused has read the cursor and the size by now, and no
      const int64_t cursor = (int64_t)ctr[B2RL_CTR_CURSOR];  // kernel before the next critic step reads the storage
      const int rs = A.fmt.row_stride, chunks = rs >> 2;
      float* sto = const_cast<float*>(A.storage) + (size_t)agent * A.storage_agent_stride;
      for (int i = t; i < A.n_new * chunks; i += blockDim.x) {
        const int r = i / chunks, c4 = i - r * chunks;
        const int64_t d = (cursor + r) % A.capacity;
        st_stream4(sto + (size_t)d * rs + 4 * c4, ld_stream4(A.new_rows + (size_t)r * rs + 4 * c4));
      }
      __syncthreads();
      if (t == 0) {
        ctr[B2RL_CTR_CURSOR] = (uint64_t)((cursor + A.n_new) % A.capacity);
        const uint64_t size = ctr[B2RL_CTR_SIZE] + (uint64_t)A.n_new;
        ctr[B2RL_CTR_SIZE] = size < (uint64_t)A.capacity ? size : (uint64_t)A.capacity;
      }
    }
    return;
  }
  setup();
  ticket_finish<OPT>(W, agent, cnt, ticket);
}

cudaError_t init_wgrad() {
  cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WgradSmem));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WgradSmem));
  return e;
}

cudaError_t launch_wgrad(const b2rl_update_args_t& a, int actor_step, int bump_counter, const b2rl_adam_args_t* opt,
                         cudaStream_t st, int skip_vec) {
  WgradArgs W;
  W.skip_vec = skip_vec;
  W.u = a;
  W.actor_step = actor_step;
  W.bump_counter = bump_counter;
  W.has_opt = opt != nullptr;
  if (opt) W.opt = *opt;
  int ctas = 0;
  const int n_nets = actor_step ? 1 : 2;
  for (int n = 0; n < n_nets; ++n) ctas += tiles_of_net(actor_step ? a.actor : a.critic[n]) + VEC_CTAS;
  W.work_ctas = ctas;
  if (opt && opt->n_seg > 1) ctas += XTRA_CTAS;
  if (!opt) ctas += 1;  // the counter-bump CTA of the plain kernel
  dim3 grid(ctas, a.n_agents);
  if (opt) return launch_k(wgrad_kernel<true>, grid, dim3(WT), -1, sizeof(WgradSmem), st, W);
  return launch_k(wgrad_kernel<false>, grid, dim3(WT), -1, sizeof(WgradSmem), st, W);
}

}  // namespace b2rl
