// critic.cu — fused critic step: next action -> twin target Q -> TD target -> twin online Q ->
// MSE -> backward (dX path), for 4 batch rows per CTA PAIR. Replaces agents/agent.py:186-235
// (Agent.update_qnets up to qf_loss.backward()); the weight-gradient contraction over the batch
// is wgrad.cu, the optimizer step is adam.cu.
//
// The twin critics are independent except for min(Q'_1, Q'_2) in the TD target, so the two CTAs of a
// thread-block cluster each take one critic (target pass, online pass, backward) for the same 4 rows
// and exchange one float4 (their target Q for the 4 rows) through distributed shared memory. Both
// recompute the (cheap, identical) next-action pass. Critical path: 4 network passes instead of 7,
// on 2*B/4 = 128 SMs instead of 64.
#include <cooperative_groups.h>

#include "mlp_rows.cuh"
#include "policy.cuh"
#include "rng.cuh"

namespace b2rl {

struct CriticSmem {
  float4 x[XMAX];  // [next_obs | a'] for the target pass, then [obs | act]
  Acts a;
  Scratch s;
  NetStage nsA, nsT, nsQ;      // staged small tensors of the actor, target critic k, online critic k
  float rowbuf[ROWS * RS_CAP]; // this CTA's 4 transition rows
  float lo[MAX_OUT / 2], hi[MAX_OUT / 2], eps[ROWS][MAX_OUT / 2];
  float4 logpi, qn[2], y;      // per-row scalars
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1)
critic_fused_kernel(const __grid_constant__ b2rl_update_args_t A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CriticSmem& M = *reinterpret_cast<CriticSmem*>(smem_raw);
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int k = (int)cluster.block_rank();  // the critic this CTA owns
  const int t = threadIdx.x, w = t >> 5, l = t & 31;
  const int agent = blockIdx.y, rb = blockIdx.x >> 1, b0 = rb * ROWS;
  const int O = A.fmt.ob_dim, AD = A.fmt.ac_dim, rs = A.fmt.row_stride, B = A.batch;
  const uint32_t gid = (uint32_t)(A.agent_base + agent);  // global agent id: keys the Philox streams
  const bool td3 = A.hp.td3 != 0;

  const float* P = A.arena + (size_t)agent * A.arena_agent_stride;  // region 0: online
  const float* T = P + A.region_stride;                             // region 1: target
  const float* rows = A.rows + (size_t)agent * A.rows_agent_stride;
  const uint64_t step = A.counters[(size_t)agent * 8 + B2RL_CTR_Q];
  float* wsb = A.workspace + (size_t)agent * A.workspace_agent_stride;

  // ---- one asynchronous burst at kernel start: the 4 rows and every small tensor the kernel will touch
  B2RL_TICK(0);
  stage_rows(rows, rs, b0, M.rowbuf);
  const Net act = stage_net(td3 ? T : P, A.actor, M.nsA);  // SAC samples from the ONLINE actor (agent.py:205),
  const Net qt = stage_net(T, A.critic[k], M.nsT);         // TD3 uses the TARGET actor (agent.py:194-202)
  const Net qo = stage_net(P, A.critic[k], M.nsQ);
  if (t < AD) {
    M.lo[t] = __ldg(A.min_ac + t);
    M.hi[t] = __ldg(A.max_ac + t);
  }
  if (t < ROWS * AD) {
    const int r = t / AD, a = t - r * AD;
    const int64_t e = ((int64_t)agent * B + b0 + r) * AD + a;
    const bool need = !td3 || A.hp.targ_smoothing;
    const float z = need ? noise_at(A.eps, e, A.hp.seed, b0 + r, a, step, gid, STREAM_CRITIC_EPS) : 0.f;
    M.eps[r][a] = z;
    if (need && A.eps_out && k == 0) A.eps_out[e] = z;
  }
  cp_async_wait_all();
  __syncthreads();

  // ---- next action
  tile_from_rows(M.rowbuf, rs, O + AD + 2, O, M.x, 0);
  __syncthreads();
  {
    trunk_fwd(act, M.x, M.a, M.s, nullptr, nullptr, b0, 1);
    rowdot(act.w3, act.b3, act.out_dim, M.a.h2, M.s.u);
    B2RL_TICK(10);
    __syncthreads();
    if (w < ROWS) {  // warp r <-> batch row b0+r, lane <-> action dim
      const int r = w;
      float lp = 0.f;
      if (l < AD) {
        const float lo = M.lo[l], hi = M.hi[l];
        const float scale = (hi - lo) * 0.5f, bias = (hi + lo) * 0.5f;  // agents/nets.py:200-204
        const float z = M.eps[r][l];
        float act_v;
        if (td3) {
          float th;
          act_v = td3_action(f4get(M.s.u[l], r), scale, bias, th);
          if (A.hp.targ_smoothing) {
            float n = __fmul_rn(z, A.hp.td3_std);
            n = fminf(fmaxf(n, -A.hp.td3_c), A.hp.td3_c);
            act_v = fminf(fmaxf(__fadd_rn(act_v, n), lo), hi);
          }
        } else {
          const GaussSample g = gauss_sample(f4get(M.s.u[l], r), f4get(M.s.u[AD + l], r), z, scale, bias);
          act_v = g.action;
          lp = g.logp;
        }
        reinterpret_cast<float*>(&M.x[O + l])[r] = act_v;
      }
      lp = warp_sum(lp);
      if (l == 0) reinterpret_cast<float*>(&M.logpi)[r] = lp;
    }
    __syncthreads();
  }

  // ---- target Q_k on (next_obs, a')  (agent.py:208-210), then swap the 4 values with the peer CTA
  {
    B2RL_TICK(11);
    trunk_fwd(qt, M.x, M.a, M.s, nullptr, nullptr, b0, 12);
    rowdot(qt.w3, qt.b3, 1, M.a.h2, &M.qn[k]);
    B2RL_TICK(21);
    __syncthreads();
    if (t == 0) *cluster.map_shared_rank(&M.qn[k], k ^ 1) = M.qn[k];
    cluster.sync();
    B2RL_TICK(22);
  }

  // ---- TD target (agent.py:212-228)
  if (t < ROWS) {
    const int r = t;
    const float q0 = f4get(M.qn[0], r), q1 = f4get(M.qn[1], r);
    const float qmin = fminf(q0, q1);
    float qp = A.hp.bcq_mix ? __fadd_rn(__fmul_rn(0.75f, qmin), __fmul_rn(0.25f, fmaxf(q0, q1))) : qmin;
    if (!td3) {
      const float alpha = expf(A.log_alpha[(size_t)agent * 5]);
      qp = __fsub_rn(qp, __fmul_rn(alpha, f4get(M.logpi, r)));
    }
    const float rew = M.rowbuf[r * rs + O + AD], done = M.rowbuf[r * rs + O + AD + 1];
    const float y = __fadd_rn(rew, __fmul_rn(__fmul_rn(1.0f - done, A.hp.gamma), qp));
    reinterpret_cast<float*>(&M.y)[r] = y;
    if (A.dbg_targ_q && k == 0) A.dbg_targ_q[(size_t)agent * B + b0 + r] = y;
  }
  tile_from_rows(M.rowbuf, rs, 0, O + AD, M.x, 0);  // [obs | act] is contiguous in the row
  __syncthreads();

  // ---- online Q_k, loss, backward (agent.py:230-235)
  {
    const Workspace ws = ws_carve(wsb, B, k);
    float* part = ws.part + (size_t)rb * PART_LEN;
    B2RL_TICK(23);
    trunk_fwd(qo, M.x, M.a, M.s, ws.h1, ws.h2, b0, 24);
    rowdot(qo.w3, qo.b3, 1, M.a.h2, &M.s.u[0]);
    B2RL_TICK(33);
    __syncthreads();
    const float4 qv = M.s.u[0], yv = M.y;
    const float4 dlt = make_float4(qv.x - yv.x, qv.y - yv.y, qv.z - yv.z, qv.w - yv.w);
    const float sc = 2.0f / (float)B;  // d mean((q-y)^2) / dq
    const float4 dq = make_float4(dlt.x * sc, dlt.y * sc, dlt.z * sc, dlt.w * sc);
    if (t < ROWS) {
      ws.dz3[(size_t)(b0 + t) * MAX_OUT] = f4get(dq, t);
      if (A.dbg_q) A.dbg_q[((size_t)agent * 2 + k) * B + b0 + t] = f4get(qv, t);
    }
    if (t == 0) {
      part[PART_DB3] = dq.x + dq.y + dq.z + dq.w;
      part[PART_SCAL] = dlt.x * dlt.x + dlt.y * dlt.y + dlt.z * dlt.z + dlt.w * dlt.w;
    }
    const float w3 = t < ET ? qo.w3[t] : 0.f;
    const float4 dh2 = make_float4(dq.x * w3, dq.y * w3, dq.z * w3, dq.w * w3);
    B2RL_TICK(34);
    trunk_bwd(qo, dh2, M.a, M.s, ws.dz1, ws.dz2, part, b0);
    B2RL_TICK(35);
  }
}

size_t critic_smem_bytes() { return sizeof(CriticSmem); }

#ifdef B2RL_TIMING
extern "C" int b2rl_debug_timing(long long* host64) {  // critic_fused_kernel's phase timestamps (debug builds)
  return cudaMemcpyFromSymbol(host64, g_b2rl_timing, sizeof(long long) * 64) == cudaSuccess ? 0 : -2;
}
#endif

// loads the kernel (CUDA loads lazily; a first launch inside stream capture would fail) and opts in
// to > 48 KB of dynamic shared memory
cudaError_t init_critic() {
  return cudaFuncSetAttribute(critic_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)sizeof(CriticSmem));
}

cudaError_t launch_critic_fused(const b2rl_update_args_t& a, cudaStream_t st) {
  dim3 grid(2 * (a.batch / ROWS), a.n_agents);  // clusters of 2 along x: (row block, critic)
  critic_fused_kernel<<<grid, NT, sizeof(CriticSmem), st>>>(a);
  return cudaGetLastError();
}

}  // namespace b2rl
