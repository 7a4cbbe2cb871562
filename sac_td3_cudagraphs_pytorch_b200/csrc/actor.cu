// actor.cu — fused actor step (forward, policy loss, backward through the critics' inputs into the
// actor; agents/agent.py:247-283), the SAC temperature step (agents/agent.py:295-303) and the
// inference policy (agents/agent.py:172-181), 4 batch rows per CTA.
#include <cooperative_groups.h>

#include "mlp_rows.cuh"
#include "policy.cuh"
#include "rng.cuh"

namespace b2rl {

struct ActorSmem {
  float4 x[XMAX];  // [obs | a_pi]
  Acts pi;         // actor activations (kept for its backward pass)
  Acts q;          // this CTA's critic activations
  Scratch s;
  NetStage nsA, nsQ;
  float rowbuf[ROWS * RS_CAP];
  float4 da_peer[MAX_OUT];  // dLoss/da through the peer CTA's critic: [action dim] -> 4 rows
  float4 qv[2], logpi, dq[2];
  float lo[MAX_OUT / 2], hi[MAX_OUT / 2];
  float eps[ROWS][MAX_OUT / 2], sg[ROWS][MAX_OUT / 2], yy[ROWS][MAX_OUT / 2], th[ROWS][MAX_OUT / 2];
};

// SAC: the two CTAs of a cluster share 4 batch rows; CTA k evaluates and differentiates critic k
// (both recompute the identical actor forward), they swap Q_k (one float4) to agree on the arg-min,
// CTA 1 hands its dQ/da to CTA 0 through distributed shared memory and retires; CTA 0 runs the actor's
// backward pass. TD3's loss uses critic 0 only (agent.py:274-275): CTA 1 retires at once.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1)
actor_fused_kernel(const __grid_constant__ b2rl_update_args_t A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ActorSmem& M = *reinterpret_cast<ActorSmem*>(smem_raw);
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int k = (int)cluster.block_rank();
  const int t = threadIdx.x, w = t >> 5, l = t & 31;
  const int agent = blockIdx.y, rb = blockIdx.x >> 1, b0 = rb * ROWS;
  const int O = A.fmt.ob_dim, AD = A.fmt.ac_dim, rs = A.fmt.row_stride, B = A.batch;
  const uint32_t gid = (uint32_t)(A.agent_base + agent);  // global agent id: keys the Philox streams
  const bool td3 = A.hp.td3 != 0;
  if (td3 && k == 1) return;  // the TD3 path never touches the cluster barrier

  const float* P = A.arena + (size_t)agent * A.arena_agent_stride;
  const float* rows = A.rows + (size_t)agent * A.rows_agent_stride;
  const uint64_t step = A.counters[(size_t)agent * 8 + B2RL_CTR_PI];
  float* wsb = A.workspace + (size_t)agent * A.workspace_agent_stride;
  const Workspace ws = ws_carve(wsb, B, 0);
  float* part = ws.part + (size_t)rb * PART_LEN;
  const float alpha = td3 ? 0.f : expf(A.log_alpha[(size_t)agent * 5]);
  const float invB = 1.0f / (float)B;

  // ---- asynchronous burst: rows and small tensors
  stage_rows(rows, rs, b0, M.rowbuf);
  const Net act = stage_net(P, A.actor, M.nsA);
  const Net q = stage_net(P, A.critic[k], M.nsQ);
  if (t < AD) {
    M.lo[t] = __ldg(A.min_ac + t);
    M.hi[t] = __ldg(A.max_ac + t);
  }
  if (!td3 && t < ROWS * AD) {
    const int r = t / AD, a = t - r * AD;
    const int64_t e = ((int64_t)agent * B + b0 + r) * AD + a;
    const float z = noise_at(A.eps, e, A.hp.seed, b0 + r, a, step, gid, STREAM_ACTOR_EPS);
    M.eps[r][a] = z;
    if (A.eps_out && k == 0) A.eps_out[e] = z;
  }
  cp_async_wait_all();
  __syncthreads();

  // ---- actor forward on obs (agent.py:251 / :254-255)
  tile_from_rows(M.rowbuf, rs, 0, O, M.x, 0);
  __syncthreads();
  trunk_fwd(act, M.x, M.pi, M.s, k == 0 ? ws.h1 : nullptr, k == 0 ? ws.h2 : nullptr, b0);
  rowdot(act.w3, act.b3, act.out_dim, M.pi.h2, M.s.u);
  __syncthreads();
  if (w < ROWS) {
    const int r = w;
    float lp = 0.f;
    if (l < AD) {
      const float lo = M.lo[l], hi = M.hi[l];
      const float scale = (hi - lo) * 0.5f, bias = (hi + lo) * 0.5f;
      float act_v;
      if (td3) {
        float th;
        act_v = td3_action(f4get(M.s.u[l], r), scale, bias, th);
        M.th[r][l] = th;
      } else {
        const GaussSample g = gauss_sample(f4get(M.s.u[l], r), f4get(M.s.u[AD + l], r), M.eps[r][l], scale, bias);
        act_v = g.action;
        lp = g.logp;
        M.sg[r][l] = g.sigma; M.yy[r][l] = g.y; M.th[r][l] = g.th;
      }
      reinterpret_cast<float*>(&M.x[O + l])[r] = act_v;
    }
    lp = warp_sum(lp);
    if (l == 0) reinterpret_cast<float*>(&M.logpi)[r] = lp;
  }
  __syncthreads();

  // ---- Q_k(obs, a_pi) with the critic's parameters held constant (agent.py:272-278)
  trunk_fwd(q, M.x, M.q, M.s, nullptr, nullptr, b0);
  rowdot(q.w3, q.b3, 1, M.q.h2, &M.qv[k]);
  __syncthreads();
  if (!td3) {
    if (t == 0) *cluster.map_shared_rank(&M.qv[k], k ^ 1) = M.qv[k];
    cluster.sync();
  }

  // ---- loss and dLoss/dQ_k per row: SAC  mean(alpha*logpi - min_k Q_k), TD3  mean(-Q_0)
  if (t < ROWS) {
    const int r = t;
    const float q0 = f4get(M.qv[0], r);
    float lossr, d0 = -invB, d1 = 0.f;
    if (td3) {
      lossr = -q0;
    } else {
      const float q1 = f4get(M.qv[1], r);
      const bool first = q0 <= q1;  // torch.min(0) returns the first minimal index on ties
      d0 = first ? -invB : 0.f;
      d1 = first ? 0.f : -invB;
      lossr = __fsub_rn(__fmul_rn(alpha, f4get(M.logpi, r)), first ? q0 : q1);
    }
    reinterpret_cast<float*>(&M.dq[0])[r] = d0;
    reinterpret_cast<float*>(&M.dq[1])[r] = d1;
    reinterpret_cast<float*>(&M.s.u[0])[r] = lossr;
  }
  __syncthreads();
  if (t == 0 && k == 0) {
    const float4 lr4 = M.s.u[0], lp4 = M.logpi;
    part[PART_SCAL] = lr4.x + lr4.y + lr4.z + lr4.w;
    part[PART_SCAL + 1] = lp4.x + lp4.y + lp4.z + lp4.w;
  }

  // ---- backward through critic k down to its action inputs
  {
    const float w3 = t < ET ? q.w3[t] : 0.f;
    const float4 dq = M.dq[k];
    const float4 dh2 = make_float4(dq.x * w3, dq.y * w3, dq.z * w3, dq.w * w3);
    trunk_bwd(q, dh2, M.q, M.s, nullptr, nullptr, nullptr, b0);
    rowdot(q.w1t + (size_t)O * HID, nullptr, AD, M.s.d, M.s.u);  // dQ/da_i = sum_j dz1_j * W1[j][O+i]
    __syncthreads();
  }
  if (!td3) {
    if (k == 1 && t < AD) *cluster.map_shared_rank(&M.da_peer[t], 0) = M.s.u[t];
    cluster.sync();
    if (k == 1) return;
  }

  // ---- backward through the action head (CTA 0)
  if (w < ROWS) {
    const int r = w;
    float g_a = 0.f, g_b = 0.f;
    if (l < AD) {
      const float scale = (M.hi[l] - M.lo[l]) * 0.5f;
      float ga = f4get(M.s.u[l], r);                    // through critic 0
      if (!td3) ga += f4get(M.da_peer[l], r);           // + through critic 1
      if (td3) {
        const float th = M.th[r][l];
        g_a = ga * scale * (1.0f - th * th);
      } else {
        GaussSample g;
        g.y = M.yy[r][l]; g.sigma = M.sg[r][l]; g.th = M.th[r][l];
        gauss_backward(g, M.eps[r][l], scale, ga, alpha * invB, g_a, g_b);
      }
      reinterpret_cast<float*>(&M.s.du[l])[r] = g_a;
      ws.dz3[(size_t)(b0 + r) * MAX_OUT + l] = g_a;
      if (!td3) {
        reinterpret_cast<float*>(&M.s.du[AD + l])[r] = g_b;
        ws.dz3[(size_t)(b0 + r) * MAX_OUT + AD + l] = g_b;
      }
    }
  }
  __syncthreads();
  if (t < act.out_dim) {
    const float4 d = M.s.du[t];
    part[PART_DB3 + t] = d.x + d.y + d.z + d.w;
  }
  const float4 dh2 = head_bwd(act.w3, act.out_dim, M.s.du);
  trunk_bwd(act, dh2, M.pi, M.s, ws.dz1, ws.dz2, part, b0);
}

// ---- SAC temperature step -------------------------------------------------------------------
// Every CTA: log-prob of a fresh sample from the UPDATED actor for its 4 rows; the last CTA to
// finish (ticket in counters[4]) sums the per-CTA partials in a fixed order, forms
//   alpha_loss = mean(alpha * (-logpi - targ_ent)),  d/dlog_alpha = same value,
// and applies torch's capturable Adam step to the scalar.
struct AlphaSmem {
  float4 x[XMAX];
  Acts pi;
  Scratch s;
  NetStage ns;
  float4 logpi;
  int last;
};

__device__ __forceinline__ void adam_scalar(float* st /*{p,g,m,v}*/, float g, float lr, float t, float b1, float b2,
                                            float eps) {
  float p = st[0], m = st[2], v = st[3];
  m = m + (1.0f - b1) * (g - m);
  v = v * b2 + (1.0f - b2) * g * g;
  const float bc1 = 1.0f - powf(b1, t), bc2 = 1.0f - powf(b2, t);
  const float ssn = -(lr / bc1);
  const float denom = sqrtf(v) / (sqrtf(bc2) * ssn) + eps / ssn;
  p += m / denom;
  st[0] = p; st[1] = g; st[2] = m; st[3] = v;
}

__global__ void __launch_bounds__(NT, 1) alpha_kernel(const __grid_constant__ b2rl_update_args_t A, float lr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AlphaSmem& M = *reinterpret_cast<AlphaSmem*>(smem_raw);
  const int t = threadIdx.x, w = t >> 5, l = t & 31;
  const int agent = blockIdx.y, rb = blockIdx.x, b0 = rb * ROWS;
  const int O = A.fmt.ob_dim, AD = A.fmt.ac_dim, rs = A.fmt.row_stride, B = A.batch;
  const uint32_t gid = (uint32_t)(A.agent_base + agent);
  const float* P = A.arena + (size_t)agent * A.arena_agent_stride;
  const float* rows = A.rows + (size_t)agent * A.rows_agent_stride;
  uint64_t* ctr = A.counters + (size_t)agent * 8;
  const uint64_t step = ctr[B2RL_CTR_PI];
  float* wsb = A.workspace + (size_t)agent * A.workspace_agent_stride;
  float* part = ws_carve(wsb, B, 0).part;

  const Net act = stage_net(P, A.actor, M.ns);
  load_x(rows, rs, b0, 0, O, M.x, 0);
  cp_async_wait_all();
  __syncthreads();
  trunk_fwd(act, M.x, M.pi, M.s, nullptr, nullptr, b0);
  rowdot(act.w3, act.b3, act.out_dim, M.pi.h2, M.s.u);
  __syncthreads();
  if (w < ROWS) {
    const int r = w;
    float lp = 0.f;
    if (l < AD) {
      const float lo = __ldg(A.min_ac + l), hi = __ldg(A.max_ac + l);
      const int64_t e = ((int64_t)agent * B + b0 + r) * AD + l;
      const float z = noise_at(A.eps2, e, A.hp.seed, b0 + r, l, step, gid, STREAM_ALPHA_EPS);
      if (A.eps2_out) A.eps2_out[e] = z;
      lp = gauss_sample(f4get(M.s.u[l], r), f4get(M.s.u[AD + l], r), z, (hi - lo) * 0.5f, (hi + lo) * 0.5f).logp;
    }
    lp = warp_sum(lp);
    if (l == 0) reinterpret_cast<float*>(&M.logpi)[r] = lp;
  }
  __syncthreads();
  if (t == 0) {
    const float4 lp4 = M.logpi;
    const float te = A.hp.targ_ent;
    // sum_r (-logpi_r - targ_ent), agent.py:300
    part[(size_t)rb * PART_LEN + PART_SCAL + 2] = (-lp4.x - te) + (-lp4.y - te) + (-lp4.z - te) + (-lp4.w - te);
    __threadfence();
    const unsigned long long ticket = atomicAdd((unsigned long long*)&ctr[B2RL_CTR_TICKET], 1ULL);
    M.last = (ticket == (unsigned long long)gridDim.x - 1);
  }
  __syncthreads();
  if (!M.last) return;
  if (w == 0) {  // the last CTA: lane-strided sum of the per-CTA partials, fixed shuffle tree
    __threadfence();
    float s = 0.f;
    for (int i = l; i < (int)gridDim.x; i += 32) s += __ldcg(&part[(size_t)i * PART_LEN + PART_SCAL + 2]);
    s = warp_sum(s);
    if (l == 0) {
      float* st = A.log_alpha + (size_t)agent * 5;
      const float alpha = expf(st[0]);
      const float loss = alpha * (s / (float)B);
      if (lr > 0.f) {
        const uint64_t tstep = ctr[B2RL_CTR_ALPHA] + 1;
        adam_scalar(st, loss, lr, (float)tstep, 0.9f, 0.999f, 1e-8f);
        ctr[B2RL_CTR_ALPHA] = tstep;
      } else {
        st[1] = loss;  // data parallel: local gradient only; b2rl_alpha_adam finishes after the all-reduce
      }
      ctr[B2RL_CTR_TICKET] = 0;  // re-arm the ticket for the next launch / graph replay
      float* out = A.out + (size_t)agent * 8;
      out[B2RL_OUT_ALPHA_LOSS] = loss;
      out[B2RL_OUT_ALPHA] = expf(st[0]);
    }
  }
}

__global__ void alpha_adam_kernel(float* log_alpha, uint64_t* counters, int n_agents, float lr, float grad_scale,
                                  float* out) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_agents) return;
  float* st = log_alpha + (size_t)a * 5;
  uint64_t* ctr = counters + (size_t)a * 8;
  const uint64_t tstep = ctr[B2RL_CTR_ALPHA] + 1;
  adam_scalar(st, st[1] * grad_scale, lr, (float)tstep, 0.9f, 0.999f, 1e-8f);
  ctr[B2RL_CTR_ALPHA] = tstep;
  if (out) {
    out[(size_t)a * 8 + B2RL_OUT_ALPHA_LOSS] = st[1];
    out[(size_t)a * 8 + B2RL_OUT_ALPHA] = expf(st[0]);
  }
}

cudaError_t launch_alpha_adam(float* log_alpha, uint64_t* counters, int n_agents, float lr, float grad_scale,
                              float* out, cudaStream_t st) {
  alpha_adam_kernel<<<(n_agents + 127) / 128, 128, 0, st>>>(log_alpha, counters, n_agents, lr, grad_scale, out);
  return cudaGetLastError();
}

// ---- inference policy ----------------------------------------------------------------------------
struct PredictSmem {
  float4 x[XMAX];
  Acts pi;
  Scratch s;
  NetStage ns;
};

__global__ void __launch_bounds__(NT, 1)
predict_kernel(const __grid_constant__ b2rl_update_args_t A, const float* __restrict__ obs, int n, int mode,
               float explore_std, uint64_t draw, float* __restrict__ actions) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PredictSmem& M = *reinterpret_cast<PredictSmem*>(smem_raw);
  const int t = threadIdx.x, w = t >> 5, l = t & 31;
  const int b0 = blockIdx.x * ROWS;
  const int O = A.fmt.ob_dim, AD = A.fmt.ac_dim;
  const bool td3 = A.hp.td3 != 0;
  const float* P = A.arena;
  const uint64_t step = draw;
  const Net act = stage_net(P, A.actor, M.ns);
  load_x(obs, O, b0, 0, O, M.x, 0, min(ROWS, n - b0));
  cp_async_wait_all();
  __syncthreads();
  trunk_fwd(act, M.x, M.pi, M.s, nullptr, nullptr, b0);
  rowdot(act.w3, act.b3, act.out_dim, M.pi.h2, M.s.u);
  __syncthreads();
  if (w < ROWS && l < AD && b0 + w < n) {
    const int r = w;
    const float lo = __ldg(A.min_ac + l), hi = __ldg(A.max_ac + l);
    const float scale = (hi - lo) * 0.5f, bias = (hi + lo) * 0.5f;
    const int64_t e = (int64_t)(b0 + r) * AD + l;
    float a;
    if (td3) {  // agents/nets.py:149-159
      float th;
      a = td3_action(f4get(M.s.u[l], r), scale, bias, th);
      if (mode == 1) a += noise_at(A.eps, e, ~A.hp.seed, b0 + r, l, step, 0, STREAM_ACTOR_EPS) * (scale * explore_std);
    } else if (mode == 1) {  // sample
      const float z = noise_at(A.eps, e, ~A.hp.seed, b0 + r, l, step, 0, STREAM_ACTOR_EPS);  // key != learner's
      a = gauss_sample(f4get(M.s.u[l], r), f4get(M.s.u[AD + l], r), z, scale, bias).action;
    } else {  // mode = tanh(mean)*scale + bias, agents/nets.py:233
      a = __fadd_rn(__fmul_rn(tanhf(f4get(M.s.u[l], r)), scale), bias);
    }
    actions[e] = a;
  }
}

cudaError_t init_actor() {
  cudaError_t e = cudaFuncSetAttribute(actor_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(ActorSmem));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(alpha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AlphaSmem));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PredictSmem));
  if (e == cudaSuccess) {
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, alpha_adam_kernel);
  }
  return e;
}

cudaError_t launch_actor_fused(const b2rl_update_args_t& a, cudaStream_t st) {
  dim3 grid(2 * (a.batch / ROWS), a.n_agents);  // clusters of 2 along x: (row block, critic)
  actor_fused_kernel<<<grid, NT, sizeof(ActorSmem), st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_alpha(const b2rl_update_args_t& a, float lr, cudaStream_t st) {
  dim3 grid(a.batch / ROWS, a.n_agents);
  alpha_kernel<<<grid, NT, sizeof(AlphaSmem), st>>>(a, lr);
  return cudaGetLastError();
}

cudaError_t launch_predict(const b2rl_update_args_t& a, const float* obs, int n, int mode, float explore_std,
                           uint64_t draw, float* actions, cudaStream_t st) {
  predict_kernel<<<(n + ROWS - 1) / ROWS, NT, sizeof(PredictSmem), st>>>(a, obs, n, mode, explore_std, draw, actions);
  return cudaGetLastError();
}

}  // namespace b2rl
