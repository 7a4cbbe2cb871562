// mlp_rows.cuh — building blocks of the fused SAC/TD3 kernels: a 2x256 MLP evaluated (and
// differentiated) for a tile of ROWS=4 batch rows by one 512-thread CTA.
//
// Why this shape: at batch 256 the update is a chain of ~20 dependent
// [256x256]x[256xB] products. Batch rows are independent through the whole forward pass and
// through the dX part of the backward pass, so a CTA that owns 4 rows can run every layer of
// every network back to back with no grid-wide synchronisation; only the weight gradients
// (a contraction over the batch) need a second kernel (wgrad.cu). Every weight element is
// used by exactly one thread of the CTA, so weights go global/L2 -> registers with coalesced
// 128-bit loads (no shared-memory staging, which would only add a write+read of the same bytes);
// activations for the 4 rows live in shared memory as one float4 per feature (a broadcast read).
//
// Arithmetic restated from agents/nets.py:66-92 (Linear -> LayerNorm -> ReLU twice, then head).
#pragma once
#include "common.cuh"

namespace b2rl {

struct Net {  // resolved pointers of one network; the small tensors may point into shared memory (NetStage)
  const float *w1t, *b1, *g1, *be1, *w2t, *b2, *g2, *be2, *w3, *b3, *w2n;
  int in_dim, out_dim, ln;
};
__device__ __forceinline__ Net resolve(const float* region, const b2rl_net_t& n) {
  Net r;
  r.w1t = region + n.w1t; r.b1 = region + n.b1; r.g1 = region + n.g1; r.be1 = region + n.be1;
  r.w2t = region + n.w2t; r.b2 = region + n.b2; r.g2 = region + n.g2; r.be2 = region + n.be2;
  r.w3 = region + n.w3;   r.b3 = region + n.b3; r.w2n = region + n.w2n;
  r.in_dim = n.in_dim; r.out_dim = n.out_dim; r.ln = n.layer_norm;
  return r;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory"); }

// Everything of a network that the epilogues and the head touch (biases, LayerNorm affine, head weights), copied
// once into shared memory with cp.async at kernel start: in the first version each epilogue began with an exposed
// L2 round trip for these (profiles/r1_phase_timing_*.txt: 20 % of the kernel was such waits).
constexpr int W3_STAGE = 16;  // head rows staged (1 for a critic, A or 2A for an actor when it fits)
struct NetStage {
  float b1[HID], g1[HID], be1[HID], b2[HID], g2[HID], be2[HID];
  float w3[W3_STAGE * HID];
  float b3[MAX_OUT];
};
// Issue the copies (all threads), return a Net whose small tensors point into `st`. Call cp_async_wait_all() and
// __syncthreads() before the first epilogue.
__device__ __forceinline__ Net stage_net(const float* region, const b2rl_net_t& d, NetStage& st) {
  Net n = resolve(region, d);
  const int t = threadIdx.x;
  const int nv = d.layer_norm ? 6 : 2;
  for (int i = t; i < nv * (HID / 4); i += NT) {  // 64 float4 per vector
    const int v = i / (HID / 4), c = (i % (HID / 4)) * 4;
    const float* src = d.layer_norm ? (v == 0 ? n.b1 : v == 1 ? n.g1 : v == 2 ? n.be1 : v == 3 ? n.b2 : v == 4 ? n.g2 : n.be2)
                                    : (v == 0 ? n.b1 : n.b2);
    float* dst = d.layer_norm ? (v == 0 ? st.b1 : v == 1 ? st.g1 : v == 2 ? st.be1 : v == 3 ? st.b2 : v == 4 ? st.g2 : st.be2)
                              : (v == 0 ? st.b1 : st.b2);
    cp_async16(dst + c, src + c);
  }
  for (int i = t; i < (d.out_dim + 3) / 4; i += NT) cp_async16(st.b3 + 4 * i, n.b3 + 4 * i);
  n.b1 = st.b1; n.b2 = st.b2; n.b3 = st.b3;
  if (d.layer_norm) { n.g1 = st.g1; n.be1 = st.be1; n.g2 = st.g2; n.be2 = st.be2; }
  if (d.out_dim <= W3_STAGE) {
    for (int i = t; i < d.out_dim * (HID / 4); i += NT) cp_async16(st.w3 + 4 * i, n.w3 + 4 * i);
    n.w3 = st.w3;
  }
  return n;
}

struct Acts {  // what one forward pass leaves behind for its backward pass
  float4 h1[HID], h2[HID];    // post-ReLU activations, [feature] -> 4 rows
  float4 xh1[HID], xh2[HID];  // LayerNorm x-hat (or the pre-activation when layer_norm is off)
  float2 st1[ROWS], st2[ROWS];  // per row (mean, rstd) of the two LayerNorms
};

constexpr int XMAX = 1024;  // max input width (O + A)

struct Scratch {
  float red[KSPLIT * ROWS * HID];  // split-K partial sums [k-slice][row][col]
  float z[ROWS][HID];              // pre-activation rows (LayerNorm statistics are taken row-wise by one warp each)
  float z2[ROWS][HID];
  float2 stat[ROWS];
  float4 u[MAX_OUT];               // head outputs / small row-dot results: [output] -> 4 rows
  float4 du[MAX_OUT];
  float4 d[HID];                   // gradient tile fed to the backward GEMM
};

// ---- the GEMM: red[w][r][j] = sum_{k in slice(w)} W[k][j] * x[k].r  -----------------------------
// W is [K][256] row-major (w1t / w2t for the forward pass, w2n for dX). Warp w owns a contiguous
// slice of K; lane l owns columns 4l..4l+3 and 128+4l..128+4l+3, so each warp-level load is a
// fully coalesced 512-byte row segment and each x[k] read is a shared-memory broadcast.
__device__ __forceinline__ void fma16(float (&acc)[4][4], const float4& w, const float4& x) {
  acc[0][0] = fmaf(w.x, x.x, acc[0][0]); acc[0][1] = fmaf(w.x, x.y, acc[0][1]);
  acc[0][2] = fmaf(w.x, x.z, acc[0][2]); acc[0][3] = fmaf(w.x, x.w, acc[0][3]);
  acc[1][0] = fmaf(w.y, x.x, acc[1][0]); acc[1][1] = fmaf(w.y, x.y, acc[1][1]);
  acc[1][2] = fmaf(w.y, x.z, acc[1][2]); acc[1][3] = fmaf(w.y, x.w, acc[1][3]);
  acc[2][0] = fmaf(w.z, x.x, acc[2][0]); acc[2][1] = fmaf(w.z, x.y, acc[2][1]);
  acc[2][2] = fmaf(w.z, x.z, acc[2][2]); acc[2][3] = fmaf(w.z, x.w, acc[2][3]);
  acc[3][0] = fmaf(w.w, x.x, acc[3][0]); acc[3][1] = fmaf(w.w, x.y, acc[3][1]);
  acc[3][2] = fmaf(w.w, x.z, acc[3][2]); acc[3][3] = fmaf(w.w, x.w, acc[3][3]);
}

// (__noinline__: one copy of the hot loop for all ~20 layer evaluations of a fused kernel.)
// 16 warps = 8 k-slices x 2 column halves: warp (s, c) accumulates k in slice s for columns
// [128c, 128c+128), lane l owning columns 128c+4l..+3 for the 4 rows: per k one coalesced LDG.128 of
// weights, one broadcast LDS.128 of x[k], 16 FFMA. Weight loads are software-pipelined one group of 4 k
// ahead through two register sets (no register copies).
#define B2RL_LOADG(buf, kk)                                         \
  _Pragma("unroll") for (int u = 0; u < 4; ++u) buf[u] = ldg4(wp + (size_t)B2RL_K(kk, u) * HID);
#define B2RL_FMAG(buf, kk)                                          \
  _Pragma("unroll") for (int u = 0; u < 4; ++u) fma16(acc, buf[u], B2RL_X(kk, u));

static __device__ __noinline__ void gemm_rows(const float* __restrict__ W, int K,
                                              const float4* __restrict__ x, float* __restrict__ red) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int s = w & (KSPLIT - 1), c = w >> 3;
  const int ks = (K + KSPLIT - 1) / KSPLIT;
  const int k0 = min(K, s * ks), n = min(K, k0 + ks) - k0;
  float acc[4][4];  // [col][row]
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[i][r] = 0.f;

  const float* wp = W + (size_t)k0 * HID + c * 128 + 4 * l;
  const float4* xp = x + k0;
  // groups of 4 k; a ragged last group re-reads the slice's last row with x = 0 (no serial tail loop:
  // for the narrow first layers, K = 11 or 14, a slice IS one ragged group and costs one L2 round trip)
  const int g = (n + 3) >> 2;
  float4 A[4], B[4];
#define B2RL_K(kk, u) min((kk) + (u), n - 1)
#define B2RL_X(kk, u) (((kk) + (u)) < n ? xp[(kk) + (u)] : make_float4(0.f, 0.f, 0.f, 0.f))
  if (g > 0) { B2RL_LOADG(A, 0) }
  int i = 0;
  for (; i + 2 <= g; i += 2) {
    B2RL_LOADG(B, 4 * i + 4)
    B2RL_FMAG(A, 4 * i)
    if (i + 2 < g) { B2RL_LOADG(A, 4 * i + 8) }
    B2RL_FMAG(B, 4 * i + 4)
  }
  if (i < g) { B2RL_FMAG(A, 4 * i) }
#undef B2RL_K
#undef B2RL_X
#pragma unroll
  for (int r = 0; r < 4; ++r)
    *reinterpret_cast<float4*>(red + (s * ROWS + r) * HID + c * 128 + 4 * l) =
        make_float4(acc[0][r], acc[1][r], acc[2][r], acc[3][r]);
}
#undef B2RL_LOADG
#undef B2RL_FMAG

// ---- row statistics: warp r < 4 reduces row r of a [4][256] shared buffer ---------------------------------
__device__ __forceinline__ void row_sums(const float (*za)[HID], const float (*zb)[HID], float2* out, bool layernorm_stats) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (w >= ROWS) return;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = za[w][l + 32 * i];
  float s = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
  s = warp_sum(s);
  if (layernorm_stats) {  // (mean, rstd) with the biased variance taken around the mean (two-pass)
    const float mean = s * (1.0f / HID);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
    q = warp_sum(q);
    if (l == 0) out[w] = make_float2(mean, 1.0f / sqrtf(q * (1.0f / HID) + LN_EPS));
  } else {  // two plain means (LayerNorm backward: mean(dx), mean(dx * xhat))
    float u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] = zb[w][l + 32 * i];
    float s2 = ((u[0] + u[1]) + (u[2] + u[3])) + ((u[4] + u[5]) + (u[6] + u[7]));
    s2 = warp_sum(s2);
    if (l == 0) out[w] = make_float2(s * (1.0f / HID), s2 * (1.0f / HID));
  }
}

// ---- forward epilogue of one layer (all NT threads; three barriers) ----------------------------------------
// E1: thread (j = t & 255, rp = t >> 8) sums the k-slice partials of rows 2rp, 2rp+1 of column j, adds the bias.
// E2: warp r takes LayerNorm statistics of row r.   E3: normalise, affine, ReLU; write h / xhat tiles (+ global H).
__device__ __forceinline__ void layer_fwd_epilogue(const float* __restrict__ b, const float* __restrict__ g,
                                                   const float* __restrict__ be, bool ln, Scratch& S, float4* hT,
                                                   float4* xhT, float2* stat_keep, float* ws_h, int b0) {
  const int t = threadIdx.x, j = t & (HID - 1), rp = t >> 8, r0 = 2 * rp;
  float z0 = S.red[r0 * HID + j], z1 = S.red[(r0 + 1) * HID + j];
#pragma unroll
  for (int w = 1; w < KSPLIT; ++w) {
    z0 += S.red[(w * ROWS + r0) * HID + j];
    z1 += S.red[(w * ROWS + r0 + 1) * HID + j];
  }
  const float bj = b[j];
  z0 += bj; z1 += bj;
  float h0, h1, x0 = z0, x1 = z1;
  if (ln) {
    S.z[r0][j] = z0; S.z[r0 + 1][j] = z1;
    __syncthreads();
    row_sums(S.z, S.z, S.stat, true);
    __syncthreads();
    const float2 s0 = S.stat[r0], s1 = S.stat[r0 + 1];
    if (t < ROWS) stat_keep[t] = S.stat[t];
    const float gj = g[j], bej = be[j];
    x0 = (z0 - s0.x) * s0.y; x1 = (z1 - s1.x) * s1.y;
    h0 = fmaxf(fmaf(x0, gj, bej), 0.f); h1 = fmaxf(fmaf(x1, gj, bej), 0.f);
  } else {
    h0 = fmaxf(z0, 0.f); h1 = fmaxf(z1, 0.f);
  }
  reinterpret_cast<float2*>(&hT[j])[rp] = make_float2(h0, h1);
  reinterpret_cast<float2*>(&xhT[j])[rp] = make_float2(x0, x1);
  if (ws_h) {
    ws_h[(size_t)(b0 + r0) * HID + j] = h0;
    ws_h[(size_t)(b0 + r0 + 1) * HID + j] = h1;
  }
  __syncthreads();
}

// ---- backward epilogue of one layer: ReLU mask, LayerNorm backward (threads t < ET hold all 4 rows of a column) ---
// dh: gradient w.r.t. the post-ReLU activation of column j. Returns dz (w.r.t. the Linear output) and writes this
// CTA's column sums {sum_r dz, sum_r dn*xhat, sum_r dn} (d bias, d ln.weight, d ln.bias) to part[0..2][j].
__device__ __forceinline__ float4 layer_bwd_epilogue(float4 dh, const float4* hT, const float4* xhT, const float2* stat,
                                                     const float* __restrict__ g, bool ln, Scratch& S, float* part3) {
  const int t = threadIdx.x, j = t;
  float4 dn = make_float4(0.f, 0.f, 0.f, 0.f), xh = dn, dx = dn, dz;
  if (t < ET) {
    const float4 h = hT[j];
    xh = xhT[j];
    dn = make_float4(h.x > 0.f ? dh.x : 0.f, h.y > 0.f ? dh.y : 0.f, h.z > 0.f ? dh.z : 0.f, h.w > 0.f ? dh.w : 0.f);
  }
  if (ln) {
    if (t < ET) {
      const float gj = g[j];
      dx = make_float4(dn.x * gj, dn.y * gj, dn.z * gj, dn.w * gj);
      S.z[0][j] = dx.x; S.z[1][j] = dx.y; S.z[2][j] = dx.z; S.z[3][j] = dx.w;
      S.z2[0][j] = dx.x * xh.x; S.z2[1][j] = dx.y * xh.y; S.z2[2][j] = dx.z * xh.z; S.z2[3][j] = dx.w * xh.w;
    }
    __syncthreads();
    row_sums(S.z, S.z2, S.stat, false);
    __syncthreads();
    const float2 c0 = S.stat[0], c1 = S.stat[1], c2 = S.stat[2], c3 = S.stat[3];
    dz.x = stat[0].y * (dx.x - c0.x - xh.x * c0.y);
    dz.y = stat[1].y * (dx.y - c1.x - xh.y * c1.y);
    dz.z = stat[2].y * (dx.z - c2.x - xh.z * c2.y);
    dz.w = stat[3].y * (dx.w - c3.x - xh.w * c3.y);
  } else {
    dz = dn;
  }
  if (part3 && t < ET) {
    part3[0 * HID + j] = dz.x + dz.y + dz.z + dz.w;
    if (ln) {
      part3[1 * HID + j] = dn.x * xh.x + dn.y * xh.y + dn.z * xh.z + dn.w * xh.w;
      part3[2 * HID + j] = dn.x + dn.y + dn.z + dn.w;
    }
  }
  return dz;
}

// thread j < ET: sum the KSPLIT partials of column j (fixed order)
__device__ __forceinline__ float4 reduce_partials(const float* __restrict__ red) {
  const int j = threadIdx.x;
  float z[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float s = red[r * HID + j];
#pragma unroll
    for (int w = 1; w < KSPLIT; ++w) s += red[(w * ROWS + r) * HID + j];
    z[r] = s;
  }
  return make_float4(z[0], z[1], z[2], z[3]);
}

// ---- small products against [n][256] row-major matrices (global or staged in shared memory) -----------------
// out[o] (4 rows) = bias[o] + sum_k W[o][k] * x[k]: warp w takes outputs w, w+NW, ...; lanes stride k.
// Used for the heads (n = 1, A or 2A) and for dQ/da = dz1 . w1t[O+a][:] in the actor step.
static __device__ __noinline__ void rowdot(const float* __restrict__ W, const float* __restrict__ bias, int n,
                                           const float4* __restrict__ x, float4* __restrict__ out) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int o = w; o < n; o += NW) {
    const float bo = (bias && l == 0) ? bias[o] : 0.f;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < HID / 32; ++i) {
      const int k = l + 32 * i;
      const float wv = W[(size_t)o * HID + k];
      const float4 xv = x[k];
      a.x = fmaf(wv, xv.x, a.x); a.y = fmaf(wv, xv.y, a.y); a.z = fmaf(wv, xv.z, a.z); a.w = fmaf(wv, xv.w, a.w);
    }
    a.x = warp_sum(a.x); a.y = warp_sum(a.y); a.z = warp_sum(a.z); a.w = warp_sum(a.w);
    if (l == 0) out[o] = make_float4(a.x + bo, a.y + bo, a.z + bo, a.w + bo);
  }
}

// dh[k = threadIdx.x] (4 rows) = sum_o du[o] * W[o][k]   (head backward)
__device__ __forceinline__ float4 head_bwd(const float* __restrict__ W, int n, const float4* __restrict__ du) {
  const int k = threadIdx.x;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k >= ET) return a;
  for (int o = 0; o < n; ++o) {
    const float wv = W[(size_t)o * HID + k];
    const float4 d = du[o];
    a.x = fmaf(wv, d.x, a.x); a.y = fmaf(wv, d.y, a.y); a.z = fmaf(wv, d.z, a.z); a.w = fmaf(wv, d.w, a.w);
  }
  return a;
}

// ---- input tiles -------------------------------------------------------------------------------------------
// from global rows: x[dst + k] = rows[b0 + r][off + k] for r = 0..3
__device__ __forceinline__ void load_x(const float* __restrict__ rows, int row_stride, int b0, int off, int len,
                                       float4* __restrict__ x, int dst, int n_valid = ROWS) {
  for (int k = threadIdx.x; k < len; k += NT) {
    float v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int rr = r < n_valid ? r : n_valid - 1;  // clamp (predict with n % 4 != 0)
      v[r] = __ldg(rows + (size_t)(b0 + rr) * row_stride + off + k);
    }
    x[dst + k] = make_float4(v[0], v[1], v[2], v[3]);
  }
}
// the CTA's 4 transition rows, copied once into shared memory (cp.async); later tiles are cut out of it
constexpr int RS_CAP = 2056;  // max row_stride (2*(O) + A + 2 with O + A <= 1024)
__device__ __forceinline__ void stage_rows(const float* __restrict__ rows, int row_stride, int b0, float* rowbuf) {
  const int chunks = row_stride >> 2;
  for (int i = threadIdx.x; i < ROWS * chunks; i += NT) {
    const int r = i / chunks, c = (i - r * chunks) * 4;
    cp_async16(rowbuf + r * row_stride + c, rows + (size_t)(b0 + r) * row_stride + c);
  }
}
__device__ __forceinline__ void tile_from_rows(const float* rowbuf, int row_stride, int off, int len, float4* x, int dst) {
  for (int k = threadIdx.x; k < len; k += NT)
    x[dst + k] = make_float4(rowbuf[off + k], rowbuf[row_stride + off + k], rowbuf[2 * row_stride + off + k],
                             rowbuf[3 * row_stride + off + k]);
}

__device__ __forceinline__ void store_rows(float* __restrict__ dst, int b0, float4 v) {  // dst [B][256]
  const int j = threadIdx.x;
  dst[(size_t)(b0 + 0) * HID + j] = v.x;
  dst[(size_t)(b0 + 1) * HID + j] = v.y;
  dst[(size_t)(b0 + 2) * HID + j] = v.z;
  dst[(size_t)(b0 + 3) * HID + j] = v.w;
}

// ---- the two hidden layers, forward. Leaves h1/h2/xh1/xh2 and the LayerNorm statistics in `A`. ---------------
// Ends with a __syncthreads: A.h2 is readable by every thread on return.
static __device__ __noinline__ void trunk_fwd(const Net& n, const float4* __restrict__ x, Acts& A, Scratch& S,
                                              float* ws_h1, float* ws_h2, int b0, int tk = 54) {
  B2RL_TICK(tk + 0);
  gemm_rows(n.w1t, n.in_dim, x, S.red);
  B2RL_TICK(tk + 1);
  __syncthreads();
  B2RL_TICK(tk + 2);
  layer_fwd_epilogue(n.b1, n.g1, n.be1, n.ln, S, A.h1, A.xh1, A.st1, ws_h1, b0);
  B2RL_TICK(tk + 4);
  gemm_rows(n.w2t, HID, A.h1, S.red);
  B2RL_TICK(tk + 5);
  __syncthreads();
  B2RL_TICK(tk + 6);
  layer_fwd_epilogue(n.b2, n.g2, n.be2, n.ln, S, A.h2, A.xh2, A.st2, ws_h2, b0);
  B2RL_TICK(tk + 8);
}

// ---- the two hidden layers, backward (dX path). dh2 = gradient w.r.t. h2 for column j (threads < ET). ------
// Writes dz2/dz1 to the workspace (for wgrad.cu) and this CTA's column partial sums when `part` != NULL.
// On return S.d holds dz1 (synchronised).
static __device__ __noinline__ void trunk_bwd(const Net& n, float4 dh2, const Acts& A, Scratch& S, float* ws_dz1,
                                              float* ws_dz2, float* part, int b0) {
  const int j = threadIdx.x;
  float4 dz = layer_bwd_epilogue(dh2, A.h2, A.xh2, A.st2, n.g2, n.ln, S, part ? part + 3 * HID : nullptr);
  if (j < ET) {
    S.d[j] = dz;
    if (ws_dz2) store_rows(ws_dz2, b0, dz);
  }
  __syncthreads();
  gemm_rows(n.w2n, HID, S.d, S.red);
  __syncthreads();
  float4 dh1 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j < ET) dh1 = reduce_partials(S.red);
  dz = layer_bwd_epilogue(dh1, A.h1, A.xh1, A.st1, n.g1, n.ln, S, part);
  if (j < ET) {
    if (ws_dz1) store_rows(ws_dz1, b0, dz);
    S.d[j] = dz;  // safe: every thread passed a barrier after gemm_rows finished reading S.d
  }
  __syncthreads();
}

}  // namespace b2rl
