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

struct Net {  // resolved pointers of one network inside one arena region
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

struct Acts {  // what one forward pass leaves behind for its backward pass
  float4 h1[HID], h2[HID];    // post-ReLU activations, [feature] -> 4 rows
  float4 xh1[HID], xh2[HID];  // LayerNorm x-hat (or the pre-activation when layer_norm is off)
};

constexpr int XMAX = 1024;  // max input width (O + A)

struct Scratch {
  float red[KSPLIT * ROWS * HID];  // split-K partial sums [k-slice][row][col]
  float4 sred[2][EW];              // block_sum4 ping-pong
  float4 u[MAX_OUT];           // head outputs / small row-dot results: [output] -> 4 rows
  float4 du[MAX_OUT];
  float4 d[HID];               // gradient tile fed to the backward GEMM
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

// ---- forward epilogue: bias, LayerNorm (biased variance, eps 1e-5, affine), ReLU -----------------
// Returns h for column j = threadIdx.x; xhat/rstd are what the backward pass needs.
__device__ __forceinline__ float4 fwd_epilogue(float4 z, float bj, float gj, float bej, bool ln, float4 (*sred)[EW],
                                               int& tog, float4& xhat, float4& rstd) {
  z.x += bj; z.y += bj; z.z += bj; z.w += bj;
  float4 n;
  if (ln) {
    const float inv = 1.0f / HID;
    float4 s = block_sum4(z, sred[tog]); tog ^= 1;
    const float4 mean = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
    const float4 d = make_float4(z.x - mean.x, z.y - mean.y, z.z - mean.z, z.w - mean.w);
    s = block_sum4(make_float4(d.x * d.x, d.y * d.y, d.z * d.z, d.w * d.w), sred[tog]); tog ^= 1;
    rstd = make_float4(1.0f / sqrtf(s.x * inv + LN_EPS), 1.0f / sqrtf(s.y * inv + LN_EPS),
                       1.0f / sqrtf(s.z * inv + LN_EPS), 1.0f / sqrtf(s.w * inv + LN_EPS));
    xhat = make_float4(d.x * rstd.x, d.y * rstd.y, d.z * rstd.z, d.w * rstd.w);
    n = make_float4(fmaf(xhat.x, gj, bej), fmaf(xhat.y, gj, bej), fmaf(xhat.z, gj, bej), fmaf(xhat.w, gj, bej));
  } else {
    rstd = make_float4(1.f, 1.f, 1.f, 1.f);
    xhat = z;
    n = z;
  }
  return make_float4(fmaxf(n.x, 0.f), fmaxf(n.y, 0.f), fmaxf(n.z, 0.f), fmaxf(n.w, 0.f));
}

// ---- backward epilogue: ReLU mask, LayerNorm backward ---------------------------------------------
// dh: gradient w.r.t. the post-ReLU activation of column j. Returns dz (gradient w.r.t. the Linear
// output). colsum = {sum_r dz, sum_r dn*xhat, sum_r dn}: this CTA's contribution to d(bias),
// d(ln.weight), d(ln.bias) for column j.
__device__ __forceinline__ float4 bwd_epilogue(float4 dh, float4 h, float4 xhat, float4 rstd, float gj, bool ln,
                                               float4 (*sred)[EW], int& tog, float (&colsum)[3]) {
  const float4 dn = make_float4(h.x > 0.f ? dh.x : 0.f, h.y > 0.f ? dh.y : 0.f, h.z > 0.f ? dh.z : 0.f,
                                h.w > 0.f ? dh.w : 0.f);
  float4 dz;
  if (ln) {
    const float inv = 1.0f / HID;
    const float4 dx = make_float4(dn.x * gj, dn.y * gj, dn.z * gj, dn.w * gj);
    float4 s1 = block_sum4(dx, sred[tog]); tog ^= 1;
    float4 s2 = block_sum4(make_float4(dx.x * xhat.x, dx.y * xhat.y, dx.z * xhat.z, dx.w * xhat.w), sred[tog]);
    tog ^= 1;
    dz.x = rstd.x * (dx.x - s1.x * inv - xhat.x * (s2.x * inv));
    dz.y = rstd.y * (dx.y - s1.y * inv - xhat.y * (s2.y * inv));
    dz.z = rstd.z * (dx.z - s1.z * inv - xhat.z * (s2.z * inv));
    dz.w = rstd.w * (dx.w - s1.w * inv - xhat.w * (s2.w * inv));
    colsum[1] = dn.x * xhat.x + dn.y * xhat.y + dn.z * xhat.z + dn.w * xhat.w;
    colsum[2] = dn.x + dn.y + dn.z + dn.w;
  } else {
    dz = dn;
    colsum[1] = colsum[2] = 0.f;
  }
  colsum[0] = dz.x + dz.y + dz.z + dz.w;
  return dz;
}

// ---- small products against [n][256] row-major matrices ----------------------------------------------
// out[o] (4 rows) = bias[o] + sum_k W[o][k] * x[k]: warp w takes outputs w, w+NW, ...; lanes stride k.
// Used for the heads (n = 1, A or 2A) and for dQ/da = dz1 . w1t[O+a][:] in the actor step.
static __device__ __noinline__ void rowdot(const float* __restrict__ W, const float* __restrict__ bias, int n,
                                       const float4* __restrict__ x, float4* __restrict__ out) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int o = w; o < n; o += NW) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < HID / 32; ++i) {
      const int k = l + 32 * i;
      const float wv = __ldg(W + (size_t)o * HID + k);
      const float4 xv = x[k];
      a.x = fmaf(wv, xv.x, a.x); a.y = fmaf(wv, xv.y, a.y); a.z = fmaf(wv, xv.z, a.z); a.w = fmaf(wv, xv.w, a.w);
    }
    a.x = warp_sum(a.x); a.y = warp_sum(a.y); a.z = warp_sum(a.z); a.w = warp_sum(a.w);
    if (l == 0) {
      const float bo = bias ? __ldg(bias + o) : 0.f;
      out[o] = make_float4(a.x + bo, a.y + bo, a.z + bo, a.w + bo);
    }
  }
}

// dh[k = threadIdx.x] (4 rows) = sum_o du[o] * W[o][k]   (head backward)
__device__ __forceinline__ float4 head_bwd(const float* __restrict__ W, int n, const float4* __restrict__ du) {
  const int k = threadIdx.x;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k >= ET) return a;
  for (int o = 0; o < n; ++o) {
    const float wv = __ldg(W + (size_t)o * HID + k);
    const float4 d = du[o];
    a.x = fmaf(wv, d.x, a.x); a.y = fmaf(wv, d.y, a.y); a.z = fmaf(wv, d.z, a.z); a.w = fmaf(wv, d.w, a.w);
  }
  return a;
}

// ---- input tile: x[dst + k] = rows[b0 + r][off + k] for r = 0..3 --------------------------------------
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

__device__ __forceinline__ void store_rows(float* __restrict__ dst, int b0, float4 v) {  // dst [B][256]
  const int j = threadIdx.x;
  dst[(size_t)(b0 + 0) * HID + j] = v.x;
  dst[(size_t)(b0 + 1) * HID + j] = v.y;
  dst[(size_t)(b0 + 2) * HID + j] = v.z;
  dst[(size_t)(b0 + 3) * HID + j] = v.w;
}

// ---- the two hidden layers, forward. Leaves h1/h2/xh1/xh2 in `A`; rstd1/rstd2 in registers. -----------
// Ends with a __syncthreads: A.h2 is readable by every thread on return.
static __device__ __noinline__ void trunk_fwd(const Net& n, const float4* __restrict__ x, Acts& A, Scratch& S, int& tog,
                                              float4& rstd1, float4& rstd2, float* ws_h1, float* ws_h2, int b0,
                                              int tk = 54) {
  const int j = threadIdx.x;
  // per-column parameters are fetched before the GEMMs so that their L2 round trip is off the epilogues
  float b1 = 0.f, g1 = 1.f, be1 = 0.f, b2 = 0.f, g2 = 1.f, be2 = 0.f;
  if (j < ET) {
    b1 = __ldg(n.b1 + j); b2 = __ldg(n.b2 + j);
    if (n.ln) { g1 = __ldg(n.g1 + j); be1 = __ldg(n.be1 + j); g2 = __ldg(n.g2 + j); be2 = __ldg(n.be2 + j); }
  }
  B2RL_TICK(tk + 0);
  gemm_rows(n.w1t, n.in_dim, x, S.red);
  B2RL_TICK(tk + 1);
  __syncthreads();
  B2RL_TICK(tk + 2);
  if (j < ET) {
    float4 xh;
    const float4 h = fwd_epilogue(reduce_partials(S.red), b1, g1, be1, n.ln, S.sred, tog, xh, rstd1);
    A.h1[j] = h;
    A.xh1[j] = xh;
    if (ws_h1) store_rows(ws_h1, b0, h);
  }
  B2RL_TICK(tk + 3);
  __syncthreads();
  B2RL_TICK(tk + 4);
  gemm_rows(n.w2t, HID, A.h1, S.red);
  B2RL_TICK(tk + 5);
  __syncthreads();
  B2RL_TICK(tk + 6);
  if (j < ET) {
    float4 xh;
    const float4 h = fwd_epilogue(reduce_partials(S.red), b2, g2, be2, n.ln, S.sred, tog, xh, rstd2);
    A.h2[j] = h;
    A.xh2[j] = xh;
    if (ws_h2) store_rows(ws_h2, b0, h);
  }
  B2RL_TICK(tk + 7);
  __syncthreads();
  B2RL_TICK(tk + 8);
}

// ---- the two hidden layers, backward (dX path). dh2 = gradient w.r.t. h2 for column j (threads < ET). ------
// Writes dz2/dz1 to the workspace (for wgrad.cu) and this CTA's column partial sums when `part` != NULL.
// On return S.d holds dz1 (synchronised).
static __device__ __noinline__ void trunk_bwd(const Net& n, float4 dh2, const Acts& A, Scratch& S, int& tog, float4 rstd1,
                                              float4 rstd2, float* ws_dz1, float* ws_dz2, float* part, int b0) {
  const int j = threadIdx.x;
  float g1 = 1.f, g2 = 1.f;
  if (j < ET && n.ln) { g1 = __ldg(n.g1 + j); g2 = __ldg(n.g2 + j); }
  if (j < ET) {
    float cs[3];
    const float4 dz = bwd_epilogue(dh2, A.h2[j], A.xh2[j], rstd2, g2, n.ln, S.sred, tog, cs);
    S.d[j] = dz;
    if (ws_dz2) store_rows(ws_dz2, b0, dz);
    if (part) {
      part[3 * HID + j] = cs[0];
      part[4 * HID + j] = cs[1];
      part[5 * HID + j] = cs[2];
    }
  }
  __syncthreads();
  gemm_rows(n.w2n, HID, S.d, S.red);
  __syncthreads();
  if (j < ET) {
    float cs[3];
    const float4 dz = bwd_epilogue(reduce_partials(S.red), A.h1[j], A.xh1[j], rstd1, g1, n.ln, S.sred, tog, cs);
    if (ws_dz1) store_rows(ws_dz1, b0, dz);
    if (part) {
      part[0 * HID + j] = cs[0];
      part[1 * HID + j] = cs[1];
      part[2 * HID + j] = cs[2];
    }
    S.d[j] = dz;  // safe: every thread passed the barrier above, i.e. finished reading S.d in gemm_rows
  }
  __syncthreads();
}

}  // namespace b2rl
