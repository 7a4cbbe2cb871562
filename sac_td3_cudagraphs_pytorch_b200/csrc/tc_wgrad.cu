// tc_wgrad.cu — weight gradients for LARGE batches on the tensor cores:  D[m][n] = sum_b A[b][m] * Bm[b][n]
// (dW1t = X0^T dZ1, dW2t = H1^T dZ2, dW3 = dZ3^T H2: the contraction over the batch that `loss.backward()` does for the
// parameters, agents/agent.py:235,283). wgrad.cu's FFMA tiles each walk the whole batch and take 1 ms at B = 65 536;
// here the batch is split over work items (split-K), each accumulated as a 128 x 256 tile in TMEM over its slice, and a
// second kernel adds the slices in a fixed order (deterministic). Stacked agents with one split per agent skip the second
// kernel: the TMEM epilogue writes the gradient tensor itself.
//
// Both operands are MN-major (the batch index is the contraction and the slow one in memory). For 32-bit MN-major
// operands tcgen05 accepts ONE shared-memory layout, SWIZZLE_128B_BASE32B (measured: with the plain 128-byte swizzle the
// MMA returns zeros): 128-byte rows of 32 consecutive m, 32-byte units XORed with (row & 3), atoms of 4 batch rows —
// in 16-byte units ((8,n),(4,k)):((1,LBO),(8,SBO)) (cute/atom/mma_traits_sm100.hpp). TMA writes exactly that with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B and boxes of 32 columns x 32 batch rows: atoms 512 B apart (SBO), 32-column chunks
// 4 KB apart (LBO). One tcgen05.mma kind::tf32 (K = 8) consumes two atoms (1 KB) of every chunk. PREC 1 = 3xTF32 as in
// tc_linear.cu (both operands are activations here: both lo parts are made in shared memory by two split warps).
#include <cuda.h>

#include "common.cuh"

namespace b2rl {

constexpr int GM = 128, GN = 256, GK = 32;
constexpr int G_A_BYTES = GM * GK * 4, G_B_BYTES = GN * GK * 4;  // 16 KB, 32 KB per stage

// PAIR (3xTF32, gradients of 129..256 rows: dW2): the two m tiles of a split are the two CTAs of a cluster running ONE
// tcgen05.mma.cta_group::2 (M = 256) per k-step and product — CTA r supplies its 128 columns of A and columns
// [128 r, 128 r + 128) of B, as in tc_linear.cu's pair form. The single-SM form made each m tile's CTA load (and split) the
// WHOLE B slab: 201 MB of operands for a 65 536-row dW2 instead of 134, 96 KB per stage instead of 64.
template <int PREC, bool PAIR>
struct __align__(1024) GSmemT {
  static constexpr int STAGES = PREC ? (PAIR ? 3 : 2) : 4;
  static constexpr int BN = PAIR ? GN / 2 : GN;  // B columns this CTA holds
  float a[STAGES][GM * GK];
  float b[STAGES][BN * GK];
  float alo[PREC ? STAGES : 1][PREC ? GM * GK : 4];
  float blo[PREC ? STAGES : 1][PREC ? BN * GK : 4];
  uint64_t full[STAGES], empty[STAGES], lo_ready[STAGES], peer_full[STAGES], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t gs32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g_mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_%=;\n bra W_%=;\n D_%=:\n}" ::"r"(gs32(b)),
      "r"(parity)
      : "memory");
}
// rank-3 maps {column, batch row, agent}: rows beyond an agent's batch are zero-filled (they add nothing to the sum)
__device__ __forceinline__ void g_tma_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(gs32(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(gs32(bar))
               : "memory");
}
// MN-major, SWIZZLE_128B_BASE32B (layout type 1): LBO = 4 KB between 32-column chunks [16,30), SBO = 512 B between 4-row atoms [32,46)
__device__ __forceinline__ uint64_t g_desc(const void* smem, int byte_off) {
  const uint64_t addr = (uint64_t)((gs32(smem) + (uint32_t)byte_off) & 0x3FFFFu) >> 4;
  return addr | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
// as TC_IDESC of tc_linear.cu, with a_major = b_major = MN (bits 15, 16)
constexpr uint32_t G_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(GN >> 3) << 17) |
                             ((uint32_t)(GM >> 4) << 24);
__device__ __forceinline__ void g_umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(
      G_IDESC), "r"(accumulate)
      : "memory");
}
// the pair forms (tc_linear.cu): one thread of cluster rank 0 drives both SMs; commits arrive in both CTAs
__device__ __forceinline__ void g_umma_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(GN >> 3) << 17) |
                             ((uint32_t)((2 * GM) >> 4) << 24);
  const uint32_t z = 0;
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
__device__ __forceinline__ void g_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(gs32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void g_arrive_rank0(uint64_t* bar) {  // the mbarrier at the same offset in cluster rank 0
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(gs32(bar)));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
__device__ __forceinline__ void g_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Work items (m tile, split, agent): item (mt, sp, ag) accumulates rows [sp * rows_per_split, ...) of agent ag's batch for
// output rows [128 mt, 128 mt + 128) and writes its slice to part[ag][sp][MA_pad][256] (or, with one split, straight to the
// gradient tensor). PERSISTENT: one CTA per SM walks items blockIdx.x, blockIdx.x + gridDim.x, ...; the TMA ring runs on
// across items and the accumulator is double-buffered in TMEM (2 x 256 columns), so the prologue (barrier init, TMEM
// allocation, descriptor fetch) is paid once per SM and the epilogue of one item overlaps the loads and MMAs of the next —
// a stacked population is 1 024-2 048 items of only 8 slabs each (the one-CTA-per-item version spent 60 % of its time
// outside the main loop: 103 us to read 268 MB). Warps: 0 TMA producer, 1 MMA issuer, 2-5 epilogue (thread <-> TMEM lane
// <-> output row m), 6-7 the operands' lo parts (3xTF32). PAIR: see GSmemT — the walkers then run per cluster and rank 0's
// elected thread issues for both SMs.
constexpr int G_THREADS_P = 256;
template <int PREC, bool PAIR>
__global__ void __launch_bounds__(G_THREADS_P, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int Bn, int rows_per_split,
                int MA_pad, int MA, float* __restrict__ part, float* __restrict__ Cd, float* __restrict__ Ctd,
                unsigned long long* bump, long long ps, long long cs, int n_mt, int n_sp, int n_ag) {
  extern __shared__ unsigned char g_raw[];
  using Smem = GSmemT<PREC, PAIR>;
  constexpr int ST = Smem::STAGES;
  constexpr uint32_t STAGE_BYTES = G_A_BYTES + Smem::BN * GK * 4;
  Smem& S = *reinterpret_cast<Smem*>(g_raw + ((1024u - (gs32(g_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t crank = 0;
  if constexpr (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  // PAIR: an item is (split, agent) and the cluster's CTA r takes m tile r; the walkers below are per cluster then
  const int n_items = PAIR ? n_sp * n_ag : n_mt * n_sp * n_ag;
  const int w0 = PAIR ? (int)blockIdx.x >> 1 : (int)blockIdx.x, wstep = PAIR ? (int)gridDim.x >> 1 : (int)gridDim.x;
  // item -> (mt, sp, ag, first batch row, slabs)
  auto item_of = [&](int i, int& mt, int& sp, int& ag, int& kb0, int& KB) {
    if constexpr (PAIR) {
      mt = (int)crank;
      sp = i % n_sp;
      ag = i / n_sp;
    } else {
      mt = i % n_mt;
      sp = (i / n_mt) % n_sp;
      ag = i / (n_mt * n_sp);
    }
    kb0 = sp * rows_per_split;
    const int kb1 = min(Bn, kb0 + rows_per_split);
    KB = (max(kb1 - kb0, 0) + GK - 1) / GK;
  };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < ST; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gs32(&S.full[s])), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gs32(&S.empty[s])), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gs32(&S.lo_ready[s])), "r"(PAIR ? 128 : 64) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gs32(&S.peer_full[s])), "r"(1) : "memory");
    }
    for (int b = 0; b < 2; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gs32(&S.acc_full[b])), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gs32(&S.acc_empty[b])), "r"(PAIR ? 256 : 128) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // two accumulator tiles of 256 fp32 columns x 128 lanes (all of TMEM: one CTA per SM)
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gs32(&S.tmem_base)), "n"(2 * GN) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gs32(&S.tmem_base)), "n"(2 * GN) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if constexpr (PAIR) g_cluster_sync();  // the peer's barriers are initialised before anything of ours can land on them
  const uint32_t tmem = S.tmem_base;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer: 4 + 8 boxes of 32 columns x 32 batch rows per stage; the ring runs on across items
      int it = 0;
      for (int i = w0; i < n_items; i += wstep) {
        int mt, sp, ag, kb0, KB;
        item_of(i, mt, sp, ag, kb0, KB);
        const int m0 = mt * GM, n0 = PAIR ? (int)crank * (GN / 2) : 0;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % ST;
          if (it >= ST) g_mbar_wait(&S.empty[s], ((it / ST) - 1) & 1);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gs32(&S.full[s])), "r"(STAGE_BYTES) : "memory");
          const int row = kb0 + kb * GK;  // (rows beyond the batch are zero-filled by TMA: they add nothing)
#pragma unroll
          for (int c = 0; c < GM / 32; ++c) g_tma_3d(S.a[s] + c * 1024, &mapA, m0 + 32 * c, row, ag, &S.full[s]);
#pragma unroll
          for (int c = 0; c < Smem::BN / 32; ++c) g_tma_3d(S.b[s] + c * 1024, &mapB, n0 + 32 * c, row, ag, &S.full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (PAIR && lane == 0 && crank == 1) {  // rank 1: tell the issuer when this CTA's slabs have landed
      int it = 0;
      for (int i = w0; i < n_items; i += wstep) {
        int mt, sp, ag, kb0, KB;
        item_of(i, mt, sp, ag, kb0, KB);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          g_mbar_wait(&S.full[it % ST], (it / ST) & 1);
          g_arrive_rank0(&S.peer_full[it % ST]);
        }
      }
    } else if (PAIR && lane == 0) {  // ===== MMA issuer of the pair: every instruction drives both SMs
      int it = 0, nb = 0;
      for (int i = w0; i < n_items; i += wstep) {
        int mt, sp, ag, kb0, KB;
        item_of(i, mt, sp, ag, kb0, KB);
        if (KB == 0) continue;
        const int buf = nb & 1;
        const uint32_t acc = tmem + buf * GN;
        if (nb >= 2) {  // BOTH epilogues must have drained this tile (item nb - 2)
          g_mbar_wait(&S.acc_empty[buf], ((nb >> 1) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % ST;
          g_mbar_wait(&S.full[s], (it / ST) & 1);
          g_mbar_wait(&S.peer_full[s], (it / ST) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int k = 0; k < GK / 8; ++k) g_umma_pair(acc, g_desc(S.a[s], k * 1024), g_desc(S.b[s], k * 1024), (kb | k) != 0);
          g_mbar_wait(&S.lo_ready[s], (it / ST) & 1);  // both CTAs' split warps are done with this slab
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int k = 0; k < GK / 8; ++k) {
            g_umma_pair(acc, g_desc(S.alo[s], k * 1024), g_desc(S.b[s], k * 1024), 1);
            g_umma_pair(acc, g_desc(S.a[s], k * 1024), g_desc(S.blo[s], k * 1024), 1);
          }
          g_commit_pair(&S.empty[s]);
        }
        g_commit_pair(&S.acc_full[buf]);
        ++nb;
      }
    } else if (!PAIR && lane == 0) {  // ===== MMA issuer
      int it = 0, nb = 0;  // nb: non-empty items so far (they alternate between the two accumulator tiles)
      for (int i = w0; i < n_items; i += wstep) {
        int mt, sp, ag, kb0, KB;
        item_of(i, mt, sp, ag, kb0, KB);
        if (KB == 0) continue;
        const int buf = nb & 1;
        const uint32_t acc = tmem + buf * GN;
        if (nb >= 2) {  // the epilogue must have drained this tile (item nb - 2)
          g_mbar_wait(&S.acc_empty[buf], ((nb >> 1) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % ST;
          g_mbar_wait(&S.full[s], (it / ST) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int k = 0; k < GK / 8; ++k)  // 8 batch rows (two atoms, 1 KB) of every chunk per instruction
            g_umma(acc, g_desc(S.a[s], k * 1024), g_desc(S.b[s], k * 1024), (kb | k) != 0);
          if constexpr (PREC == 1) {  // (the hi.hi product runs while the lo parts are made)
            g_mbar_wait(&S.lo_ready[s], (it / ST) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int k = 0; k < GK / 8; ++k) {
              g_umma(acc, g_desc(S.alo[s], k * 1024), g_desc(S.b[s], k * 1024), 1);
              g_umma(acc, g_desc(S.a[s], k * 1024), g_desc(S.blo[s], k * 1024), 1);
            }
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(gs32(&S.empty[s])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(gs32(&S.acc_full[buf])) : "memory");
        ++nb;
      }
    }
  } else if (warp >= 6) {  // ===== (3xTF32) both operands' lo parts, slab by slab as the ring fills
    if constexpr (PREC == 1) {
      const int lt = threadIdx.x - 192;  // 0..63
      int it = 0;
      for (int i = w0; i < n_items; i += wstep) {
        int mt, sp, ag, kb0, KB;
        item_of(i, mt, sp, ag, kb0, KB);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % ST;
          g_mbar_wait(&S.full[s], (it / ST) & 1);
          const float4* a4 = reinterpret_cast<const float4*>(S.a[s]);
          float4* al4 = reinterpret_cast<float4*>(S.alo[s]);
#pragma unroll 4
          for (int q = 0; q < GM * GK / 4 / 64; ++q) al4[lt + 64 * q] = tf32_lo4(a4[lt + 64 * q]);
          const float4* b4 = reinterpret_cast<const float4*>(S.b[s]);
          float4* bl4 = reinterpret_cast<float4*>(S.blo[s]);
#pragma unroll 4
          for (int q = 0; q < Smem::BN * GK / 4 / 64; ++q) bl4[lt + 64 * q] = tf32_lo4(b4[lt + 64 * q]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          if constexpr (PAIR) g_arrive_rank0(&S.lo_ready[s]);
          else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gs32(&S.lo_ready[s])) : "memory");
        }
      }
    }
  } else {  // ===== epilogue warps 2-5: TMEM -> this item's slice (or the gradient tensor itself)
    const int et = threadIdx.x - 64, lg = warp & 3;
    int nb = 0;
    for (int i = w0; i < n_items; i += wstep) {
      int mt, sp, ag, kb0, KB;
      item_of(i, mt, sp, ag, kb0, KB);
      const int m0 = mt * GM;
      // Cd != NULL (one split: nothing to add up): the tile goes straight to the gradient tensor C [MA][256] of this agent
      // (thread <-> row m: 128 contiguous bytes per chunk) and to its transpose Ct [256][MA] (lanes <-> consecutive m:
      // coalesced), and the item (mt 0, agent) advances the agent's update counter — no reduce launch, no scratch round trip.
      const int mrow = m0 + 32 * lg + lane;
      float* dst = Cd ? Cd + (size_t)ag * ps + (size_t)mrow * GN
                      : part + (((size_t)ag * n_sp + sp) * MA_pad + mrow) * GN;
      float* dstT = (Cd && Ctd) ? Ctd + (size_t)ag * ps + mrow : nullptr;
      if (Cd && bump && mt == 0 && et == 0) bump[(size_t)ag * cs] += 1ull;
      const bool live = mrow < MA;  // only rows that exist are written (dW1: 14 of 128, a critic's dW3: 1)
      if (KB > 0) {
        const int buf = nb & 1;
        // (every epilogue warp waits, also one none of whose rows exists: the four warps then stay within one item of each
        // other and of the MMA issuer, which the arrival counts of acc_empty rely on)
        g_mbar_wait(&S.acc_full[buf], (nb >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (m0 + 32 * lg < MA) {  // (else nothing to read back or write)
          const uint32_t tl = tmem + buf * GN + ((uint32_t)(32 * lg) << 16);
          for (int c = 0; c < GN / 32; ++c) {
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(tl + c * 32));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
            if (live) {
#pragma unroll
              for (int q = 0; q < 8; ++q) d4[q] = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
              if (dstT) {
#pragma unroll
                for (int q = 0; q < 32; ++q) dstT[(size_t)(c * 32 + q) * MA] = __uint_as_float(r[q]);
              }
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // hand the tile back to the MMA issuer
        if constexpr (PAIR) g_arrive_rank0(&S.acc_empty[buf]);
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gs32(&S.acc_empty[buf])) : "memory");
        ++nb;
      } else if (live) {  // an empty slice (more splits than slabs): zeros
        for (int c = 0; c < GN / 4; ++c) reinterpret_cast<float4*>(dst)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dstT)
          for (int n = 0; n < GN; ++n) dstT[(size_t)n * MA] = 0.f;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (PAIR) g_cluster_sync();  // (the peer's last commits arrive on this CTA's barriers: do not exit under them)
  if (warp == 1) {
    if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * GN) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * GN) : "memory");
  }
}

// bump, if given, is incremented once (the update counter that wgrad.cu's extra CTA advances).
// C[m][n] = sum_sp part[sp][m][n] for m < MA; Ct [256][MA] = its transpose when given (MA = 256: the w2n shadow).
// A CTA owns 8 rows x 32 columns; its 8 warps each add every 8th slice (8 rows = 8 loads in flight), then the 8 partial
// sums are added in warp order: a fixed order, so the result is deterministic, and a 1-row gradient (dW3 of a critic)
// still has 8 warps sharing its 148 slices.
__global__ void __launch_bounds__(256)
tc_wgrad_reduce_kernel(const float* __restrict__ part, int S, int MA, int MA_pad, float* __restrict__ C, float* __restrict__ Ct,
                       unsigned long long* bump, long long ps, long long cs) {
  __shared__ float red[8][8][33];  // [split lane][row][column]
  {  // stacked agents: blockIdx.z = agent
    const size_t ag = blockIdx.z;
    part += ag * S * MA_pad * GN, C += ag * ps;
    if (Ct) Ct += ag * ps;
    if (bump) bump += ag * cs;
  }
  if (bump && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *bump += 1ull;  // the step's update counter (wgrad.cu's bump CTA)
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int m0 = blockIdx.y * 8, n = blockIdx.x * 32 + tx;
  const size_t stride = (size_t)MA_pad * GN;
  float acc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = 0.f;
  // (four slices = up to 32 loads per thread in flight: the loop is a chain of L2 round trips otherwise — 17 us for the
  //  74 slices of a 256 x 256 gradient, against ~4 us of traffic)
  for (int sp0 = ty; sp0 < S; sp0 += 32) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int sp = sp0 + 8 * u;
      const float* src = part + (size_t)sp * stride + (size_t)m0 * GN + n;
#pragma unroll
      for (int r = 0; r < 8; ++r) v[u][r] = (sp < S && m0 + r < MA) ? __ldg(src + (size_t)r * GN) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[r] += v[u][r];
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) red[ty][r][tx] = acc[r];
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int l = 0; l < 8; ++l) s += red[l][ty][tx];  // thread (tx, ty) finishes row m0 + ty
  if (m0 + ty < MA) C[(size_t)(m0 + ty) * GN + n] = s;
  if (Ct) {
    __syncthreads();
    red[0][ty][tx] = s;
    __syncthreads();
    const int mm = threadIdx.x & 7, nn = threadIdx.x >> 3;  // 8 consecutive m per 32-byte segment
    if (m0 + mm < MA) Ct[(size_t)(blockIdx.x * 32 + nn) * MA + m0 + mm] = red[0][mm][nn];
  }
}

typedef CUresult (*GEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static GEncodeFn g_encode() {
  static GEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (GEncodeFn)p;
  }();
  return fn;
}
// [n_agents][rows][cols] fp32, pitch ld floats, agents rows * ld apart; boxes of 32 columns x 32 rows x 1 agent
static bool g_map(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int n_agents) {
  GEncodeFn fn = g_encode();
  if (!fn) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(n_agents < 1 ? 1 : n_agents)};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(float), (cuuint64_t)rows * ld * sizeof(float)};
  const cuuint32_t box[3] = {32, (cuuint32_t)GK, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t init_tc_wgrad() {
  cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GSmemT<0, false>) + 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(tc_wgrad_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GSmemT<1, false>) + 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(tc_wgrad_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GSmemT<1, true>) + 1024);
  cudaFuncAttributes fa;
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, tc_wgrad_reduce_kernel);
  return e;
}

// stacked = 0: one learner — one wave of 148 CTAs over (m tiles x splits), at least 4 slabs per CTA.
// stacked = 1: a population — the agents fill the machine; the split depends on (Bn, MA) ONLY, never on how many agents a
// launch holds, so that gradients are bitwise the same however a population is sharded: one split per 1024 batch rows.
int tc_wgrad_splits(int Bn, int MA, int stacked) {
  const int mt = (MA + GM - 1) / GM;
  const int cap = (148 + mt - 1) / mt;
  int s = stacked ? (Bn + 1023) / 1024 : (Bn + 4 * GK - 1) / (4 * GK);
  return s < 1 ? 1 : (s > cap ? cap : s);
}

// A [Bn][lda] (columns 0..MA-1 used, a_cols columns exist), Bm [Bn][256] -> C [MA][256] (+ Ct [256][MA]); scratch >=
// n_agents * splits * MA_pad * 256 floats, MA_pad = MA rounded up to 128
cudaError_t launch_tc_wgrad(const float* A, int64_t lda, int a_cols, int MA, const float* Bm, int Bn, float* C, float* Ct,
                            float* scratch, int x3, unsigned long long* bump, const Stk& k, int stacked, cudaStream_t st) {
  CUtensorMap ma, mb;
  if (!g_map(&ma, A, Bn, a_cols, lda, k.n) || !g_map(&mb, Bm, Bn, GN, GN, k.n)) return cudaErrorInvalidValue;
  const int mt = (MA + GM - 1) / GM, MA_pad = mt * GM, S = tc_wgrad_splits(Bn, MA, stacked);
  const int rps = (((Bn + S - 1) / S) + GK - 1) / GK * GK;
  const bool direct = stacked && S == 1;  // one split per agent: the epilogue writes C / Ct itself (and bumps the counter)
  float* cd = direct ? C : nullptr;
  float* ctd = direct ? Ct : nullptr;
  static int sms = [] {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  const int n_items = mt * S * k.n, grid = n_items < sms ? n_items : sms;  // persistent: one CTA per SM
  cudaError_t e;
  if (x3 && mt == 2) {  // 3xTF32, 129..256 rows: the two m tiles of a split as one cluster on 2-SM MMAs
    const int pairs = S * k.n, clusters = pairs < sms / 2 ? pairs : sms / 2;
    e = launch_k(tc_wgrad_kernel<1, true>, dim3(2 * clusters), dim3(G_THREADS_P), -2, sizeof(GSmemT<1, true>) + 1024, st, ma, mb, Bn, rps,
                 MA_pad, MA, scratch, cd, ctd, bump, (long long)k.ps, (long long)k.cs, mt, S, k.n);
  } else {
    if (x3) tc_wgrad_kernel<1, false><<<grid, G_THREADS_P, sizeof(GSmemT<1, false>) + 1024, st>>>(ma, mb, Bn, rps, MA_pad, MA, scratch, cd, ctd, bump, k.ps, k.cs, mt, S, k.n);
    else tc_wgrad_kernel<0, false><<<grid, G_THREADS_P, sizeof(GSmemT<0, false>) + 1024, st>>>(ma, mb, Bn, rps, MA_pad, MA, scratch, cd, ctd, bump, k.ps, k.cs, mt, S, k.n);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess || direct) return e;
  tc_wgrad_reduce_kernel<<<dim3(GN / 32, (MA + 7) / 8, k.n), 256, 0, st>>>(scratch, S, MA, MA_pad, C, Ct, bump, k.ps, k.cs);
  return cudaGetLastError();
}

}  // namespace b2rl
