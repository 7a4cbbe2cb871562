// tc_linear.cu — the hidden-layer product for LARGE batches on the 5th-generation tensor cores:
//     Z = X · Wᵀ + b ;  H = ReLU(LayerNorm(Z))        X [M][256], W [256][256] (natural torch layout [out][in])
// agents/nets.py:66-82 (fc_block_2: Linear -> LayerNorm -> ReLU) for M in the tens of thousands (BASELINE.json
// config 5, batch 65 536; stacked populations), where the row-group kernels of mlp_cluster.cuh stream every layer's
// weights from L2 once per 8 rows and top out at ~10 TFLOP/s.
//
// One CTA per 128-row tile, warp-specialised (192 threads):
//   warp 0   TMA producer: cp.async.bulk.tensor 2D loads of the A tile (128 x 32 fp32) and of the whole weight
//            k-slab (256 x 32 fp32) into a 4-stage shared-memory ring, 128-byte swizzle, completion on mbarriers;
//   warp 1   TMEM allocation (256 columns) and the MMA issuer: one elected thread issues
//            tcgen05.mma.cta_group::1.kind::tf32  M=128 N=256 K=8, four per k-slab, accumulating in TMEM;
//            tcgen05.commit releases each ring slot back to the producer and finally signals the epilogue;
//   warps 2-5 epilogue: thread <-> TMEM lane <-> one output ROW, so bias, the LayerNorm statistics (two passes over
//            the row, all in registers: no shuffles, no shared memory), affine and ReLU are thread-local; rows are
//            written with 128-bit stores.
// Operands are fp32 in memory and are read by the tensor core as TF32 (10-bit mantissa, truncated): products carry
// ~1e-3 relative error — the "looser stated bound" of the north star for tensor-core modes; the accumulation, the
// LayerNorm and everything downstream are fp32. Both operands are K-major: X rows and the natural-layout weight
// rows are contiguous along the contraction (the dX product uses the w2t copy the same way).
#include <cuda.h>

#include "common.cuh"

namespace b2rl {

constexpr int TCM = 128, TCN = 256, TCK = 32, TC_STAGES = 4, TC_THREADS = 192;
constexpr uint32_t TC_STAGE_BYTES = (TCM + TCN) * TCK * sizeof(float);  // 48 KB

struct __align__(1024) TcSmem {
  float a[TC_STAGES][TCM * TCK];  // 16 KB per stage, 128-byte rows, swizzle-128B
  float b[TC_STAGES][TCN * TCK];  // 32 KB per stage
  uint64_t full[TC_STAGES], empty[TC_STAGES], acc_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect_(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_%=;\n bra W_%=;\n D_%=:\n}" ::"r"(s32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(s32(bar))
               : "memory");
}
// K-major, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor fields:
// start address [0,14), stride byte offset [32,46), version = 1 [46,48), layout SWIZZLE_128B = 2 [61,64))
__device__ __forceinline__ uint64_t umma_desc(const void* smem, int byte_off) {
  const uint64_t addr = (uint64_t)((s32(smem) + (uint32_t)byte_off) & 0x3FFFFu) >> 4;
  return addr | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor: c_format F32 = 1 [4,6), a/b format TF32 = 2 [7,10) [10,13), K-major both, N>>3 [17,23), M>>4 [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TCN >> 3) << 17) | ((uint32_t)(TCM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_linear_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int M,
                 const float* __restrict__ bias, const float* __restrict__ g, const float* __restrict__ be, int ln, int relu,
                 float* __restrict__ H, float* __restrict__ XH, float2* __restrict__ stat) {
  extern __shared__ unsigned char tc_raw[];  // (the swizzle atoms need 1024-byte alignment: align by hand)
  TcSmem& S = *reinterpret_cast<TcSmem*>(tc_raw + ((1024u - (s32(tc_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TCM;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init_(&S.full[s], 1); mbar_init_(&S.empty[s], 1); }
    mbar_init_(&S.acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 256 fp32 columns x 128 lanes for the accumulator tile
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&S.tmem_base)), "n"(TCN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;
  constexpr int KB = HID / TCK;  // 8 k-slabs

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % TC_STAGES;
        if (kb >= TC_STAGES) mbar_wait_(&S.empty[s], ((kb / TC_STAGES) - 1) & 1);
        mbar_expect_(&S.full[s], TC_STAGE_BYTES);
        tma_load_2d(S.a[s], &mapA, kb * TCK, m0, &S.full[s]);
        tma_load_2d(S.b[s], &mapB, kb * TCK, 0, &S.full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % TC_STAGES;
        mbar_wait_(&S.full[s], (kb / TC_STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int k = 0; k < TCK / 8; ++k)  // UMMA K = 8 tf32 = 32 bytes: advance inside the 128-byte swizzle row
          umma_tf32(tmem, umma_desc(S.a[s], k * 32), umma_desc(S.b[s], k * 32), (kb | k) != 0);
        umma_commit(&S.empty[s]);  // (implies tcgen05.fence::before_thread_sync)
      }
      umma_commit(&S.acc_full);
    }
  } else {  // ===== epilogue: warp w may touch TMEM lanes 32*(w % 4) .. +31
    const int lg = warp & 3;
    const int row = m0 + 32 * lg + lane;
    mbar_wait_(&S.acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tl = tmem + ((uint32_t)(32 * lg) << 16);
    float v[32];
    float mean = 0.f, rstd = 1.f;
    if (ln) {
      float s1 = 0.f;
      for (int c = 0; c < TCN / 32; ++c) {
        tmem_ld32(tl + c * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) s1 += v[i] + __ldg(bias + c * 32 + i);
      }
      mean = s1 * (1.0f / TCN);
      float s2 = 0.f;
      for (int c = 0; c < TCN / 32; ++c) {
        tmem_ld32(tl + c * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float d = v[i] + __ldg(bias + c * 32 + i) - mean;
          s2 = fmaf(d, d, s2);
        }
      }
      rstd = 1.0f / sqrtf(s2 * (1.0f / TCN) + LN_EPS);
      if (stat && row < M) stat[row] = make_float2(mean, rstd);
    }
    for (int c = 0; c < TCN / 32; ++c) {
      tmem_ld32(tl + c * 32, v);
      float h[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int j = c * 32 + i;
        float x = v[i] + __ldg(bias + j);
        if (ln) {
          x = (x - mean) * rstd;
          v[i] = x;  // x-hat
          x = fmaf(x, __ldg(g + j), __ldg(be + j));
        } else {
          v[i] = x;  // pre-activation
        }
        h[i] = relu ? fmaxf(x, 0.f) : x;
      }
      if (row < M) {
        float4* hp = reinterpret_cast<float4*>(H + (size_t)row * TCN + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) hp[i] = make_float4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
        if (XH) {
          float4* xp = reinterpret_cast<float4*>(XH + (size_t)row * TCN + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) xp[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TCN) : "memory");
}

// ---- host side: tensor maps (driver entry point fetched through the runtime: libb2rl links no libcuda) -------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}
// [rows][cols] fp32, row pitch `ld` floats; box = 32 columns (128 bytes = the swizzle span) x box_rows rows
static bool make_map(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)TCK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t init_tc() {
  return cudaFuncSetAttribute(tc_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcSmem) + 1024);
}

cudaError_t launch_tc_linear(const float* X, int64_t ldx, int M, const float* W, const float* bias, const float* g,
                             const float* be, int ln, int relu, float* H, float* XH, float* stat, cudaStream_t st) {
  CUtensorMap ma, mb;
  if (!make_map(&ma, X, M, HID, ldx, TCM) || !make_map(&mb, W, HID, HID, HID, TCN)) return cudaErrorInvalidValue;
  tc_linear_kernel<<<(M + TCM - 1) / TCM, TC_THREADS, sizeof(TcSmem) + 1024, st>>>(ma, mb, M, bias, g, be, ln, relu, H, XH,
                                                                                  reinterpret_cast<float2*>(stat));
  return cudaGetLastError();
}

}  // namespace b2rl
