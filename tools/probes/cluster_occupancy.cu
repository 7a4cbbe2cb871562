// Probe: how many clusters of size S (256 threads, given dynamic smem per CTA) can be co-resident on this GPU?
// Also FFMA vs FFMA2 issue throughput. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_occupancy cluster_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* p) { extern __shared__ float s[]; if (p) p[0] = s[0]; }

__global__ void __launch_bounds__(256) ffma_k(float* sink, int iters, float m0) {
  float a[32];
  for (int i = 0; i < 32; ++i) a[i] = 1.0f + 1e-3f * (threadIdx.x + i);
  float m = m0 + threadIdx.x * 1e-9f, c = m0 * 1e-3f + threadIdx.x * 1e-9f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = fmaf(a[i], m, c);
  }
  float s = 0.f;
  for (int i = 0; i < 32; ++i) s += a[i];
  if (s == 12345.678f) sink[0] = s;
}
__global__ void __launch_bounds__(256) ffma2_k(float* sink, int iters, float m0) {
  float2 a[16];
  for (int i = 0; i < 16; ++i) a[i] = make_float2(1.0f + 1e-3f * (threadIdx.x + i), 1.0f + 2e-3f * (threadIdx.x + i));
  float m = m0 + threadIdx.x * 1e-9f;
  float2 c = make_float2(m0 * 1e-3f + threadIdx.x * 1e-9f, m0 * 2e-3f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float2 mm = make_float2(m, m);
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(*reinterpret_cast<unsigned long long*>(&a[i]))
                   : "l"(*reinterpret_cast<unsigned long long*>(&mm)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
    }
  }
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += a[i].x + a[i].y;
  if (s == 12345.678f) sink[0] = s;
}

int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int smem : {100 * 1024, 170 * 1024, 220 * 1024}) {
    for (int cs : {1, 2, 4, 8, 16}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
      printf("smem %3d KB cluster %2d: max active clusters %d (%d CTAs) %s\n", smem / 1024, cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  float* sink; cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, ctas = 148 * 8;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); ffma_k<<<ctas, 256>>>(sink, iters, 0.999f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("FFMA  (3-reg): %.1f TFLOP/s\n", 2.0 * 32 * iters * 256.0 * ctas / ms / 1e9);
    cudaEventRecord(e0); ffma2_k<<<ctas, 256>>>(sink, iters, 0.999f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("FFMA2 (3-reg): %.1f TFLOP/s\n", 2.0 * 32 * iters * 256.0 * ctas / ms / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
