// Probe behind the decision on a single-launch "persistent" update (north_star bullet 6, SURVEY §8(b) update_persistent_*):
// what does a PHASE BOUNDARY cost on a B200, as (a) a kernel boundary with programmatic dependent launch inside a CUDA graph —
// what the engine's 3-launch iteration pays today — and (b) a grid-wide barrier inside one persistent launch (one CTA per SM,
// 148 x 256 threads, the shape of the fused update kernels) — what a persistent iteration would pay instead?
// Both variants run the SAME phases: every CTA reads what another CTA wrote in the previous phase (so the boundary must also
// order memory, as critic_fused -> wgrad -> Adam do through the L2-resident workspace), does `work` dependent FMA steps, and
// writes. The per-phase time difference is the boundary cost difference; everything else is identical by construction.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o persistent_probe persistent_probe.cu ; run: ./persistent_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float phase_work(const float* in, int work, int cta, int n_cta) {
  // read the slot the NEXT cta wrote in the previous phase (a cross-SM dependency through L2)
  float x = __ldcg(in + ((cta + 1) % n_cta) * 32 + (threadIdx.x & 31));
  for (int i = 0; i < work; ++i) x = fmaf(x, 0.999f, 1e-3f);
  return x;
}

// (a) one phase per launch, programmatic dependent launch: launch_dependents + wait first thing, like pdl_enter()
__global__ void __launch_bounds__(256) phase_kernel(const float* in, float* out, int work) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const float x = phase_work(in, work, blockIdx.x, gridDim.x);
  if (threadIdx.x < 32) out[blockIdx.x * 32 + threadIdx.x] = x;
}

// (b) all phases in one launch; grid barrier = one atomic arrive per CTA on a monotonic counter + acquire spin
__global__ void __launch_bounds__(256) persistent_kernel(float* a, float* b, int work, int phases, unsigned int* counter) {
  float* in = a;
  float* out = b;
  for (int p = 0; p < phases; ++p) {
    const float x = phase_work(in, work, blockIdx.x, gridDim.x);
    if (threadIdx.x < 32) out[blockIdx.x * 32 + threadIdx.x] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(counter, 1u);
      const unsigned int target = (unsigned int)(p + 1) * gridDim.x;
      unsigned int v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      } while (v < target);
    }
    __syncthreads();
    float* t = in; in = out; out = t;
  }
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  float *a, *b;
  unsigned int* ctr;
  cudaMalloc(&a, sms * 32 * 4); cudaMalloc(&b, sms * 32 * 4); cudaMalloc(&ctr, 4);
  cudaMemset(a, 0, sms * 32 * 4); cudaMemset(b, 0, sms * 32 * 4);
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int phases = 300, reps = 20;
  printf("B200 phase-boundary probe: %d CTAs x 256 threads (one per SM), %d phases per graph / launch, %d repetitions\n", sms, phases, reps);
  printf("%10s %22s %22s %22s\n", "work/phase", "graph+PDL us/phase", "graph no-PDL us/phase", "persistent us/phase");
  for (int work : {0, 2000, 8000, 32000}) {
    float t_pdl = 0.f, t_plain = 0.f, t_pers = 0.f;
    for (int pdl = 1; pdl >= 0; --pdl) {
      cudaGraph_t g; cudaGraphExec_t ge;
      cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
      for (int p = 0; p < phases; ++p) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sms); cfg.blockDim = dim3(256); cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
        cudaLaunchKernelEx(&cfg, phase_kernel, (const float*)((p & 1) ? b : a), (p & 1) ? a : b, work);
      }
      cudaStreamEndCapture(st, &g);
      cudaGraphInstantiate(&ge, g, 0);
      for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, st);
      cudaStreamSynchronize(st);
      cudaEventRecord(e0, st);
      for (int i = 0; i < reps; ++i) cudaGraphLaunch(ge, st);
      cudaEventRecord(e1, st);
      cudaStreamSynchronize(st);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      (pdl ? t_pdl : t_plain) = ms * 1e3f / (reps * phases);
      cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
    }
    {
      void* args[] = {&a, &b, (void*)&work, (void*)&phases, &ctr};
      for (int i = 0; i < 3; ++i) { cudaMemsetAsync(ctr, 0, 4, st); cudaLaunchCooperativeKernel((void*)persistent_kernel, dim3(sms), dim3(256), args, 0, st); }
      cudaStreamSynchronize(st);
      float tot = 0.f;
      for (int i = 0; i < reps; ++i) {
        cudaMemsetAsync(ctr, 0, 4, st);
        cudaEventRecord(e0, st);
        cudaLaunchCooperativeKernel((void*)persistent_kernel, dim3(sms), dim3(256), args, 0, st);
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        tot += ms;
      }
      t_pers = tot * 1e3f / (reps * phases);
    }
    printf("%10d %22.3f %22.3f %22.3f\n", work, t_pdl, t_plain, t_pers);
  }
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
