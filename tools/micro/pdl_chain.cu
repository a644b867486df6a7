// Cost of one dependent-kernel boundary on B200: a chain of tiny kernels (32 CTAs x 256 threads, one dependent
// load + store each) launched with programmatic dependent launch inside a CUDA graph.
//   mode 0: griddepcontrol.wait at the top (what every kernel of the decode step does)
//   mode 1: the consumer polls a global counter its predecessor's CTAs bump after their stores (release/acquire),
//           and only calls griddepcontrol.wait at the very end (keeps completion transitive)
//   mode 2: no PDL attribute at all (plain stream order)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void link_kernel(const float* __restrict__ in, float* __restrict__ out, unsigned* cnt_in, unsigned* cnt_out,
                            unsigned target, int mode) {
  asm volatile("griddepcontrol.launch_dependents;");
  if (mode == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (mode == 1 && cnt_in) {
    if (threadIdx.x == 0) {
      unsigned v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt_in) : "memory"); } while (v < target);
    }
    __syncthreads();
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  out[i] = __ldcg(in + i) + 1.0f;
  if (mode == 1) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(cnt_out, 1u);
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
}
int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0, n = 1000, G = argc > 2 ? atoi(argv[2]) : 32;
  float *a, *b; unsigned* cnt;
  cudaMalloc(&a, G * 256 * 4); cudaMalloc(&b, G * 256 * 4); cudaMalloc(&cnt, (n + 1) * 4);
  cudaMemset(a, 0, G * 256 * 4); cudaMemset(cnt, 0, (n + 1) * 4);
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
  for (int k = 0; k < n; ++k) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G); cfg.blockDim = dim3(256); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = mode == 2 ? 0 : 1;
    cudaLaunchKernelEx(&cfg, link_kernel, (const float*)((k & 1) ? b : a), (k & 1) ? a : b, k ? cnt + k - 1 : (unsigned*)nullptr, cnt + k,
                       (unsigned)G, mode);
  }
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int it = 0; it < 5; ++it) {
    cudaMemsetAsync(cnt, 0, (n + 1) * 4, st);
    cudaEventRecord(e0, st); cudaGraphLaunch(ge, st); cudaEventRecord(e1, st); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  printf("mode %d grid %d: %.3f us per dependent kernel (%s)\n", mode, G, best * 1e3 / n, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
