// Issue cost of cp.async.bulk from one thread: clock64 after each of 12 back-to-back copies (no waits in between).
// Usage: tma_issue <copy_bytes> <grid>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../plangen_b200/csrc/common.cuh"
using namespace pg;
__global__ void __launch_bounds__(128) issue_kernel(const uint8_t* __restrict__ a, int copy_bytes, long long* out, int with_expect) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[16];
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&full[i], 1); mbar_fence_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint64_t pol = policy_evict_first();
    long long t[14];
    const uint8_t* src = a + (size_t)blockIdx.x * 12 * copy_bytes;
    t[0] = clock64();
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      if (with_expect) mbar_expect_tx(&full[i], copy_bytes);
      bulk_copy_g2s(ring + (size_t)i * copy_bytes, src + (size_t)i * copy_bytes, copy_bytes, &full[i], pol);
      t[i + 1] = clock64();
    }
    if (!with_expect) for (int i = 0; i < 12; ++i) mbar_expect_tx(&full[i], copy_bytes);
    for (int i = 0; i < 12; ++i) mbar_wait(&full[i], 0);
    t[13] = clock64();
    if (blockIdx.x == 0) for (int i = 0; i < 14; ++i) out[i] = t[i] - t[0];
  }
}
int main(int argc, char** argv) {
  const int copy_bytes = argc > 1 ? atoi(argv[1]) : 16384, G = argc > 2 ? atoi(argv[2]) : 1;
  uint8_t* a; long long* out;
  cudaMalloc(&a, (size_t)G * 12 * copy_bytes); cudaMalloc(&out, 14 * 8);
  cudaMemset(a, 1, (size_t)G * 12 * copy_bytes);
  const size_t smem = 12 * (size_t)copy_bytes + 1024;
  cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int we = 1; we >= 0; --we) {
    for (int it = 0; it < 2; ++it) issue_kernel<<<G, 128, smem>>>(a, copy_bytes, out, we);
    cudaDeviceSynchronize();
    long long h[14]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("copy %d B grid %d expect_tx %s: cycles after each issue:", copy_bytes, G, we ? "interleaved" : "after");
    for (int i = 1; i <= 12; ++i) printf(" %lld", h[i]);
    printf(" | all landed %lld (%s)\n", h[13], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
