// Microbenchmark: achievable HBM read bandwidth of TMA bulk copies (cp.async.bulk) into smem rings.
// Usage: stream_bench <copy_bytes> <stages_per_producer> <producers_per_cta> <ctas_per_sm> <stride_mode>
//   stride_mode 0: each producer streams one contiguous region; 1: chunks of 2 copies from two arrays (K/V style)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../plangen_b200/csrc/common.cuh"
using namespace pg;

__global__ void __launch_bounds__(256) stream_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t bytes_per_producer,
                                                     int copy_bytes, int stages, int producers, int mode, unsigned long long* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[64], empty[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stage_bytes = mode ? 2 * copy_bytes : copy_bytes;
  if (tid == 0) { for (int i = 0; i < producers * stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } mbar_fence_init(); }
  __syncthreads();
  const size_t n_copies = bytes_per_producer / stage_bytes;
  if (warp < producers) {
    if (lane == 0) {
      const int p = warp;
      const size_t gp = ((size_t)blockIdx.x * producers + p);
      const uint8_t* src_a = a + gp * bytes_per_producer / (mode ? 2 : 1);
      const uint8_t* src_b = b + gp * bytes_per_producer / 2;
      const uint64_t pol = policy_evict_first();
      for (size_t i = 0; i < n_copies; ++i) {
        const int s = p * stages + (int)(i % stages);
        mbar_wait(&empty[s], (((uint32_t)(i / stages)) & 1u) ^ 1u);
        mbar_expect_tx(&full[s], stage_bytes);
        bulk_copy_g2s(ring + (size_t)s * stage_bytes, src_a + i * copy_bytes, copy_bytes, &full[s], pol);
        if (mode) bulk_copy_g2s(ring + (size_t)s * stage_bytes + copy_bytes, src_b + i * copy_bytes, copy_bytes, &full[s], pol);
      }
    }
  } else if (warp < 2 * producers) {
    const int p = warp - producers;
    unsigned long long acc = 0;
    for (size_t i = 0; i < n_copies; ++i) {
      const int s = p * stages + (int)(i % stages);
      mbar_wait(&full[s], ((uint32_t)(i / stages)) & 1u);
      acc += ring[(size_t)s * stage_bytes + lane];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (acc == 0x12345678ull) *sink = acc;
  }
}

int main(int argc, char** argv) {
  const int copy_bytes = argc > 1 ? atoi(argv[1]) : 8192, stages = argc > 2 ? atoi(argv[2]) : 3, producers = argc > 3 ? atoi(argv[3]) : 4;
  const int ctas_per_sm = argc > 4 ? atoi(argv[4]) : 1, mode = argc > 5 ? atoi(argv[5]) : 0;
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int G = argc > 6 ? atoi(argv[6]) : prop.multiProcessorCount * ctas_per_sm;   // optional: grid size override
  const int stage_bytes = mode ? 2 * copy_bytes : copy_bytes;
  const size_t per_prod = ((size_t)4 << 30) / ((size_t)G * producers) / stage_bytes * stage_bytes;   // ~4 GB total
  const size_t total = per_prod * G * producers;
  uint8_t *a, *b; unsigned long long* sink;
  cudaMalloc(&a, total); cudaMalloc(&b, total / 2 + 1024); cudaMalloc(&sink, 8);
  cudaMemset(a, 1, total); cudaMemset(b, 1, total / 2 + 1024);
  const size_t smem = (size_t)producers * stages * stage_bytes + 1024;
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int it = 0; it < 5; ++it) {
    cudaEventRecord(e0);
    stream_kernel<<<G, 256, smem>>>(a, b, per_prod, copy_bytes, stages, producers, mode, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  cudaError_t err = cudaGetLastError();
  printf("copy %6d B x%d stages x%d producers grid %d mode %d smem %zu KB: %.1f GB/s = %.1f GB/s per CTA (%s)\n", copy_bytes, stages, producers, G, mode,
         smem / 1024, total / (best * 1e-3) / 1e9, total / (best * 1e-3) / 1e9 / G, cudaGetErrorString(err));
  return 0;
}
