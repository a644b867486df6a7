// Linear-layer contractions of the decode path:  C[s][m][n] = sum_{k in split s} X[m][k] * W[n][k]
// (X = activations, K contiguous; W = nn.Linear weight [out, in], K contiguous).  Output is always
// fp32 split-K partials [splits][M][N]; the consumer kernel (norm / rope / swiglu / gelu / sampler)
// reduces the splits, so the reduction order is fixed and the result deterministic.
//
//   gemm_tc_kernel   : bf16 operands on the 5th-gen tensor cores.  "swap-AB": the WEIGHT tile is
//                      the MMA A operand (M = 128 weight rows) and the token tile is the MMA B
//                      operand (N = NT tokens, 16..256), so a decode step with R = 2..128 rows
//                      streams weights at full TMA rate without padding rows to 128.
//                      TMA (SWIZZLE_128B) -> smem ring -> tcgen05.mma (fp32 accum in TMEM) ->
//                      tcgen05.ld epilogue.  Warp-specialised: warp0 TMA, warp1 MMA, warps2-5 epilogue.
//                      With PDL the weight tiles of all stages are requested BEFORE
//                      griddepcontrol.wait, so the HBM stream does not drain at kernel boundaries.
//   gemm_simt_kernel : fp32 (check mode) / bf16 CUDA-core fallback used for parity tests and for
//                      shapes the TMA path does not take (K % 8 != 0).
#pragma once
#include "common.cuh"

namespace pg {

// ------------------------------------------------------------------------ tcgen05 path
constexpr int TC_BM = 128;   // weight rows per CTA (MMA M)
constexpr int TC_BK = 64;    // bf16 elements per k-block = one 128-byte swizzle row
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;

template <int NT>
struct TcCfg {
  static constexpr int B_BYTES = NT * TC_BK * 2;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = NT <= 32 ? 32 : NT <= 64 ? 64 : NT <= 128 ? 128 : 256;
  static constexpr int smem_bytes(int stages) { return stages * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/; }
};

template <int NT>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
               float* __restrict__ C, int M, int N, int K, int kb_per_split, int num_stages, int use_pdl,
               const uint8_t* __restrict__ w_tiled, Prof prof, bf16* __restrict__ swiglu_out) {
  using Cfg = TcCfg<NT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + num_stages * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + num_stages;
  uint64_t* tmem_full_bar = bars + 2 * num_stages;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * num_stages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * TC_BM;
  const int m0 = blockIdx.y * NT;
  const int split = blockIdx.z;
  const int num_kb = (K + TC_BK - 1) / TC_BK;
  const int kb_begin = split * kb_per_split;
  const int kb_end = min(kb_begin + kb_per_split, num_kb);
  const int nkb = max(kb_end - kb_begin, 0);

  if (use_pdl & 1) pdl_launch_dependents();
  prof_begin(prof);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_x);
    for (int i = 0; i < num_stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      const uint64_t pol_w = policy_evict_first();   // weights are streamed once per step
      const uint64_t pol_x = policy_evict_last();    // activations are re-read by every CTA
      const int pre = min(nkb, num_stages);
      // use_pdl bit 1: the W operand is a constant weight, so its tiles may be requested before the
      // previous kernel has finished (it is an activation in the VQ attention contractions)
      // A tile source: TMA tensor load from the row-major weight, or - when the weight was packed tile-major
      // and pre-swizzled (w_tiled) - ONE contiguous 16 KB bulk copy per tile (full DRAM bursts, one issue).
      auto load_a = [&](int stage, int kb) {
        if (w_tiled)
          bulk_copy_g2s(smem + stage * Cfg::STAGE_BYTES, w_tiled + ((size_t)blockIdx.x * num_kb + kb) * TC_A_BYTES, TC_A_BYTES,
                        &full_bar[stage], pol_w);
        else
          tma_load_2d(smem + stage * Cfg::STAGE_BYTES, &map_w, &full_bar[stage], kb * TC_BK, n0, pol_w);
      };
      if ((use_pdl & 3) == 1) pdl_wait();
      for (int i = 0; i < pre; ++i) {
        mbar_expect_tx(&full_bar[i], Cfg::STAGE_BYTES);
        load_a(i, kb_begin + i);
      }
      if ((use_pdl & 3) == 3) pdl_wait();
      for (int i = 0; i < pre; ++i)
        tma_load_2d(smem + i * Cfg::STAGE_BYTES + TC_A_BYTES, &map_x, &full_bar[i], (kb_begin + i) * TC_BK, m0, pol_x);
      for (int i = pre; i < nkb; ++i) {
        const int s = i % num_stages;
        const uint32_t round = (uint32_t)(i / num_stages);
        mbar_wait(&empty_bar[s], (round & 1u) ^ 1u, 1);
        mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        load_a(s, kb_begin + i);
        tma_load_2d(smem + s * Cfg::STAGE_BYTES + TC_A_BYTES, &map_x, &full_bar[s], (kb_begin + i) * TC_BK, m0, pol_x);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(TC_BM, NT);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % num_stages;
        const uint32_t round = (uint32_t)(i / num_stages);
        mbar_wait(&full_bar[s], round & 1u, 2);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint64_t da = umma_desc_k_sw128(a_addr);
        const uint64_t db = umma_desc_k_sw128(a_addr + TC_A_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k)   // UMMA_K = 16 bf16 = 32 bytes = +2 in the encoded address
          umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((i | k) != 0));
        umma_commit(&empty_bar[s]);            // frees the smem stage once these MMAs retire
      }
      umma_commit(tmem_full_bar);              // accumulator complete -> epilogue
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global (4 warps) =====================
    const int quarter = warp & 3;              // a warp may only touch TMEM lanes 32*(warp%4)..+31
    const int n = n0 + quarter * 32 + lane;
    float* out = C + (size_t)split * M * N;
    if (use_pdl & 1) pdl_wait();
    if (swiglu_out != nullptr) {
      // Fused SwiGLU epilogue (no split-K): the weight rows are interleaved in blocks of 64, so lanes 0-63 of
      // the tile hold gate(f0 .. f0+63) and lanes 64-127 hold up(f0 .. f0+63).  The two "up" warps park
      // their accumulators in the (now idle) first pipeline stage, the two "gate" warps combine
      //   h = rnd(rnd(silu(rnd(g))) * rnd(u))        (HF LlamaMLP :182-184, bf16 rounding points of autocast)
      // and store h[m][f] directly as the bf16 operand of the down projection.
      mbar_wait(tmem_full_bar, 0, 3);
      tc_fence_after();
      float* xch = reinterpret_cast<float*>(smem);            // [64 lanes][17] fp32 per 16-column chunk
      const int F = N / 2;
      const int f = blockIdx.x * 64 + (quarter & 1) * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        if (quarter >= 2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) xch[((quarter - 2) * 32 + lane) * 17 + j] = __uint_as_float(v[j]);
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (quarter < 2 && f < F) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = m0 + c0 + j;
            if (m < M) {
              const float g = bf16_round(__uint_as_float(v[j]));
              const float u = bf16_round(xch[(quarter * 32 + lane) * 17 + j]);
              const float sg = bf16_round(g / (1.0f + expf(-g)));
              swiglu_out[(size_t)m * F + f] = __float2bfloat16_rn(sg * u);
            }
          }
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
      }
    } else if (nkb > 0) {
      mbar_wait(tmem_full_bar, 0, 3);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        if (n < N) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = m0 + c0 + j;
            if (m < M) out[(size_t)m * N + n] = __uint_as_float(v[j]);   // 32 lanes -> 128 B contiguous
          }
        }
      }
    } else if (n < N) {
      for (int j = 0; j < NT; ++j) {
        const int m = m0 + j;
        if (m < M) out[(size_t)m * N + n] = 0.f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  prof_end(prof);
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// --------------------------------------------------------------------------- SIMT path
template <typename T> struct Vec4Ld;
template <> struct Vec4Ld<float> {
  static PG_DEVINL void ld(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <> struct Vec4Ld<bf16> {
  static PG_DEVINL void ld(const bf16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    v[0] = bf16lo(t.x); v[1] = bf16hi(t.x); v[2] = bf16lo(t.y); v[3] = bf16hi(t.y);
  }
};

// 64x64 output tile, BK = 16, 256 threads, 4x4 micro-tile, fp32 accumulate (k ascending).
template <typename T>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ X, const T* __restrict__ W, float* __restrict__ C, int M, int N, int K,
                 int k_per_split) {
  __shared__ float Xs[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  pdl_launch_dependents();
  pdl_wait();          // every kernel of a PDL chain must wait, or completion is no longer transitive
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  const int lr = tid >> 2, lk = (tid & 3) * 4;       // loader: row within tile, k offset
  const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4; // compute: micro-tile origin
  float acc[4][4] = {};
  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
    float xv[4] = {0, 0, 0, 0}, wv[4] = {0, 0, 0, 0};
    const int k = k0 + lk;
    if (k < k_end) {
      if (m0 + lr < M) Vec4Ld<T>::ld(X + (size_t)(m0 + lr) * K + k, xv);
      if (n0 + lr < N) Vec4Ld<T>::ld(W + (size_t)(n0 + lr) * K + k, wv);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { Xs[lk + j][lr] = xv[j]; Ws[lk + j][lr] = wv[j]; }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = Xs[kk][tm + i]; b[i] = Ws[kk][tn + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = C + (size_t)blockIdx.z * M * N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tm + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn + j;
      if (n < N) out[(size_t)m * N + n] = acc[i][j];
    }
  }
}

}  // namespace pg
