// Wide-tile contraction on a CTA PAIR (tcgen05 cta_group::2):  C[m][n] = sum_k X[m][k] * W[n][k]  for the tensor-bound
// shapes (prompt prefill, SigLIP tower, aligner, VQ 1x1 convolutions): token tiles of 256 rows, no split-K.
//
// Why: gemm_tc_kernel<256> tops out at 52-56 % of the tensor pipe (profiles/r02_gemm_prefill.full.txt).  At full rate one
// SM's MMAs read their A (128 x 16) and B (256 x 16) operand slices from shared memory at 96 B/clk while TMA refills the
// ring at another 96 B/clk - against a 128 B/clk shared-memory port.  Two SMs of a TPC issuing ONE 256 x 256 x 16 MMA
// between them (swap-AB: 256 weight rows x 256 tokens) each hold their own 128 weight rows (A) and only HALF of the
// token tile (B; the tensor cores fetch the other half from the peer's shared memory): per SM 32 KB instead of 48 KB per
// k-block land in shared memory, and B is read once per pair instead of once per SM.
//
// Structure per CTA (cluster of 2 along the weight-tile axis, rank 0 = leader):
//   warp 0 (one thread)  TMA producer: its A tile + its half of the token tile, 2-SM loads whose completion bytes are
//                        counted on the LEADER's full barrier (the leader alone posts the expected byte count);
//   warp 1 (one thread)  leader only: tcgen05.mma.cta_group::2, commits multicast to both CTAs' empty / accumulator
//                        barriers; both CTAs: cta_group::2 TMEM allocation;
//   warps 2-5            epilogue of the CTA's own 128 weight rows x 256 tokens (accumulator in its own TMEM): fp32
//                        partial store, or the fused row epilogues of gemm.cuh (EpiFuse).
#pragma once
#include "gemm.cuh"

namespace pg {

constexpr int TC2_NT = 256;                        // tokens per pair tile (MMA N)
constexpr int TC2_BHALF = (TC2_NT / 2) * TC_BK * 2;  // one CTA's half of the token tile: 16 KB
constexpr int TC2_STAGE_BYTES = TC_A_BYTES + TC2_BHALF;
constexpr uint32_t TC2_PEER_MASK = 0xFEFFFFFFu;    // clears the CTA-rank bit of a shared::cluster address: CTA 0's copy
static constexpr int tc2_smem_bytes(int stages) { return stages * TC2_STAGE_BYTES + 1024 + 256; }

PG_DEVINL void tma2_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar) & TC2_PEER_MASK), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
PG_DEVINL void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs of this thread completed) on the barrier at this offset in BOTH CTAs of the pair
PG_DEVINL void umma2_commit_mc(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
PG_DEVINL void tmem2_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
PG_DEVINL void tmem2_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
PG_DEVINL void tmem2_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, float* __restrict__ C, int M, int N,
                int K, int num_stages, int use_pdl, Prof prof, EpiFuse ep) {
  extern __shared__ uint8_t smem_raw2[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw2 + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + num_stages * TC2_STAGE_BYTES);
  uint64_t* full_bar = bars;                       // used in the leader only
  uint64_t* empty_bar = bars + num_stages;
  uint64_t* tmem_full_bar = bars + 2 * num_stages;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * num_stages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n0 = blockIdx.x * TC_BM;               // this CTA's weight rows
  const int m0 = blockIdx.y * TC2_NT;              // the pair's token tile
  const int nkb = (K + TC_BK - 1) / TC_BK;

  if (use_pdl) pdl_launch_dependents();
  prof_begin(prof);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_x);
    for (int i = 0; i < num_stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem2_alloc(tmem_slot, 256);
    tmem2_relinquish();
  }
  tc_fence_before();
  cluster_sync_all();                               // barriers of both CTAs initialised before any remote completion / commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (one thread per CTA) =====================
    if (lane == 0) {
      const uint64_t pol_w = policy_evict_last(), pol_x = policy_evict_last();   // both operands are re-read by other tiles
      if (use_pdl) pdl_wait();
      for (int i = 0; i < nkb; ++i) {
        const int s = i % num_stages;
        mbar_wait(&empty_bar[s], (((uint32_t)(i / num_stages)) & 1u) ^ 1u, 81);
        // the leader posts the bytes of BOTH CTAs' loads of this stage; the peer's loads count on the same barrier
        if (leader) mbar_expect_tx(&full_bar[s], 2 * TC2_STAGE_BYTES);
        tma2_load_2d(smem + s * TC2_STAGE_BYTES, &map_w, &full_bar[s], i * TC_BK, n0, pol_w);
        tma2_load_2d(smem + s * TC2_STAGE_BYTES + TC_A_BYTES, &map_x, &full_bar[s], i * TC_BK, m0 + (int)rank * (TC2_NT / 2), pol_x);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only, one thread) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(2 * TC_BM, TC2_NT);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % num_stages;
        mbar_wait(&full_bar[s], ((uint32_t)(i / num_stages)) & 1u, 82);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * TC2_STAGE_BYTES);
        const uint64_t da = umma_desc_k_sw128(a_addr);
        const uint64_t db = umma_desc_k_sw128(a_addr + TC_A_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k)
          umma2_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((i | k) != 0));
        umma2_commit_mc(&empty_bar[s]);             // frees the stage in both CTAs once these MMAs retire
      }
      umma2_commit_mc(tmem_full_bar);               // accumulators of both CTAs complete
    }
  } else {
    // ===================== epilogue: own 128 weight rows x 256 tokens =====================
    const int quarter = warp & 3;
    const int n = n0 + quarter * 32 + lane;
    if (use_pdl) pdl_wait();
    mbar_wait(tmem_full_bar, 0, 83);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    if (ep.out != nullptr) {
      const float bias_n = (ep.bias && n < N) ? ep.bias[n] : 0.f;
      const uint32_t stg = smem_u32(smem);                   // [256][128] bf16 = 64 KB over the (now idle) ring
      const int nl = quarter * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < TC2_NT; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float t = bf16_round(__uint_as_float(v[j]) + bias_n);
          if (ep.gelu) t = t * 0.5f * (1.0f + erf_fast(t * 0.70710678118654752440f));
          const unsigned short hb = __bfloat16_as_ushort(__float2bfloat16_rn(t));
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(stg + (uint32_t)((c0 + j) * 256 + nl * 2)), "h"(hb) : "memory");
        }
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      const int t128 = threadIdx.x - 64;
      const int ch = t128 & 15, r0 = t128 >> 4;
      if (n0 + ch * 8 < N) {
#pragma unroll 4
        for (int r = r0; r < TC2_NT; r += 8) {
          const size_t m = (size_t)m0 + r;
          if (m < (size_t)M) {
            uint4 q;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                         : "r"(stg + (uint32_t)(r * 256 + ch * 16)));
            if (ep.add16 != nullptr) {
              const uint4 a = *reinterpret_cast<const uint4*>(ep.add16 + m * N + n0 + ch * 8);
              uint32_t qs[4] = {q.x, q.y, q.z, q.w};
              const uint32_t as[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const __nv_bfloat162 o = __floats2bfloat162_rn(bf16lo(qs[u]) + bf16lo(as[u]), bf16hi(qs[u]) + bf16hi(as[u]));
                qs[u] = *reinterpret_cast<const uint32_t*>(&o);
              }
              q = make_uint4(qs[0], qs[1], qs[2], qs[3]);
            }
            *reinterpret_cast<uint4*>(ep.out + m * N + n0 + ch * 8) = q;
          }
        }
      }
    } else if (ep.resid != nullptr) {
      const float bias_n = (ep.bias && n < N) ? ep.bias[n] : 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < TC2_NT; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        if (n < N) {
          float xo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = m0 + c0 + j;
            xo[j] = (m < M) ? ep.resid[(size_t)m * N + n] : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = m0 + c0 + j;
            if (m < M) ep.resid[(size_t)m * N + n] = xo[j] + bf16_round(__uint_as_float(v[j]) + bias_n);
          }
        }
      }
    } else {
#pragma unroll 1
      for (int c0 = 0; c0 < TC2_NT; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        if (n < N) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = m0 + c0 + j;
            if (m < M) C[(size_t)m * N + n] = __uint_as_float(v[j]);
          }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();                               // nobody leaves (or frees TMEM) while the pair can still touch it
  prof_end(prof);
  if (warp == 1) tmem2_dealloc(tmem_base, 256);
}

}  // namespace pg
