// Stream-K variant of the decode-step gate|up contraction with the fused SwiGLU epilogue (a4.3, HF LlamaMLP :182-184).
//
// Why: the SwiGLU epilogue is non-linear, so gemm_tc_kernel gives each 128-row weight tile (64 gate + 64 up features) its
// whole K range in ONE CTA: 2F / 128 tiles = 88 CTAs for Janus-1.3B (60 of the 148 SMs idle while 46 MB of weights
// stream; at the measured ~48 GB/s per producing CTA that caps the launch at 4.3 TB/s) and 172 CTAs = two waves, the
// second one 24 tiles wide, for Janus-Pro-7B.  Here the (tile, k-block) unit space is cut into one contiguous, equal
// range per SM.  A range covers the tail of one tile, possibly whole tiles, and the head of the next one:
//   * a fragment that does not contain its tile's last k-block is a PARTIAL: its fp32 accumulators go to an L2 scratch
//     tile and a per-tile arrival counter is bumped (release);
//   * the fragment with the tile's last k-block is the FINISHER: it adds the partials of the earlier CTAs in k order
//     (deterministic), then runs the SwiGLU combine and stores h[m][f] as bf16.
// A CTA processes its fragments in REVERSE range order (head of the last tile first, tail of the first tile last), so a
// partial is published long before its finisher - which only reaches that tile at the very end of its own range - asks
// for it.  Two TMEM accumulator sets alternate between fragments, so the epilogue of one fragment overlaps the main loop
// of the next; the weight / token rings never drain at a fragment boundary.
// Ten warps: W producers (even / odd k-blocks), X producers (even / odd), two MMA issuers (even / odd, own accumulator
// blocks, as gemm_tc_kernel), four epilogue warps.  Arithmetic per element is that of gemm_tc_kernel's SwiGLU epilogue.
#pragma once
#include "gemm.cuh"

namespace pg {

constexpr int SKG_THREADS = 320;
constexpr int SKG_XCH_BYTES = 2 * 128 * 9 * 4;          // SwiGLU exchange buffers (two, alternating per 16-column chunk)
constexpr int SKG_MAX_FRAGS = 4;

template <int NT>
struct SkgCfg {
  static constexpr int B_BYTES = NT * TC_BK * 2;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr int SET_COLS = 4 * NT;                // four accumulator blocks per set (two per issuer)
  static constexpr int TMEM_COLS = 2 * SET_COLS < 32 ? 32 : 2 * SET_COLS;
  static constexpr int smem_bytes(int stages) { return stages * STAGE_BYTES + SKG_XCH_BYTES + 1024 + 512; }
};

struct SkgFrag { int tile, kb_lo, kb_hi, base; };         // base: index of the fragment's first k-block in the CTA's sequence

// CTA `c` of `G` owns units [c*U/G, (c+1)*U/G) of the flat (tile, k-block) space
PG_DEVINL long skg_u0(int c, long U, int G) { return ((long)c * U) / G; }
PG_DEVINL int skg_owner(long u, long U, int G) {
  int c = (int)((u * G) / U);
  while (c + 1 < G && skg_u0(c + 1, U, G) <= u) ++c;
  while (c > 0 && skg_u0(c, U, G) > u) --c;
  return c;
}

template <int NT>
__global__ void __launch_bounds__(SKG_THREADS, 1)
gemm_swiglu_sk_kernel(const __grid_constant__ CUtensorMap map_x, const uint8_t* __restrict__ w_tiled, int M, int F, int n_tiles,
                      int num_kb, int num_stages, int use_pdl, float* __restrict__ scratch, int max_contrib,
                      int* __restrict__ counters, bf16* __restrict__ h_out, Prof prof) {
  using Cfg = SkgCfg<NT>;
  extern __shared__ uint8_t skg_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)skg_smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* xch = smem + num_stages * Cfg::STAGE_BYTES;
  uint64_t* bars = (uint64_t*)(xch + SKG_XCH_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + num_stages;
  uint64_t* acc_full = bars + 2 * num_stages;            // [2]
  uint64_t* acc_empty = bars + 2 * num_stages + 2;       // [2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * num_stages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (use_pdl) pdl_launch_dependents();
  prof_begin(prof);

  // ---- this CTA's fragments, in processing order (reverse range order)
  const long U = (long)n_tiles * num_kb;
  const int G = (int)gridDim.x, c = (int)blockIdx.x;
  const long u0 = skg_u0(c, U, G), u1 = skg_u0(c + 1, U, G);
  SkgFrag fr[SKG_MAX_FRAGS];
  int nfrag = 0, total = 0;
  if (u1 > u0) {
    const int t_first = (int)(u0 / num_kb), t_last = (int)((u1 - 1) / num_kb);
    for (int t = t_last; t >= t_first && nfrag < SKG_MAX_FRAGS; --t) {
      const long lo = (long)t * num_kb, hi = lo + num_kb;
      fr[nfrag].tile = t;
      fr[nfrag].kb_lo = (int)((u0 > lo ? u0 : lo) - lo);
      fr[nfrag].kb_hi = (int)((u1 < hi ? u1 : hi) - lo);
      fr[nfrag].base = total;
      total += fr[nfrag].kb_hi - fr[nfrag].kb_lo;
      ++nfrag;
    }
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    for (int i = 0; i < num_stages; ++i) { mbar_init(&full_bar[i], 2); mbar_init(&empty_bar[i], 1); }
    mbar_init(&acc_full[0], 2); mbar_init(&acc_full[1], 2);
    mbar_init(&acc_empty[0], 4); mbar_init(&acc_empty[1], 4);
    mbar_fence_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // k-block i of the CTA's sequence -> (fragment, kb)
  auto locate = [&](int i, int& j, int& kb) {
    j = 0;
    while (j + 1 < nfrag && fr[j + 1].base <= i) ++j;
    kb = fr[j].kb_lo + (i - fr[j].base);
  };

  if (warp < 2) {
    // ===================== weight producers: even / odd k-blocks, tile-major 16 KB bulk copies =====================
    if (lane == 0) {
      const uint64_t pol_w = policy_evict_first();
      for (int i = warp; i < total; i += 2) {
        int j, kb;
        locate(i, j, kb);
        const int s = i % num_stages;
        mbar_wait(&empty_bar[s], (((uint32_t)(i / num_stages)) & 1u) ^ 1u, 71);
        mbar_expect_tx(&full_bar[s], TC_A_BYTES);
        bulk_copy_g2s(smem + s * Cfg::STAGE_BYTES, w_tiled + ((size_t)fr[j].tile * num_kb + kb) * TC_A_BYTES, TC_A_BYTES, &full_bar[s], pol_w);
      }
    }
  } else if (warp < 4) {
    // ===================== token-tile producers =====================
    if (lane == 0) {
      const uint64_t pol_x = policy_evict_last();
      if (use_pdl) pdl_wait();
      for (int i = warp - 2; i < total; i += 2) {
        int j, kb;
        locate(i, j, kb);
        const int s = i % num_stages;
        mbar_wait(&empty_bar[s], (((uint32_t)(i / num_stages)) & 1u) ^ 1u, 72);
        mbar_expect_tx(&full_bar[s], Cfg::B_BYTES);
        tma_load_2d(smem + s * Cfg::STAGE_BYTES + TC_A_BYTES, &map_x, &full_bar[s], kb * TC_BK, 0, pol_x);
      }
    }
  } else if (warp < 6) {
    // ===================== MMA issuers: issuer p takes k-blocks i = p (mod 2), accumulator blocks 2p, 2p+1 of the set =====================
    if (lane == 0) {
      const int p = warp - 4;
      const uint32_t idesc = umma_idesc_bf16(TC_BM, NT);
      for (int j = 0; j < nfrag; ++j) {
        const int set = j & 1, b0 = fr[j].base, b1 = b0 + (fr[j].kb_hi - fr[j].kb_lo);
        // the epilogue must have drained this set (fragment j - 2) before it is overwritten
        mbar_wait(&acc_empty[set], (((uint32_t)(j >> 1)) & 1u) ^ 1u, 73);
        tc_fence_after();
        int first = b0 + ((b0 & 1) == p ? 0 : 1);
        bool any = false;
        for (int i = first; i < b1; i += 2) {
          const int s = i % num_stages;
          mbar_wait(&full_bar[s], ((uint32_t)(i / num_stages)) & 1u, 74);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint64_t da = umma_desc_k_sw128(a_addr);
          const uint64_t db = umma_desc_k_sw128(a_addr + TC_A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_bf16(tmem_base + (uint32_t)(set * Cfg::SET_COLS + (2 * p + (k & 1)) * NT), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                      (uint32_t)(any || k >= 2));
          umma_commit(&empty_bar[s]);
          any = true;
        }
        if (any) umma_commit(&acc_full[set]);
        else mbar_arrive(&acc_full[set]);              // no k-block of this fragment fell to this issuer
      }
    }
  } else {
    // ===================== epilogue warps (quarter = TMEM lane group) =====================
    const int quarter = warp & 3;
    const int t128 = threadIdx.x - 192;
    if (use_pdl) pdl_wait();
    int chunk_ctr = 0;                                    // SwiGLU exchange buffers alternate per 16-column chunk ACROSS fragments
    for (int j = 0; j < nfrag; ++j) {
      const int set = j & 1, b0 = fr[j].base, len = fr[j].kb_hi - fr[j].kb_lo, tile = fr[j].tile;
      const int cnt0 = (len + ((b0 & 1) == 0 ? 1 : 0)) / 2, cnt1 = len - cnt0;       // k-blocks of issuer 0 / 1
      const int blk0 = cnt0 > 0 ? 0 : 2, n_used = (cnt0 > 0 ? 2 : 0) + (cnt1 > 0 ? 2 : 0);
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * Cfg::SET_COLS + blk0 * NT);
      const bool finisher = fr[j].kb_hi == num_kb;
      mbar_wait(&acc_full[set], ((uint32_t)(j >> 1)) & 1u, 75);
      tc_fence_after();
      if (!finisher) {
        // ---- PARTIAL: accumulators -> scratch tile [tile][rank][NT][128], then publish
        const int rank = c - skg_owner((long)tile * num_kb, U, G);
        float* pt = scratch + (((size_t)tile * max_contrib + rank) * NT) * TC_BM + quarter * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 16) {
          uint32_t v[16];
          tmem_ld_acc_sum<4, NT>(taddr + (uint32_t)c0, v, n_used);
#pragma unroll
          for (int q = 0; q < 16; ++q) pt[(size_t)(c0 + q) * TC_BM] = __uint_as_float(v[q]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[set]);
        __threadfence();                                   // this thread's partial stores are visible before the arrival below
        asm volatile("bar.sync 3, 128;" ::: "memory");
        if (t128 == 0) atomicAdd(counters + tile, 1);
      } else {
        // ---- FINISHER: earlier contributors' partials (k order) + own accumulators, SwiGLU, store h
        const int nc = fr[j].kb_lo > 0 ? c - skg_owner((long)tile * num_kb, U, G) : 0;
        if (nc > 0) {
          if (t128 == 0) {
            uint32_t spins = 0;
            uint64_t t0 = 0;
            for (;;) {
              int got;
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(got) : "l"(counters + tile) : "memory");
              if (got >= nc) break;
              if ((++spins & 0xFFFu) == 0) {
                const uint64_t now = global_timer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 4000000000ull) { printf("plangen_b200: stream-K partial wait timed out (cta %d tile %d)\n", c, tile); __trap(); }
              }
            }
            counters[tile] = 0;                            // re-arm for the next launch / graph replay
          }
          asm volatile("bar.sync 3, 128;" ::: "memory");
        }
        const float* ps = scratch + ((size_t)tile * max_contrib * NT) * TC_BM + quarter * 32 + lane;
        const uint32_t xch_s = smem_u32(xch);
        const int pair = quarter & 1;                           // f-block within the tile
        const int f = tile * 64 + pair * 32 + lane;
        const bool is_gate = quarter < 2;
        const int jbase = is_gate ? 0 : 8;                      // columns this warp finishes
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 16) {
          uint32_t v[16];
          float add[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) add[q] = 0.f;
          if (nc > 0) {
            // contributor 0 holds the first k-blocks of the tile: sum in rank order, the finisher's own range last
            float pv[2][16];
            for (int r = 0; r < nc; r += 2) {
#pragma unroll
              for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                for (int q = 0; q < 16; ++q)
                  pv[rr][q] = (r + rr < nc) ? __ldcg(ps + ((size_t)(r + rr) * NT + c0 + q) * TC_BM) : 0.f;
#pragma unroll
              for (int q = 0; q < 16; ++q) { add[q] += pv[0][q]; add[q] += pv[1][q]; }
            }
          }
          tmem_ld_acc_sum<4, NT>(taddr + (uint32_t)c0, v, n_used);
          if (c0 + 16 >= NT) {                                   // last TMEM read of this fragment: the set may be reused
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[set]);
          }
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = __float_as_uint(add[q] + __uint_as_float(v[q]));
          // gate warps (quarters 0, 1) pair with the up warps two quarters above; per 16-column chunk the gate warp hands
          // g of columns 8..15 to the up warp and receives u of columns 0..7 (see gemm_tc_kernel)
          // (a thread may only write buffer b again after the barrier of the chunk in between, which every reader of
          // b has passed - also when the previous chunk belonged to the previous fragment)
          const uint32_t buf = xch_s + (uint32_t)(chunk_ctr++ & 1) * (128u * 9u * 4u);
          const uint32_t mine = buf + (uint32_t)(quarter * 32 + lane) * 36u;
          const uint32_t theirs = buf + (uint32_t)((quarter ^ 2) * 32 + lane) * 36u;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(mine + 4u * q), "r"(v[(is_gate ? 8 : 0) + q]) : "memory");
          asm volatile("bar.sync 2, 128;" ::: "memory");
          float h[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            uint32_t ow;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(ow) : "r"(theirs + 4u * q) : "memory");
            const float own = __uint_as_float(v[jbase + q]), other = __uint_as_float(ow);
            const float g = bf16_round(is_gate ? own : other);
            const float u = bf16_round(is_gate ? other : own);
            float ex, rc;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-g * 1.4426950408889634f));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(1.0f + ex));
            const float sg = bf16_round(g * rc);
            h[q] = sg * u;
          }
          if (f < F) {
            bf16* dst = h_out + (size_t)(c0 + jbase) * F + f;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (c0 + jbase + q < M) dst[(size_t)q * F] = __float2bfloat16_rn(h[q]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  prof_end(prof);
  if (warp == 4) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace pg
