// One decode step of the LM (all layers) as ONE persistent kernel: 1 CTA per SM, no launches between ops.
//
// Why (B200): a 1.3B decode step moves ~2.5 GB of weights + the KV cache and does almost no math, so
// the only resource that matters is the HBM stream.  With one kernel per op the stream drains at every
// kernel boundary (~8 per layer x 24 layers) and each op pays its own launch / prologue / epilogue
// latency.  Here every SM runs a warp-specialised CTA for the whole step:
//
//   warp 16  weight producer : streams THIS CTA's share of every contraction's weight tiles (TMA tensor
//                         loads, SWIZZLE_128B) through a ring of 16 KB stages, in program order, for ALL
//                         layers.  Weights do not depend on the step's activations, so this warp never waits
//                         for a grid barrier: while the others wait for a dependency it keeps HBM busy.
//   warp 17  operand producer: streams the small activation operand tiles (L2-resident) of the
//                         contractions through a second ring, after the grid barrier of the producing phase.
//   warp 18  MMA issuer : tcgen05.mma (swap-AB: weight tile = M 128, token tile = N 32), fp32 accumulators
//                         double-buffered in TMEM.
//   warp 19  K/V producers (lanes 0-3, one per attention group): the cached K/V tiles of the group's share
//                         of the attention, TMA bulk copies into the group's private stages; cached tokens are
//                         step-independent too, so they are prefetched during the QKV contraction.
//   warps 0-15 workers  : TMEM epilogues (warps 0-3), the KV-cache attention (4 independent groups x 4 warps,
//                         attn_tma.cuh), split-K reductions fused with residual + RMSNorm and SwiGLU.
//
// Phases per layer (grid barrier after each; arithmetic identical to the per-op kernels):
//   0 QKV contraction -> 1 attention (+RoPE, KV append) -> 2 O contraction -> 3 residual + RMSNorm
//   -> 4 gate|up contraction -> 5 SwiGLU -> 6 down contraction -> 7 residual + RMSNorm (next layer / final)
#pragma once
#include "common.cuh"
#include "lm_kernels.cuh"
#include "attn_tma.cuh"

namespace pg {

constexpr int SK_NSA = 4;                       // weight ring: stages of 16 KB
constexpr int SK_NSB = 4;                       // operand ring: activation tiles
constexpr int SK_SPG = 2;                       // K/V stages per attention group
constexpr int SK_NKV = AT_NG * SK_SPG;          // K/V ring stages of 16 KB
constexpr int SK_NT = 32;                       // token tile (rows R <= 32)
constexpr int SK_WORKERS = 16;
constexpr int SK_WTHREADS = SK_WORKERS * 32;
constexpr int SK_THREADS = 32 * (SK_WORKERS + 4);
constexpr int SK_A_BYTES = 16384;
constexpr int SK_B_BYTES = SK_NT * TC_BK * 2;   // 4 KB
constexpr int SK_SMEM = (SK_NSA + SK_NKV) * SK_A_BYTES + SK_NSB * SK_B_BYTES + 1024;
constexpr int SK_PHASES = 8;
constexpr int SK_MAXS = 10;                     // split-K slabs per contraction in the step kernel (sched_for cap)
constexpr int SK_RNK = 8;                       // RMSNorm elements per worker thread: D <= 8 * 512
static_assert(SK_A_BYTES == 2 * AT_TILE_BYTES && SK_A_BYTES == TC_A_BYTES, "one ring serves weights and K/V");

struct GemmSched { int n_tiles, splits, kb_per_split, num_kb; };

struct StepParams {
  int R, H, D, HD, F, L, Tmax;
  float eps, scale;
  const CUtensorMap* wmaps;      // [L][4]: qkv, o, gu, d      (global memory)
  const CUtensorMap* amaps;      // [3]: xn [R,D], attn_out [R,HD], h [R,F]
  const float* ln;               // [L][2][D]  input_layernorm, post_attention_layernorm
  const float* norm_w;           // [D] final norm
  float* x;                      // residual stream fp32 [R, D]
  bf16 *xn, *attn_out, *h, *hidden_t;
  float* hidden_f;
  float *part_qkv, *part_o, *part_gu, *part_d;
  bf16* kv;                      // [L][2][R][H][Tmax][128]
  const int32_t* kv_start;
  const float *cosT, *sinT;
  float* attn_ws;
  int* attn_cnt;
  unsigned long long* grid_bar;  // monotonic arrival counter
  unsigned long long* epoch;     // number of step-kernel launches so far (device side)
  int pos_base;
  int* step_ptr;                 // nullable: pos = pos_base + *step_ptr
  int inc_step;
  GemmSched g_qkv, g_o, g_gu, g_d;
  unsigned long long* prof;      // nullable: [gridDim][L*8+1] phase-end timestamps (ns) of each CTA's worker thread 0
};

PG_DEVINL unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
PG_DEVINL void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// spin until the grid-wide arrival counter reaches `target` (bounded: trap instead of hanging the GPU)
PG_DEVINL void grid_wait(const unsigned long long* bar, unsigned long long target, int site) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (ld_acquire_u64(bar) < target) {
    if ((++spins & 0xFFFu) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {
        printf("plangen_b200: grid barrier timed out (site %d, block %d thread %d, have %llu want %llu)\n", site,
               blockIdx.x, threadIdx.x, ld_acquire_u64(bar), target);
        __trap();
      }
    }
  }
}


__global__ void __launch_bounds__(SK_THREADS, 1)
decode_step_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ringA = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ringKV = ringA + SK_NSA * SK_A_BYTES;
  uint8_t* ringB = ringKV + SK_NKV * SK_A_BYTES;
  __shared__ uint64_t fullA[SK_NSA], emptyA[SK_NSA], fullB[SK_NSB], emptyB[SK_NSB], tmem_full[2], tmem_empty[2];
  __shared__ uint64_t kv_full[SK_NKV], kv_empty[SK_NKV];
  __shared__ uint32_t tmem_slot;
  __shared__ int row_units[AT_MAX_ROWS + 1];
  __shared__ int stage_tab[AT_NG * SK_SPG];
  __shared__ AttnGroupSmem gsm[AT_NG];
  __shared__ float red_s[32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, c = blockIdx.x;
  const int R = p.R, H = p.H, D = p.D, HD = p.HD, F = p.F, L = p.L;
  const int pos = p.pos_base + (p.step_ptr ? *p.step_ptr : 0);
  const unsigned long long epoch = *p.epoch;
  const unsigned long long bar_base = epoch * (unsigned long long)(SK_PHASES * L) * (unsigned long long)G;
  auto bar_target = [&](int phase_index) { return bar_base + (unsigned long long)(phase_index + 1) * (unsigned long long)G; };

  if (tid == 0) {
    for (int i = 0; i < SK_NSA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < SK_NSB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    for (int i = 0; i < SK_NKV; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], AT_GW); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    mbar_fence_init();
  }
  if (tid < SK_NKV) stage_tab[(tid % AT_NG) * SK_SPG + tid / AT_NG] = tid;     // group g owns K/V stages g, g+4
  if (warp == SK_WORKERS + 2) { tmem_alloc(&tmem_slot, 2 * SK_NT); tmem_relinquish(); }
  if (warp == 0) build_row_units(row_units, p.kv_start, R, pos, lane);     // attention schedule (same for every layer)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  AttnCut cut;
  cut.U = row_units[R] * H;
  cut.per = max(1, (cut.U + G - 1) / G);
  cut.sub = (cut.per + AT_NG - 1) / AT_NG;
  const size_t kv_layer = (size_t)2 * R * H * p.Tmax * HEAD_DIM;     // elements per layer (k then v)
  const size_t kv_half = (size_t)R * H * p.Tmax * HEAD_DIM;

  // contraction items of this CTA: it = c, c + G, ...; tile = it / splits, split = it % splits
  struct Item { int tile, split, kb0, kb1; };
  auto item_of = [&](const GemmSched& g, int it) {
    Item r;
    r.tile = it / g.splits; r.split = it % g.splits;
    r.kb0 = r.split * g.kb_per_split; r.kb1 = min(r.kb0 + g.kb_per_split, g.num_kb);
    return r;
  };

  if (warp == SK_WORKERS) {
    // =========================================================== weight producer
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      int ja = 0;
      auto stream_weights = [&](const GemmSched& g, const CUtensorMap* map) {
        for (int it = c; it < g.n_tiles * g.splits; it += G) {
          const Item im = item_of(g, it);
          for (int kb = im.kb0; kb < im.kb1; ++kb, ++ja) {
            const int s = ja % SK_NSA;
            mbar_wait(&emptyA[s], (((uint32_t)(ja / SK_NSA)) & 1u) ^ 1u, 21, ja);
            mbar_expect_tx(&fullA[s], SK_A_BYTES);
            tma_load_2d(ringA + (size_t)s * SK_A_BYTES, map, &fullA[s], kb * TC_BK, im.tile * TC_BM, pol);
          }
        }
      };
      for (int l = 0; l < L; ++l) {
        const CUtensorMap* wm = p.wmaps + (size_t)l * 4;
        stream_weights(p.g_qkv, wm + 0);
        stream_weights(p.g_o, wm + 1);
        stream_weights(p.g_gu, wm + 2);
        stream_weights(p.g_d, wm + 3);
      }
    }
  } else if (warp == SK_WORKERS + 1) {
    // =========================================================== operand producer: activation tiles
    if (lane == 0) {
      const uint64_t pol = policy_evict_last();
      int jb = 0;
      auto stream_act = [&](const GemmSched& g, const CUtensorMap* map, int phase_index) {
        if (phase_index > 0) grid_wait(p.grid_bar, bar_target(phase_index - 1), 31);
        fence_proxy_async_global();                            // generic-proxy stores of other CTAs -> TMA reads
        for (int it = c; it < g.n_tiles * g.splits; it += G) {
          const Item im = item_of(g, it);
          for (int kb = im.kb0; kb < im.kb1; ++kb, ++jb) {
            const int s = jb % SK_NSB;
            mbar_wait(&emptyB[s], (((uint32_t)(jb / SK_NSB)) & 1u) ^ 1u, 22, jb);
            mbar_expect_tx(&fullB[s], SK_B_BYTES);
            tma_load_2d(ringB + (size_t)s * SK_B_BYTES, map, &fullB[s], kb * TC_BK, 0, pol);
          }
        }
      };
      for (int l = 0; l < L; ++l) {
        stream_act(p.g_qkv, p.amaps + 0, l * SK_PHASES + 0);
        stream_act(p.g_o, p.amaps + 1, l * SK_PHASES + 2);
        stream_act(p.g_gu, p.amaps + 0, l * SK_PHASES + 4);
        stream_act(p.g_d, p.amaps + 2, l * SK_PHASES + 6);
      }
    }
  } else if (warp == SK_WORKERS + 2) {
    // =========================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(TC_BM, SK_NT);
      int ja = 0, jb = 0, ji = 0;                              // ring positions and accumulator-buffer counter
      auto run_gemm = [&](const GemmSched& g) {
        for (int it = c; it < g.n_tiles * g.splits; it += G) {
          const Item im = item_of(g, it);
          const int buf = ji & 1;
          mbar_wait(&tmem_empty[buf], (((uint32_t)(ji >> 1)) & 1u) ^ 1u, 23, ji);   // epilogue drained this buffer
          tc_fence_after();
          const uint32_t tacc = tmem_base + (uint32_t)(buf * SK_NT);
          for (int kb = im.kb0; kb < im.kb1; ++kb, ++ja, ++jb) {
            const int sb = jb % SK_NSB, sa = ja % SK_NSA;
            mbar_wait(&fullB[sb], (uint32_t)(jb / SK_NSB) & 1u, 24, jb);
            mbar_wait(&fullA[sa], (uint32_t)(ja / SK_NSA) & 1u, 25, ja);
            tc_fence_after();
            const uint64_t da = umma_desc_k_sw128(smem_u32(ringA + (size_t)sa * SK_A_BYTES));
            const uint64_t db = umma_desc_k_sw128(smem_u32(ringB + (size_t)sb * SK_B_BYTES));
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)
              umma_bf16(tacc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((kb > im.kb0) | (k > 0)));
            umma_commit(&emptyA[sa]);
            umma_commit(&emptyB[sb]);
          }
          umma_commit(&tmem_full[buf]);
          ++ji;
        }
      };
      for (int l = 0; l < L; ++l) {
        run_gemm(p.g_qkv);
        run_gemm(p.g_o);
        run_gemm(p.g_gu);
        run_gemm(p.g_d);
      }
    }
  } else if (warp == SK_WORKERS + 3) {
    // =========================================================== K/V producers: lane g feeds attention group g
    if (lane < AT_NG) {
      const int g = lane;
      int gb, ge;
      cut.group_range(c, g, gb, ge);
      int kload = 0;
      const uint64_t pol = policy_evict_first();
      for (int l = 0; l < L; ++l) {
        const bf16* kc = p.kv + (size_t)l * kv_layer;
        attn_produce_group<SK_SPG>(gb, ge, row_units, R, H, p.Tmax, p.kv_start, kc, kc + kv_half, ringKV, SK_A_BYTES,
                                   stage_tab + g * SK_SPG, kv_full, kv_empty, kload, SK_A_BYTES, pol);
      }
    }
  } else {
    // =========================================================== workers (16 warps)
    int ji = 0;                                                // mirror of the accumulator-buffer counter
    int kc_att = 0;                                            // K/V tiles consumed by this warp's attention group
    int prof_i = 0;
    unsigned long long* prof = p.prof ? p.prof + (size_t)c * (SK_PHASES * L + 1) : nullptr;
    if (prof && tid == 0) prof[prof_i++] = global_timer_ns();
    auto phase_done = [&]() {                                  // all of this CTA's global writes of the phase are done
      fence_proxy_async_global();                              // generic-proxy stores -> later TMA (async-proxy) reads
      __threadfence();                                         // (also invalidates L1: next phase re-reads from L2)
      named_bar_sync(1, SK_WTHREADS);
      if (tid == 0) {
        atomicAdd(p.grid_bar, 1ull);
        if (prof) prof[prof_i++] = global_timer_ns();
      }
    };
    auto phase_wait = [&](int phase_index) {                   // previous phase complete on every CTA
      if (tid == 0) grid_wait(p.grid_bar, bar_target(phase_index - 1), 32);
      named_bar_sync(1, SK_WTHREADS);
    };
    // TMEM epilogue of one contraction phase: warps 0-3, partials -> out[split][m][n]
    auto epilogue = [&](const GemmSched& g, float* out, int N) {
      for (int it = c; it < g.n_tiles * g.splits; it += G) {
        const Item im = item_of(g, it);
        const int buf = ji & 1;
        if (warp < 4) {
          mbar_wait(&tmem_full[buf], ((uint32_t)(ji >> 1)) & 1u, 26, ji);
          tc_fence_after();
          const int n = im.tile * TC_BM + warp * 32 + lane;
          float* o = out + (size_t)im.split * R * N;
#pragma unroll
          for (int c0 = 0; c0 < SK_NT; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * SK_NT + c0), v);
            tmem_ld_wait();
            if (n < N) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < R) o[(size_t)(c0 + j) * N + n] = __uint_as_float(v[j]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        }
        ++ji;
      }
    };
    // fixed-order split-K sum with L2 (cache-global) loads: the slabs were written by other CTAs

    // residual + RMSNorm of row `c` (CTAs >= R idle): x += rnd(sum part); xn = w * (x * r)
    unsigned long long* fine = (p.prof && c == 0) ? p.prof + (size_t)G * (SK_PHASES * L + 1) : nullptr;
    int fine_on = 0;
    auto stamp = [&](int i) { if (fine && fine_on && tid == 0) fine[i] = global_timer_ns(); };
    auto resid_norm = [&](const float* part, int S, const float* w, bf16* xn_out, float* y_out, bf16* y_out_t) {
      if (c < R) {
        stamp(1);
        float* xr = p.x + (size_t)c * D;
        float v[SK_RNK], a[SK_RNK];
        // all loads of two elements first (the row and every split slab), so they are in flight together
#pragma unroll
        for (int k0 = 0; k0 < SK_RNK; k0 += 2) {
          float b[2][SK_MAXS];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int d = tid + (k0 + j) * SK_WTHREADS;
            v[k0 + j] = 0.f;
            if (d < D) { v[k0 + j] = __ldcg(xr + d); load_splits(part + (size_t)c * D + d, S, (size_t)R * D, b[j]); }
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int d = tid + (k0 + j) * SK_WTHREADS;
            a[k0 + j] = (d < D) ? sum_loaded(b[j], S) : 0.f;
          }
        }
        float ss = 0.f;
        if (fine && fine_on) { float keep = 0.f;
#pragma unroll
          for (int k = 0; k < SK_RNK; ++k) keep += a[k] + v[k];
          if (keep == 123.456f) red_s[31] = keep;               // force the loads to complete before the stamp
          stamp(2); }
#pragma unroll
        for (int k = 0; k < SK_RNK; ++k) {
          const int d = tid + k * SK_WTHREADS;
          if (d < D) {
            const float t = bf16_round(v[k] + bf16_round(a[k]));   // bf16 residual stream of the decode steps
            xr[d] = t;
            v[k] = t;
            ss += t * t;
          }
        }
        ss = warp_sum(ss);
        if (lane == 0) red_s[warp] = ss;
        named_bar_sync(1, SK_WTHREADS);
        float tot = (lane < SK_WORKERS) ? red_s[lane] : 0.f;
        tot = warp_sum(tot);
        const float r = rsqrtf(tot / (float)D + p.eps);
        stamp(3);
#pragma unroll
        for (int k = 0; k < SK_RNK; ++k) {
          const int d = tid + k * SK_WTHREADS;
          if (d < D) {
            const float hn = bf16_round(v[k] * r);
            const float y = w[d] * hn;
            if (xn_out) xn_out[(size_t)c * D + d] = __float2bfloat16_rn(y);
            if (y_out) y_out[(size_t)c * D + d] = y;
            if (y_out_t) y_out_t[(size_t)c * D + d] = __float2bfloat16_rn(y);
          }
        }
      }
    };

    const int grp = warp / AT_GW, tg = tid - grp * AT_GT;
    int gb, ge;
    cut.group_range(c, grp, gb, ge);
    for (int l = 0; l < L; ++l) {
      const int pb = l * SK_PHASES;
      bf16* kc = p.kv + (size_t)l * kv_layer;
      bf16* vc = kc + kv_half;
      // The single monotonic arrival counter is only a barrier if nobody arrives for phase k+1 before everyone
      // arrived for phase k: EVERY phase therefore starts by waiting for the previous one, also the
      // contraction phases whose data dependencies are already enforced through the operand rings.
      // ---------------- phase 0: QKV contraction
      if (l > 0) phase_wait(pb + 0);
      epilogue(p.g_qkv, p.part_qkv, 3 * HD);
      phase_done();
      // ---------------- phase 1: attention, 4 independent group streams
      phase_wait(pb + 1);
      attn_group_stream<SK_SPG>(tg, gb, ge, cut, row_units, R, H, p.Tmax, pos, p.part_qkv, p.g_qkv.splits,
                                (size_t)R * 3 * HD, p.cosT, p.sinT, kc, vc, p.kv_start, p.attn_out, p.attn_ws, p.attn_cnt,
                                p.scale, true, ringKV, SK_A_BYTES, stage_tab + grp * SK_SPG, kv_full, kv_empty, kc_att,
                                gsm[grp], c * AT_NG + grp, 2 + grp);
      phase_done();
      // ---------------- phase 2: O contraction
      phase_wait(pb + 2);
      epilogue(p.g_o, p.part_o, D);
      phase_done();
      // ---------------- phase 3: residual + post-attention RMSNorm
      fine_on = (l == 12);
      stamp(0);
      phase_wait(pb + 3);
      resid_norm(p.part_o, p.g_o.splits, p.ln + ((size_t)l * 2 + 1) * D, p.xn, nullptr, nullptr);
      stamp(4);
      phase_done();
      stamp(5);
      fine_on = 0;
      // ---------------- phase 4: gate|up contraction
      phase_wait(pb + 4);
      epilogue(p.g_gu, p.part_gu, 2 * F);
      phase_done();
      // ---------------- phase 5: SwiGLU
      phase_wait(pb + 5);
      {
        const int S = p.g_gu.splits;
        const size_t sstride = (size_t)R * 2 * F, total = (size_t)R * F;
        constexpr int UN = 2;
        for (size_t i0 = (size_t)c * SK_WTHREADS + tid; i0 < total; i0 += (size_t)G * SK_WTHREADS * UN) {
          float gsum[UN], usum[UN];
          float bg[UN][SK_MAXS], bu[UN][SK_MAXS];
#pragma unroll
          for (int q = 0; q < UN; ++q) {                        // every load of the UN elements before any add / store
            const size_t i = i0 + (size_t)q * G * SK_WTHREADS;
            if (i < total) {
              const size_t tok = i / F, f = i % F;
              const size_t gi = (f / 64) * 128 + f % 64;       // gate|up rows interleaved in blocks of 64 (weights.py)
              load_splits(p.part_gu + tok * 2 * F + gi, S, sstride, bg[q]);
              load_splits(p.part_gu + tok * 2 * F + gi + 64, S, sstride, bu[q]);
            }
          }
#pragma unroll
          for (int q = 0; q < UN; ++q) {
            const size_t i = i0 + (size_t)q * G * SK_WTHREADS;
            gsum[q] = (i < total) ? sum_loaded(bg[q], S) : 0.f;
            usum[q] = (i < total) ? sum_loaded(bu[q], S) : 0.f;
          }
#pragma unroll
          for (int q = 0; q < UN; ++q) {
            const size_t i = i0 + (size_t)q * G * SK_WTHREADS;
            if (i < total) {
              const float g = bf16_round(gsum[q]), uu = bf16_round(usum[q]);
              const float sg = bf16_round(g / (1.0f + expf(-g)));
              p.h[i] = __float2bfloat16_rn(sg * uu);
            }
          }
        }
      }
      phase_done();
      // ---------------- phase 6: down contraction
      phase_wait(pb + 6);
      epilogue(p.g_d, p.part_d, D);
      phase_done();
      // ---------------- phase 7: residual + next layer's input RMSNorm (or the final norm)
      phase_wait(pb + 7);
      if (l + 1 < L) resid_norm(p.part_d, p.g_d.splits, p.ln + ((size_t)(l + 1) * 2) * D, p.xn, nullptr, nullptr);
      else resid_norm(p.part_d, p.g_d.splits, p.norm_w, nullptr, p.hidden_f, p.hidden_t);
      phase_done();
    }
    // every CTA has passed the first barrier long ago: safe to advance the launch epoch / step counter
    if (c == 0 && tid == 0) {
      grid_wait(p.grid_bar, bar_target(SK_PHASES * L - 1), 33);
      *p.epoch = epoch + 1;
      if (p.inc_step && p.step_ptr) *p.step_ptr += 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == SK_WORKERS + 2) tmem_dealloc(tmem_base, 2 * SK_NT);
}

}  // namespace pg
