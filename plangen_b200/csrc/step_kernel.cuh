// One decode step of the LM (all layers) as ONE persistent kernel: 1 CTA per SM, no launches between ops.
//
// Why (B200): a 1.3B decode step moves ~2.5 GB of weights + the KV cache and does almost no math, so
// the only resource that matters is the HBM stream.  With one kernel per op the stream drains at every
// kernel boundary (~8 per layer x 24 layers) and each op pays its own launch / prologue / epilogue
// latency.  Here every SM runs a warp-specialised CTA for the whole step:
//
//   warp 16  producer A : streams THIS CTA's share of every op's weight tiles (TMA tensor loads,
//                         SWIZZLE_128B) and KV-cache tiles (TMA bulk copies) through one ring of 16 KB
//                         stages, in program order, for ALL layers.  Weights and cached K/V do not
//                         depend on the step's activations, so this warp never waits for a grid barrier:
//                         while the other warps wait for a dependency, it keeps the ring (and HBM) busy.
//   warp 17  producer B : streams the small activation operand tiles (L2-resident) of the contractions
//                         through a second ring, after the grid barrier of the producing phase.
//   warp 18  MMA issuer : tcgen05.mma (swap-AB: weight tile = M 128, token tile = N 32), fp32 accumulators
//                         double-buffered in TMEM.
//   warps 0-15 workers  : TMEM epilogues (warps 0-3), the KV-cache attention (4 groups x 4 warps, fp32
//                         online softmax on CUDA cores straight from the ring), split-K reductions fused
//                         with residual + RMSNorm and SwiGLU.
//
// Phases per layer (grid barrier after each; arithmetic identical to the per-op kernels):
//   0 QKV contraction -> 1 attention (+RoPE, KV append) -> 2 O contraction -> 3 residual + RMSNorm
//   -> 4 gate|up contraction -> 5 SwiGLU -> 6 down contraction -> 7 residual + RMSNorm (next layer / final)
#pragma once
#include "common.cuh"
#include "lm_kernels.cuh"
#include "attn_tma.cuh"

namespace pg {

constexpr int SK_NSA = 8;                       // A ring: weight / KV stages of 16 KB
constexpr int SK_NSB = 8;                       // B ring: activation tiles
constexpr int SK_NT = 32;                       // token tile (rows R <= 32)
constexpr int SK_WORKERS = 16;
constexpr int SK_WTHREADS = SK_WORKERS * 32;
constexpr int SK_THREADS = 32 * (SK_WORKERS + 3);
constexpr int SK_A_BYTES = 16384;
constexpr int SK_B_BYTES = SK_NT * TC_BK * 2;   // 4 KB
constexpr int SK_SMEM = SK_NSA * SK_A_BYTES + SK_NSB * SK_B_BYTES + 1024;
constexpr int SK_PHASES = 8;
constexpr int SK_RNK = 8;                       // RMSNorm elements per worker thread: D <= 8 * 512
static_assert(SK_NSA % AT_NG == 0, "attention stage ownership");
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

PG_DEVINL void umma_commit4(uint64_t* bar) {   // the A ring's empty barriers expect 4 arrivals (see attention groups)
  umma_commit(bar); umma_commit(bar); umma_commit(bar); umma_commit(bar);
}

__global__ void __launch_bounds__(SK_THREADS, 1)
decode_step_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ringA = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ringB = ringA + SK_NSA * SK_A_BYTES;
  __shared__ uint64_t fullA[SK_NSA], emptyA[SK_NSA], fullB[SK_NSB], emptyB[SK_NSB], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ int row_units[AT_MAX_ROWS + 1];
  __shared__ float q_s[HEAD_DIM], k_s[HEAD_DIM], v_s[HEAD_DIM];
  __shared__ float m_s[SK_WORKERS], l_s[SK_WORKERS], o_s[SK_WORKERS][HEAD_DIM];
  __shared__ float red_s[32];
  __shared__ int is_last_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, c = blockIdx.x;
  const int R = p.R, H = p.H, D = p.D, HD = p.HD, F = p.F, L = p.L;
  const int pos = p.pos_base + (p.step_ptr ? *p.step_ptr : 0);
  const unsigned long long epoch = *p.epoch;
  const unsigned long long bar_base = epoch * (unsigned long long)(SK_PHASES * L) * (unsigned long long)G;
  auto bar_target = [&](int phase_index) { return bar_base + (unsigned long long)(phase_index + 1) * (unsigned long long)G; };

  if (tid == 0) {
    for (int i = 0; i < SK_NSA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], AT_GW); }
    for (int i = 0; i < SK_NSB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    mbar_fence_init();
  }
  if (warp == SK_WORKERS + 2) { tmem_alloc(&tmem_slot, 2 * SK_NT); tmem_relinquish(); }
  // attention schedule of this step (same for every layer): units per row -> exclusive prefix
  if (warp == 0) {
    int carry = 0;
    for (int r0 = 0; r0 < R; r0 += 32) {
      const int r = r0 + lane;
      int u = (r < R) ? row_tiles(p.kv_start[r], pos) + 1 : 0;
      int incl = u;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (r < R) row_units[r] = carry + incl - u;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) row_units[R] = carry;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int U = row_units[R] * H;
  const int per = (U + G - 1) / G;
  const int u_begin = min(c * per, U), u_end = min(u_begin + per, U);
  const size_t kv_layer = (size_t)2 * R * H * p.Tmax * HEAD_DIM;     // elements per layer (k then v)
  const size_t kv_half = (size_t)R * H * p.Tmax * HEAD_DIM;

  // ------------------------------------------------------------------ helpers shared by the roles
  // contraction items of this CTA: it = c, c + G, ...; tile = it / splits, split = it % splits
  struct Item { int tile, split, kb0, kb1; };
  auto item_of = [&](const GemmSched& g, int it) {
    Item r;
    r.tile = it / g.splits; r.split = it % g.splits;
    r.kb0 = r.split * g.kb_per_split; r.kb1 = min(r.kb0 + g.kb_per_split, g.num_kb);
    return r;
  };

  if (warp == SK_WORKERS) {
    // =========================================================== producer A: weights + cached K/V
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      int ja = 0;                                              // A-ring tile counter of this CTA
      auto stage_acquire = [&]() {
        const int s = ja % SK_NSA;
        const uint32_t round = (uint32_t)(ja / SK_NSA);
        mbar_wait(&emptyA[s], (round & 1u) ^ 1u, 21, ja);
        mbar_expect_tx(&fullA[s], SK_A_BYTES);
        ++ja;
        return s;
      };
      auto stream_weights = [&](const GemmSched& g, const CUtensorMap* map) {
        for (int it = c; it < g.n_tiles * g.splits; it += G) {
          const Item im = item_of(g, it);
          for (int kb = im.kb0; kb < im.kb1; ++kb) {
            const int s = stage_acquire();
            tma_load_2d(ringA + (size_t)s * SK_A_BYTES, map, &fullA[s], kb * TC_BK, im.tile * TC_BM, pol);
          }
        }
      };
      for (int l = 0; l < L; ++l) {
        const CUtensorMap* wm = p.wmaps + (size_t)l * 4;
        stream_weights(p.g_qkv, wm + 0);
        // cached K/V tiles of this CTA's attention units
        {
          const bf16* kc = p.kv + (size_t)l * kv_layer;
          const bf16* vc = kc + kv_half;
          int r = 0;
          while (r + 1 < R && row_units[r + 1] * H <= u_begin) ++r;
          for (int u = u_begin; u < u_end; ++u) {
            while (row_units[r + 1] * H <= u) ++r;
            const int ur = row_units[r + 1] - row_units[r];
            const int local = u - row_units[r] * H;
            const int h = local / ur, k = local % ur;
            if (k == ur - 1) continue;
            const int t0 = (p.kv_start[r] / AT_TILE + k) * AT_TILE;
            const int s = stage_acquire();
            const size_t off = (((size_t)r * H + h) * p.Tmax + t0) * HEAD_DIM;
            bulk_load(ringA + (size_t)s * SK_A_BYTES, kc + off, AT_TILE_BYTES, &fullA[s], pol);
            bulk_load(ringA + (size_t)s * SK_A_BYTES + AT_TILE_BYTES, vc + off, AT_TILE_BYTES, &fullA[s], pol);
          }
        }
        stream_weights(p.g_o, wm + 1);
        stream_weights(p.g_gu, wm + 2);
        stream_weights(p.g_d, wm + 3);
      }
    }
  } else if (warp == SK_WORKERS + 1) {
    // =========================================================== producer B: activation operand tiles
    if (lane == 0) {
      const uint64_t pol = policy_evict_last();
      int jb = 0;
      auto stream_act = [&](const GemmSched& g, const CUtensorMap* map, int phase_index) {
        if (phase_index > 0) grid_wait(p.grid_bar, bar_target(phase_index - 1), 31);
        fence_proxy_async_global();                            // generic-proxy stores of other CTAs -> TMA reads
        for (int it = c; it < g.n_tiles * g.splits; it += G) {
          const Item im = item_of(g, it);
          for (int kb = im.kb0; kb < im.kb1; ++kb) {
            const int s = jb % SK_NSB;
            const uint32_t round = (uint32_t)(jb / SK_NSB);
            mbar_wait(&emptyB[s], (round & 1u) ^ 1u, 22, jb);
            mbar_expect_tx(&fullB[s], SK_B_BYTES);
            tma_load_2d(ringB + (size_t)s * SK_B_BYTES, map, &fullB[s], kb * TC_BK, 0, pol);
            ++jb;
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
          for (int kb = im.kb0; kb < im.kb1; ++kb) {
            const int sb = jb % SK_NSB, sa = ja % SK_NSA;
            // B first: its arrival implies this CTA finished the previous phase, hence every earlier use of
            // the A stage (possibly by an attention group) has completed -> no parity aliasing on fullA
            mbar_wait(&fullB[sb], (uint32_t)(jb / SK_NSB) & 1u, 24, jb);
            mbar_wait(&fullA[sa], (uint32_t)(ja / SK_NSA) & 1u, 25, ja);
            tc_fence_after();
            const uint64_t da = umma_desc_k_sw128(smem_u32(ringA + (size_t)sa * SK_A_BYTES));
            const uint64_t db = umma_desc_k_sw128(smem_u32(ringB + (size_t)sb * SK_B_BYTES));
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)
              umma_bf16(tacc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((kb > im.kb0) | (k > 0)));
            umma_commit4(&emptyA[sa]);
            umma_commit(&emptyB[sb]);
            ++ja; ++jb;
          }
          umma_commit(&tmem_full[buf]);
          ++ji;
        }
      };
      for (int l = 0; l < L; ++l) {
        run_gemm(p.g_qkv);
        // the attention phase consumes this CTA's K/V tiles from the A ring: skip over them
        {
          int r = 0;
          while (r + 1 < R && row_units[r + 1] * H <= u_begin) ++r;
          for (int u = u_begin; u < u_end; ++u) {
            while (row_units[r + 1] * H <= u) ++r;
            const int ur = row_units[r + 1] - row_units[r];
            if ((u - row_units[r] * H) % ur != ur - 1) ++ja;
          }
        }
        run_gemm(p.g_o);
        run_gemm(p.g_gu);
        run_gemm(p.g_d);
      }
    }
  } else {
    // =========================================================== workers (16 warps)
    int ja = 0, ji = 0;                                        // mirrors of the A-ring / accumulator counters
    const float LOG2E = 1.4426950408889634f;
    int prof_i = 0;
    unsigned long long* prof = p.prof ? p.prof + (size_t)c * (SK_PHASES * L + 1) : nullptr;
    if (prof && tid == 0) prof[prof_i++] = global_timer_ns();
    auto phase_done = [&]() {                                  // all of this CTA's global writes of the phase are done
      fence_proxy_async_global();                              // generic-proxy stores -> later TMA (async-proxy) reads
      __threadfence();
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
        ja += im.kb1 - im.kb0;
        ++ji;
      }
    };
    // residual + RMSNorm of row `c` (CTAs >= R idle): x += rnd(sum part); xn = w * (x * r)
    auto resid_norm = [&](const float* part, int S, const float* w, bf16* xn_out, float* y_out, bf16* y_out_t) {
      if (c < R) {
        float* xr = p.x + (size_t)c * D;
        float v[SK_RNK];
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < SK_RNK; ++k) {
          const int d = tid + k * SK_WTHREADS;
          v[k] = 0.f;
          if (d < D) {
            float t = xr[d];
            const float a = bf16_round(reduce_splits(part, S, (size_t)R * D, (size_t)c * D + d));
            t = bf16_round(t + a);                              // bf16 residual stream of the decode steps
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
      // ---------------- phase 1: attention over this CTA's units
      phase_wait(pb + 1);
      {
        const int S = p.g_qkv.splits;
        const size_t sstride = (size_t)R * 3 * HD;
        int u = u_begin;
        int r = 0;
        while (r + 1 < R && row_units[r + 1] * H <= u_begin) ++r;
        while (u < u_end) {
          while (row_units[r + 1] * H <= u) ++r;
          const int ur = row_units[r + 1] - row_units[r];
          const int item_base = row_units[r] * H;
          const int h = (u - item_base) / ur;
          const int item_lo = item_base + h * ur, item_hi = item_lo + ur;
          const int seg_lo = u, seg_hi = min(item_hi, u_end);
          const int start = p.kv_start[r];
          const bool owns_new = (seg_hi == item_hi);
          const int n_tiles_seg = (seg_hi - seg_lo) - (owns_new ? 1 : 0);
          {
            const float* row = p.part_qkv + (size_t)r * 3 * HD;
            const int jj = tid & 63;
            const float cs = p.cosT[pos * 64 + jj], sn = p.sinT[pos * 64 + jj];
            if (tid < 64) {
              const float x1 = bf16_round(reduce_splits(row, S, sstride, (size_t)h * HEAD_DIM + jj));
              const float x2 = bf16_round(reduce_splits(row, S, sstride, (size_t)h * HEAD_DIM + jj + 64));
              float a, b;
              rope_pair<bf16>(x1, x2, cs, sn, true, a, b);
              q_s[jj] = a * (p.scale * LOG2E); q_s[jj + 64] = b * (p.scale * LOG2E);
            } else if (owns_new && tid < 128) {
              const float x1 = bf16_round(reduce_splits(row, S, sstride, (size_t)HD + h * HEAD_DIM + jj));
              const float x2 = bf16_round(reduce_splits(row, S, sstride, (size_t)HD + h * HEAD_DIM + jj + 64));
              float a, b;
              rope_pair<bf16>(x1, x2, cs, sn, true, a, b);
              const float v1 = bf16_round(reduce_splits(row, S, sstride, (size_t)2 * HD + h * HEAD_DIM + jj));
              const float v2 = bf16_round(reduce_splits(row, S, sstride, (size_t)2 * HD + h * HEAD_DIM + jj + 64));
              k_s[jj] = a; k_s[jj + 64] = b; v_s[jj] = v1; v_s[jj + 64] = v2;
              const size_t cidx = (((size_t)r * H + h) * p.Tmax + pos) * HEAD_DIM + jj;
              kc[cidx] = __float2bfloat16_rn(a); kc[cidx + 64] = __float2bfloat16_rn(b);
              vc[cidx] = __float2bfloat16_rn(v1); vc[cidx + 64] = __float2bfloat16_rn(v2);
            }
          }
          named_bar_sync(1, SK_WTHREADS);
          float qv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) qv[i] = q_s[lane * 4 + i];
          float m = -INFINITY, lsum = 0.f, o[4] = {0.f, 0.f, 0.f, 0.f};
          const int first_tile_k = seg_lo - item_lo;
          const int grp = warp / AT_GW, wig = warp % AT_GW;
          for (int jt = ja + ((grp - ja) % AT_NG + AT_NG) % AT_NG; jt < ja + n_tiles_seg; jt += AT_NG) {
            const int t = jt - ja;
            const int s = jt % SK_NSA;
            mbar_wait(&fullA[s], (uint32_t)(jt / SK_NSA) & 1u, 27, jt);
            const uint8_t* kt = ringA + (size_t)s * SK_A_BYTES + wig * AT_TW * (HEAD_DIM * 2);
            const uint8_t* vt = kt + AT_TILE_BYTES;
            const int t0 = (start / AT_TILE + first_tile_k + t) * AT_TILE + wig * AT_TW;
            float sc[AT_TW];
#pragma unroll
            for (int i = 0; i < AT_TW; ++i) {
              const uint2 kk = *reinterpret_cast<const uint2*>(kt + i * (HEAD_DIM * 2) + lane * 8);
              float d = bf16lo(kk.x) * qv[0];
              d = fmaf(bf16hi(kk.x), qv[1], d); d = fmaf(bf16lo(kk.y), qv[2], d); d = fmaf(bf16hi(kk.y), qv[3], d);
              sc[i] = d;
            }
#pragma unroll
            for (int off = 16, n = AT_TW; off >= 4; off >>= 1, n >>= 1) {
              const bool upper = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < n / 2; ++i) {
                const float send = upper ? sc[i] : sc[i + n / 2];
                const float recv = __shfl_xor_sync(0xffffffffu, send, off);
                sc[i] = (upper ? sc[i + n / 2] : sc[i]) + recv;
              }
            }
            sc[0] += __shfl_xor_sync(0xffffffffu, sc[0], 2);
            sc[0] += __shfl_xor_sync(0xffffffffu, sc[0], 1);
            const int tok = t0 + (lane >> 2);
            const bool valid = (tok >= start) && (tok < pos);
            const float sv = valid ? sc[0] : -INFINITY;
            const float mx = fmaxf(m, warp_max(sv));
            const float pr = valid ? exp2f(sv - mx) : 0.f;
            const float corr = (mx == -INFINITY) ? 1.f : exp2f(m - mx);
            lsum = lsum * corr + 0.25f * warp_sum(pr);
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] *= corr;
#pragma unroll
            for (int i = 0; i < AT_TW; ++i) {
              const float pi = __shfl_sync(0xffffffffu, pr, i * 4);
              const uint2 vv = *reinterpret_cast<const uint2*>(vt + i * (HEAD_DIM * 2) + lane * 8);
              o[0] = fmaf(pi, bf16lo(vv.x), o[0]); o[1] = fmaf(pi, bf16hi(vv.x), o[1]);
              o[2] = fmaf(pi, bf16lo(vv.y), o[2]); o[3] = fmaf(pi, bf16hi(vv.y), o[3]);
            }
            m = mx;
            __syncwarp();
            if (lane == 0) mbar_arrive(&emptyA[s]);
          }
          if (owns_new && warp == 0) {
            float d = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) d = fmaf(k_s[lane * 4 + i], qv[i], d);
            d = warp_sum(d);
            const float mx = fmaxf(m, d);
            const float corr = (m == -INFINITY) ? 0.f : exp2f(m - mx);
            const float pr = exp2f(d - mx);
            lsum = lsum * corr + pr;
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = fmaf(pr, v_s[lane * 4 + i], o[i] * corr);
            m = mx;
          }
          if (lane == 0) { m_s[warp] = m; l_s[warp] = lsum; }
#pragma unroll
          for (int i = 0; i < 4; ++i) o_s[warp][lane * 4 + i] = o[i];
          named_bar_sync(1, SK_WTHREADS);
          const int c_first = item_lo / per, c_last = (item_hi - 1) / per;
          const int n_contrib = c_last - c_first + 1;
          const int it = r * H + h;
          const bool out_thread = tid < HEAD_DIM;
          float M = -INFINITY, Ltot = 0.f, acc = 0.f;
          if (out_thread) {
#pragma unroll
            for (int w = 0; w < SK_WORKERS; ++w) M = fmaxf(M, m_s[w]);
#pragma unroll
            for (int w = 0; w < SK_WORKERS; ++w) {
              const float f = (m_s[w] == -INFINITY) ? 0.f : exp2f(m_s[w] - M);
              Ltot += l_s[w] * f;
              acc += o_s[w][tid] * f;
            }
          }
          const size_t oidx = (size_t)r * HD + h * HEAD_DIM + tid;
          if (n_contrib == 1) {
            if (out_thread) p.attn_out[oidx] = __float2bfloat16_rn(acc / Ltot);
            named_bar_sync(1, SK_WTHREADS);
          } else {
            if (out_thread) {
              float* wp = p.attn_ws + ((size_t)it * AT_MAX_SLOTS + (c - c_first)) * (HEAD_DIM + 2);
              wp[tid] = acc;
              if (tid == 0) { wp[HEAD_DIM] = M; wp[HEAD_DIM + 1] = Ltot; }
              __threadfence();
            }
            named_bar_sync(1, SK_WTHREADS);
            if (tid == 0) {
              const int prev = atomicAdd(p.attn_cnt + it, 1);
              is_last_s = (prev == n_contrib - 1);
            }
            named_bar_sync(1, SK_WTHREADS);
            if (is_last_s && out_thread) {
              __threadfence();
              const float* wb = p.attn_ws + (size_t)it * AT_MAX_SLOTS * (HEAD_DIM + 2);
              float Mg = -INFINITY;
              for (int s2 = 0; s2 < n_contrib; ++s2) Mg = fmaxf(Mg, __ldcg(wb + (size_t)s2 * (HEAD_DIM + 2) + HEAD_DIM));
              float Lg = 0.f, og = 0.f;
              for (int s2 = 0; s2 < n_contrib; ++s2) {
                const float ms = __ldcg(wb + (size_t)s2 * (HEAD_DIM + 2) + HEAD_DIM);
                const float f = (ms == -INFINITY) ? 0.f : exp2f(ms - Mg);
                Lg += __ldcg(wb + (size_t)s2 * (HEAD_DIM + 2) + HEAD_DIM + 1) * f;
                og += __ldcg(wb + (size_t)s2 * (HEAD_DIM + 2) + tid) * f;
              }
              p.attn_out[oidx] = __float2bfloat16_rn(og / Lg);
              if (tid == 0) p.attn_cnt[it] = 0;
            }
            named_bar_sync(1, SK_WTHREADS);
          }
          ja += n_tiles_seg;
          u = seg_hi;
        }
      }
      phase_done();
      // ---------------- phase 2: O contraction
      phase_wait(pb + 2);
      epilogue(p.g_o, p.part_o, D);
      phase_done();
      // ---------------- phase 3: residual + post-attention RMSNorm
      phase_wait(pb + 3);
      resid_norm(p.part_o, p.g_o.splits, p.ln + ((size_t)l * 2 + 1) * D, p.xn, nullptr, nullptr);
      phase_done();
      // ---------------- phase 4: gate|up contraction
      phase_wait(pb + 4);
      epilogue(p.g_gu, p.part_gu, 2 * F);
      phase_done();
      // ---------------- phase 5: SwiGLU
      phase_wait(pb + 5);
      {
        const int S = p.g_gu.splits;
        const size_t sstride = (size_t)R * 2 * F, total = (size_t)R * F;
        for (size_t i = (size_t)c * SK_WTHREADS + tid; i < total; i += (size_t)G * SK_WTHREADS) {
          const size_t tok = i / F, f = i % F;
          const float g = bf16_round(reduce_splits(p.part_gu, S, sstride, tok * 2 * F + f));
          const float uu = bf16_round(reduce_splits(p.part_gu, S, sstride, tok * 2 * F + F + f));
          const float sg = bf16_round(g / (1.0f + expf(-g)));
          p.h[i] = __float2bfloat16_rn(sg * uu);
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
