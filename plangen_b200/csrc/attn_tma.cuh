// KV-cache decode attention for the paired cond/uncond CFG batch, bf16 cache, one launch per layer.
//
// Design (B200): the kernel is a persistent HBM streamer.  Work = "units" of 32 cached tokens of one
// (row, head) item plus one extra unit per item for the token being decoded.  The flat unit space
// [0, U) is cut into equal contiguous ranges, one per CTA (one CTA per SM, 192 KB of K/V in flight each), so rows with long prompts
// (cond) and short prompts (uncond) are balanced exactly.  In every CTA one producer thread streams
// the K and V tiles of its range with TMA bulk copies (cp.async.bulk, 8 KB + 8 KB per stage) through a
// ring of mbarrier-guarded shared-memory stages; 4 consumer warps take the tiles round-robin and run
// an fp32 online softmax: scores by a transposed warp reduction, P.V with one row broadcast per token.  Cached tokens do not depend on the current step, so with PDL the
// ring is filled BEFORE griddepcontrol.wait (while the QKV contraction is still running); only q and
// the new k/v wait.  A CTA that finishes an item alone writes the output; otherwise partial
// (m, l, o) records are merged by the last CTA to arrive (threadfence + counter, deterministic order).
#pragma once
#include "common.cuh"
#include "lm_kernels.cuh"

namespace pg {

constexpr int AT_TILE = 32;                         // tokens per unit
constexpr int AT_TILE_BYTES = AT_TILE * HEAD_DIM * 2;   // 8 KB (K) and 8 KB (V)
constexpr int AT_NG = 4;                            // consumer warp groups (one tile at a time each)
constexpr int AT_GW = 4;                            // warps per group: each takes AT_TILE / AT_GW tokens of the tile
constexpr int AT_NW = AT_NG * AT_GW;                // consumer warps
constexpr int AT_TW = AT_TILE / AT_GW;              // tokens per warp per tile (8)
constexpr int AT_STAGES = 12;                      // MUST be a multiple of AT_NG (see the consumer loop)
constexpr int AT_THREADS = 32 * (AT_NW + 1);
constexpr int AT_MAX_ROWS = 256;
constexpr int AT_MAX_SLOTS = 64;                    // partial records per item
constexpr int AT_SMEM = AT_STAGES * 2 * AT_TILE_BYTES + 1024;

PG_DEVINL void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
PG_DEVINL void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// units of one row: tiles covering cached tokens [start, pos) on a 32-token grid, plus the new token
PG_DEVINL int row_tiles(int start, int pos) { return pos > start ? ((pos - 1) / AT_TILE - start / AT_TILE + 1) : 0; }

static_assert(AT_STAGES % AT_NG == 0, "each ring stage must belong to exactly one consumer warp group");
static_assert(AT_TW == 8, "the transposed reduction below is written for 8 tokens per warp");

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_decode_tma_kernel(const float* __restrict__ part, int S, size_t split_stride, const float* __restrict__ cosT,
                       const float* __restrict__ sinT, bf16* __restrict__ kcache, bf16* __restrict__ vcache,
                       const int32_t* __restrict__ kv_start, bf16* __restrict__ out, float* __restrict__ ws_part,
                       int* __restrict__ ws_count, int R, int H, int Tmax, int pos_base,
                       const int* __restrict__ step_ptr, float scale, int bf16_trig, int early_trigger) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[AT_STAGES], empty_bar[AT_STAGES];
  __shared__ int row_units[AT_MAX_ROWS + 1];        // exclusive prefix of units per row (per head)
  __shared__ float q_s[HEAD_DIM], k_s[HEAD_DIM], v_s[HEAD_DIM];
  __shared__ float m_s[AT_NW], l_s[AT_NW], o_s[AT_NW][HEAD_DIM];
  __shared__ int is_last_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int HD = H * HEAD_DIM;
  if (early_trigger) pdl_launch_dependents();
  // `pos` comes from the device-side step counter, which the PREVIOUS decode step's last kernel
  // incremented - long complete by now (it precedes this step's sampler and QKV kernels), but to keep the
  // PDL chain transitive we only read it after the dependency wait in the consumer path; the producer
  // needs it early, so it reads it here: the counter is only written by the final kernel of a step, and
  // every kernel of this step transitively waited for that kernel.
  const int pos = pos_base + (step_ptr ? *step_ptr : 0);

  if (tid == 0) {
    for (int i = 0; i < AT_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], AT_GW); }
    mbar_fence_init();
  }
  // per-row unit counts -> exclusive prefix (R <= 256: one pass by warp 0)
  if (warp == 0) {
    int carry = 0;
    for (int r0 = 0; r0 < R; r0 += 32) {
      const int r = r0 + lane;
      int u = (r < R) ? row_tiles(kv_start[r], pos) + 1 : 0;
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
  __syncthreads();
  const int U = row_units[R] * H;                         // all units, item-major: item = r * H + h
  const int per = (U + gridDim.x - 1) / gridDim.x;
  const int u_begin = min(blockIdx.x * per, U), u_end = min(u_begin + per, U);
  if (u_begin >= u_end) { pdl_wait(); return; }

  // locate the row containing a flat unit index (rows are few: linear scan)
  auto find_row = [&](int u) {
    int r = 0;
    while (r + 1 < R && row_units[r + 1] * H <= u) ++r;
    return r;
  };

  if (warp == AT_NW) {
    // ============================== producer: stream K/V tiles of [u_begin, u_end) ==============================
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      int r = find_row(u_begin);
      int j = 0;                                          // tile counter (skips the new-token units)
      for (int u = u_begin; u < u_end; ++u) {
        while (row_units[r + 1] * H <= u) ++r;
        const int ur = row_units[r + 1] - row_units[r];   // units of this row (tiles + 1)
        const int local = u - row_units[r] * H;
        const int h = local / ur, k = local % ur;
        if (k == ur - 1) continue;                        // the new-token unit has no cached tile
        const int t0 = (kv_start[r] / AT_TILE + k) * AT_TILE;
        const int s = j % AT_STAGES;
        const uint32_t round = (uint32_t)(j / AT_STAGES);
        mbar_wait(&empty_bar[s], (round & 1u) ^ 1u, 11, j * 100000 + (u_end - u_begin) * 100 + (pos % 100));
        mbar_expect_tx(&full_bar[s], 2 * AT_TILE_BYTES);
        const size_t off = (((size_t)r * H + h) * Tmax + t0) * HEAD_DIM;
        bulk_load(ring + (size_t)s * 2 * AT_TILE_BYTES, kcache + off, AT_TILE_BYTES, &full_bar[s], pol);
        bulk_load(ring + (size_t)s * 2 * AT_TILE_BYTES + AT_TILE_BYTES, vcache + off, AT_TILE_BYTES, &full_bar[s], pol);
        ++j;
      }
    }
    pdl_wait();
    return;
  }

  // ============================== consumers (AT_NW warps) ==============================
  pdl_wait();                                             // QKV partials of this step are now visible
  const float LOG2E = 1.4426950408889634f;
  int j = 0;                                              // CTA-wide tile counter, identical in all warps
  int u = u_begin;
  int r = find_row(u_begin);
  while (u < u_end) {
    while (row_units[r + 1] * H <= u) ++r;
    const int ur = row_units[r + 1] - row_units[r];
    const int item_base = row_units[r] * H;
    const int h = (u - item_base) / ur;
    const int item_lo = item_base + h * ur, item_hi = item_lo + ur;       // flat units of this item
    const int seg_lo = u, seg_hi = min(item_hi, u_end);
    const int start = kv_start[r];
    const bool owns_new = (seg_hi == item_hi);                            // this CTA holds the new-token unit
    const int n_tiles_seg = (seg_hi - seg_lo) - (owns_new ? 1 : 0);
    // ---- q (all), k/v of the new token (owner): reduce split-K partials, RoPE
    {
      const float* row = part + (size_t)r * 3 * HD;
      const int jj = tid & 63;
      const float c = cosT[pos * 64 + jj], sn = sinT[pos * 64 + jj];
      if (tid < 64) {
        const float x1 = bf16_round(reduce_splits(row, S, split_stride, (size_t)h * HEAD_DIM + jj));
        const float x2 = bf16_round(reduce_splits(row, S, split_stride, (size_t)h * HEAD_DIM + jj + 64));
        float a, b;
        rope_pair<bf16>(x1, x2, c, sn, bf16_trig != 0, a, b);
        q_s[jj] = a * (scale * LOG2E); q_s[jj + 64] = b * (scale * LOG2E);
      } else if (owns_new && tid < 128) {
        const float x1 = bf16_round(reduce_splits(row, S, split_stride, (size_t)HD + h * HEAD_DIM + jj));
        const float x2 = bf16_round(reduce_splits(row, S, split_stride, (size_t)HD + h * HEAD_DIM + jj + 64));
        float a, b;
        rope_pair<bf16>(x1, x2, c, sn, bf16_trig != 0, a, b);
        const float v1 = bf16_round(reduce_splits(row, S, split_stride, (size_t)2 * HD + h * HEAD_DIM + jj));
        const float v2 = bf16_round(reduce_splits(row, S, split_stride, (size_t)2 * HD + h * HEAD_DIM + jj + 64));
        k_s[jj] = a; k_s[jj + 64] = b; v_s[jj] = v1; v_s[jj + 64] = v2;
        const size_t cidx = (((size_t)r * H + h) * Tmax + pos) * HEAD_DIM + jj;
        kcache[cidx] = __float2bfloat16_rn(a); kcache[cidx + 64] = __float2bfloat16_rn(b);
        vcache[cidx] = __float2bfloat16_rn(v1); vcache[cidx + 64] = __float2bfloat16_rn(v2);
      }
    }
    named_bar_sync(1, AT_NW * 32);
    float qv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) qv[i] = q_s[lane * 4 + i];
    float m = -INFINITY, l = 0.f, o[4] = {0.f, 0.f, 0.f, 0.f};
    const int first_tile_k = seg_lo - item_lo;                            // tile index within the item
    // Tile jt of the CTA's stream lives in stage jt % AT_STAGES and is consumed by warp GROUP jt % AT_NG
    // (4 warps, 8 tokens of the tile each, every warp with its own online-softmax state).  With AT_STAGES a
    // multiple of AT_NG every stage belongs to exactly ONE group, whose warps therefore observe every phase
    // of that stage's mbarrier in order.  (TMA loads complete out of order: a warp that waited on a stage
    // whose previous phase it had not itself observed would see parity aliasing and read stale data.)
    const int grp = warp / AT_GW, wig = warp % AT_GW;
    for (int jt = j + ((grp - j) % AT_NG + AT_NG) % AT_NG; jt < j + n_tiles_seg; jt += AT_NG) {
      const int t = jt - j;
      const int s = jt % AT_STAGES;
      const uint32_t round = (uint32_t)(jt / AT_STAGES);
      mbar_wait(&full_bar[s], round & 1u, 12, jt * 100000 + n_tiles_seg * 100 + (pos % 100));
      const uint8_t* kt = ring + (size_t)s * 2 * AT_TILE_BYTES + wig * AT_TW * (HEAD_DIM * 2);
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
      // transposed reduction over lane bits 4,3,2 (8 -> 1 value per lane), then plain butterflies over bits
      // 1,0: afterwards every lane holds the full dot product of token (lane >> 2)
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
      const float p = valid ? exp2f(sv - mx) : 0.f;
      const float corr = (mx == -INFINITY) ? 1.f : exp2f(m - mx);
      // every token's p is replicated on 4 lanes: sum over the warp counts it 4 times
      l = l * corr + 0.25f * warp_sum(p);
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] *= corr;
#pragma unroll
      for (int i = 0; i < AT_TW; ++i) {
        const float pi = __shfl_sync(0xffffffffu, p, i * 4);
        const uint2 vv = *reinterpret_cast<const uint2*>(vt + i * (HEAD_DIM * 2) + lane * 8);
        o[0] = fmaf(pi, bf16lo(vv.x), o[0]); o[1] = fmaf(pi, bf16hi(vv.x), o[1]);
        o[2] = fmaf(pi, bf16lo(vv.y), o[2]); o[3] = fmaf(pi, bf16hi(vv.y), o[3]);
      }
      m = mx;
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);                          // 4 arrivals (one per warp of the group) free the stage
    }
    // ---- the token being decoded (owner CTA, warp 0), straight from shared memory
    if (owns_new && warp == 0) {
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) d = fmaf(k_s[lane * 4 + i], qv[i], d);
      d = warp_sum(d);
      const float mx = fmaxf(m, d);
      const float corr = (m == -INFINITY) ? 0.f : exp2f(m - mx);
      const float p = exp2f(d - mx);
      l = l * corr + p;
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = fmaf(p, v_s[lane * 4 + i], o[i] * corr);
      m = mx;
    }
    // ---- merge the consumer warps of this CTA
    if (lane == 0) { m_s[warp] = m; l_s[warp] = l; }
#pragma unroll
    for (int i = 0; i < 4; ++i) o_s[warp][lane * 4 + i] = o[i];
    named_bar_sync(1, AT_NW * 32);
    const int c_first = item_lo / per, c_last = (item_hi - 1) / per;      // CTAs that touch this item
    const int n_contrib = c_last - c_first + 1;
    const int it = r * H + h;
    const bool out_thread = tid < HEAD_DIM;                               // one output dim each
    float M = -INFINITY, Ltot = 0.f, acc = 0.f;
    if (out_thread) {
#pragma unroll
      for (int w = 0; w < AT_NW; ++w) M = fmaxf(M, m_s[w]);
#pragma unroll
      for (int w = 0; w < AT_NW; ++w) {
        const float f = (m_s[w] == -INFINITY) ? 0.f : exp2f(m_s[w] - M);
        Ltot += l_s[w] * f;
        acc += o_s[w][tid] * f;
      }
    }
    const size_t oidx = (size_t)r * HD + h * HEAD_DIM + tid;
    if (n_contrib == 1) {
      if (out_thread) out[oidx] = __float2bfloat16_rn(acc / Ltot);
      named_bar_sync(1, AT_NW * 32);                                      // o_s / m_s are rewritten by the next segment
    } else {
      if (out_thread) {
        float* wp = ws_part + ((size_t)it * AT_MAX_SLOTS + (blockIdx.x - c_first)) * (HEAD_DIM + 2);
        wp[tid] = acc;
        if (tid == 0) { wp[HEAD_DIM] = M; wp[HEAD_DIM + 1] = Ltot; }
        __threadfence();
      }
      named_bar_sync(1, AT_NW * 32);
      if (tid == 0) {
        const int prev = atomicAdd(ws_count + it, 1);
        is_last_s = (prev == n_contrib - 1);
      }
      named_bar_sync(1, AT_NW * 32);
      if (is_last_s && out_thread) {
        __threadfence();
        const float* wb = ws_part + (size_t)it * AT_MAX_SLOTS * (HEAD_DIM + 2);
        float Mg = -INFINITY;
        for (int s2 = 0; s2 < n_contrib; ++s2) Mg = fmaxf(Mg, __ldcg(wb + (size_t)s2 * (HEAD_DIM + 2) + HEAD_DIM));
        float Lg = 0.f, og = 0.f;
        for (int s2 = 0; s2 < n_contrib; ++s2) {
          const float ms = __ldcg(wb + (size_t)s2 * (HEAD_DIM + 2) + HEAD_DIM);
          const float f = (ms == -INFINITY) ? 0.f : exp2f(ms - Mg);
          Lg += __ldcg(wb + (size_t)s2 * (HEAD_DIM + 2) + HEAD_DIM + 1) * f;
          og += __ldcg(wb + (size_t)s2 * (HEAD_DIM + 2) + tid) * f;
        }
        out[oidx] = __float2bfloat16_rn(og / Lg);
        if (tid == 0) ws_count[it] = 0;                                   // re-arm for the next launch
      }
      named_bar_sync(1, AT_NW * 32);
    }
    j += n_tiles_seg;
    u = seg_hi;
  }
}

}  // namespace pg
