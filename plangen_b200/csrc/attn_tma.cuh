// KV-cache decode attention for the paired cond/uncond CFG batch, bf16 cache, one launch per layer.
//
// Design (B200): the kernel is a persistent HBM streamer, one CTA per SM.  Work = "units" of 32 cached
// tokens of one (row, head) item plus one extra unit per item for the token being decoded.  The flat
// unit space [0, U) is cut into equal contiguous ranges, one per CTA, and every CTA range again into 4
// contiguous sub-ranges, one per consumer GROUP (4 warps), so rows with long prompts (cond) and short
// prompts (uncond) are balanced exactly.  Each group is an independent stream:
//   * it owns AT_SPG private ring stages (K 8 KB + V 8 KB each) fed by its own producer warp with TMA bulk
//     copies (cp.async.bulk), in order;
//   * its 4 warps each take 8 of a tile's 32 tokens and keep their own fp32 online-softmax state (scores
//     by a transposed warp reduction, P.V with one row broadcast per token);
//   * item boundaries (RoPE'd q of the next item, merge of the 4 warps, hand-off to other contributors)
//     only synchronise the group's 128 threads - the other groups keep streaming.
// Because a stage always belongs to the same group, its warps observe every phase of the stage's
// mbarriers in order (TMA loads complete out of order; a consumer that skipped a phase would alias
// parities).  Cached tokens do not depend on the current step, so with PDL the rings are filled BEFORE
// griddepcontrol.wait (while the QKV contraction is still running); only q and the new k/v wait.
// An item finished by a single group is written directly; otherwise partial (m, l, o) records are merged
// by the last contributor to arrive (threadfence + counter; fixed summation order -> deterministic).
#pragma once
#include "common.cuh"
#include "lm_kernels.cuh"

namespace pg {

constexpr int AT_TILE = 32;                             // tokens per unit
constexpr int AT_TILE_BYTES = AT_TILE * HEAD_DIM * 2;   // 8 KB (K) and 8 KB (V)
constexpr int AT_NG = 4;                                // consumer groups per CTA
constexpr int AT_GW = 4;                                // warps per group
constexpr int AT_GT = AT_GW * 32;                       // threads per group
constexpr int AT_NW = AT_NG * AT_GW;
constexpr int AT_TW = AT_TILE / AT_GW;                  // tokens per warp per tile (8)
#ifndef PG_AT_SPG
#define PG_AT_SPG 3
#endif
constexpr int AT_SPG = PG_AT_SPG;                      // ring stages per group
constexpr int AT_STAGES = AT_NG * AT_SPG;
constexpr int AT_THREADS = 32 * (AT_NW + AT_NG);      // 16 consumer warps + one producer warp per group
constexpr int AT_MAX_ROWS = 256;
constexpr int AT_MAX_SLOTS = 64;                        // partial records per item
constexpr int AT_SMEM = AT_STAGES * 2 * AT_TILE_BYTES + 1024;
static_assert(AT_TW == 8, "the transposed reduction below is written for 8 tokens per warp");

PG_DEVINL void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// explicit shared-space 8-byte load (generic loads through a function-argument pointer are not proven to
// be shared memory by the compiler and take the slow generic path)
PG_DEVINL uint2 lds_v2(uint32_t saddr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
  return v;
}
PG_DEVINL void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
PG_DEVINL bool mbar_test_wait(uint64_t* bar, uint32_t parity) {       // non-blocking probe
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// debug timeline: stamp k of group g of this CTA (16 stamps per group); dbg == nullptr in production
PG_DEVINL void at_stamp(unsigned long long* dbg, int g, int k) {
  if (dbg) dbg[((size_t)blockIdx.x * AT_NG + g) * 16 + k] = global_timer_ns();
}

// units of one row: tiles covering cached tokens [start, pos) on a 32-token grid, plus the new token
PG_DEVINL int row_tiles(int start, int pos) { return pos > start ? ((pos - 1) / AT_TILE - start / AT_TILE + 1) : 0; }

// exclusive prefix of units per row into row_units[0..R] (one warp)
PG_DEVINL void build_row_units(int* row_units, const int32_t* kv_start, int R, int pos, int lane,
                                int* row_start = nullptr, int new_token_unit = 1) {
  int carry = 0;
  for (int r0 = 0; r0 < R; r0 += 32) {
    const int r = r0 + lane;
    const int st = (r < R) ? kv_start[r] : 0;
    if (row_start && r < R) row_start[r] = st;
    int u = (r < R) ? row_tiles(st, pos) + new_token_unit : 0;
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

// How the flat unit space is cut: CTA c gets [c*per, (c+1)*per), its group g gets [.. + g*sub, .. + (g+1)*sub)
struct AttnCut {
  int U, per, sub;
  PG_DEVINL void group_range(int c, int g, int& lo, int& hi) const {
    const int cb = min(c * per, U), ce = min(cb + per, U);
    lo = min(cb + g * sub, ce);
    hi = min(lo + sub, ce);
  }
  PG_DEVINL int slot_of(int u) const { return (u / per) * AT_NG + (u % per) / sub; }
  PG_DEVINL bool slot_nonempty(int slot) const {
    int lo, hi;
    group_range(slot / AT_NG, slot % AT_NG, lo, hi);
    return lo < hi;
  }
};

struct AttnGroupSmem {
  float q[HEAD_DIM], k[HEAD_DIM], v[HEAD_DIM];
  float m[AT_GW], l[AT_GW], o[AT_GW][HEAD_DIM];
  int is_last;
};

// One consumer group's whole stream: units [gb, ge) of the flat space.  The group's t-th tile lives in
// ring stage stage_of[(kc + t) % SPG] with parity ((kc + t) / SPG) & 1.  All 128 threads of the group call this.
template <int SPG>
PG_DEVINL void attn_group_stream(int tg, int gb, int ge, const AttnCut& cut, const int* row_units, int R, int H,
                                 int Tmax, int pos, const float* __restrict__ part, int S, size_t split_stride,
                                 const float* __restrict__ cosT, const float* __restrict__ sinT,
                                 bf16* __restrict__ kcache, bf16* __restrict__ vcache,
                                 const int32_t* __restrict__ kv_start, bf16* __restrict__ out,
                                 float* __restrict__ ws_part, int* __restrict__ ws_count, float scale, bool bf16_trig,
                                 uint8_t* ring, int stage_stride_bytes, const int* stage_of, uint64_t* full_bar,
                                 uint64_t* empty_bar, int& kc, AttnGroupSmem& sm, int my_slot, int bar_id,
                                 int dbg_skip_math = 0, unsigned long long* dbg = nullptr) {
  const int lane = tg & 31, wig = tg >> 5;
  const int HD = H * HEAD_DIM;
  const float LOG2E = 1.4426950408889634f;
  const uint32_t ring_s = smem_u32(ring);
  int u = gb;
  int r = 0;
  while (r + 1 < R && row_units[r + 1] * H <= u) ++r;
  while (u < ge) {
    while (row_units[r + 1] * H <= u) ++r;
    const int ur = row_units[r + 1] - row_units[r];
    const int item_base = row_units[r] * H;
    const int h = (u - item_base) / ur;
    const int item_lo = item_base + h * ur, item_hi = item_lo + ur;       // flat units of this item
    const int seg_lo = u, seg_hi = min(item_hi, ge);
    const int start = kv_start[r];
    const bool owns_new = (seg_hi == item_hi);                            // this group holds the new-token unit
    const int n_tiles_seg = (seg_hi - seg_lo) - (owns_new ? 1 : 0);
    // ---- q (all), k/v of the new token (owner): reduce split-K partials, RoPE
    {
      const float* row = part + (size_t)r * 3 * HD;
      const int jj = tg & 63;
      const float c = cosT[pos * 64 + jj], sn = sinT[pos * 64 + jj];
      if (tg < 64) {
        const float x1 = bf16_round(reduce_splits(row, S, split_stride, (size_t)h * HEAD_DIM + jj));
        const float x2 = bf16_round(reduce_splits(row, S, split_stride, (size_t)h * HEAD_DIM + jj + 64));
        float a, b;
        rope_pair<bf16>(x1, x2, c, sn, bf16_trig, a, b);
        sm.q[jj] = a * (scale * LOG2E); sm.q[jj + 64] = b * (scale * LOG2E);
      } else if (owns_new) {
        const float x1 = bf16_round(reduce_splits(row, S, split_stride, (size_t)HD + h * HEAD_DIM + jj));
        const float x2 = bf16_round(reduce_splits(row, S, split_stride, (size_t)HD + h * HEAD_DIM + jj + 64));
        float a, b;
        rope_pair<bf16>(x1, x2, c, sn, bf16_trig, a, b);
        const float v1 = bf16_round(reduce_splits(row, S, split_stride, (size_t)2 * HD + h * HEAD_DIM + jj));
        const float v2 = bf16_round(reduce_splits(row, S, split_stride, (size_t)2 * HD + h * HEAD_DIM + jj + 64));
        sm.k[jj] = a; sm.k[jj + 64] = b; sm.v[jj] = v1; sm.v[jj + 64] = v2;
        const size_t cidx = (((size_t)r * H + h) * Tmax + pos) * HEAD_DIM + jj;
        kcache[cidx] = __float2bfloat16_rn(a); kcache[cidx + 64] = __float2bfloat16_rn(b);
        vcache[cidx] = __float2bfloat16_rn(v1); vcache[cidx + 64] = __float2bfloat16_rn(v2);
      }
    }
    named_bar_sync(bar_id, AT_GT);
    if (tg == 0 && u == gb) at_stamp(dbg, bar_id - 1, 2);
    float qv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) qv[i] = sm.q[lane * 4 + i];
    float m = -INFINITY, l = 0.f, o[4] = {0.f, 0.f, 0.f, 0.f};
    const int first_tile_k = seg_lo - item_lo;                            // tile index within the item
    for (int t = 0; t < n_tiles_seg; ++t, ++kc) {
      const int s = stage_of[kc % SPG];
      mbar_wait(&full_bar[s], (uint32_t)(kc / SPG) & 1u, 12, kc);
      if (tg == 0 && kc == 0) at_stamp(dbg, bar_id - 1, 3);
      if (dbg_skip_math) {                                                // profiling aid: measure the pure stream rate
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
        continue;
      }
      const uint32_t kt = ring_s + (uint32_t)(s * stage_stride_bytes + wig * AT_TW * (HEAD_DIM * 2) + lane * 8);
      const uint32_t vt = kt + AT_TILE_BYTES;
      const int t0 = (start / AT_TILE + first_tile_k + t) * AT_TILE + wig * AT_TW;
      float sc[AT_TW];
#pragma unroll
      for (int i = 0; i < AT_TW; ++i) {
        const uint2 kk = lds_v2(kt + i * (HEAD_DIM * 2));
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
      l = l * corr + 0.25f * warp_sum(p);                                 // every token's p sits on 4 lanes
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] *= corr;
#pragma unroll
      for (int i = 0; i < AT_TW; ++i) {
        const float pi = __shfl_sync(0xffffffffu, p, i * 4);
        const uint2 vv = lds_v2(vt + i * (HEAD_DIM * 2));
        o[0] = fmaf(pi, bf16lo(vv.x), o[0]); o[1] = fmaf(pi, bf16hi(vv.x), o[1]);
        o[2] = fmaf(pi, bf16lo(vv.y), o[2]); o[3] = fmaf(pi, bf16hi(vv.y), o[3]);
      }
      m = mx;
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);                          // AT_GW arrivals free the stage
    }
    if (tg == 0 && seg_hi == ge) at_stamp(dbg, bar_id - 1, 4);
    // ---- the token being decoded (owner, warp 0 of the group), straight from shared memory
    if (owns_new && wig == 0) {
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) d = fmaf(sm.k[lane * 4 + i], qv[i], d);
      d = warp_sum(d);
      const float mx = fmaxf(m, d);
      const float corr = (m == -INFINITY) ? 0.f : exp2f(m - mx);
      const float p = exp2f(d - mx);
      l = l * corr + p;
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = fmaf(p, sm.v[lane * 4 + i], o[i] * corr);
      m = mx;
    }
    // ---- merge the 4 warps of the group
    if (lane == 0) { sm.m[wig] = m; sm.l[wig] = l; }
#pragma unroll
    for (int i = 0; i < 4; ++i) sm.o[wig][lane * 4 + i] = o[i];
    named_bar_sync(bar_id, AT_GT);
    float M = -INFINITY, Ltot = 0.f, acc = 0.f;
#pragma unroll
    for (int w = 0; w < AT_GW; ++w) M = fmaxf(M, sm.m[w]);
#pragma unroll
    for (int w = 0; w < AT_GW; ++w) {
      const float f = (sm.m[w] == -INFINITY) ? 0.f : exp2f(sm.m[w] - M);
      Ltot += sm.l[w] * f;
      acc += sm.o[w][tg] * f;                                             // tg < 128: one output dim each
    }
    // contributors of this item = non-empty group slots intersecting [item_lo, item_hi)
    const int s_first = cut.slot_of(item_lo), s_last = cut.slot_of(item_hi - 1);
    int n_contrib = 0, my_rank = 0;
    for (int sl = s_first; sl <= s_last; ++sl) {
      if (cut.slot_nonempty(sl)) {
        if (sl < my_slot) ++my_rank;
        ++n_contrib;
      }
    }
    const int it = r * H + h;
    const size_t oidx = (size_t)r * HD + h * HEAD_DIM + tg;
    if (n_contrib == 1) {
      out[oidx] = __float2bfloat16_rn(acc / Ltot);
    } else {
      float* wp = ws_part + ((size_t)it * AT_MAX_SLOTS + my_rank) * (HEAD_DIM + 2);
      wp[tg] = acc;
      if (tg == 0) { wp[HEAD_DIM] = M; wp[HEAD_DIM + 1] = Ltot; }
      __threadfence();
      named_bar_sync(bar_id, AT_GT);
      if (tg == 0) {
        const int prev = atomicAdd(ws_count + it, 1);
        sm.is_last = (prev == n_contrib - 1);
      }
      named_bar_sync(bar_id, AT_GT);
      if (sm.is_last) {
        __threadfence();
        const float* wb = ws_part + (size_t)it * AT_MAX_SLOTS * (HEAD_DIM + 2);
        float Mg = -INFINITY;
        for (int s2 = 0; s2 < n_contrib; ++s2) Mg = fmaxf(Mg, __ldcg(wb + (size_t)s2 * (HEAD_DIM + 2) + HEAD_DIM));
        float Lg = 0.f, og = 0.f;
        for (int s2 = 0; s2 < n_contrib; ++s2) {
          const float ms = __ldcg(wb + (size_t)s2 * (HEAD_DIM + 2) + HEAD_DIM);
          const float f = (ms == -INFINITY) ? 0.f : exp2f(ms - Mg);
          Lg += __ldcg(wb + (size_t)s2 * (HEAD_DIM + 2) + HEAD_DIM + 1) * f;
          og += __ldcg(wb + (size_t)s2 * (HEAD_DIM + 2) + tg) * f;
        }
        out[oidx] = __float2bfloat16_rn(og / Lg);
        if (tg == 0) ws_count[it] = 0;                                    // re-arm for the next launch
      }
    }
    named_bar_sync(bar_id, AT_GT);                                        // sm.* is rewritten by the next item
    u = seg_hi;
  }
}

// Producer side of ONE group stream (one dedicated producer warp per group, lane 0): in-order TMA bulk
// copies of the group's K/V tiles into its private stages.  A single thread feeding all four groups
// (~1300 cycles of address arithmetic, barrier probe and two bulk-copy issues per tile) was the bottleneck
// of the kernel; four independent issuers are not.
template <int SPG, bool NEW_TOKEN_UNIT = true>
PG_DEVINL void attn_produce_group(int gb, int ge, const int* row_units, int R, int H, int Tmax,
                                  const int32_t* __restrict__ kv_start, const bf16* __restrict__ kcache,
                                  const bf16* __restrict__ vcache, uint8_t* ring, int stage_stride_bytes,
                                  const int* stage_of, uint64_t* full_bar, uint64_t* empty_bar, int& kload,
                                  uint32_t tx_bytes, uint64_t pol, int r_hint = 0) {
  int r = r_hint;
  while (r + 1 < R && row_units[r + 1] * H <= gb) ++r;
  int ur = row_units[r + 1] - row_units[r];
  int local = gb - row_units[r] * H;
  int h = local / ur, k = local % ur;
  for (int u = gb; u < ge; ++u) {
    if (!NEW_TOKEN_UNIT || k != ur - 1) {                // the new-token unit has no cached tile
      const int s = stage_of[kload % SPG];
      mbar_wait(&empty_bar[s], (((uint32_t)(kload / SPG)) & 1u) ^ 1u, 11, kload);
      mbar_expect_tx(&full_bar[s], tx_bytes);
      const int t0k = (kv_start[r] / AT_TILE + k) * AT_TILE;
      const size_t off = (((size_t)r * H + h) * Tmax + t0k) * HEAD_DIM;
      bulk_load(ring + (size_t)s * stage_stride_bytes, kcache + off, AT_TILE_BYTES, &full_bar[s], pol);
      bulk_load(ring + (size_t)s * stage_stride_bytes + AT_TILE_BYTES, vcache + off, AT_TILE_BYTES, &full_bar[s], pol);
      ++kload;
    }
    // advance (r, h, k) to the next unit without divisions
    if (++k == ur) {
      k = 0;
      if (++h == H) {
        h = 0;
        ++r;
        if (r < R) ur = row_units[r + 1] - row_units[r];
      }
    }
  }
}

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_decode_tma_kernel(const float* __restrict__ part, int S, size_t split_stride, const float* __restrict__ cosT,
                       const float* __restrict__ sinT, bf16* __restrict__ kcache, bf16* __restrict__ vcache,
                       const int32_t* __restrict__ kv_start, bf16* __restrict__ out, float* __restrict__ ws_part,
                       int* __restrict__ ws_count, int R, int H, int Tmax, int pos_base,
                       const int* __restrict__ step_ptr, float scale, int bf16_trig, int early_trigger, Prof prof,
                       unsigned long long* dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[AT_STAGES], empty_bar[AT_STAGES];
  __shared__ int row_units[AT_MAX_ROWS + 1];        // exclusive prefix of units per row (per head)
  __shared__ AttnGroupSmem gsm[AT_NG];
  __shared__ int stage_tab[AT_NG * AT_SPG];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < AT_NG) at_stamp(dbg, tid, 0);
  if (early_trigger & 1) pdl_launch_dependents();
  prof_begin(prof);
  // The step counter is only written by the last kernel of a decode step; graph replays are fully ordered,
  // and with plain launches the host passes the position explicitly (step_ptr == nullptr), so reading it
  // before the PDL wait is safe.
  const int pos = pos_base + (step_ptr ? *step_ptr : 0);

  if (tid == 0) {
    for (int i = 0; i < AT_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], AT_GW); }
    mbar_fence_init();
  }
  if (tid < AT_STAGES) stage_tab[(tid % AT_NG) * AT_SPG + tid / AT_NG] = tid;   // group g owns stages g, g+4, g+8
  if (warp == 0) build_row_units(row_units, kv_start, R, pos, lane);
  __syncthreads();
  if (tid < AT_NG) at_stamp(dbg, tid, 1);
  AttnCut cut;
  cut.U = row_units[R] * H;
  cut.per = max(1, (cut.U + (int)gridDim.x - 1) / (int)gridDim.x);
  cut.sub = (cut.per + AT_NG - 1) / AT_NG;
  const int c = blockIdx.x;

  if (warp >= AT_NW) {
    // ============================== producers: one warp (lane 0) per group stream ==============================
    if (lane == 0) {
      const int g = warp - AT_NW;
      int gb, ge;
      cut.group_range(c, g, gb, ge);
      int kload = 0;
      attn_produce_group<AT_SPG>(gb, ge, row_units, R, H, Tmax, kv_start, kcache, vcache, ring, 2 * AT_TILE_BYTES,
                                 stage_tab + g * AT_SPG, full_bar, empty_bar, kload, 2 * AT_TILE_BYTES,
                                 policy_evict_first());
      at_stamp(dbg, g, 7);
    }
    pdl_wait();
    return;
  }
  // ============================== consumers: 4 independent groups ==============================
  pdl_wait();                                             // QKV partials of this step are now visible
  const int g = warp / AT_GW, tg = tid - g * AT_GT;
  int gb, ge;
  cut.group_range(c, g, gb, ge);
  int kc = 0;
  attn_group_stream<AT_SPG>(tg, gb, ge, cut, row_units, R, H, Tmax, pos, part, S, split_stride, cosT, sinT, kcache,
                            vcache, kv_start, out, ws_part, ws_count, scale, (bf16_trig & 1) != 0, ring, 2 * AT_TILE_BYTES,
                            stage_tab + g * AT_SPG, full_bar, empty_bar, kc, gsm[g], c * AT_NG + g, 1 + g, early_trigger & 2, dbg);
  if (tg == 0) at_stamp(dbg, g, 5);
  prof_end(prof);
}

}  // namespace pg
