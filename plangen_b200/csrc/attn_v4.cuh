// Decode attention, generation 4: the group streams of attn_tma.cuh with every item-boundary latency moved
// off the streaming warps.
//
// Measured on B200 (tools/attn_scaling.py): the generation-3 kernel costs  t = 15.7 us + bytes / 7.6 TB/s ,
// i.e. the marginal rate is already the HBM rate and the loss is a fixed ~16 us per launch.  That fixed part
// is the serial work each consumer group does between tiles: reducing the split-K QKV partials + RoPE for
// the next item's q (one L2 round trip), and handing a finished segment to the other contributors
// (threadfence + atomic + last-arriver merge, three more round trips) - twice per group per launch, with
// the group's 48 KB ring only able to hide ~1 us of it.
//
// Here each group gets a HELPER warp next to its TMA producer warp:
//   * the helper walks the group's segments ahead of the consumers and fetches the split-K QKV partials of
//     up to 4 segments as ONE batch of loads (under the saturated KV stream an L2 round trip costs 4-6 us,
//     so dependent round trips are what must be avoided), RoPEs q into a 4-deep smem queue (q_full /
//     q_empty), then the new token's k/v (kv_full; also appended to the cache);
//   * the consumers only wait on q_full (normally ready), stream their tiles, merge their 4 warps through
//     shared memory and drop the segment result (M, L, o[128]) into a 2-deep outbox (o_full / o_empty);
//   * the helper drains the outbox: direct bf16 store for single-contributor items; otherwise a partial
//     record + release flag, fire and forget.  The LAST contributor of an item (it meets the item first in
//     its range and only waits for earlier groups) keeps its record in registers and merges all records in
//     rank order at the end of its stream.
// The token being decoded rides on the item's last tile instead of being a unit of its own, so no segment
// is ever empty.  Per-tile arithmetic is that of generation 3; the cut (hence the merge order) differs.
#pragma once
#include "attn_tma.cuh"

namespace pg {

constexpr int A4_NQ = 4;                                   // q queue depth per group
constexpr int A4_NO = 2;                                   // outbox depth per group
constexpr int A4_THREADS = 32 * (AT_NW + 2 * AT_NG);      // 16 consumer + 4 TMA producer + 4 helper warps
constexpr int A4_SMEM = AT_STAGES * 2 * AT_TILE_BYTES + 128;

struct A4Seg {
  int r;                                                   // row, -1 = end of stream
  int h, start, first_tile, n_tiles, owns_new;
};
struct A4QSlot {
  float q[HEAD_DIM];
  bf16 k[HEAD_DIM], v[HEAD_DIM];                           // bf16-rounded values: exact
  A4Seg seg;
};
struct A4Out {
  float o[HEAD_DIM];
  float M, L;
};
struct A4Group {
  A4QSlot qs[A4_NQ];
  A4Out ob[A4_NO];
  float m[AT_GW], l[AT_GW], o[AT_GW][HEAD_DIM];
  uint64_t q_full[A4_NQ], kv_full[A4_NQ], q_empty[A4_NQ], o_full[A4_NO], o_empty[A4_NO], start;
};

// first row whose units reach past flat unit u (whole warp; rows are few, a linear scan by one lane costs ~0.6 us)
PG_DEVINL int a4_row_of(const int* row_units, int R, int H, int u, int lane) {
  int r = 0;
  for (int r0 = 0; r0 < R; r0 += 32) {
    const int rr = r0 + lane;
    const bool below = (rr + 1 < R) && (row_units[rr + 1] * H <= u);
    r += __popc(__ballot_sync(0xffffffffu, below));
  }
  return r;
}

// ------------------------------------------------------------------ helper warp: segment walk
struct A4Walk {
  int u, r;
};
// describe the segment starting at unit w.u (w.u < ge) and advance w.u to its end (shared memory reads and
// one division; the contributor count is left to a4_contributors, off the start-up path)
PG_DEVINL A4Seg a4_next_segment(A4Walk& w, int ge, const int* row_units, int H, const int* row_start) {
  while (row_units[w.r + 1] * H <= w.u) ++w.r;
  const int r = w.r;
  const int ur = row_units[r + 1] - row_units[r];
  const int item_base = row_units[r] * H;
  const int h = (w.u - item_base) / ur;
  const int item_lo = item_base + h * ur, item_hi = item_lo + ur;
  const int seg_hi = min(item_hi, ge);
  A4Seg s;
  s.r = r; s.h = h; s.start = row_start[r];
  s.first_tile = w.u - item_lo;
  s.owns_new = (seg_hi == item_hi) ? 1 : 0;
  s.n_tiles = seg_hi - w.u;                                // every unit is a tile; the new token rides on the last one
  w.u = seg_hi;
  return s;
}
// contributors of item (r, h) = non-empty group slots intersecting its unit range; rank of my_slot among them
PG_DEVINL void a4_contributors(int r, int h, const AttnCut& cut, const int* row_units, int H, int my_slot,
                               int& n_contrib, int& my_rank) {
  const int ur = row_units[r + 1] - row_units[r];
  const int item_lo = row_units[r] * H + h * ur, item_hi = item_lo + ur;
  const int s_first = cut.slot_of(item_lo), s_last = cut.slot_of(item_hi - 1);
  n_contrib = 0; my_rank = 0;
  for (int sl = s_first; sl <= s_last; ++sl) {
    if (cut.slot_nonempty(sl)) {
      if (sl < my_slot) ++my_rank;
      ++n_contrib;
    }
  }
}

// Split-K slabs of up to 4 segments x NV values, summed left to right (the order of reduce_splits), with
// the loads of 3 slabs in flight per value: under the saturated KV stream one L2 round trip costs 4-6 us,
// so the helper never issues dependent loads one at a time.
template <int NV>
PG_DEVINL void a4_reduce_batch(const float* const (&base)[4], const bool (&on)[4], const int (&off)[NV], int S,
                               size_t split_stride, float (&acc)[4][NV], uint64_t* issued_bar = nullptr) {
  for (int s0 = 0; s0 < S; s0 += 3) {
    float t[3][4][NV];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int v = 0; v < NV; ++v)
          t[k][i][v] = (on[i] && s0 + k < S) ? __ldcg(base[i] + (size_t)(s0 + k) * split_stride + off[v]) : 0.f;
    if (issued_bar) { mbar_arrive(issued_bar); issued_bar = nullptr; }   // loads are in the pipe: release the KV stream
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (s0 + k < S) acc[i][v] = (s0 + k == 0) ? t[k][i][v] : acc[i][v] + t[k][i][v];
  }
}

// Segments j0 .. j0+n-1 (n <= 4): q of all of them in one batch of loads -> q_full, then the new token's k/v
// of the owner segments, two segments per batch -> kv_full (and appended to the cache).  Whole warp.
PG_DEVINL void a4_produce_batch(int lane, A4Group& G, int j0, int n, const A4Seg (&sg)[4], int H, int Tmax, int pos,
                                const float* __restrict__ part, int S, size_t split_stride,
                                const float (&cs)[2], const float (&sn)[2],
                                bf16* __restrict__ kcache, bf16* __restrict__ vcache, float scale, bool bf16_trig,
                                uint64_t* issued_bar = nullptr) {
  const float LOG2E = 1.4426950408889634f;
  const int HD = H * HEAD_DIM;
  const float* base[4];
  bool on[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    on[i] = i < n;
    base[i] = part + (size_t)(on[i] ? sg[i].r : 0) * 3 * HD + (on[i] ? sg[i].h : 0) * HEAD_DIM;
  }
  {
    // a lane owns the rotary pairs (lane, lane + 64) and (lane + 32, lane + 96)
    const int off[4] = {lane, lane + 64, lane + 32, lane + 96};
    float acc[4][4];
    a4_reduce_batch<4>(base, on, off, S, split_stride, acc, issued_bar);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (!on[i]) continue;
      const int slot = (j0 + i) % A4_NQ;
      mbar_wait(&G.q_empty[slot], (((uint32_t)((j0 + i) / A4_NQ)) & 1u) ^ 1u, 21, j0 + i);
      A4QSlot& Q = G.qs[slot];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        float a, b;
        rope_pair<bf16>(bf16_round(acc[i][2 * p]), bf16_round(acc[i][2 * p + 1]), cs[p], sn[p], bf16_trig, a, b);
        Q.q[lane + 32 * p] = a * (scale * LOG2E); Q.q[lane + 32 * p + 64] = b * (scale * LOG2E);
      }
      if (lane == 0) Q.seg = sg[i];
      mbar_arrive(&G.q_full[slot]);                        // 32 arrivals (release) publish q + descriptor
    }
  }
#pragma unroll
  for (int i0 = 0; i0 < 4; i0 += 2) {
    if (i0 >= n) break;
    bool own[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) own[i] = (i >= i0 && i < i0 + 2 && on[i] && sg[i].owns_new);
    if (own[i0] || own[i0 + 1]) {
      const int off[8] = {HD + lane, HD + lane + 64, HD + lane + 32, HD + lane + 96,
                          2 * HD + lane, 2 * HD + lane + 64, 2 * HD + lane + 32, 2 * HD + lane + 96};
      float acc[4][8];
      a4_reduce_batch<8>(base, own, off, S, split_stride, acc);
#pragma unroll
      for (int i = i0; i < i0 + 2; ++i) {
        if (!own[i]) continue;
        A4QSlot& Q = G.qs[(j0 + i) % A4_NQ];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const int jj = lane + 32 * p;
          float a, b;
          rope_pair<bf16>(bf16_round(acc[i][2 * p]), bf16_round(acc[i][2 * p + 1]), cs[p], sn[p], bf16_trig, a, b);
          const float v1 = bf16_round(acc[i][4 + 2 * p]), v2 = bf16_round(acc[i][4 + 2 * p + 1]);
          Q.k[jj] = __float2bfloat16_rn(a); Q.k[jj + 64] = __float2bfloat16_rn(b);
          Q.v[jj] = __float2bfloat16_rn(v1); Q.v[jj + 64] = __float2bfloat16_rn(v2);
          const size_t cidx = (((size_t)sg[i].r * H + sg[i].h) * Tmax + pos) * HEAD_DIM + jj;
          kcache[cidx] = __float2bfloat16_rn(a); kcache[cidx + 64] = __float2bfloat16_rn(b);
          vcache[cidx] = __float2bfloat16_rn(v1); vcache[cidx + 64] = __float2bfloat16_rn(v2);
        }
      }
    }
#pragma unroll
    for (int i = i0; i < i0 + 2; ++i)
      if (on[i]) mbar_arrive(&G.kv_full[(j0 + i) % A4_NQ]);   // every use of a slot completes one kv phase
  }
}
// end-of-stream marker in queue slot j
PG_DEVINL void a4_publish_end(int lane, A4Group& G, int j) {
  const int slot = j % A4_NQ;
  mbar_wait(&G.q_empty[slot], (((uint32_t)(j / A4_NQ)) & 1u) ^ 1u, 25, j);
  if (lane == 0) G.qs[slot].seg.r = -1;
  mbar_arrive(&G.q_full[slot]);
  mbar_arrive(&G.kv_full[slot]);
}

PG_DEVINL void a4_store_out(bf16* orow, const float (&o)[4], float L) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(o[0] / L, o[1] / L), b = __floats2bfloat162_rn(o[2] / L, o[3] / L);
  uint2 t; t.x = *reinterpret_cast<const uint32_t*>(&a); t.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(orow) = t;
}

// A segment result that still needs the other contributors' records (kept in the helper's registers)
struct A4Deferred {
  float o[4], M, L;
  int it, n_contrib;                                       // n_contrib == 0: nothing deferred
};
constexpr int A4_REC = HEAD_DIM + 2;                       // 130 floats (M, L after the 128 outputs); 8-byte aligned

// finished segment j: outbox slot -> output row, or partial record + flag.  An item cut over several groups
// is finished by its LAST contributor (it only ever waits for earlier groups - no residency assumption);
// that group meets the item first in its range, so the merge is deferred to the end of the helper.
PG_DEVINL void a4_drain(int lane, A4Group& G, int j, int it, int n_contrib, int my_rank, bf16* __restrict__ out,
                        float* __restrict__ ws_part, int* __restrict__ flags, A4Deferred& df) {
  const int slot = j % A4_NO;
  mbar_wait(&G.o_full[slot], ((uint32_t)(j / A4_NO)) & 1u, 24, j);
  const A4Out& O = G.ob[slot];
  float acc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = O.o[lane * 4 + i];
  const float M = O.M, L = O.L;
  mbar_arrive(&G.o_empty[slot]);                           // 32 arrivals: everything is in registers
  if (n_contrib == 1) {
    a4_store_out(out + (size_t)it * HEAD_DIM + lane * 4, acc, L);
    return;
  }
  if (my_rank == n_contrib - 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) df.o[i] = acc[i];
    df.M = M; df.L = L; df.it = it; df.n_contrib = n_contrib;
    return;
  }
  float* wp = ws_part + ((size_t)it * AT_MAX_SLOTS + my_rank) * A4_REC;
  *reinterpret_cast<float2*>(wp + lane * 4) = make_float2(acc[0], acc[1]);
  *reinterpret_cast<float2*>(wp + lane * 4 + 2) = make_float2(acc[2], acc[3]);
  if (lane == 0) *reinterpret_cast<float2*>(wp + HEAD_DIM) = make_float2(M, L);
  __syncwarp();
  if (lane == 0) {                                         // release (cumulative over the warp's record stores)
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flags + (size_t)it * AT_MAX_SLOTS + my_rank), "r"(1) : "memory");
  }
}
// the deferred merge: wait for the records of ranks 0 .. n-2, combine in rank order (own record last)
PG_DEVINL void a4_finish_deferred(int lane, const A4Deferred& df, bf16* __restrict__ out,
                                  const float* __restrict__ ws_part, int* __restrict__ flags) {
  if (df.n_contrib == 0) return;
  const int n_other = df.n_contrib - 1;
  int* fl = flags + (size_t)df.it * AT_MAX_SLOTS;
  unsigned long long t0 = 0;
  for (int b0 = 0; b0 < n_other; b0 += 32) {               // lane s polls the flag of rank b0 + s
    const int rk = b0 + lane;
    for (;;) {
      int f = 1;
      if (rk < n_other) asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(fl + rk) : "memory");
      if (__all_sync(0xffffffffu, f != 0)) break;
      if (t0 == 0) t0 = global_timer_ns();
      else if (global_timer_ns() - t0 > 4000000000ull) {
        if (lane == 0) printf("attn v4: hand-off flag timeout item %d (cta %d)\n", df.it, (int)blockIdx.x);
        __trap();
      }
    }
    if (rk < n_other) fl[rk] = 0;                          // re-arm for the next launch
  }
  const float* wb = ws_part + (size_t)df.it * AT_MAX_SLOTS * A4_REC;
  float Mg = df.M;
  for (int s2 = 0; s2 < n_other; ++s2) Mg = fmaxf(Mg, __ldcg(wb + (size_t)s2 * A4_REC + HEAD_DIM));
  float Lg = 0.f, og[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s2 = 0; s2 < n_other; ++s2) {
    const float2 ml = __ldcg(reinterpret_cast<const float2*>(wb + (size_t)s2 * A4_REC + HEAD_DIM));
    const float2 o01 = __ldcg(reinterpret_cast<const float2*>(wb + (size_t)s2 * A4_REC + lane * 4));
    const float2 o23 = __ldcg(reinterpret_cast<const float2*>(wb + (size_t)s2 * A4_REC + lane * 4 + 2));
    const float f = (ml.x == -INFINITY) ? 0.f : exp2f(ml.x - Mg);
    Lg += ml.y * f;
    og[0] += o01.x * f; og[1] += o01.y * f; og[2] += o23.x * f; og[3] += o23.y * f;
  }
  {
    const float f = (df.M == -INFINITY) ? 0.f : exp2f(df.M - Mg);
    Lg += df.L * f;
#pragma unroll
    for (int i = 0; i < 4; ++i) og[i] += df.o[i] * f;
  }
  a4_store_out(out + (size_t)df.it * HEAD_DIM + lane * 4, og, Lg);
}

// ------------------------------------------------------------------ consumers (128 threads of one group)
template <int SPG>
PG_DEVINL void a4_consume(int tg, A4Group& G, int H, int pos, uint8_t* ring, int stage_stride_bytes,
                          const int* stage_of, uint64_t* full_bar, uint64_t* empty_bar, int bar_id, int dbg_skip_math,
                          unsigned long long* dbg) {
  const int lane = tg & 31, wig = tg >> 5;
  const uint32_t ring_s = smem_u32(ring);
  int kc = 0;
  for (int j = 0;; ++j) {
    const int qslot = j % A4_NQ;
    mbar_wait(&G.q_full[qslot], ((uint32_t)(j / A4_NQ)) & 1u, 22, j);
    const A4QSlot& Q = G.qs[qslot];
    const A4Seg seg = Q.seg;
    if (seg.r < 0) break;
    if (tg == 0 && j == 0) at_stamp(dbg, bar_id - 1, 2);
    float qv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) qv[i] = Q.q[lane * 4 + i];
    float m = -INFINITY, l = 0.f, o[4] = {0.f, 0.f, 0.f, 0.f};
    const int start = seg.start;
    for (int t = 0; t < seg.n_tiles; ++t, ++kc) {
      const int s = stage_of[kc % SPG];
      mbar_wait(&full_bar[s], (uint32_t)(kc / SPG) & 1u, 12, kc);
      if (tg == 0 && kc == 0) at_stamp(dbg, bar_id - 1, 3);
      if (dbg_skip_math) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
        continue;
      }
      const uint32_t kt = ring_s + (uint32_t)(s * stage_stride_bytes + wig * AT_TW * (HEAD_DIM * 2) + lane * 8);
      const uint32_t vt = kt + AT_TILE_BYTES;
      const int t0 = (start / AT_TILE + seg.first_tile + t) * AT_TILE + wig * AT_TW;
      float sc[AT_TW];
#pragma unroll
      for (int i = 0; i < AT_TW; ++i) {
        const uint2 kk = lds_v2(kt + i * (HEAD_DIM * 2));
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
      const float p = valid ? exp2f(sv - mx) : 0.f;
      const float corr = (mx == -INFINITY) ? 1.f : exp2f(m - mx);
      l = l * corr + 0.25f * warp_sum(p);
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
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    if (tg == 0) at_stamp(dbg, bar_id - 1, 4);             // (last write wins: end of the group's stream)
    if (seg.owns_new && wig == 0) {                        // the token being decoded, from the queue slot
      mbar_wait(&G.kv_full[qslot], ((uint32_t)(j / A4_NQ)) & 1u, 26, j);
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) d = fmaf(__bfloat162float(Q.k[lane * 4 + i]), qv[i], d);
      d = warp_sum(d);
      const float mx = fmaxf(m, d);
      const float corr = (m == -INFINITY) ? 0.f : exp2f(m - mx);
      const float p = exp2f(d - mx);
      l = l * corr + p;
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = fmaf(p, __bfloat162float(Q.v[lane * 4 + i]), o[i] * corr);
      m = mx;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&G.q_empty[qslot]);         // AT_GW arrivals free the queue slot
    // ---- merge the 4 warps of the group
    if (lane == 0) { G.m[wig] = m; G.l[wig] = l; }
#pragma unroll
    for (int i = 0; i < 4; ++i) G.o[wig][lane * 4 + i] = o[i];
    named_bar_sync(bar_id, AT_GT);
    float M = -INFINITY, Ltot = 0.f, acc = 0.f;
#pragma unroll
    for (int w = 0; w < AT_GW; ++w) M = fmaxf(M, G.m[w]);
#pragma unroll
    for (int w = 0; w < AT_GW; ++w) {
      const float f = (G.m[w] == -INFINITY) ? 0.f : exp2f(G.m[w] - M);
      Ltot += G.l[w] * f;
      acc += G.o[w][tg] * f;
    }
    const int oslot = j % A4_NO;
    mbar_wait(&G.o_empty[oslot], (((uint32_t)(j / A4_NO)) & 1u) ^ 1u, 23, j);
    A4Out& O = G.ob[oslot];
    O.o[tg] = acc;
    if (tg == 0) { O.M = M; O.L = Ltot; }
    mbar_arrive(&G.o_full[oslot]);                         // 128 arrivals publish the record
    named_bar_sync(bar_id, AT_GT);                         // G.m/l/o are rewritten by the next segment
  }
}

__global__ void __launch_bounds__(A4_THREADS, 1)
attn_decode_v4_kernel(const float* __restrict__ part, int S, size_t split_stride, const float* __restrict__ cosT,
                      const float* __restrict__ sinT, bf16* __restrict__ kcache, bf16* __restrict__ vcache,
                      const int32_t* __restrict__ kv_start, bf16* __restrict__ out, float* __restrict__ ws_part,
                      int* __restrict__ flags, int R, int H, int Tmax, int pos_base,
                      const int* __restrict__ step_ptr, float scale, int bf16_trig, int early_trigger, Prof prof,
                      unsigned long long* dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  __shared__ uint64_t full_bar[AT_STAGES], empty_bar[AT_STAGES];
  __shared__ int row_units[AT_MAX_ROWS + 1], row_start[AT_MAX_ROWS];
  __shared__ A4Group gsm[AT_NG];
  __shared__ int stage_tab[AT_NG * AT_SPG];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < AT_NG) at_stamp(dbg, tid, 0);
  if (early_trigger & 1) pdl_launch_dependents();
  prof_begin(prof);
  const int pos = pos_base + (step_ptr ? *step_ptr : 0);    // see attn_decode_tma_kernel

  if (tid == 0) {
    for (int i = 0; i < AT_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], AT_GW); }
    for (int g = 0; g < AT_NG; ++g) {
      for (int i = 0; i < A4_NQ; ++i) { mbar_init(&gsm[g].q_full[i], 32); mbar_init(&gsm[g].kv_full[i], 32); mbar_init(&gsm[g].q_empty[i], AT_GW); }
      for (int i = 0; i < A4_NO; ++i) { mbar_init(&gsm[g].o_full[i], AT_GT); mbar_init(&gsm[g].o_empty[i], 32); }
      mbar_init(&gsm[g].start, 32);
    }
    mbar_fence_init();
  }
  if (tid < AT_STAGES) stage_tab[(tid % AT_NG) * AT_SPG + tid / AT_NG] = tid;
  if (warp == 0) build_row_units(row_units, kv_start, R, pos, lane, row_start, 0);
  __syncthreads();
  if (tid < AT_NG) at_stamp(dbg, tid, 1);
  AttnCut cut;
  cut.U = row_units[R] * H;
  cut.per = max(1, (cut.U + (int)gridDim.x - 1) / (int)gridDim.x);
  cut.sub = (cut.per + AT_NG - 1) / AT_NG;
  const int c = blockIdx.x;
  const bool hold_stream = (early_trigger & 4) == 0;

  if (warp >= AT_NW + AT_NG) {
    // ============================== helpers: q ahead of the consumers, results behind them ==============================
    const int g = warp - AT_NW - AT_NG;
    A4Group& G = gsm[g];
    int gb, ge;
    cut.group_range(c, g, gb, ge);
    const int my_slot = c * AT_NG + g;
    float cs[2], sn[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) { cs[p] = cosT[pos * 64 + lane + 32 * p]; sn[p] = sinT[pos * 64 + lane + 32 * p]; }
    A4Walk w{gb, a4_row_of(row_units, R, H, gb, lane)};
    A4Walk wd = w;                                          // second walker, for the drains
    // the first A4_NQ segments are walked before the wait and fetched as one batch right after it
    A4Seg sg[4];
    int n0 = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sg[i] = A4Seg{};
      if (i < A4_NQ && w.u < ge) { sg[i] = a4_next_segment(w, ge, row_units, H, row_start); n0 = i + 1; }
    }
    pdl_wait();                                             // QKV partials of this step are now visible
    if (lane == 0) at_stamp(dbg, g, 9);
    int jq = 0;                                             // queue entries published so far (segments, then the end marker)
    bool ended = false;
    if (n0 > 0) {
      a4_produce_batch(lane, G, 0, n0, sg, H, Tmax, pos, part, S, split_stride, cs, sn, kcache, vcache, scale, (bf16_trig & 1) != 0,
                       hold_stream ? &G.start : nullptr);
      jq = n0;
    } else if (hold_stream) {
      mbar_arrive(&G.start);
    }
    if (lane == 0) at_stamp(dbg, g, 8);
    if (w.u >= ge && jq < A4_NQ) { a4_publish_end(lane, G, jq); ++jq; ended = true; }
    A4Deferred df;
    df.n_contrib = 0;
    for (int j = 0; j < jq - (ended ? 1 : 0); ++j) {
      const A4Seg sd = a4_next_segment(wd, ge, row_units, H, row_start);
      int n_contrib, my_rank;
      a4_contributors(sd.r, sd.h, cut, row_units, H, my_slot, n_contrib, my_rank);
      a4_drain(lane, G, j, sd.r * H + sd.h, n_contrib, my_rank, out, ws_part, flags, df);
      if (!ended) {
        if (w.u >= ge) { a4_publish_end(lane, G, jq); ended = true; }
        else {
          sg[0] = a4_next_segment(w, ge, row_units, H, row_start);
          a4_produce_batch(lane, G, jq, 1, sg, H, Tmax, pos, part, S, split_stride, cs, sn, kcache, vcache, scale, (bf16_trig & 1) != 0);
        }
        ++jq;
      }
    }
    a4_finish_deferred(lane, df, out, ws_part, flags);
    if (prof.buf && lane == 0) atomicMax(&prof.buf[PROF_SLOTS + prof.slot], (unsigned long long)global_timer_ns());
    if (lane == 0) at_stamp(dbg, g, 6);
    return;
  }
  if (warp >= AT_NW) {
    // ============================== TMA producers: one warp (lane 0) per group stream ==============================
    const int g = warp - AT_NW;
    int gb, ge;
    cut.group_range(c, g, gb, ge);
    const int r_hint = a4_row_of(row_units, R, H, gb, lane);
    if (lane == 0) {
      // the helper's q loads go into the memory pipe ahead of the KV flood (a round trip behind ~20 MB of
      // queued tile requests costs 4-5 us and would stall the consumers' start)
      if (hold_stream) mbar_wait(&gsm[g].start, 0, 27, g);
      int kload = 0;
      attn_produce_group<AT_SPG, false>(gb, ge, row_units, R, H, Tmax, row_start, kcache, vcache, ring, 2 * AT_TILE_BYTES,
                                 stage_tab + g * AT_SPG, full_bar, empty_bar, kload, 2 * AT_TILE_BYTES,
                                 policy_evict_first(), r_hint);
      at_stamp(dbg, g, 7);
    }
    pdl_wait();
    return;
  }
  // ============================== consumers: 4 independent groups ==============================
  pdl_wait();
  const int g = warp / AT_GW, tg = tid - g * AT_GT;
  a4_consume<AT_SPG>(tg, gsm[g], H, pos, ring, 2 * AT_TILE_BYTES, stage_tab + g * AT_SPG, full_bar, empty_bar, 1 + g,
                     early_trigger & 2, dbg);
  if (tg == 0) at_stamp(dbg, g, 5);
}

}  // namespace pg
