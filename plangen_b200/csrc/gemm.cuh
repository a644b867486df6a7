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
//                      tcgen05.ld epilogue.  Warp-specialised, six warps: warp 0 weight producer (even
//                      k-blocks), warp 1 MMA issuer, warps 2-5 epilogue - whose lane 0 first serves as second
//                      weight producer (warp 2), token-tile producers (warps 3, 4) and second MMA issuer (warp 5).
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

// 16 accumulator columns of this warp's 32 TMEM lanes, summed over the NACC column blocks (block 0 first)
template <int NACC, int NT>
PG_DEVINL void tmem_ld_acc_sum(uint32_t taddr, uint32_t (&v)[16], int n_used) {
  if constexpr (NACC == 1) {
    tmem_ld_32x32b_x16(taddr, v);
    tmem_ld_wait();
  } else {
    // all column blocks are requested before the single wait (a tcgen05.ld round trip is ~100+ cycles)
    uint32_t w[NACC - 1][16];
    tmem_ld_32x32b_x16(taddr, v);
#pragma unroll
    for (int a = 1; a < NACC; ++a)
      if (a < n_used) tmem_ld_32x32b_x16(taddr + (uint32_t)(a * NT), w[a - 1]);
    tmem_ld_wait();
#pragma unroll
    for (int a = 1; a < NACC; ++a) {
      if (a >= n_used) break;
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[a - 1][j]));
    }
  }
}

template <int NT>
struct TcCfg {
  static constexpr int B_BYTES = NT * TC_BK * 2;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  // Successive tcgen05.mma into ONE accumulator form a dependent chain: with N <= 64 an MMA occupies the
  // tensor pipe for 8-32 cycles but each link of the chain costs the full pipeline latency (~180 cycles,
  // measured: 0.38 us per 4-MMA k-block), which paced every weight-streaming contraction of the decode step.
  // The K-steps of a k-block therefore go round-robin into NACC independent accumulators (TMEM column
  // blocks) that the epilogue adds up in a fixed order.
  static constexpr int NACC = NT <= 64 ? 4 : NT <= 128 ? 2 : 1;
  static constexpr int ACC_COLS = NACC * NT;
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  static constexpr int smem_bytes(int stages) { return stages * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/; }
};

// Tile-major weight copy for the streaming contractions: tile (nt, kb) = rows nt*128.., k kb*64.. as one
// contiguous 16 KB block already in the SWIZZLE_128B shared-memory layout (16-byte chunk c of row r sits at
// r*128 + ((c ^ (r & 7)) * 16)), zero-padded at the edges.  A row-major weight makes every TMA tile 128
// separate 128-byte pieces at a stride of K*2 bytes (a different DRAM page each); measured on the gate|up
// contraction that holds the stream at ~2.6 TB/s.
__global__ void __launch_bounds__(256)
tile_weight_kernel(const bf16* __restrict__ W, uint8_t* __restrict__ out, int N, int K, int num_kb, size_t n_chunks) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_chunks; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx & 7), r = (int)((idx >> 3) & 127);
    const size_t tile = idx >> 10;
    const int kb = (int)(tile % num_kb);
    const size_t nt = tile / num_kb;
    const size_t n = nt * TC_BM + r;
    const int k = kb * TC_BK + c * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (n < (size_t)N && k + 8 <= K) v = *reinterpret_cast<const uint4*>(W + n * K + k);
    *reinterpret_cast<uint4*>(out + tile * TC_A_BYTES + r * 128 + ((c ^ (r & 7)) * 16)) = v;
  }
}

// debug timeline (tools/gemm_timeline.py): 8 %globaltimer stamps per CTA for launches whose weight-row count
// equals g_gemm_dbg_n; nullptr in production
__device__ unsigned long long* g_gemm_dbg = nullptr;
__device__ int g_gemm_dbg_n = 0;
PG_DEVINL void gemm_stamp(unsigned long long* dbg, int k) {
  if (dbg) dbg[((size_t)(blockIdx.z * gridDim.x + blockIdx.x)) * 8 + k] = global_timer_ns();
}

// Implicit-GEMM 3x3 convolution (pad 1) over a channels-last activation [B][H][W][Cin]: the "token" tile of
// CTA y is a bw x bh block of output pixels of one image, and the operand tile of k-block kb = (tap, cin0)
// is that pixel block shifted by the tap offset - one 4-D TMA box {64 channels, bw, bh, 1} whose
// out-of-bounds rows/columns are zero-filled by the hardware (the conv padding).  No im2col matrix exists.
struct ConvGeom {
  int enabled, H, W, Cin, bw, bh, tiles_x, tiles_y;
};

// Fused row epilogue (no split-K) for the wide-tile contractions of the vision tower / aligner (and any other
// Linear with a bias): instead of fp32 partials for a row kernel to re-read,
//   out  : out[m][n]   = bf16( act( bf16(acc + bias[n]) ) )         act = identity | exact-erf GELU
//          the 128 x NT tile is staged in the (idle) operand ring as [m][128 n] bf16 and written with 16-byte stores,
//          256 contiguous bytes per token row;
//          with add16: out[m][n] = bf16( bf16(acc + bias[n]) + add16[m][n] )  (VQ ResnetBlock / AttnBlock skip connection,
//          vq_model.py:352,390; add16 may alias out: a thread reads its 16 bytes before it writes them).  In the
//          implicit-GEMM convolution the tile's pixel block is a contiguous run of rows of the NHWC output (bw == W or
//          bh == 1), so the same staged store serves it;
//   resid: resid[m][n] += bf16(acc + bias[n])                        fp32 residual stream updated in place
//          (128 contiguous bytes per warp access, the pattern of the partial store).
// Rounding points are those of bias_act_kernel / vit_resid_ln_kernel / conv_epilogue_kernel (autocast: Linear and conv
// outputs in bf16).
// erf for the fused GELU epilogue: Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7) with ex2.approx / rcp.approx, a dozen
// instructions instead of erff's ~40 - the epilogue of the 4096-wide fc1 contraction of the vision tower was bound by
// erff (tensor pipe 26 %).  The value is rounded to bf16 (8 mantissa bits) right after; the row kernels and the fp32 check
// mode keep erff.  A deliberate approximation, like the SwiGLU epilogue's silu.
PG_DEVINL float erf_fast(float x) {
  const float ax = fabsf(x);
  float t, ex;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-ax * ax * 1.4426950408889634f));
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  const float r = fmaf(-poly, ex, 1.0f);
  return copysignf(r, x);
}

struct EpiFuse {
  const float* bias;   // [N] or nullptr
  bf16* out;           // [M][N] row-major, or nullptr
  float* resid;        // [M][N] fp32, or nullptr
  int gelu;
  const bf16* add16;   // [M][N] bf16 added after the first rounding (out mode), or nullptr
};

// Prefill QKV projection finished inside the CTA-pair contraction (gemm_tc2.cuh, EpiFuse::gelu == 3): a staged tile is one
// (q | k | v, head) x 256 tokens, so the row-store pass rotates q / k (RoPE pairs d, d + 64 are in the same staged row) and
// scatters q to [tok][H*128], k / v to the cache strips - what qkv_rope_store_bf16_kernel did from a bf16 round trip.
struct QkvEpi {
  const float* cosT; const float* sinT;      // [position][64] fp32
  bf16* q_out; bf16* kcache; bf16* vcache;   // q [tok][H*128]; caches [R][H][Tmax][128] of this layer
  const int32_t* row_off;                    // packed stream offsets (nullptr: padded [R][P] block)
  const int32_t* kv_start;
  const int32_t* rope_start;                 // RoPE position = column - rope_start[row] (text prefill), nullptr: the column
  int R, P, H, Tmax;
};

template <int NT>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
               float* __restrict__ C, int M, int N, int K, int kb_per_split, int num_stages, int use_pdl,
               const uint8_t* __restrict__ w_tiled, Prof prof, bf16* __restrict__ swiglu_out, ConvGeom cg, EpiFuse ep) {
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
  int cv_b = 0, cv_y0 = 0, cv_x0 = 0;
  if (cg.enabled) {
    const int per_img = cg.tiles_x * cg.tiles_y;
    cv_b = blockIdx.y / per_img;
    const int r = blockIdx.y % per_img;
    cv_y0 = (r / cg.tiles_x) * cg.bh;
    cv_x0 = (r % cg.tiles_x) * cg.bw;
  }
  const int split = blockIdx.z;
  const int num_kb = (K + TC_BK - 1) / TC_BK;
  const int kb_begin = split * kb_per_split;
  const int kb_end = min(kb_begin + kb_per_split, num_kb);
  const int nkb = max(kb_end - kb_begin, 0);

  if (use_pdl & 1) pdl_launch_dependents();
  prof_begin(prof);
  unsigned long long* dbg = (g_gemm_dbg_n == N && blockIdx.y == 0) ? g_gemm_dbg : nullptr;
  if (threadIdx.x == 0) gemm_stamp(dbg, 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_x);
    for (int i = 0; i < num_stages; ++i) { mbar_init(&full_bar[i], 2); mbar_init(&empty_bar[i], 1); }   // full: W + X producer
    mbar_init(tmem_full_bar, Cfg::NACC >= 2 ? 2 : 1);
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

  // ===================== TMA producers (three threads) =====================
  // One thread issuing cp.async.bulk / TMA tensor loads sustains only ~3.6 operations/us (measured,
  // tools/micro/stream_bench.cu: 16 KB copies from one thread top out at 59 GB/s per CTA whatever the ring
  // depth), which bounded every weight-streaming contraction at ~30 GB/s per CTA with one producer doing
  // W + X per k-block.  The W tiles are therefore split over two issuing threads (even / odd k-blocks: warp 0
  // and epilogue warp 2, idle until the accumulator is complete) and the X tiles over two (warps 3, 4).
  // Each stage's full barrier takes two arrive.expect_tx (W bytes, X bytes).
  // INVARIANT: every role that is split over two threads strides the k-blocks by 2 and the ring depth is EVEN
  // whenever stages are reused (launch_tc), so a stage always belongs to the same thread of each role and
  // that thread observes every phase of the stage's barriers - a thread that skipped a phase would see the
  // parity of an older phase as "complete" (mbarrier parity aliasing) and run ahead of the data.
  // use_pdl bit 1: the W operand is a constant weight, so its tiles may be requested before the previous
  // kernel has finished (it is an activation in the VQ attention contractions).
  constexpr bool DUAL = Cfg::NACC >= 2;        // two MMA issuers (below)
  auto produce_w = [&](int first) {
    const uint64_t pol_w = policy_evict_first();     // weights are streamed once per step
    if ((use_pdl & 3) == 1) pdl_wait();
    for (int i = first; i < nkb; i += 2) {
      const int s = i % num_stages;
      mbar_wait(&empty_bar[s], (((uint32_t)(i / num_stages)) & 1u) ^ 1u, 1);
      if (dbg && blockIdx.x == 5 && blockIdx.z == 0 && i < 64) dbg[16384 + 64 + i] = global_timer_ns();
      mbar_expect_tx(&full_bar[s], TC_A_BYTES);
      if (w_tiled)   // weight packed tile-major and pre-swizzled: one contiguous 16 KB bulk copy per tile
        bulk_copy_g2s(smem + s * Cfg::STAGE_BYTES, w_tiled + ((size_t)blockIdx.x * num_kb + kb_begin + i) * TC_A_BYTES, TC_A_BYTES,
                      &full_bar[s], pol_w);
      else
        tma_load_2d(smem + s * Cfg::STAGE_BYTES, &map_w, &full_bar[s], (kb_begin + i) * TC_BK, n0, pol_w);
    }
  };
  auto produce_x = [&](int first) {
    const uint64_t pol_x = policy_evict_last();      // activations are re-read by every CTA
    if (use_pdl & 1) pdl_wait();
    if (first == 0) gemm_stamp(dbg, 1);
    for (int i = first; i < nkb; i += 2) {
      const int s = i % num_stages;
      mbar_wait(&empty_bar[s], (((uint32_t)(i / num_stages)) & 1u) ^ 1u, 1);
      if (dbg && blockIdx.x == 5 && blockIdx.z == 0 && i < 64) dbg[16384 + 128 + i] = global_timer_ns();
      mbar_expect_tx(&full_bar[s], Cfg::B_BYTES);
      if (cg.enabled) {
        const int k0 = (kb_begin + i) * TC_BK, tap = k0 / cg.Cin;
        tma_load_4d(smem + s * Cfg::STAGE_BYTES + TC_A_BYTES, &map_x, &full_bar[s], k0 - tap * cg.Cin, cv_x0 + tap % 3 - 1,
                    cv_y0 + tap / 3 - 1, cv_b, pol_x);
      } else {
        tma_load_2d(smem + s * Cfg::STAGE_BYTES + TC_A_BYTES, &map_x, &full_bar[s], (kb_begin + i) * TC_BK, m0, pol_x);
      }
    }
  };
  // ===================== MMA issuers (one or two threads) =====================
  // Measured (clock64 around the issue loop): one thread spends 256 cycles issuing the four K=16 MMAs of a
  // k-block (the tensor pipe reads the 4 KB A slice of each from shared memory at 64 B/clk - that is the
  // floor of streaming weights as the A operand), then 115 cycles in tcgen05.commit and ~100 in the next
  // full-barrier wait during which the pipe drains: ~50% utilisation, 0.25-0.38 us per k-block.  Two
  // issuers on alternate k-blocks, each with its own accumulator blocks, keep the pipe fed.
  auto issue_mma = [&](int first, int stride, int acc0, int nacc) {
    const uint32_t idesc = umma_idesc_bf16(TC_BM, NT);
    for (int i = first; i < nkb; i += stride) {
      const int s = i % num_stages;
      mbar_wait(&full_bar[s], ((uint32_t)(i / num_stages)) & 1u, 2);
      if (i == 0) gemm_stamp(dbg, 2);
      if (i == nkb - 1) gemm_stamp(dbg, 3);
      if (dbg && blockIdx.x == 5 && blockIdx.z == 0 && i < 64) dbg[16384 + i] = global_timer_ns();
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
      const uint64_t da = umma_desc_k_sw128(a_addr);
      const uint64_t db = umma_desc_k_sw128(a_addr + TC_A_BYTES);
#pragma unroll
      for (int k = 0; k < TC_BK / 16; ++k)     // UMMA_K = 16 bf16 = 32 bytes = +2 in the encoded address
        umma_bf16(tmem_base + (uint32_t)((acc0 + k % nacc) * NT), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                  (uint32_t)((i != first) || (k / nacc != 0)));
      umma_commit(&empty_bar[s]);              // frees the smem stage once these MMAs retire
    }
    umma_commit(tmem_full_bar);                // this issuer's accumulators complete -> epilogue
  };
  if (warp == 0) {
    if (lane == 0) produce_w(0);
  } else if (warp == 1) {
    if (lane == 0) issue_mma(0, DUAL ? 2 : 1, 0, DUAL ? Cfg::NACC / 2 : Cfg::NACC);
  } else {
    // ===================== epilogue: TMEM -> registers -> global (4 warps) =====================
    const int quarter = warp & 3;              // a warp may only touch TMEM lanes 32*(warp%4)..+31
    const int n = n0 + quarter * 32 + lane;
    float* out = C + (size_t)split * M * N;
    const int acc_used = (DUAL && nkb < 2) ? Cfg::NACC / 2 : Cfg::NACC;   // the second issuer's blocks stay unwritten
    if (lane == 0) {                            // producer / second issuer duty first (see above), then the epilogue
      if (warp == 2) produce_w(1);
      else if (warp == 5) { if (DUAL) issue_mma(1, 2, Cfg::NACC / 2, Cfg::NACC / 2); }
      else produce_x(warp - 3);
    }
    __syncwarp();
    if (use_pdl & 1) pdl_wait();
    if (swiglu_out != nullptr) {
      // Fused SwiGLU epilogue (no split-K): the weight rows are interleaved in blocks of 64, so lanes 0-63 of
      // the tile hold gate(f0 .. f0+63) and lanes 64-127 hold up(f0 .. f0+63).  Gate and up warps swap half of
      // their accumulators through the (now idle) first pipeline stage and combine
      //   h = rnd(rnd(silu(rnd(g))) * rnd(u))        (HF LlamaMLP :182-184, bf16 rounding points of autocast)
      // storing h[m][f] directly as the bf16 operand of the down projection.
      mbar_wait(tmem_full_bar, 0, 3);
      if (warp == 2 && lane == 0) gemm_stamp(dbg, 4);
      tc_fence_after();
      // Each gate warp (quarters 0, 1) pairs with the up warp two quarters above it.  Per 16-column chunk the
      // gate warp hands g of columns 8..15 to the up warp and receives u of columns 0..7, so all four warps
      // compute (8 tokens each) and the 8 results of a thread are independent chains the scheduler can overlap.
      // exchange buffers in the (now idle) first pipeline stage, addressed in the shared window: two of
      // [128 lanes][9] fp32, alternating per 16-column chunk so one named barrier per chunk is enough (a thread can
      // only write buffer b again after the barrier of the chunk in between, which every reader of b has passed)
      const uint32_t xch_s = smem_u32(smem);
      const int F = N / 2;
      const int pair = quarter & 1;                           // f-block within the tile
      const int f = blockIdx.x * 64 + pair * 32 + lane;
      const bool is_gate = quarter < 2;
      const int jbase = is_gate ? 0 : 8;                      // columns this warp finishes
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 16) {
        uint32_t v[16];
        tmem_ld_acc_sum<Cfg::NACC, NT>(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v, acc_used);
        if (c0 == 0 && warp == 2 && lane == 0) gemm_stamp(dbg, 7);
        const uint32_t buf = xch_s + (uint32_t)((c0 >> 4) & 1) * (128u * 9u * 4u);
        const uint32_t mine = buf + (uint32_t)(quarter * 32 + lane) * 36u;           // what I give away
        const uint32_t theirs = buf + (uint32_t)((quarter ^ 2) * 32 + lane) * 36u;   // what my partner gives me
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(mine + 4u * j), "r"(v[(is_gate ? 8 : 0) + j]) : "memory");
        asm volatile("bar.sync 2, 128;" ::: "memory");
        float h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t ow;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(ow) : "r"(theirs + 4u * j) : "memory");
          const float own = __uint_as_float(v[jbase + j]), other = __uint_as_float(ow);
          const float g = bf16_round(is_gate ? own : other);
          const float u = bf16_round(is_gate ? other : own);
          // silu(g) = g / (1 + 2^(-g log2 e)) with ex2.approx / rcp.approx (flush-to-zero forms: no denormal
          // fix-up code): ~1e-6 relative error against a result kept to 8 mantissa bits; the exact expf + IEEE
          // division cost ~100 dependent instructions per element on warps that have their scheduler to themselves
          float ex, rc;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-g * 1.4426950408889634f));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(1.0f + ex));
          const float sg = bf16_round(g * rc);
          h[j] = sg * u;
        }
        if (f < F) {
          bf16* dst = swiglu_out + (size_t)(m0 + c0 + jbase) * F + f;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (m0 + c0 + jbase + j < M) dst[(size_t)j * F] = __float2bfloat16_rn(h[j]);
        }
      }
    } else if (nkb > 0) {
      mbar_wait(tmem_full_bar, 0, 3);
      if (warp == 2 && lane == 0) gemm_stamp(dbg, 4);
      tc_fence_after();
      if (ep.out != nullptr) {
        // ---- fused bias (+ GELU) -> bf16 row-major, staged through the idle operand ring
        const float bias_n = (ep.bias && n < N) ? ep.bias[n] : 0.f;
        const uint32_t stg = smem_u32(smem);                   // [NT][128] bf16 = NT * 256 bytes
        const int nl = quarter * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 16) {
          uint32_t v[16];
          tmem_ld_acc_sum<Cfg::NACC, NT>(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v, acc_used);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float t = bf16_round(__uint_as_float(v[j]) + bias_n);
            if (ep.gelu) t = t * 0.5f * (1.0f + erf_fast(t * 0.70710678118654752440f));
            const unsigned short hb = __bfloat16_as_ushort(__float2bfloat16_rn(t));
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(stg + (uint32_t)((c0 + j) * 256 + nl * 2)), "h"(hb) : "memory");
          }
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        const int t128 = threadIdx.x - 64;                      // 0..127 over the epilogue warps 2..5
        const int ch = t128 & 15, r0 = t128 >> 4;              // 16-byte chunk of the 256-byte row, first row
        // rows of the output this tile covers: token rows m0.., or (convolution) the pixel block's run of NHWC rows, which
        // must not run past the end of its image (a partial block at the bottom edge)
        const size_t row0 = cg.enabled ? ((size_t)cv_b * cg.H + cv_y0) * cg.W + cv_x0 : (size_t)m0;
        const size_t row_end = cg.enabled ? (size_t)(cv_b + 1) * cg.H * cg.W : (size_t)M;
        if (n0 + ch * 8 < N) {
#pragma unroll 4
          for (int r = r0; r < NT; r += 8) {
            const size_t m = row0 + r;
            if (m < row_end) {
              uint4 q;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                           : "r"(stg + (uint32_t)(r * 256 + ch * 16)));
              bf16* dst = ep.out + m * N + n0 + ch * 8;
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
              *reinterpret_cast<uint4*>(dst) = q;
            }
          }
        }
      } else if (ep.resid != nullptr) {
        // ---- fused bias + residual add into the fp32 stream
        const float bias_n = (ep.bias && n < N) ? ep.bias[n] : 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 16) {
          uint32_t v[16];
          tmem_ld_acc_sum<Cfg::NACC, NT>(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v, acc_used);
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
      } else if (!cg.enabled) {
        // plain split-K partial store (the decode step's hot epilogue: keep it branch-free)
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 16) {
          uint32_t v[16];
          tmem_ld_acc_sum<Cfg::NACC, NT>(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v, acc_used);
          if (n < N) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int m = m0 + c0 + j;
              if (m < M) out[(size_t)m * N + n] = __uint_as_float(v[j]);   // 32 lanes -> 128 B contiguous
            }
          }
        }
      } else {
        // convolution: pixel (c0 + j) of the bw x bh block -> flat NHWC pixel index
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 16) {
          uint32_t v[16];
          tmem_ld_acc_sum<Cfg::NACC, NT>(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v, acc_used);
          if (n < N) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int yy = (c0 + j) / cg.bw, y = cv_y0 + yy;
              if (y < cg.H) {
                const size_t o = (size_t)((cv_b * cg.H + y) * cg.W + cv_x0 + (c0 + j) - yy * cg.bw) * N + n;
                out[o] = __uint_as_float(v[j]);
              }
            }
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
  if (warp == 2 && lane == 0) gemm_stamp(dbg, 5);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) gemm_stamp(dbg, 6);
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
