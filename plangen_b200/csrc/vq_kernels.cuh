// VQ-16 decode side (three_party/Janus/janus/models/vq_model.py:505-508 -> :284-299, :500-503,
// :193-214): codebook gather with L2 normalisation, post_quant_conv, and the conv decoder.
// Activations are kept channels-last (NHWC, [B*H*W, C]) so every convolution is the TN contraction
// of gemm.cuh: C[pixel][cout] = sum_k col[pixel][k] * Wc[cout][k], k = (ky, kx, cin).  The im2col
// producer fuses GroupNorm(32, eps 1e-6) + swish (x * sigmoid(x)) and the nearest x2 upsample of
// `Upsample.forward` (:417-427); zero padding is applied after the activation, as conv2d does.
#pragma once
#include "common.cuh"

namespace pg {

// ---------------------------------------------------------------- codebook + post_quant_conv
// z_q = F.normalize(codebook)[code] (vq_model.py:286-290, every call), then 1x1 conv 8 -> Z.
template <typename T>
__global__ void vq_codebook_pqc_kernel(const int32_t* __restrict__ codes, const float* __restrict__ codebook,
                                       const T* __restrict__ w, const float* __restrict__ bias, T* __restrict__ out,
                                       int code_dim, int Z, int V, size_t n_pix) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t pix = blockIdx.x;
  if (pix >= n_pix) return;
  __shared__ float e[32];
  const int code = min(max(codes[pix], 0), V - 1);
  if (threadIdx.x == 0) {
    float ss = 0.f;
    for (int k = 0; k < code_dim; ++k) { const float v = codebook[(size_t)code * code_dim + k]; ss += v * v; }
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    for (int k = 0; k < code_dim; ++k) e[k] = Act<T>::rnd(codebook[(size_t)code * code_dim + k] / denom);
  }
  __syncthreads();
  for (int z = threadIdx.x; z < Z; z += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < code_dim; ++k) acc = fmaf(e[k], Act<T>::ld(w + (size_t)z * code_dim + k), acc);
    Act<T>::st(out + pix * Z + z, acc + bias[z]);
  }
}

// ------------------------------------------------------------------------- GroupNorm statistics
// stage 1: per (image, pixel chunk) partial sum / sum of squares for all 32 groups.
template <typename T>
__global__ void __launch_bounds__(512)
gn_partial_kernel(const T* __restrict__ x, float* __restrict__ partial, int HW, int C, int chunk_pix) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float t_sum[512], t_sq[512];
  const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  const int cpg = C / 32;
  const int p0 = chunk * chunk_pix, p1 = min(HW, p0 + chunk_pix);
  // thread -> a fixed channel residue (tid % min(C,256)); pixels strided; every thread's (sum, sumsq) goes to
  // shared memory and the 32 group totals are formed in a FIXED order (deterministic, batch-independent)
  const int nt = (int)blockDim.x;
  const int cols = min(C, nt);                     // distinct channels handled per pass
  const int cidx = threadIdx.x % cols, prow = threadIdx.x / cols, pstride = max(1, nt / cols);
  float a = 0.f, q = 0.f;
  int my_group = -1;
  if (C <= nt) {
    my_group = cidx / cpg;
    if (prow < pstride)
      for (int p = p0 + prow; p < p1; p += pstride) {
        const float v = Act<T>::ld(x + ((size_t)b * HW + p) * C + cidx);
        a += v; q += v * v;
      }
  }
  t_sum[threadIdx.x] = a; t_sq[threadIdx.x] = q;
  __syncthreads();
  if (C <= nt) {
    if (threadIdx.x < 32) {
      const int g = threadIdx.x;
      float sa = 0.f, sq = 0.f;
      for (int t = 0; t < nt; ++t) {
        if ((t % cols) / cpg == g && t / cols < pstride) { sa += t_sum[t]; sq += t_sq[t]; }
      }
      float* o = partial + (((size_t)b * nchunks + chunk) * 32 + g) * 2;
      o[0] = sa; o[1] = sq;
    }
  }
  (void)my_group;
}
// stage 2: combine chunks in double -> (mean, rstd) per (image, group)
__global__ void gn_finalize_kernel(const float* __restrict__ partial, float* __restrict__ stats, int nchunks,
                                   double inv_count, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, g = threadIdx.x;
  if (g >= 32) return;
  double a = 0.0, q = 0.0;
  for (int c = 0; c < nchunks; ++c) {
    const float* o = partial + (((size_t)b * nchunks + c) * 32 + g) * 2;
    a += (double)o[0]; q += (double)o[1];
  }
  const double mean = a * inv_count;
  double var = q * inv_count - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[((size_t)b * 32 + g) * 2 + 0] = (float)mean;
  stats[((size_t)b * 32 + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// ----------------------------------------------------- im2col (+ GroupNorm + swish + upsample x2)
// in:  [B][Hi][Wi][C] (T), out col: [B*Ho*Wo][ks*ks*C] (T), Ho = Hi * up, pad = ks / 2.
// stats == nullptr -> no normalisation; swish only applies together with stats.
// One warp per output pixel; a lane owns 4 consecutive channels of every 128-channel slab, so each tap is
// one coalesced 256-byte read and one coalesced 256-byte write per warp and the GroupNorm scale/shift of
// the lane's channels are loaded once per pixel.
template <typename T> struct Ch4;
template <> struct Ch4<float> {
  static PG_DEVINL void ld(const float* p, float (&v)[4]) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  static PG_DEVINL void st(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Ch4<bf16> {
  static PG_DEVINL void ld(const bf16* p, float (&v)[4]) { const uint2 t = *reinterpret_cast<const uint2*>(p); v[0] = bf16lo(t.x); v[1] = bf16hi(t.x); v[2] = bf16lo(t.y); v[3] = bf16hi(t.y); }
  static PG_DEVINL void st(bf16* p, const float (&v)[4]) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t; t.x = *reinterpret_cast<const uint32_t*>(&a); t.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

// bf16 fast path of stage 1 (C % 128 == 0: a thread owns 4 consecutive channels of one group, 8-byte
// loads, four pixels in flight): same output layout and a fixed summation order, so results do not depend on
// the batch composition.
__global__ void __launch_bounds__(256)
gn_partial_vec_kernel(const bf16* __restrict__ x, float* __restrict__ partial, int HW, int C, int chunk_pix) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float t_sum[256], t_sq[256];
  const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  const int cvn = C / 4;                               // channel vectors per pixel (32 .. 128)
  const int cv = threadIdx.x % cvn, pl = threadIdx.x / cvn, PL = 256 / cvn;
  const int p0 = chunk * chunk_pix, p1 = min(HW, p0 + chunk_pix);
  const bf16* base = x + (size_t)b * HW * C + cv * 4;
  float a = 0.f, q = 0.f;
  int p = p0 + pl;
  for (; p + 3 * PL < p1; p += 4 * PL) {
    float v[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) Ch4<bf16>::ld(base + (size_t)(p + u * PL) * C, v[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) { a += v[u][j]; q += v[u][j] * v[u][j]; }
  }
  for (; p < p1; p += PL) {
    float v[4];
    Ch4<bf16>::ld(base + (size_t)p * C, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) { a += v[j]; q += v[j] * v[j]; }
  }
  t_sum[threadIdx.x] = a; t_sq[threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int g = threadIdx.x, vpg = cvn / 32;         // channel vectors per group
    float sa = 0.f, sq = 0.f;
    for (int l = 0; l < PL; ++l)
      for (int j = 0; j < vpg; ++j) { sa += t_sum[l * cvn + g * vpg + j]; sq += t_sq[l * cvn + g * vpg + j]; }
    float* o = partial + (((size_t)b * nchunks + chunk) * 32 + g) * 2;
    o[0] = sa; o[1] = sq;
  }
}

// GroupNorm apply (+ swish) (+ nearest x2 upsample), channels-last bf16 -> bf16, 16 bytes per thread.  The
// activation a 3x3 implicit-GEMM convolution reads (engine.cu run_conv_gemm); same arithmetic as im2col_kernel.
__global__ void __launch_bounds__(256)
gn_apply_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, const float* __restrict__ stats,
                const float* __restrict__ gamma, const float* __restrict__ beta, int Hi, int Wi, int C, int up, int swish) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int cvn = C / 8, cv = threadIdx.x % cvn, pl = threadIdx.x / cvn, PL = 256 / cvn;
  const int Ho = Hi * up, Wo = Wi * up, cpg = C / 32;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j;
    sc[j] = 1.f; sh[j] = 0.f;
    if (stats) {
      const float mean = stats[((size_t)b * 32 + c / cpg) * 2], rstd = stats[((size_t)b * 32 + c / cpg) * 2 + 1];
      sc[j] = rstd * gamma[c];
      sh[j] = beta[c] - mean * sc[j];
    }
  }
  const int n_pix = Ho * Wo;
  for (int p = blockIdx.x * PL + pl; p < n_pix; p += gridDim.x * PL) {
    const int oy = p / Wo, ox = p - oy * Wo;
    const uint4 t = *reinterpret_cast<const uint4*>(in + (((size_t)b * Hi + oy / up) * Wi + ox / up) * C + cv * 8);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float y0 = bf16lo(w[k]), y1 = bf16hi(w[k]);
      if (stats) {
        y0 = fmaf(y0, sc[2 * k], sh[2 * k]); y1 = fmaf(y1, sc[2 * k + 1], sh[2 * k + 1]);
        if (swish) { y0 = y0 / (1.0f + expf(-y0)); y1 = y1 / (1.0f + expf(-y1)); }
      }
      const __nv_bfloat162 o = __floats2bfloat162_rn(y0, y1);
      r[k] = *reinterpret_cast<const uint32_t*>(&o);
    }
    *reinterpret_cast<uint4*>(out + ((size_t)b * n_pix + p) * C + cv * 8) = make_uint4(r[0], r[1], r[2], r[3]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
im2col_kernel(const T* __restrict__ in, T* __restrict__ col, const float* __restrict__ stats,
              const float* __restrict__ gamma, const float* __restrict__ beta, int Hi, int Wi, int C, int ks, int up,
              int swish, size_t n_pix /* B*Ho*Wo */) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int Ho = Hi * up, Wo = Wi * up, pad = ks / 2, cpg = C / 32, taps = ks * ks;
  const size_t warp_id = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const size_t n_warps = (size_t)gridDim.x * (blockDim.x >> 5);
  for (size_t pix = warp_id; pix < n_pix; pix += n_warps) {
    const int ox = (int)(pix % Wo);
    const int oy = (int)((pix / Wo) % Ho);
    const int b = (int)(pix / ((size_t)Wo * Ho));
    T* dst_row = col + pix * (size_t)taps * C;
    for (int c0 = lane * 4; c0 < C; c0 += 128) {
      float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
      if (stats) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c0 + j;
          const float mean = stats[((size_t)b * 32 + c / cpg) * 2], rstd = stats[((size_t)b * 32 + c / cpg) * 2 + 1];
          sc[j] = rstd * gamma[c];                        // y = (x - mean) * rstd * gamma + beta  (fp32, as group_norm under autocast)
          sh[j] = beta[c] - mean * sc[j];
        }
      }
      for (int tap = 0; tap < taps; ++tap) {
        const int iy = oy + tap / ks - pad, ix = ox + tap % ks - pad;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (iy >= 0 && iy < Ho && ix >= 0 && ix < Wo) {
          Ch4<T>::ld(in + (((size_t)b * Hi + iy / up) * Wi + ix / up) * C + c0, v);
          if (stats) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // same operation order as the reference: ((x - mean) * rstd) * gamma + beta
              float y = fmaf(v[j], sc[j], sh[j]);
              if (swish) y = y / (1.0f + expf(-y));       // x * sigmoid(x)
              v[j] = y;
            }
          }
        }
        Ch4<T>::st(dst_row + (size_t)tap * C + c0, v);
      }
    }
  }
}

// ------------------------------------------------------------------ VQ encode side (editing path)
// image fp32 NCHW [B][3][H][W] -> NHWC [B][H][W][Cp] in the activation type, channels 3..Cp-1 zero (the padded
// conv_in weight has zero taps there), so conv_in runs on the same im2col + contraction path as every other conv
template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_pad_kernel(const float* __restrict__ img, T* __restrict__ out, int C, int Cp, int HW, size_t total /* B*HW*Cp */) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    const size_t pix = i / Cp, b = pix / HW, p = pix % HW;
    Act<T>::st(out + i, c < C ? img[(b * C + c) * HW + p] : 0.f);
  }
}

// Downsample (vq_model.py:440-445): F.pad(x, (0,1,0,1)) then 3x3 stride-2 conv without padding:
// col[b, oy, ox][tap * C + c] = in[b, 2 oy + tap / 3, 2 ox + tap % 3, c]  (zero past the bottom / right edge)
template <typename T>
__global__ void __launch_bounds__(256)
im2col_down_kernel(const T* __restrict__ in, T* __restrict__ col, int Hi, int Wi, int C, size_t n_pix /* B*Ho*Wo */) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int Ho = Hi / 2, Wo = Wi / 2;
  const size_t warp_id = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const size_t n_warps = (size_t)gridDim.x * (blockDim.x >> 5);
  for (size_t pix = warp_id; pix < n_pix; pix += n_warps) {
    const int ox = (int)(pix % Wo);
    const int oy = (int)((pix / Wo) % Ho);
    const int b = (int)(pix / ((size_t)Wo * Ho));
    T* dst_row = col + pix * (size_t)9 * C;
    for (int c0 = lane * 4; c0 < C; c0 += 128) {
      for (int tap = 0; tap < 9; ++tap) {
        const int iy = 2 * oy + tap / 3, ix = 2 * ox + tap % 3;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (iy < Hi && ix < Wi) Ch4<T>::ld(in + (((size_t)b * Hi + iy) * Wi + ix) * C + c0, v);
        Ch4<T>::st(dst_row + (size_t)tap * C + c0, v);
      }
    }
  }
}

// VectorQuantizer.forward, inference branch with l2_norm (vq_model.py:236-262).
// Pre-pass: en[v][:] = codebook[v] / max(||codebook[v]||, 1e-12), e2[v] = sum(en[v]^2)   (every call, as the reference)
__global__ void vq_codebook_norm_kernel(const float* __restrict__ codebook, float* __restrict__ en, float* __restrict__ e2,
                                        int V, int Cd) {
  pdl_launch_dependents();
  pdl_wait();
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  float ss = 0.f;
  for (int k = 0; k < Cd; ++k) ss += codebook[(size_t)v * Cd + k] * codebook[(size_t)v * Cd + k];
  const float den = fmaxf(sqrtf(ss), 1e-12f);
  float s2 = 0.f;
  for (int k = 0; k < Cd; ++k) {
    const float x = codebook[(size_t)v * Cd + k] / den;
    en[(size_t)v * Cd + k] = x;
    s2 += x * x;
  }
  e2[v] = s2;
}
// One CTA per position: zn = z / max(||z||, 1e-12) (fp32), d[v] = (sum(zn^2) + e2[v]) - 2 * (zn . en[v]) with the
// product evaluated as the reference's matmul is (bf16 operands / bf16 result under autocast, fp32 in check mode),
// index of the smallest distance, FIRST index on ties (torch.argmin).
constexpr int VQ_CD_MAX = 16;
template <typename T>
__global__ void __launch_bounds__(256)
vq_quantize_kernel(const T* __restrict__ z, const float* __restrict__ en, const float* __restrict__ e2,
                   int32_t* __restrict__ idx_out, int V, int Cd) {
  __shared__ float bd[8];
  __shared__ int bi[8];
  pdl_launch_dependents();
  pdl_wait();
  const size_t pix = blockIdx.x;
  float zn[VQ_CD_MAX];
  float ss = 0.f;
  for (int k = 0; k < Cd; ++k) { zn[k] = Act<T>::ld(z + pix * Cd + k); ss += zn[k] * zn[k]; }
  const float den = fmaxf(sqrtf(ss), 1e-12f);
  float z2 = 0.f;
  for (int k = 0; k < Cd; ++k) { zn[k] = zn[k] / den; z2 += zn[k] * zn[k]; }
  float zq[VQ_CD_MAX];
  for (int k = 0; k < Cd; ++k) zq[k] = Act<T>::rnd(zn[k]);           // matmul operand cast (autocast) / identity (fp32)
  float best = INFINITY;
  int besti = 0x7fffffff;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    float dot = 0.f;
    for (int k = 0; k < Cd; ++k) dot = fmaf(zq[k], Act<T>::rnd(en[(size_t)v * Cd + k]), dot);
    dot = Act<T>::rnd(dot);
    const float dist = (z2 + e2[v]) - 2.0f * dot;
    if (dist < best) { best = dist; besti = v; }                     // ascending v per thread: first index kept on ties
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (od < best || (od == best && oi < besti)) { best = od; besti = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { bd[warp] = best; bi[warp] = besti; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (bd[w] < best || (bd[w] == best && bi[w] < besti)) { best = bd[w]; besti = bi[w]; }
    idx_out[pix] = besti < V ? besti : 0;
  }
}

// -------------------------------------------------------- conv epilogue: + bias (+ residual)
// out[pix][co] = rnd(rnd(part + bias) + residual)   NHWC; or NCHW fp32 for the final conv_out.
template <typename T>
__global__ void __launch_bounds__(256)
conv_epilogue_kernel(const float* __restrict__ part, const float* __restrict__ bias, const T* __restrict__ residual,
                     T* __restrict__ out, float* __restrict__ out_nchw, int Cout, int HW, size_t total,
                     int bias_per_row) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i / Cout;
    const int co = (int)(i % Cout);
    float v = part[i];
    if (bias) v += bias_per_row ? bias[pix] : bias[co];
    v = Act<T>::rnd(v);
    if (residual) v = Act<T>::rnd(v + Act<T>::ld(residual + i));
    if (out) Act<T>::st(out + i, v);
    if (out_nchw) {
      const size_t b = pix / HW, p = pix % HW;
      out_nchw[(b * Cout + co) * HW + p] = v;
    }
  }
}

// --------------------------------------------- AttnBlock softmax: P = softmax(scores * C^-0.5, dim=-1)
template <typename T>
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ scores, T* __restrict__ P, int n, float scale) {
  __shared__ float red[32];
  pdl_launch_dependents();
  pdl_wait();
  const size_t row = blockIdx.x;
  const float* s = scores + row * n;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < n; j += blockDim.x) mx = fmaxf(mx, Act<T>::rnd(Act<T>::rnd(s[j]) * scale));
  mx = block_max(mx, red);
  float sum = 0.f;
  for (int j = threadIdx.x; j < n; j += blockDim.x) sum += expf(Act<T>::rnd(Act<T>::rnd(s[j]) * scale) - mx);
  sum = block_sum(sum, red);
  for (int j = threadIdx.x; j < n; j += blockDim.x)
    Act<T>::st(P + row * n + j, expf(Act<T>::rnd(Act<T>::rnd(s[j]) * scale) - mx) / sum);
}

}  // namespace pg
