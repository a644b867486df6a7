// Stage-1 layout-text decode: System.x2t (plangen_base.py:513-523) =
// language_model.generate(inputs_embeds=, attention_mask=, pad_token_id=eos, eos_token_id=eos,
// max_new_tokens=512, do_sample=False, use_cache=True), i.e. HF GenerationMixin greedy search.
//
// One step of the loop after the lm_head contraction, in ONE kernel per step (one CTA per row):
//   next_token_logits = logits[:, -1, :].float()          (bf16-rounded logits under autocast)
//   next_tokens = argmax(next_token_logits, -1)           (first index wins ties, torch.argmax)
//   next_tokens = next_tokens * unfinished + pad_token_id * (1 - unfinished)
//   unfinished &= next_tokens != eos_token_id             (EosTokenCriteria; MaxLengthCriteria is the loop bound)
//   this_peer_finished = unfinished.max() == 0            -> *n_gen = step + 1 the first time it holds
//   next inputs_embeds = embed_tokens(next_tokens)        (fp32 table: nn.Embedding is not autocast)
//   + the first RMSNorm of the next decode step (input_layernorm of layer 0)
// No logits round trip and no per-step host sync (HF syncs every step for `unfinished.max() == 0`; the host
// here polls *n_unfinished every few steps).
#pragma once
#include "lm_kernels.cuh"

namespace pg {

struct GreedyState {
  int* unfinished;      // [R] 1 while the row has not produced eos
  int* n_unfinished;    // rows still unfinished
  int* n_gen;           // number of steps after which every row had finished (initialised to max_new_tokens)
};

constexpr int TXT_THREADS = 1024;
constexpr int TXT_MAX_PER_THREAD = 8;     // D <= 8 * TXT_THREADS

template <typename T>
__global__ void __launch_bounds__(TXT_THREADS)
lm_argmax_embed_kernel(const float* __restrict__ part, int S, size_t split_stride, int V, int eos_id, int pad_id,
                       GreedyState gs, int step_base, const int* __restrict__ step_ptr, int max_new,
                       int32_t* __restrict__ tokens_out, const float* __restrict__ embed_table, int vocab_rows, int D,
                       float* __restrict__ x_next, const float* __restrict__ next_norm_w, T* __restrict__ xn_next,
                       float eps, float* __restrict__ dbg_logits) {
  __shared__ float red[32];
  __shared__ float bestv[32];
  __shared__ int besti[32];
  __shared__ int tok_s;
  pdl_launch_dependents();
  // the norm scale of the next step's first layer is a constant: fetch before the dependency wait
  float wv[TXT_MAX_PER_THREAD];
#pragma unroll
  for (int k = 0; k < TXT_MAX_PER_THREAD; ++k) {
    const int d = threadIdx.x + k * blockDim.x;
    wv[k] = (next_norm_w != nullptr && d < D) ? next_norm_w[d] : 0.f;
  }
  pdl_wait();
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int step = step_base + (step_ptr ? *step_ptr : 0);
  const float* row = part + (size_t)r * V;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  // 8 elements per thread in flight (independent L2 round trips), index-ascending per thread
  for (int v0 = tid; v0 < V; v0 += 8 * TXT_THREADS) {
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int v = v0 + j * TXT_THREADS;
      x[j] = 0.f;
      if (v < V) x[j] = (S == 1) ? __ldcg(row + v) : reduce_splits(row, S, split_stride, (size_t)v);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int v = v0 + j * TXT_THREADS;
      if (v < V) {
        const float t = Act<T>::rnd(x[j]);          // the Linear's output dtype under autocast, then `.float()`
        if (dbg_logits) dbg_logits[((size_t)step * gridDim.x + r) * V + v] = t;
        if (t > bv || (t == bv && v < bi)) { bv = t; bi = v; }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) { bestv[warp] = bv; besti[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    bv = lane < nw ? bestv[lane] : -INFINITY;
    bi = lane < nw ? besti[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      const int was = gs.unfinished[r];
      int tok = was ? bi : pad_id;
      if (tok < 0 || tok >= V) tok = 0;              // all-NaN row: keep the gather in bounds
      tokens_out[(size_t)r * max_new + step] = tok;
      if (was && tok == eos_id) {
        gs.unfinished[r] = 0;
        if (atomicSub(gs.n_unfinished, 1) == 1) *gs.n_gen = step + 1;
      }
      tok_s = tok;
    }
  }
  __syncthreads();
  if (x_next == nullptr) return;                     // last step of the loop: nothing follows
  const int tok = min(tok_s, vocab_rows - 1);
  const float* src = embed_table + (size_t)tok * D;
  float v[TXT_MAX_PER_THREAD];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < TXT_MAX_PER_THREAD; ++k) {
    const int d = tid + k * blockDim.x;
    v[k] = 0.f;
    if (d < D) {
      v[k] = src[d];
      x_next[(size_t)r * D + d] = v[k];
      ss += v[k] * v[k];
    }
  }
  if (xn_next == nullptr) return;
  ss = block_sum(ss, red);
  const float rs = rsqrtf(ss / (float)D + eps);
#pragma unroll
  for (int k = 0; k < TXT_MAX_PER_THREAD; ++k) {
    const int d = tid + k * blockDim.x;
    if (d < D) Act<T>::st(xn_next + (size_t)r * D + d, wv[k] * (v[k] * rs));   // fp32 residual stream: no rounding of hn
  }
}

__global__ void greedy_state_init_kernel(GreedyState gs, int R, int max_new) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R) gs.unfinished[i] = 1;
  if (i == 0) { *gs.n_unfinished = R; *gs.n_gen = max_new; }
}

}  // namespace pg
