"""ORACLE (test infrastructure, NOT product code).

numpy restatement of what `torch.multinomial(probs, 1, generator=cuda_gen)`
(plangen_base.py:591) executes on a CUDA device, so the sampled token ids can be
reproduced bit-for-bit on a CPU and compared with the fused CUDA sampler.

Third-party pieces restated (torch pinned ==2.5.1, requirements.txt:289; 2.11.0
installed; the relevant code is unchanged between them):

  * aten/src/ATen/native/Distributions.cpp `multinomial_out`, n_sample == 1 fast
    path:  q = empty_like(p).exponential_(1, gen);  result = argmax(p / q, -1)
  * torch/include/ATen/native/cuda/DistributionTemplates.h
      calc_execution_policy :50-62   (grid / philox offset, a function of the SM
                                      count and maxThreadsPerMultiProcessor)
      distribution_elementwise_grid_stride_kernel :67-89 (thread -> element map)
      exponential_kernel :562-573, uniform_and_transform :429-441
  * torch/include/ATen/core/TransformationHelper.h `exponential` :129-146 (CUDA
    branch: -log(u), with u >= 1-eps/2 mapped to eps/2)
  * curand_kernel.h / curand_philox4x32_x.h: Philox4x32-10, `curand_init(seed,
    subsequence, offset)`, `curand_uniform4` (u = x * 2^-32 + 2^-33).

Pinned on the GPU box by tests/test_sampler_gpu.py against the real
torch.multinomial on cuda:0 (there is no such check possible on a CPU-only box).
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Counters/keys: uint32 arrays (broadcastable)."""
    c0 = np.asarray(c0, dtype=np.uint32); c1 = np.asarray(c1, dtype=np.uint32)
    c2 = np.asarray(c2, dtype=np.uint32); c3 = np.asarray(c3, dtype=np.uint32)
    k0 = np.asarray(k0, dtype=np.uint32); k1 = np.asarray(k1, dtype=np.uint32)
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = c0.astype(np.uint64) * PHILOX_M0
            p1 = c2.astype(np.uint64) * PHILOX_M1
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32); lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32); lo1 = (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            if r < 9:
                k0 = k0 + PHILOX_W0
                k1 = k1 + PHILOX_W1
    return c0, c1, c2, c3


def execution_policy(numel: int, num_sms: int, max_threads_per_sm: int, unroll: int = 4):
    """calc_execution_policy: (counter_offset, grid, block)."""
    block = 256
    grid = (numel + block - 1) // block
    grid = min(num_sms * (max_threads_per_sm // block), grid)
    counter_offset = ((numel - 1) // (block * grid * unroll) + 1) * 4
    return counter_offset, grid, block


def uniform_for_elements(numel: int, seed: int, offset: int, num_sms: int,
                         max_threads_per_sm: int) -> np.ndarray:
    """The float32 uniform in (0,1] that element `li` of a `numel`-element tensor
    receives from distribution_elementwise_grid_stride_kernel (unroll 4)."""
    _, grid, block = execution_policy(numel, num_sms, max_threads_per_sm)
    stride = grid * block
    li = np.arange(numel, dtype=np.int64)
    idx = li % stride                 # thread id (Philox subsequence)
    slot = li // stride               # ii + 4 * loop iteration
    ii = (slot % 4).astype(np.int64)
    it = (slot // 4).astype(np.uint64)
    # curand_init(seed, subsequence=idx, offset): ctr = (offset/4 [+it], 0, idx_lo, idx_hi)
    ctr = np.uint64(offset // 4) + it
    assert offset % 4 == 0
    c0 = (ctr & _MASK32).astype(np.uint32)
    c1 = (ctr >> np.uint64(32)).astype(np.uint32)
    c2 = (idx.astype(np.uint64) & _MASK32).astype(np.uint32)
    c3 = (idx.astype(np.uint64) >> np.uint64(32)).astype(np.uint32)
    k0 = np.uint32(seed & 0xFFFFFFFF)
    k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
    r = np.stack(philox4x32_10(c0, c1, c2, c3, k0, k1), axis=0)      # (4, numel)
    x = r[ii, np.arange(numel)]
    # _curand_uniform (curand_uniform.h:69-72): `x * 2^-32f + 2^-33f` in float:
    # uint32 -> float conversion rounds (RN) first; the multiply by a power of two
    # is exact, so FFMA contraction or mul+add give the same single rounding.
    xf = x.astype(np.float32)
    u = (xf.astype(np.float64) * 2.0 ** -32 + 2.0 ** -33).astype(np.float32)
    return u


def exponential_from_uniform(u: np.ndarray) -> np.ndarray:
    """transformation::exponential<float>, CUDA branch, lambda = 1."""
    eps = np.float32(np.finfo(np.float32).eps)
    with np.errstate(divide="ignore"):
        lg = np.where(u >= np.float32(1.0) - eps / np.float32(2), -eps / np.float32(2),
                      np.log(u.astype(np.float32)))
    return (np.float32(-1.0) * lg.astype(np.float32)).astype(np.float32)


def cuda_multinomial1(probs: np.ndarray, seed: int, offset: int, num_sms: int,
                      max_threads_per_sm: int = 2048):
    """Returns (token ids (B,), new philox offset).  probs: float32 (B, V)."""
    probs = np.ascontiguousarray(probs, dtype=np.float32)
    numel = probs.size
    counter_offset, _, _ = execution_policy(numel, num_sms, max_threads_per_sm)
    u = uniform_for_elements(numel, seed, offset, num_sms, max_threads_per_sm)
    q = exponential_from_uniform(u).reshape(probs.shape)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = (probs / q).astype(np.float32)
    return ratio.argmax(axis=-1).astype(np.int64), offset + counter_offset


class PhiloxSampler:
    """Drop-in `sampler` for oracle.sample_image that mimics a CUDA generator
    reseeded to (seed, offset 0) (plangen_base.py:526)."""

    def __init__(self, seed: int, num_sms: int, max_threads_per_sm: int = 2048):
        self.seed, self.offset = seed, 0
        self.num_sms, self.max_threads_per_sm = num_sms, max_threads_per_sm

    def __call__(self, probs, step):
        import torch
        tok, self.offset = cuda_multinomial1(probs.detach().float().cpu().numpy(), self.seed,
                                             self.offset, self.num_sms, self.max_threads_per_sm)
        return torch.from_numpy(tok).to(probs.device).unsqueeze(-1)
