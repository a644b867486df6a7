"""Shared helpers for the -m gpu parity tests (CUDA engine vs oracle)."""
import numpy as np
import torch

from oracle import janus_oracle as O


def product_dims(d):
    from plangen_b200.config import Dims
    return Dims.from_any(d)


_ENGINES = {}


def get_engine(d, mode, seed=0, max_batch=4, max_prompt=64, with_vq=True, options=None, sd=None, cache=True):
    from plangen_b200.engine import FastJanus
    key = (d.name, mode, seed, max_batch, max_prompt, with_vq, tuple(sorted((options or {}).items())))
    if cache and key in _ENGINES:
        return _ENGINES[key]
    if sd is None:
        sd = O.init_state_dict(d, seed=seed, with_vq=with_vq)
    eng = FastJanus(sd, product_dims(d), mode=mode, max_batch=max_batch, max_prompt=max_prompt, with_vq=with_vq,
                    options=options)
    if cache:
        _ENGINES[key] = eng
    return eng


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def assert_close(a, b, rtol, atol_frac, what=""):
    """|a-b| <= rtol*|b| + atol_frac*max|b|  (logits of a random-init model are near zero, so a pure
    relative test is meaningless for the smallest entries)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    tol = rtol * np.abs(b) + atol_frac * np.abs(b).max()
    bad = np.abs(a - b) > tol
    if bad.any():
        i = np.unravel_index(np.argmax(np.abs(a - b) - tol), a.shape)
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.size} elements off; worst at {i}: got {a[i]} want {b[i]} "
                             f"(max|want|={np.abs(b).max():.4g}, max abs err={np.abs(a-b).max():.4g})")
