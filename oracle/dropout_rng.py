"""numpy replica of plbert_b200/csrc/rng.cuh — the counter-based generator behind the CUDA path's dropout masks.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): tests build the exact masks the kernels use and hand them to the
fp64 oracle / a PyTorch reference, so dropout parity is checked element for element, not only statistically.
Reference semantics: torch.nn.Dropout — keep with probability 1 - p, scale kept values by 1 / (1 - p)
(HF modeling_albert.py:109,134,201).
"""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFF)


def rng32(seed: int, stream: int, ctr: np.ndarray) -> np.ndarray:
    """32 random bits per counter (uint32 array in, uint32 array out)."""
    lo, hi = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    x = (np.asarray(ctr, dtype=np.uint64) * np.uint64(0x9E3779B1) + lo) & _M
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x85EBCA6B)) & _M
    x ^= (np.uint64(stream & 0xFFFFFFFF) * np.uint64(0xC2B2AE35) + hi) & _M
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE3D)) & _M
    x ^= x >> np.uint64(16)
    return x.astype(np.uint32)


def thresh16(p: float) -> int:
    return int((1.0 - float(p)) * 65536.0 + 0.5)


def _draws(seed: int, stream: int, pair_ctr: np.ndarray, odd: np.ndarray) -> np.ndarray:
    x = rng32(seed, stream, pair_ctr)
    return np.where(odd, x >> np.uint32(16), x & np.uint32(0xFFFF))


def hidden_mask(rows: int, cols: int, p: float, seed: int, stream: int) -> np.ndarray:
    """Multipliers (0 or 1/(1-p)) of plb_dropout over a [rows, cols] matrix: element index = r * cols + c."""
    if p == 0.0:
        return np.ones((rows, cols))
    idx = np.arange(rows * cols, dtype=np.uint64).reshape(rows, cols)
    keep = _draws(seed, stream, (idx >> np.uint64(1)).astype(np.uint32), (idx & np.uint64(1)) != 0) < thresh16(p)
    return keep / (1.0 - p)


def attention_mask(items: int, q_len: int, k_len: int, p: float, seed: int, stream: int) -> np.ndarray:
    """Multipliers of the attention probabilities, [items, q_len, k_len]; item = sequence * heads + head
    (rng.cuh attn_dropout_ctr: pair (k, k + 1) of query row q)."""
    if p == 0.0:
        return np.ones((items, q_len, k_len))
    it = np.arange(items, dtype=np.uint64)[:, None, None]
    q = np.arange(q_len, dtype=np.uint64)[None, :, None]
    k = np.arange(k_len, dtype=np.uint64)[None, None, :]
    ctr = (((it * np.uint64(512) + q) * np.uint64(256) + (k >> np.uint64(1))) & _M).astype(np.uint32)
    keep = _draws(seed, stream, ctr, np.broadcast_to((k & np.uint64(1)) != 0, ctr.shape)) < thresh16(p)
    return keep / (1.0 - p)
