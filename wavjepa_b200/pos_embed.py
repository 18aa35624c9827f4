"""Fixed 1-D sin-cos positional table (reference wavjepa/pos_embed.py:75-93, used by wavjepa/jepa.py:163-180).

Init-time only: omega_i = 10000^(-i / (D/2)), table[p] = [sin(p * omega) || cos(p * omega)] (halves concatenated, not
interleaved), evaluated in float64 and stored as float32.
"""
from __future__ import annotations

import numpy as np


def get_1d_sincos_pos_embed_from_grid(embed_dim: int, pos: np.ndarray) -> np.ndarray:
    if embed_dim % 2 != 0:
        raise ValueError("embed_dim must be even")
    half = embed_dim // 2
    omega = 1.0 / np.power(10000.0, np.arange(half, dtype=np.float64) / float(half))
    angles = np.asarray(pos, dtype=np.float64).reshape(-1, 1) * omega.reshape(1, -1)
    return np.concatenate([np.sin(angles), np.cos(angles)], axis=1)
