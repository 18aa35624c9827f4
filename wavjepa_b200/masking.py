"""Context / target mask generators with the reference's class API, running as one integer CUDA kernel (warp per row).

Mirrors wavjepa/masking.py: `TimeInverseBlockMasker` (:46-128) and `SpeechMasker` (:131-207); both sit on
`compute_mask_indices` (wavjepa/audio_masking.py:5-194) and numpy's default_rng stream, which the kernel
(csrc/masks.cu) reproduces bit for bit.

Differences a maintainer must know about:
  * The reference masks are UNSEEDED (np.random.default_rng(None), audio_masking.py:59-64).  Here every row is
    reproducible: call c (0 = context, 1..G = target groups) of rejection-loop attempt a for global row r draws from
    numpy.random.default_rng([seed, r, a*8 + c]).  `r` starts at `row0` and advances by `batch_size` per call, so a
    data-parallel rank passes row0 = rank * rows_per_rank (or calls set_row) to get disjoint streams.
  * Outputs live on the CUDA device (the reference returns CPU tensors from DataLoader workers); `.cpu()` them if a
    CPU pipeline needs them.  There is no CPU implementation in this package.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops


class _SeededMasker(nn.Module):
    kind = -1

    def __init__(self, seed: int = 0, row0: int = 0, device: str | torch.device = "cuda"):
        super().__init__()
        self.seed = int(seed)
        self.next_row = int(row0)
        self.device = torch.device(device)
        self.last_attempts = None

    def set_row(self, row0: int) -> None:
        self.next_row = int(row0)

    def _generate(self, batch_size, n_times, in_channels, ctx_prob, ctx_len, min_context_len):
        ctx, tgt, vis, attempts, err = ops.masks_generate(
            self.kind, batch_size, n_times, in_channels, bool(self.channel_based_masking),
            self.target_masks_per_context, ctx_prob, ctx_len, self.target_prob, self.target_length,
            self.ratio_cutoff, min_context_len, self.seed, self.next_row, self.device)
        self.next_row += batch_size
        self.last_attempts = attempts
        self._err = err   # checked lazily (reading it is a host sync); see check()
        return ctx, tgt, vis

    def check(self) -> None:
        """Raises if the last call hit the reference's `num_mask == 0` ValueError path (audio_masking.py:101-103)."""
        if self._err is not None and int(self._err.item()) != 0:
            raise ValueError("mask generation failed: a span mask had zero spans (the reference raises here too)")


class TimeInverseBlockMasker(_SeededMasker):
    """reference wavjepa/masking.py:46-128 (AudioSet masker: context = inverse of a span mask minus all targets)."""
    kind = 0

    def __init__(self, target_masks_per_context: int = 4, context_mask_prob: float = 0.3,
                 context_mask_length: int = 10, target_prob: float = 0.2, target_length: int = 20,
                 ratio_cutoff: float = 0.05, channel_based_masking: bool = False, seed: int = 0, row0: int = 0,
                 device: str | torch.device = "cuda", **kwargs):
        super().__init__(seed, row0, device)
        self.target_masks_per_context = target_masks_per_context
        self.context_mask_prob = context_mask_prob
        self.context_mask_length = context_mask_length
        self.target_prob = target_prob
        self.target_length = target_length
        self.ratio_cutoff = ratio_cutoff
        self.channel_based_masking = channel_based_masking

    def forward(self, batch_size: int, n_times: int, in_channels: int):
        """-> (final_context_mask [B,T] True = hidden, target_positions [B,G,T] True = predict,
        context_and_target_mask [B,G,T] True = hidden from the predictor), all torch.bool on the device."""
        return self._generate(batch_size, n_times, in_channels, self.context_mask_prob, self.context_mask_length, 0)


class SpeechMasker(_SeededMasker):
    """reference wavjepa/masking.py:131-207 (LibriSpeech masker: context = complement of the targets with short
    visible runs removed by filter_small_clusters)."""
    kind = 1

    def __init__(self, target_masks_per_context: int = 4, target_prob: float = 0.25, target_length: int = 5,
                 ratio_cutoff: float = 0.3, min_context_len: int = 5, channel_based_masking: bool = False,
                 seed: int = 0, row0: int = 0, device: str | torch.device = "cuda", **kwargs):
        super().__init__(seed, row0, device)
        self.target_masks_per_context = target_masks_per_context
        self.target_prob = target_prob
        self.target_length = target_length
        self.ratio_cutoff = ratio_cutoff
        self.channel_based_masking = channel_based_masking
        self.min_context_len = min_context_len

    def forward(self, batch_size: int, n_times: int, in_channels: int):
        return self._generate(batch_size, n_times, in_channels, 0.0, 1, self.min_context_len)
