"""The data module's tuple contract on the GPU (reference data_modules/WebAudioDataModule.py:43-74).

The reference's DataLoader workers turn every decoded clip into
    (audio [160000] fp32 @16 kHz, context_mask [S, T], target_indices [S, G, T], ctx_and_target_masks [S, G, T])
(`_retrieve_sample`: first channel -> Kaiser-sinc resample -> pre_process (RMS -14 dBFS, pad / crop to 10 s) -> one masker
call per clip with batch_size = nr_samples_per_audio), and `.batched(batch_size)` stacks them into the 4-tuple that
`JEPA.on_after_batch_transfer` receives.  Decoding stays where it is (CPU, WebDataset is out of scope); everything after
it is two kernels per clip + one masker launch per BATCH here: `GpuBatchAssembler(masker)(samples)` takes the list of
decoded `(waveform, sample_rate)` pairs of one batch and returns the same 4-tuple, already on the device.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from .preprocess import GpuAudioPipeline


class GpuBatchAssembler:
    def __init__(self, masker, nr_samples_per_audio: int = 8, nr_time_points: int = 200, in_channels: int = 1,
                 sr: int = 16000, seconds: int = 10, device: str | torch.device = "cuda"):
        self.masker, self.S, self.T, self.C = masker, nr_samples_per_audio, nr_time_points, in_channels
        self.pipe = GpuAudioPipeline(sr=sr, seconds=seconds, device=device)

    @torch.no_grad()
    def __call__(self, samples: Sequence[Tuple[torch.Tensor, int]]):
        """samples: [(waveform [L] or [C, L] fp32, sample_rate), ...] of one batch (what `.decode(wds.torch_audio)
        .to_tuple("flac")` yields, WebAudioDataModule.py:88-94).  Returns (audio [n, 1, sr*seconds] fp32,
        ctx_masks [n, S, T'], target_indices [n, S, G, T'], ctx_and_target_masks [n, S, G, T']) on the device."""
        n = len(samples)
        audio = self.pipe([w for w, _ in samples], [int(r) for _, r in samples])
        ctx, tgt, vis = self.masker(batch_size=n * self.S, n_times=self.T, in_channels=self.C)
        t_out = ctx.shape[-1]
        g = tgt.shape[1]
        return audio, ctx.view(n, self.S, t_out), tgt.view(n, self.S, g, t_out), vis.view(n, self.S, g, t_out)
