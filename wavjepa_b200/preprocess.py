"""On-GPU input pipeline: what the reference's DataLoader workers do on the CPU for every clip
(data_modules/WebAudioDataModule.py:43-61, data_modules/dataset_functions.py:92-114), as two kernel launches per clip:

    first channel -> Kaiser-windowed sinc resampling to 16 kHz -> RMS loudness normalisation to -14 dBFS over the whole
    resampled clip -> zero-pad / crop to 10 s  ->  clips [n, 1, 160000] fp32 on the device (the input contract of
    JEPA.on_after_batch_transfer, wavjepa/jepa.py:275-316)

The polyphase filter table is torchaudio's (third-party, pinned 2.7.x in requirements.txt; the container has 2.11 with
the same formula): `sinc_resample_table` restates _get_sinc_resample_kernel for the reference's arguments
(lowpass_filter_width=64, rolloff=0.9475937167399596, sinc_interp_kaiser, beta=14.769656459379492, computed in the
audio dtype like `Resample(dtype=audio.dtype)`).  It is built once per rate pair on the host (init-time, like the
positional tables); the convolution, the reduction and the gain run in csrc/resample.cu.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch

from . import ops

LOWPASS_FILTER_WIDTH = 64
ROLLOFF = 0.9475937167399596
BETA = 14.769656459379492


def sinc_resample_table(orig_freq: int, new_freq: int, lowpass_filter_width: int = LOWPASS_FILTER_WIDTH,
                        rolloff: float = ROLLOFF, beta: float = BETA,
                        dtype: torch.dtype = torch.float32) -> Tuple[torch.Tensor, int, int, int]:
    """-> (table [new, K] in `dtype`, width, orig, new) with orig/new reduced by their gcd; K = 2*width + orig."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=dtype)[None, :] / orig
    t = torch.arange(0, -new, -1, dtype=dtype)[:, None] / new + idx
    t = t * base_freq
    t = t.clamp(-lowpass_filter_width, lowpass_filter_width)
    window = torch.i0(torch.tensor(float(beta)) * torch.sqrt(1 - (t / lowpass_filter_width) ** 2)) / torch.i0(torch.tensor(float(beta)))
    t = t * math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0, dtype=t.dtype), t.sin() / t)
    kernels = kernels * (window * scale)
    return kernels.to(torch.float32), width, orig, new


class GpuAudioPipeline:
    """clips = GpuAudioPipeline(sr=16000, seconds=10)(waveforms, sample_rates)

    waveforms: sequence of fp32 tensors, [L] or [C, L] (the first channel is used, WebAudioDataModule.py:49); they may live
    on the host (pinned memory recommended) or on the device.  Returns [n, 1, sr*seconds] fp32 on the device."""

    def __init__(self, sr: int = 16000, seconds: int = 10, device: str | torch.device = "cuda", target_dbfs: float = -14.0):
        self.sr, self.seconds, self.device, self.target_dbfs = sr, seconds, torch.device(device), target_dbfs
        self._tables: Dict[int, Tuple[torch.Tensor, int, int, int]] = {}

    def table(self, audio_sr: int):
        if audio_sr not in self._tables:
            if audio_sr == self.sr:   # identity "filter": one tap, one phase
                self._tables[audio_sr] = (torch.ones(1, 1, device=self.device), 0, 1, 1)
            else:
                k, width, orig, new = sinc_resample_table(audio_sr, self.sr)
                self._tables[audio_sr] = (k.t().contiguous().to(self.device), width, orig, new)   # [K, new]
        return self._tables[audio_sr]

    @torch.no_grad()
    def __call__(self, waveforms: Sequence[torch.Tensor], sample_rates: Sequence[int]) -> torch.Tensor:
        n = len(waveforms)
        row = self.sr * self.seconds
        clips = torch.empty(n, 1, row, device=self.device, dtype=torch.float32)
        sumsq = torch.zeros(n, device=self.device, dtype=torch.float64)
        counts: List[int] = []
        for i, (wv, asr) in enumerate(zip(waveforms, sample_rates)):
            wv = wv[0] if wv.dim() > 1 else wv
            x = wv.to(self.device, torch.float32, non_blocking=True).contiguous()
            table_t, width, orig, new = self.table(int(asr))
            length = x.numel()
            target = -((-new * length) // orig)     # ceil(new * length / orig)
            ops.resample_sinc(x, table_t, orig, new, width, target, clips[i, 0], sumsq[i:i + 1])
            counts.append(target)
        cnt = torch.tensor(counts, device=self.device, dtype=torch.int64)
        ops.rms_gain_rows(clips.view(n, row), sumsq, cnt, self.target_dbfs)
        return clips
