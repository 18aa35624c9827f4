"""Timing record of the denoiser stage (not the north-star metric): batch preparation (scene
generation at 32 kHz for 32 clips of 10 s with 2 s RIRs, resampling, 16 crops per clip) and Denoiser.train_step on
512 instances per side.      PYTHONPATH=. python scripts/bench_denoiser.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch  # noqa: E402

import wavjepa_b200 as w  # noqa: E402
from wavjepa_b200 import _lib  # noqa: E402
from wavjepa_b200.denoiser import Denoiser  # noqa: E402

dev = "cuda"
SPEC = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)]
torch.manual_seed(0)
ext = w.ConvFeatureExtractor(conv_layers_spec=SPEC, in_channels=1)
m = Denoiser(feature_extractor=ext, transformer_encoder_layers_cfg=w.TransformerLayerCFG.create(),
             transformer_encoder_cfg=w.TransformerEncoderCFG.create(), nr_samples_per_audio=16, alpha=0.25)
teacher = w.JEPA(feature_extractor=w.ConvFeatureExtractor(conv_layers_spec=SPEC, in_channels=1),
                 transformer_encoder_cfg=w.TransformerEncoderCFG.create(),
                 transformer_encoder_layers_cfg=w.TransformerLayerCFG.create(),
                 transformer_decoder_cfg=w.TransformerEncoderCFG.create(),
                 transformer_decoder_layers_cfg=w.TransformerLayerCFG.create(d_model=384), process_audio_seconds=2.01)
m._set_teacher({"state_dict": teacher.state_dict()})     # random-init weights (there is no network for checkpoints)
m.to(dev)
m.global_step = 5000

B, T32, R, S = 32, 320000, 64000, 2
g = torch.Generator(device=dev).manual_seed(0)
audio = torch.rand(B, T32, device=dev, generator=g) * 2 - 1
rir = torch.randn(B, 2, R, device=dev, generator=g) * torch.exp(-torch.arange(R, device=dev) / 8000.0) * 0.05
noise = torch.randn(B, T32, device=dev, generator=g) * 0.3
nrirs = torch.randn(B, S, 2, R, device=dev, generator=g) * torch.exp(-torch.arange(R, device=dev) / 8000.0) * 0.05
batch = (audio, rir, noise, torch.full((B,), T32 // 2, device=dev), torch.full((B,), 1000, device=dev), nrirs,
         torch.full((B,), 10.0, device=dev))


def timed(fn, reps, warmup=1):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out


ms_prep, (gen16, clean16) = timed(lambda: m.on_after_batch_transfer(batch, 0), 5, warmup=3)
for _ in range(3):
    m.train_step(gen16, clean16)
n0 = _lib.kernel_launches()
ms_step, out = timed(lambda: m.train_step(gen16, clean16), 5)
launches = (_lib.kernel_launches() - n0) // 6
N = gen16.shape[0]
print(json.dumps({"workload": f"denoiser stage: {B} clips x 10 s @ 32 kHz -> {N} instances per side (clean + generated)",
                  "prep_ms": round(ms_prep, 2), "train_step_ms": round(ms_step, 2),
                  "instances_per_s": round(N / ms_step * 1e3, 1), "kernel_launches_per_step": int(launches),
                  "loss": float(out["loss"]), "reserved_GiB": round(torch.cuda.memory_reserved() / 2**30, 1)}))
