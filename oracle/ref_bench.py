"""TEST / BENCH INFRASTRUCTURE ONLY -- times the EXECUTED, UNMODIFIED reference (labhamlet/wavjepa) on this machine.

Used by `bench.py --impl reference`, `bench.py`'s `cpu_baseline` leg and its `gpu_torch_baseline` field; never
imported by the product package.  The reference is imported through oracle/ref_loader.py from $WAVJEPA_REF,
/root/reference or the staged copy baseline/_ref/ (oracle/stage_reference.py) with in-process stand-ins for the two
packages this image lacks (pytorch_lightning, webdataset).

One "step" is the reference's own training step as its Lightning loop would run it
(train.py:164-180 + wavjepa/jepa.py:275-333, 215-228):

    batch = model.on_after_batch_transfer((clips, ctx_masks, target_indices, ctx_and_target_masks), 0)   # crop + normalise
    out = model.training_step(batch, 0)            # forward (student, predictor, EMA teacher, loss) + EMA update
    out["loss"].backward(); clip_grad_norm_(5.0); optimizer.step(); scheduler.step(); zero_grad

on BASELINE.json config 1 (2 clips x 8 crops = 16 instances of 2.01 s, N(0,1) noise at 16 kHz, AudioSet masker,
WavJEPA-base random init under torch.manual_seed(0), top-k 8).  CPU runs are fp32 without autocast on all host cores;
the GPU run is torch eager under bf16 autocast (configs/trainer/default_trainer.yaml:5 `bf16-mixed`).

    python -m oracle.ref_bench --device cpu|cuda [--clips 2] [--steps 3] [--warmup 1] [--masks] [--hear]
prints one JSON object.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import types

import torch

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

from oracle import ref_loader  # noqa: E402

MASKER = dict(target_masks_per_context=4, context_mask_prob=0.65, context_mask_length=10, target_prob=0.25,
              target_length=10, ratio_cutoff=0.1)   # configs/masker/AudioSet.yaml
CROPS, CLIP_LEN = 8, 160000


def available() -> bool:
    return ref_loader.find_reference() is not None


def _build(device: str, crops: int = CROPS):
    ref = ref_loader.load_reference()
    torch.manual_seed(0)
    ext = ref.ConvFeatureExtractor(conv_layers_spec=ref_loader.BASE_SPEC, in_channels=1)
    model = ref.JEPA(feature_extractor=ext,
                     transformer_encoder_cfg=ref.TransformerEncoderCFG.create(),
                     transformer_encoder_layers_cfg=ref.TransformerLayerCFG.create(),
                     transformer_decoder_cfg=ref.TransformerEncoderCFG.create(),
                     transformer_decoder_layers_cfg=ref.TransformerLayerCFG.create(d_model=384),
                     lr=4e-4, adam_betas=(0.9, 0.98), adam_weight_decay=0.04, resample_sr=16000,
                     process_audio_seconds=2.01, nr_samples_per_audio=crops, average_top_k_layers=8,
                     compile_modules=False)      # configs/trainer/default_trainer.yaml
    model.to(device)
    model.train()
    model.trainer = types.SimpleNamespace(max_steps=375000)   # configure_optimizers reads trainer.max_steps (:225)
    model.global_step = 1000        # inside the warm-up so that lr > 0, like bench.py's native arm
    return ref, model


def train_step_fn(device: str, n_clips: int, crops: int = CROPS):
    """-> (step callable returning the loss as a float, instances per step).  Everything the reference's Lightning loop
    does per optimisation step except logging and the DataLoader (masks are made by the reference's own masker on the
    CPU before the timed region, as its DataLoader workers would, data_modules/WebAudioDataModule.py:63-67)."""
    ref, model = _build(device, crops)
    cfgd = model.configure_optimizers()
    opt, sched = cfgd["optimizer"], cfgd["lr_scheduler"]["scheduler"]
    import warnings
    with warnings.catch_warnings():     # position the LambdaLR at global_step (lr = 4e-4 * 1000 / 100000)
        warnings.simplefilter("ignore")
        sched.last_epoch = model.global_step - 1
        sched.step()
    masker = ref.TimeInverseBlockMasker(**MASKER)
    T = model.total_patches
    g = torch.Generator().manual_seed(1234)
    pool = []
    for _ in range(2):
        clips = torch.randn(n_clips, 1, CLIP_LEN, generator=g)
        ms = [masker(batch_size=crops, n_times=T, in_channels=1) for _ in range(n_clips)]
        c = torch.stack([m[0] for m in ms]); t = torch.stack([m[1] for m in ms]); v = torch.stack([m[2] for m in ms])
        pool.append(tuple(x.to(device) for x in (clips, c, t, v)))
    params = [p for p in model.parameters() if p.requires_grad]
    use_amp = device != "cpu"
    state = {"i": 0}

    def step() -> float:
        batch = pool[state["i"] % len(pool)]
        state["i"] += 1
        batch = model.on_after_batch_transfer(batch, 0)
        if not use_amp:   # the hook emits bf16 for CUDA autocast (:311); the fp32 CPU path takes the same values as fp32
            batch = (batch[0].float(),) + tuple(batch[1:])
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=use_amp):
            out = model.training_step(batch, 0)
        out["loss"].backward()
        torch.nn.utils.clip_grad_norm_(params, 5.0)     # train.py:177-178
        opt.step()
        sched.step()
        opt.zero_grad(set_to_none=True)
        model.global_step += 1
        return float(out["loss"].detach())

    return step, n_clips * crops


def time_train(device: str, n_clips: int, steps: int, warmup: int):
    step, n_inst = train_step_fn(device, n_clips)
    sync = torch.cuda.synchronize if device != "cpu" else (lambda: None)
    loss = None
    for _ in range(warmup):
        loss = step()
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = step()
    sync()
    dt = (time.perf_counter() - t0) / steps
    return {"instances_per_s": n_inst / dt, "s_per_step": dt, "instances_per_step": n_inst, "loss": loss}


def time_masks(rows: int = 256):
    """Reference CPU mask generation, one thread (BASELINE.md 4 iii): rows/s of TimeInverseBlockMasker / SpeechMasker."""
    ref = ref_loader.load_reference()
    out = {}
    for name, mk in (("time_inverse", ref.TimeInverseBlockMasker(**MASKER)),
                     ("speech", ref.SpeechMasker(target_masks_per_context=4, target_prob=0.1, target_length=10,
                                                 ratio_cutoff=0.5, min_context_len=5))):
        mk(batch_size=8, n_times=200, in_channels=1)
        t0 = time.perf_counter()
        for _ in range(rows // 8):
            mk(batch_size=8, n_times=200, in_channels=1)
        out[name + "_rows_per_s"] = rows / (time.perf_counter() - t0)
    return out


def time_hear(n_clips: int = 2, n_samples: int = 160000):
    """Reference HEAR `get_timestamp_embeddings` on CPU, fp32 (BASELINE.md 4 iv)."""
    ref = ref_loader.load_reference()
    import hear_api.feature_helper as fh
    import hear_api.runtime as rt
    fh.FeatureExtractor.forward = lambda self, x: self._wav2feature(x)   # the reference hard-codes .cuda() (:87)
    torch.manual_seed(0)
    jepa = ref_loader.build_reference_jepa(ref)
    ext = ref.ConvFeatureExtractor(conv_layers_spec=ref_loader.BASE_SPEC, in_channels=1)
    model = rt.RuntimeJEPA(in_channels=1, weights={"state_dict": jepa.state_dict()}, is_spectrogram=False,
                           process_seconds=2.01, extractor=ext, model_size="base", sr=16000)
    model.model.cpu()     # (the reference's constructor moves the model to CUDA when there is one: this is the CPU arm)
    audio = torch.rand(n_clips, n_samples, generator=torch.Generator().manual_seed(1234)) * 2 - 1
    with torch.no_grad():
        model.get_timestamp_embeddings(audio)
        t0 = time.perf_counter()
        emb, _ = model.get_timestamp_embeddings(audio)
        dt = time.perf_counter() - t0
    return {"clips_per_s": n_clips / dt, "s": dt, "shape": list(emb.shape)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--clips", type=int, default=2)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--masks", action="store_true")
    ap.add_argument("--hear", action="store_true")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    out = {"reference_root": ref_loader.find_reference(), "device": args.device, "threads": torch.get_num_threads(),
           "host_cpus": os.cpu_count(), "torch": torch.__version__}
    if args.device != "cpu":
        torch.cuda.set_device(0)
        out["gpu"] = torch.cuda.get_device_name(0)
        n = args.clips
        while True:       # the dense reference keeps every activation: halve the batch until it fits
            try:
                out["train"] = time_train("cuda", n, args.steps, args.warmup)
                out["train"]["clips"] = n
                out["train"]["peak_mem_gib"] = torch.cuda.max_memory_allocated() / 2 ** 30
                break
            except torch.cuda.OutOfMemoryError:
                torch.cuda.empty_cache()
                if n <= 1:
                    out["train"] = {"error": "out of memory at 1 clip"}
                    break
                n //= 2
    else:
        out["train"] = time_train("cpu", args.clips, args.steps, args.warmup)
        if args.masks:
            out["masks"] = time_masks()
        if args.hear:
            out["hear"] = time_hear()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
