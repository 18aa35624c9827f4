"""Secondary measurements (BASELINE.json configs[3] and configs[4]); bench.py stays on configs[1].
    python scripts/bench_hear.py            # HEAR feature extraction: 256 x 10 s clips -> [256, 996, 768]
    python scripts/bench_hear.py --nat      # WavJEPA-Nat (binaural, T = 400) training step, 64 clips x 8 crops, 1 GPU"""
import json
import os
import sys

os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import wavjepa_b200 as w  # noqa: E402
from wavjepa_b200 import _lib, hear  # noqa: E402

dev = torch.device("cuda", 0)
SPEC = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)]


def timed(fn, warmup=2, reps=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if "--pipeline" in sys.argv:
    from wavjepa_b200.preprocess import GpuAudioPipeline
    pipe = GpuAudioPipeline(device=dev)
    for asr in (44100, 48000, 16000):
        waves = [torch.randn(asr * 10, device=dev) * 0.1 for _ in range(64)]
        ms = timed(lambda: pipe(waves, [asr] * 64))
        print(json.dumps({"workload": f"8(f)-2 input pipeline: 64 clips x 10 s @ {asr} Hz -> [64, 1, 160000] (resample + RMS + pad)",
                          "ms_per_batch": round(ms, 3), "clips_per_s": round(64 / ms * 1e3, 0)}))
    sys.exit(0)
if "--nat" in sys.argv:
    torch.manual_seed(0)
    ex = w.ConvChannelFeatureExtractor(conv_layers_spec=SPEC, in_channels=2, share_weights_over_channels=False)
    model = w.JEPA(feature_extractor=ex, transformer_encoder_cfg=w.TransformerEncoderCFG.create(),
                   transformer_encoder_layers_cfg=w.TransformerLayerCFG.create(),
                   transformer_decoder_cfg=w.TransformerEncoderCFG.create(),
                   transformer_decoder_layers_cfg=w.TransformerLayerCFG.create(d_model=384), lr=4e-4,
                   adam_weight_decay=0.04, process_audio_seconds=2.01, nr_samples_per_audio=8,
                   average_top_k_layers=8).to(dev)
    model.global_step = 1000
    n_clips, crops, T = 64, 8, model.total_patches
    B = n_clips * crops
    masker = w.TimeInverseBlockMasker(4, 0.65, 10, 0.25, 10, 0.1, channel_based_masking=True, seed=5, device=dev)
    clips = torch.randn(n_clips, 2, 160000, device=dev)

    def step():
        ctx, tgt, vis = masker(batch_size=B, n_times=T, in_channels=2)
        x16, _, _, _ = model.on_after_batch_transfer((clips, ctx.view(n_clips, crops, T), tgt.view(n_clips, crops, 4, T),
                                                      vis.view(n_clips, crops, 4, T)), 0)
        return model.train_step(x16, ctx, tgt, vis)

    ms = timed(step)
    print(json.dumps({"workload": "configs[4] WavJEPA-Nat binaural pre-training step, 64 clips x 8 crops, T=400, 1 GPU",
                      "ms_per_step": round(ms, 2), "instances_per_s": round(B / ms * 1e3, 1),
                      "reserved_GiB": round(torch.cuda.memory_reserved() / 2**30, 1), "loss": step().item()}))
else:
    torch.manual_seed(0)
    import wavjepa_b200 as w
    init = w.JEPA(feature_extractor=w.ConvFeatureExtractor(conv_layers_spec=hear.BASE_SPEC, in_channels=1),
                  transformer_encoder_cfg=w.TransformerEncoderCFG.create(),
                  transformer_encoder_layers_cfg=w.TransformerLayerCFG.create(),
                  transformer_decoder_cfg=w.TransformerEncoderCFG.create(),
                  transformer_decoder_layers_cfg=w.TransformerLayerCFG.create(d_model=384), process_audio_seconds=2.01)
    model = hear.load_model({"state_dict": init.state_dict()})     # random-init weights (no network for checkpoints)
    n, L = 256, 160000
    audio = (torch.rand(n, L, device=dev) * 2 - 1)
    k0 = _lib.kernel_launches()
    emb, ts = hear.get_timestamp_embeddings(audio, model)
    launches = _lib.kernel_launches() - k0
    assert tuple(emb.shape) == (n, 996, 768) and tuple(ts.shape) == (n, 996) and torch.isfinite(emb).all()
    ms = timed(lambda: hear.get_timestamp_embeddings(audio, model))
    gflop = 45.35 * 5 * n   # 45.35 GFLOP per 2.01 s chunk, dense
    print(json.dumps({"workload": "configs[3] HEAR get_timestamp_embeddings, 256 clips x 10 s -> [256, 996, 768]",
                      "ms_per_call": round(ms, 2), "clips_per_s": round(n / ms * 1e3, 1),
                      "audio_seconds_per_s": round(n * 10 / ms * 1e3, 0), "dense_tflops": round(gflop / ms, 1),
                      "kernel_launches": launches, "reserved_GiB": round(torch.cuda.memory_reserved() / 2**30, 1)}))
