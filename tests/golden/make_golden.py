"""Generates the golden fixtures in this directory by EXECUTING THE UNMODIFIED REFERENCE (labhamlet/wavjepa) on CPU.

Run in the build container only (needs /root/reference or $WAVJEPA_REF):   python tests/golden/make_golden.py
The reference ships no tests or golden vectors, so these files are the pins: the oracle restatement
(oracle/*.py) must reproduce them (tests/test_oracle_cpu.py) and the CUDA path is checked against both.

Weights: oracle.jepa_oracle.make_state_dict (deterministic per-tensor torch CPU generators) loaded into the reference
with load_state_dict(strict=True).  Inputs: oracle.inputs (seeded).  Tensors are stored as deterministic strided
subsamples + norms (oracle.inputs.subsample) to keep the fixtures small.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import inputs as oi          # noqa: E402
from oracle import jepa_oracle as jo     # noqa: E402
from oracle import masks_oracle as mo    # noqa: E402
from oracle import ref_loader            # noqa: E402

torch.set_num_threads(os.cpu_count())
ref = ref_loader.load_reference()


def stats(t):
    t = t.detach().float()
    return dict(shape=list(t.shape), l2=float(t.norm()), sum=float(t.double().sum()))


def ref_masks(masker, rows, n_times, in_channels, seed, cpa, first):
    """Runs the reference masker row by row under the seed contract (monkey-patched default_rng)."""
    outs, attempts = [], []
    with ref_loader.seeded_default_rng(seed, cpa, first) as rng:
        for r in rows:
            rng.set_row(r)
            outs.append(masker(batch_size=1, n_times=n_times, in_channels=in_channels))
            attempts.append(rng.call // cpa)
    c = torch.cat([o[0] for o in outs]).numpy()
    t = torch.cat([o[1] for o in outs]).numpy()
    v = torch.cat([o[2] for o in outs]).numpy()
    return c, t, v, np.asarray(attempts, dtype=np.int32)


def golden_masks():
    out = {}
    ti = ref.TimeInverseBlockMasker(target_masks_per_context=4, context_mask_prob=0.65, context_mask_length=10,
                                    target_prob=0.25, target_length=10, ratio_cutoff=0.1)
    c, t, v, a = ref_masks(ti, range(100, 164), 200, 1, 1234, 5, 0)
    out.update(ti_ctx=np.packbits(c), ti_tgt=np.packbits(t), ti_vis=np.packbits(v), ti_att=a)
    sp = ref.SpeechMasker(target_masks_per_context=4, target_prob=0.1, target_length=10, ratio_cutoff=0.5, min_context_len=5)
    c, t, v, a = ref_masks(sp, range(0, 64), 200, 1, 99, 4, 1)
    out.update(sp_ctx=np.packbits(c), sp_tgt=np.packbits(t), sp_vis=np.packbits(v), sp_att=a)
    nat = ref.TimeInverseBlockMasker(target_masks_per_context=4, context_mask_prob=0.65, context_mask_length=10,
                                     target_prob=0.25, target_length=10, ratio_cutoff=0.1, channel_based_masking=True)
    c, t, v, a = ref_masks(nat, range(7, 39), 400, 2, 5, 5, 0)
    out.update(nat_ctx=np.packbits(c), nat_tgt=np.packbits(t), nat_vis=np.packbits(v), nat_att=a)
    # a harder cutoff to exercise the rejection loop
    hard = ref.TimeInverseBlockMasker(target_masks_per_context=4, context_mask_prob=0.65, context_mask_length=10,
                                      target_prob=0.25, target_length=10, ratio_cutoff=0.2)
    c, t, v, a = ref_masks(hard, range(0, 64), 200, 1, 77, 5, 0)
    out.update(hard_ctx=np.packbits(c), hard_tgt=np.packbits(t), hard_vis=np.packbits(v), hard_att=a)
    np.savez_compressed(os.path.join(HERE, "masks.npz"), **out)
    print("masks.npz", {k: v.shape for k, v in out.items()})


def build_ref(cfg: jo.Cfg, sd):
    if cfg.per_channel:
        ext = ref.ConvChannelFeatureExtractor(conv_layers_spec=cfg.spec, in_channels=cfg.in_channels,
                                              share_weights_over_channels=False)
    else:
        ext = ref.ConvFeatureExtractor(conv_layers_spec=cfg.spec, in_channels=cfg.in_channels)
    m = ref.JEPA(feature_extractor=ext,
                 transformer_encoder_cfg=ref.TransformerEncoderCFG.create(num_layers=cfg.layers),
                 transformer_encoder_layers_cfg=ref.TransformerLayerCFG.create(d_model=cfg.d_model, nhead=cfg.nhead),
                 transformer_decoder_cfg=ref.TransformerEncoderCFG.create(num_layers=cfg.dec_layers),
                 transformer_decoder_layers_cfg=ref.TransformerLayerCFG.create(d_model=cfg.d_dec, nhead=cfg.dec_heads),
                 resample_sr=16000, process_audio_seconds=2.01, nr_samples_per_audio=8,
                 average_top_k_layers=cfg.top_k, compile_modules=False)
    missing = m.load_state_dict(sd, strict=True)
    print("load_state_dict:", missing)
    return m


def golden_state_dict_keys():
    cfg = jo.Cfg()
    torch.manual_seed(0)
    m = ref_loader.build_reference_jepa(ref)
    keys = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    assert [k for k, _ in keys] == list(jo.param_shapes(cfg).keys()), "oracle param order differs from the reference"
    with open(os.path.join(HERE, "ref_state_dict_keys.json"), "w") as f:
        json.dump(keys, f)
    print("ref_state_dict_keys.json", len(keys))


def golden_train(name, cfg, n_clips, crops, seed, masker="audioset"):
    sd = jo.make_state_dict(cfg, seed=3)
    m = build_ref(cfg, sd)
    m.train()
    inp = oi.training_inputs(cfg, n_clips, crops, seed=seed, masker=masker)
    out = m(inp["audio"], inp["ctx_masks"], inp["target_indices"], inp["ctx_and_target_masks"])
    out["loss"].backward()
    B, G, T = inp["target_indices"].shape
    preds_t = out["preds"].view(B, G, T, -1)[inp["target_indices"]]
    g = dict(loss=np.float64(out["loss"].item()))
    meta = {}
    for k, t in (("local_features", out["local_features"]), ("contextual_features", out["contextual_features"]),
                 ("targets", out["targets"]), ("preds_at_targets", preds_t)):
        g[k] = oi.subsample(t)
        meta[k] = stats(t)
    gn = {}
    for n_, p in m.named_parameters():
        if p.grad is not None:
            gn[n_] = float(p.grad.norm())
            if n_ in ("mask_token", "extract_audio.cnn.0.0.weight", "extract_audio.cnn.3.0.weight",
                      "encoder.layers.0.self_attn.in_proj_weight", "encoder.layers.11.linear2.weight",
                      "decoder.layers.5.linear1.weight", "decoder.layers.0.norm1.weight", "feature_norms.bias",
                      "post_extraction_mapper.weight", "decoder_to_encoder_mapper.bias",
                      "extract_audio.cnns.1.2.0.weight"):
                g["grad::" + n_] = oi.subsample(p.grad)
    # EMA at global_step 0 (jepa.py:186-198)
    m._step_teacher()
    g["teacher_after_ema::layers.3.linear1.weight"] = oi.subsample(m.teacher_encoder.layers[3].linear1.weight)
    meta["grad_norms"] = gn
    meta["n_c"] = (~inp["ctx_masks"]).sum(1).tolist()
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **g)
    with open(os.path.join(HERE, f"{name}.json"), "w") as f:
        json.dump(meta, f)
    print(name, "loss", g["loss"])


def golden_hear():
    cfg = jo.Cfg()
    sd = jo.make_state_dict(cfg, seed=3)
    import hear_api.feature_helper as fh
    import hear_api.runtime as rt
    fh.FeatureExtractor.forward = lambda self, x: self._wav2feature(x)  # the reference hard-codes .cuda() (:87)
    ext = ref.ConvFeatureExtractor(conv_layers_spec=cfg.spec, in_channels=1)
    model = rt.RuntimeJEPA(in_channels=1, weights={"state_dict": sd}, is_spectrogram=False, process_seconds=2.01,
                           extractor=ext, model_size="base", sr=16000)
    out = {}
    for tag, n in (("3s", 48000), ("10s", 160000), ("exact", 64318)):
        audio = oi.hear_inputs(2, n, seed=11)
        emb, ts = model.get_timestamp_embeddings(audio)
        out[f"{tag}_emb"] = oi.subsample(emb)
        out[f"{tag}_shape"] = np.asarray(emb.shape)
        out[f"{tag}_ts"] = ts[0].numpy()
        out[f"{tag}_l2"] = np.float64(emb.norm())
        if tag == "3s":
            out["3s_scene"] = model.get_scene_embeddings(audio).numpy()
        print("hear", tag, tuple(emb.shape))
    np.savez_compressed(os.path.join(HERE, "hear.npz"), **out)


def golden_hear_nat():
    """hear_api/runtime_natjepa.py, executed: `RuntimeNatJEPA.__init__` cannot run against the JEPA of this checkout (it
    passes in_channels= / is_spectrogram= through **kwargs into LightningModule.__init__ and later reads
    `self.model.in_channels`, which JEPA never defines -- the file targets another revision), so the instance is assembled
    by hand around an unmodified Nat JEPA and the UNMODIFIED get_timestamp_embeddings / get_scene_embeddings run on it."""
    import hear_api.feature_helper as fh
    import hear_api.runtime_natjepa as rn
    fh.FeatureExtractor.forward = lambda self, x: self._wav2feature(x)  # the reference hard-codes .cuda() (:87)
    cfg = jo.Cfg(in_channels=2, per_channel=True)
    sd = jo.make_state_dict(cfg, seed=3)
    jepa = build_ref(cfg, sd)
    jepa.eval()
    jepa.in_channels = 2                     # the attribute the runtime expects (:86, :142, :145)
    rt = object.__new__(rn.RuntimeNatJEPA)
    torch.nn.Module.__init__(rt)
    rt.sample_rate = 16000
    rt.model = jepa
    rt.unit_frames = int(2.01 * 16000)
    rt.output_steps = jepa.extract_audio.total_patches(rt.unit_frames) // 2
    rt.feature_extractor = fh.FeatureExtractor(in_channels=2)
    out = {}
    g = torch.Generator().manual_seed(17)
    for tag, audio in (("stereo", torch.rand(2, 2, 48000, generator=g) * 2 - 1), ("mono", oi.hear_inputs(2, 40000, seed=19))):
        emb, ts = rt.get_timestamp_embeddings(audio)
        out[f"{tag}_emb"] = oi.subsample(emb)
        out[f"{tag}_shape"] = np.asarray(emb.shape)
        out[f"{tag}_ts"] = ts[0].numpy()
        out[f"{tag}_l2"] = np.float64(emb.norm())
        print("hear_nat", tag, tuple(emb.shape))
    out["stereo_scene"] = rt.get_scene_embeddings(torch.rand(2, 2, 48000, generator=torch.Generator().manual_seed(17)) * 2 - 1).numpy()
    np.savez_compressed(os.path.join(HERE, "hear_nat.npz"), **out)


def golden_interop():
    """Rows of SURVEY.md 8(f)-3, executed on the unmodified reference: the ARCH wrapper (with a stub `arch_eval` base
    class -- the vendored harness needs pyannote), the 7-layer wav2vec2 extractor HEAR model (process_seconds=4.02) and
    the `size="large"` forward."""
    import types
    out = {}
    # ---- ARCH/configs/wavjepa_wrapper.py: get_embeddings on one clip
    stub = types.ModuleType("arch_eval")
    stub.Model = type("Model", (), {"__init__": lambda self, model, **kw: None})
    stub.ClassificationModel = stub.Model
    sys.modules["arch_eval"] = stub
    sys.path.insert(0, os.path.join(ref.root, "ARCH"))
    import importlib
    wrap = importlib.import_module("configs.wavjepa_wrapper")
    cfg = jo.Cfg()
    sd = jo.make_state_dict(cfg, seed=3)
    m = build_ref(cfg, sd)
    m.eval()
    w = wrap.WavJEPAModelWrapper(m, "cpu", None)
    for tag, n in (("a", 40000), ("b", 64318)):
        audio = oi.hear_inputs(1, n, seed=21)[0]
        out[f"arch_{tag}"] = w.get_embeddings(audio).numpy()
        print("arch", tag, out[f"arch_{tag}"].shape)
    # ---- hear_configs/WavJEPA_w2v2.py: 7-layer extractor, 4.02 s windows
    import hear_api.feature_helper as fh
    import hear_api.runtime as rt
    fh.FeatureExtractor.forward = lambda self, x: self._wav2feature(x)  # the reference hard-codes .cuda() (:87)
    w2v2 = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2
    cfg2 = jo.Cfg(spec=w2v2, seconds=4.02)
    sd2 = jo.make_state_dict(cfg2, seed=3)
    ext = ref.ConvFeatureExtractor(conv_layers_spec=w2v2, in_channels=1)
    model = rt.RuntimeJEPA(in_channels=1, weights={"state_dict": sd2}, is_spectrogram=False, process_seconds=4.02,
                           extractor=ext, model_size="base", sr=16000)
    audio = oi.hear_inputs(2, 100000, seed=12)
    emb, ts = model.get_timestamp_embeddings(audio)
    out["w2v2_emb"] = oi.subsample(emb)
    out["w2v2_shape"] = np.asarray(emb.shape)
    out["w2v2_ts"] = ts[0].numpy()
    out["w2v2_l2"] = np.float64(emb.norm())
    print("w2v2", tuple(emb.shape))
    # ---- size="large": 24 x 1024 x 16 heads (jepa.py:113-118), forward + loss on 2 instances
    cfgL = jo.Cfg(d_model=1024, nhead=16, layers=24)
    sdL = jo.make_state_dict(cfgL, seed=3)
    extL = ref.ConvFeatureExtractor(conv_layers_spec=cfgL.spec, in_channels=1)
    mL = ref.JEPA(feature_extractor=extL, transformer_encoder_cfg=ref.TransformerEncoderCFG.create(),
                  transformer_encoder_layers_cfg=ref.TransformerLayerCFG.create(),
                  transformer_decoder_cfg=ref.TransformerEncoderCFG.create(),
                  transformer_decoder_layers_cfg=ref.TransformerLayerCFG.create(d_model=384), resample_sr=16000,
                  process_audio_seconds=2.01, nr_samples_per_audio=8, average_top_k_layers=8, compile_modules=False,
                  size="large")
    print("large load:", mL.load_state_dict(sdL, strict=True))
    inp = oi.training_inputs(cfgL, 1, 2, seed=77, masker="audioset")
    with torch.no_grad():
        o = mL(inp["audio"], inp["ctx_masks"], inp["target_indices"], inp["ctx_and_target_masks"])
    out["large_loss"] = np.float64(o["loss"].item())
    out["large_local"] = oi.subsample(o["local_features"])
    out["large_targets"] = oi.subsample(o["targets"])
    print("large loss", out["large_loss"])
    np.savez_compressed(os.path.join(HERE, "interop.npz"), **out)


def golden_denoiser():
    """SURVEY.md 8(f)-4, executed on the unmodified reference: Denoiser.on_after_batch_transfer (scene generation,
    resampling, crops) with the random draws pinned, then Denoiser.forward + backward with a frozen WavJEPA-base
    teacher loaded through the reference's own `_set_teacher` (a temporary Lightning-style checkpoint file)."""
    import tempfile
    from wavjepa.denoiser import Denoiser
    out = {}
    cfg = jo.Cfg()
    sd_s = jo.make_state_dict(cfg, seed=11)
    sd_t = jo.make_state_dict(cfg, seed=12)
    ext = ref.ConvFeatureExtractor(conv_layers_spec=cfg.spec, in_channels=1)
    alpha = 0.25
    m = Denoiser(feature_extractor=ext, transformer_encoder_layers_cfg=ref.TransformerLayerCFG.create(),
                 transformer_encoder_cfg=ref.TransformerEncoderCFG.create(), resample_sr=16000,
                 process_audio_seconds=2.01, nr_samples_per_audio=2, size="base", alpha=alpha)
    own = {k: v for k, v in sd_s.items() if k in m.state_dict()}
    print("denoiser student load:", m.load_state_dict(own, strict=True))
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "teacher.ckpt")
        torch.save({"state_dict": sd_t}, path)
        m._set_teacher(path)
    json.dump(list(m.state_dict().keys()), open(os.path.join(HERE, "ref_denoiser_state_dict_keys.json"), "w"))
    # ---- batch preparation with the two random draws pinned
    batch = oi.denoiser_batch()
    B, nr, tl = batch[0].shape[0], 2, m.target_length
    L16 = batch[0].shape[1] // 2
    g = torch.Generator().manual_seed(3)
    starts = torch.randint(0, L16 - tl + 1, (B, nr), generator=g)
    perm = torch.randperm(B * nr, generator=g)
    # reference defect: on_after_batch_transfer reads `self.ORIGINAL_SR` (wavjepa/denoiser.py:262) but only the module
    # constant ORIGINAL_SR = 32000 (:23) exists -> AttributeError as shipped.  The intended value is supplied here.
    m.ORIGINAL_SR = 32000
    real_randint, real_randperm = torch.randint, torch.randperm
    torch.randint = lambda *a, **k: starts.clone()
    torch.randperm = lambda *a, **k: perm.clone()
    try:
        gen16, clean16 = m.on_after_batch_transfer(batch, 0)
    finally:
        torch.randint, torch.randperm = real_randint, real_randperm
    out["starts"], out["perm"] = starts.numpy(), perm.numpy()
    out["gen"] = oi.subsample(gen16.float())
    out["clean"] = oi.subsample(clean16.float())
    out["gen_l2"], out["clean_l2"] = np.float64(gen16.float().norm()), np.float64(clean16.float().norm())
    print("denoiser batch", tuple(gen16.shape), gen16.dtype)
    # ---- forward + backward (fp32 CPU, like the other training goldens)
    o = m(gen16.float(), clean16.float())
    o["loss"].backward()
    out["loss"] = np.float64(o["loss"].item())
    out["loss_clean"] = np.float64(o["loss_clean"].item())
    out["loss_dd"] = np.float64(o["loss_denoise_dereverb"].item())
    gn = {}
    for n, p in m.named_parameters():
        if p.grad is not None:
            gn[n] = float(p.grad.norm())
            if n in ("encoder.layers.11.linear2.weight", "encoder.layers.0.self_attn.in_proj_weight",
                     "post_extraction_mapper.weight", "extract_audio.cnn.0.0.weight", "extract_audio.cnn.3.0.weight",
                     "encoder.norm.weight", "feature_norms.bias"):
                out["grad:" + n] = oi.subsample(p.grad)
    json.dump(dict(alpha=alpha, grad_norms=gn), open(os.path.join(HERE, "denoiser.json"), "w"), indent=0)
    np.savez_compressed(os.path.join(HERE, "denoiser.npz"), **out)
    print("denoiser losses", out["loss"], out["loss_clean"], out["loss_dd"], "grads", len(gn))


if __name__ == "__main__":
    which = sys.argv[1:] or ["masks", "keys", "train", "nat", "hear", "hear_nat", "interop", "denoiser"]
    if "masks" in which:
        golden_masks()
    if "keys" in which:
        golden_state_dict_keys()
    if "train" in which:
        golden_train("train_c1", jo.Cfg(), n_clips=1, crops=4, seed=1234)
        golden_train("train_speech", jo.Cfg(), n_clips=1, crops=2, seed=4321, masker="librispeech")
    if "nat" in which:
        golden_train("train_nat", jo.Cfg(in_channels=2, per_channel=True), n_clips=1, crops=2, seed=55)
    if "hear" in which:
        golden_hear()
    if "hear_nat" in which:
        golden_hear_nat()
    if "interop" in which:
        golden_interop()
    if "denoiser" in which:
        golden_denoiser()
