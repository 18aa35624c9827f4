"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference (labhamlet/wavjepa) for pinning.

The reference is pure Python but imports two packages that are absent from this image
(`pytorch_lightning`, `webdataset`).  We pre-register tiny in-process stand-ins for them in
`sys.modules` and put the reference checkout on `sys.path`; nothing in the reference is edited
or copied.  Search order for the checkout: $WAVJEPA_REF, /root/reference, baseline/_ref.

Only `tests/golden/make_golden.py` (run in the build container, where /root/reference exists)
and `bench.py --impl reference` use this module.  It is never imported by the product package.
"""
from __future__ import annotations

import os
import sys
import types

import torch
from torch import nn

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference() -> str | None:
    for cand in (os.environ.get("WAVJEPA_REF"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "wavjepa")):
            return cand
    return None


class _AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _install_stubs() -> None:
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            """Minimal stand-in: the reference only uses the attributes below (wavjepa/jepa.py:97-107,186-191,328)."""

            def __init__(self, *a, **k):
                super().__init__()
                self.global_step = 0
                self.trainer = None
                self._hp = _AttrDict()

            def save_hyperparameters(self, *args, ignore=(), **kw):
                import inspect

                frame = inspect.currentframe().f_back
                loc = frame.f_locals
                init = type(self).__init__
                names = [n for n in inspect.signature(init).parameters if n not in ("self", "kwargs")]
                for n in names:
                    if n in ignore or n not in loc:
                        continue
                    self._hp[n] = loc[n]

            @property
            def hparams(self):
                return self._hp

            @property
            def device(self):
                try:
                    return next(self.parameters()).device
                except StopIteration:
                    return torch.device("cpu")

            def log_dict(self, *a, **k):
                return None

            def log(self, *a, **k):
                return None

        class LightningDataModule:
            def __init__(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        pl.LightningDataModule = LightningDataModule
        pl.Trainer = object
        pl.seed_everything = lambda s, **k: torch.manual_seed(s)
        sys.modules["pytorch_lightning"] = pl
    if "webdataset" not in sys.modules:
        wds = types.ModuleType("webdataset")
        wds.RandomMix = object
        wds.WebDataset = object
        wds.warn_and_continue = None
        wds.split_by_node = None
        sys.modules["webdataset"] = wds


def load_reference():
    """Returns a namespace with the reference's hot-path classes, imported unmodified."""
    ref = find_reference()
    if ref is None:
        raise RuntimeError("reference checkout not found (set $WAVJEPA_REF)")
    _install_stubs()
    if ref not in sys.path:
        sys.path.insert(0, ref)
    # wavjepa/__init__.py imports the denoiser (-> data_modules -> webdataset/torchaudio); the
    # stubs above make that import succeed without touching the reference.
    try:
        import wavjepa  # noqa: F401
    except Exception:
        # fall back to a bare package so that submodules import without wavjepa/__init__.py
        pkg = types.ModuleType("wavjepa")
        pkg.__path__ = [os.path.join(ref, "wavjepa")]
        sys.modules["wavjepa"] = pkg
    from wavjepa.jepa import JEPA
    from wavjepa.extractors.audio_feature_extractor import ConvFeatureExtractor
    from wavjepa.extractors.audio_channel_feature_extractor import ConvChannelFeatureExtractor
    from wavjepa.masking import TimeInverseBlockMasker, SpeechMasker
    from wavjepa.audio_masking import compute_mask_indices
    from wavjepa.types.wavjepa_configs import TransformerEncoderCFG, TransformerLayerCFG
    from wavjepa.pos_embed import get_1d_sincos_pos_embed_from_grid

    ns = types.SimpleNamespace(
        root=ref,
        JEPA=JEPA,
        ConvFeatureExtractor=ConvFeatureExtractor,
        ConvChannelFeatureExtractor=ConvChannelFeatureExtractor,
        TimeInverseBlockMasker=TimeInverseBlockMasker,
        SpeechMasker=SpeechMasker,
        compute_mask_indices=compute_mask_indices,
        TransformerEncoderCFG=TransformerEncoderCFG,
        TransformerLayerCFG=TransformerLayerCFG,
        get_1d_sincos_pos_embed_from_grid=get_1d_sincos_pos_embed_from_grid,
    )
    return ns


BASE_SPEC = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)]


def build_reference_jepa(ref, conv_spec=None, d_model=768, nhead=12, layers=12, d_dec=384, dec_heads=12,
                         dec_layers=12, top_k=8, in_channels=1, seconds=2.01, sr=16000):
    """Builds the reference JEPA exactly the way train.py:113-128 / hear_api/runtime.py:52-61 do."""
    conv_spec = conv_spec or BASE_SPEC
    ext = ref.ConvFeatureExtractor(conv_layers_spec=conv_spec, in_channels=in_channels)
    model = ref.JEPA(
        feature_extractor=ext,
        transformer_encoder_cfg=ref.TransformerEncoderCFG.create(num_layers=layers),
        transformer_encoder_layers_cfg=ref.TransformerLayerCFG.create(d_model=d_model, nhead=nhead),
        transformer_decoder_cfg=ref.TransformerEncoderCFG.create(num_layers=dec_layers),
        transformer_decoder_layers_cfg=ref.TransformerLayerCFG.create(d_model=d_dec, nhead=dec_heads),
        resample_sr=sr,
        process_audio_seconds=seconds,
        nr_samples_per_audio=8,
        average_top_k_layers=top_k,
        compile_modules=False,
    )
    return model


class seeded_default_rng:
    """Context manager realising the seed contract of SURVEY.md 8a-M1 WITHOUT editing the reference:
    every `np.random.default_rng(None)` issued by compute_mask_indices (audio_masking.py:64) is replaced by
    default_rng([base_seed, global_row, attempt*8 + call_idx]).  The caller sets `.row`; `.call` counts calls
    within the row (a rejected attempt of the TimeInverse masker uses 5 calls, of the Speech masker 4).
    """

    def __init__(self, base_seed: int, calls_per_attempt: int, first_call_idx: int):
        self.base_seed = base_seed
        self.cpa = calls_per_attempt
        self.first = first_call_idx
        self.row = 0
        self.call = 0

    def __enter__(self):
        import numpy as np

        self._np = np
        self._orig = np.random.default_rng

        def patched(seed=None):
            if seed is not None:
                return self._orig(seed)
            attempt, idx = divmod(self.call, self.cpa)
            self.call += 1
            return self._orig([self.base_seed, self.row, attempt * 8 + self.first + idx])

        np.random.default_rng = patched
        return self

    def set_row(self, row: int):
        self.row = row
        self.call = 0

    def __exit__(self, *exc):
        self._np.random.default_rng = self._orig
        return False
