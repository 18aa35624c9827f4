"""ARCH benchmark adapter (second caller of the inference entry point), same surface as the reference's
ARCH/configs/wavjepa_wrapper.py:56-156: `WavJEPAModelWrapper(model, device, max_length)` with `get_embeddings(audio)`
(one clip -> one [D] vector), the size / sampling-rate getters, and the `arch_eval.Model` base class when that package is
importable (it is a vendored third-party harness).

get_embeddings (wavjepa_wrapper.py:67-110): view the clip as [1, 1, L], loudness-normalise to -14 dBFS, zero-pad to a
multiple of the 2.01 s unit (a whole extra chunk on exact multiples), normalise and encode every chunk with the padded
frames key-masked, keep the unmasked frames, mean over all of them.  Here the chunks of the clip go through the encoder
as one packed batch (wavjepa_b200.hear.embed_chunks); results are identical at every kept frame.
"""
from __future__ import annotations

import torch

from . import ops
from .hear import embed_chunks

try:  # pragma: no cover - the harness is not part of this repository
    from arch_eval import Model as _ArchModel
except Exception:  # noqa: BLE001
    class _ArchModel:  # minimal stand-in with the constructor signature of arch_eval.Model
        def __init__(self, model, **kwargs):
            self.model = model


class WavJEPAModelWrapper(_ArchModel):
    def __init__(self, model, device, max_length):
        super().__init__(model)
        self.model = model
        self.sr = 16000
        self.model.eval()
        self.device = device
        self.max_length = max_length
        self.unit_frames = model.target_length
        self.output_steps = model.extract_audio.total_patches(self.unit_frames)

    @torch.no_grad()
    def get_embeddings(self, audio, **kwargs) -> torch.Tensor:
        a = torch.as_tensor(audio).to(self.model.device, torch.float32).reshape(1, 1, -1).contiguous()
        gain = ops.clip_gain(a, -14.0)
        emb, _ = embed_chunks(self.model, a, gain, self.unit_frames, self.output_steps, self.sr)
        return emb[0].mean(dim=0)

    def get_sequence_embeddings(self, audio, **kwargs):
        # the reference implementation calls its own resample() with keyword arguments it does not accept
        # (wavjepa_wrapper.py:113-117) and raises TypeError before computing anything; there is no behaviour to mirror
        raise TypeError("get_sequence_embeddings is not functional in the reference (wavjepa_wrapper.py:113-117)")

    def get_classification_embedding_size(self):
        return self.model.encoder_embedding_dim

    def get_token_embedding_size(self):
        return self.model.encoder_embedding_dim

    def get_sampling_rate(self):
        return self.sr

    def get_embedding_layer(self):
        return self.model.encoder_embedding_dim
