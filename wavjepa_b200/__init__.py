"""wavjepa_b200 -- B200-native (sm_100a) implementation of the WavJEPA pre-training / feature-extraction hot path.

Same module API as the reference (labhamlet/wavjepa): `JEPA`, `ConvFeatureExtractor`, `ConvChannelFeatureExtractor`,
`TimeInverseBlockMasker`, `SpeechMasker`, `TransformerLayerCFG`, `TransformerEncoderCFG`; HEAR entry points live in
`wavjepa_b200.hear`, the denoiser stage (`Denoiser`, scene generation) in `wavjepa_b200.denoiser`.  All compute goes through libwavjepa_b200.so (include/wavjepa_b200.h); there is no fallback.
"""
from .types import ForwardReturn, TransformerEncoderCFG, TransformerLayerCFG  # noqa: F401
from .extractors import ConvChannelFeatureExtractor, ConvFeatureExtractor, Extractor  # noqa: F401
from .masking import SpeechMasker, TimeInverseBlockMasker  # noqa: F401
from .jepa import JEPA  # noqa: F401
from .denoiser import Denoiser  # noqa: F401

__all__ = ["JEPA", "Denoiser", "ConvFeatureExtractor", "ConvChannelFeatureExtractor", "Extractor", "TimeInverseBlockMasker",
           "SpeechMasker", "TransformerLayerCFG", "TransformerEncoderCFG", "ForwardReturn"]
