"""Base class of `JEPA` / `Denoiser`: `pytorch_lightning.LightningModule` when Lightning is importable -- the reference's
models are LightningModules driven by `pl.Trainer` (wavjepa/jepa.py:72, train.py:164-180,244; wavjepa/denoiser.py:43,
denoise.py:112-140) -- otherwise a small nn.Module with the handful of LightningModule members the reference's code
paths touch (`save_hyperparameters`, `hparams`, `log`, `log_dict`, `trainer`), so that the same class body works in both
worlds and neither needs Lightning to run the fused `train_step` path.
"""
from __future__ import annotations

from torch import nn


class _AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _find_lightning():
    for name in ("pytorch_lightning", "lightning.pytorch"):
        try:
            mod = __import__(name, fromlist=["LightningModule"])
            return mod.LightningModule
        except Exception:  # noqa: BLE001  (not installed / broken install: fall back to the shim)
            continue
    return None


_PL = _find_lightning()
HAVE_LIGHTNING = _PL is not None


class _ShimModule(nn.Module):
    """The LightningModule surface used by the reference's hot path, without Lightning."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self._hparams = _AttrDict()
        self._trainer = None

    def save_hyperparameters(self, *args, **kwargs):
        for a in args:
            if isinstance(a, dict):
                self._hparams.update(a)
        self._hparams.update(kwargs)

    @property
    def hparams(self):
        return self._hparams

    @property
    def trainer(self):
        return self._trainer

    @trainer.setter
    def trainer(self, t):
        self._trainer = t

    def log(self, *args, **kwargs):
        return None

    def log_dict(self, *args, **kwargs):
        return None


Base = _PL if HAVE_LIGHTNING else _ShimModule


def attached_trainer(module):
    """The pl.Trainer driving `module`, or None (Lightning's own `.trainer` property raises when detached)."""
    return getattr(module, "_trainer", None)
