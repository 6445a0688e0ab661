"""Minimal stand-in for lightning.pytorch (training harness is out of scope; see oracle/stubs/README.md)."""
import torch


class LightningModule(torch.nn.Module):
    def log(self, *args, **kwargs):
        pass


class LightningDataModule:
    def __init__(self, *args, **kwargs):
        pass


class Callback:
    pass


class Trainer:
    def __init__(self, *args, **kwargs):
        raise RuntimeError("lightning stub: Trainer is not available")


class _Callbacks:
    class ModelCheckpoint:
        def __init__(self, *args, **kwargs):
            pass


callbacks = _Callbacks()
