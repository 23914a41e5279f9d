"""Minimal stand-in for torchmetrics (test infrastructure; see ../README.md)."""
import torch
import torch.nn as nn


class Metric(nn.Module):
    def __init__(self, dist_sync_on_step=False, **kwargs):
        super().__init__()
        self._defaults = {}

    def add_state(self, name, default, dist_reduce_fx=None):
        self._defaults[name] = default.clone() if torch.is_tensor(default) else default
        self.register_buffer(name, default.clone() if torch.is_tensor(default) else default)

    def forward(self, *args, **kwargs):
        self.update(*args, **kwargs)
        return self.compute()

    def reset(self):
        for k, v in self._defaults.items():
            setattr(self, k, v.clone() if torch.is_tensor(v) else v)
