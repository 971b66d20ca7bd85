"""Module base class and conventions shared by the mirrored layer library.

Boundary contract reproduced from the reference's models/BaseModels.py:12-71:
lenient ``load_state_dict`` that copies by name and reports instead of raising (:41-52), the two weight
initialisers keyed on ``isinstance(m, nn.Conv2d / nn.BatchNorm2d)`` (:17-39), ``total_parameters`` (:64-68).
"""
import math
from contextlib import contextmanager

from torch import nn


class BaseModule(nn.Module):
    def __init__(self):
        self.act_fn = None
        super().__init__()

    # -- initialisers -----------------------------------------------------------------------------
    def _trainable(self, kinds):
        return (m for m in self.modules() if isinstance(m, kinds) and m.weight is not None and m.weight.requires_grad)

    def selu_init_params(self):
        for m in self._trainable((nn.Conv2d, nn.Linear)):
            m.weight.data.normal_(0.0, 1.0 / math.sqrt(m.weight.numel()))
            if m.bias is not None:
                m.bias.data.zero_()
        for m in self._trainable(nn.BatchNorm2d):
            m.weight.data.fill_(1)
            m.bias.data.zero_()

    def initialize_weights(self):
        for m in self._trainable(nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="leaky_relu")
            if m.bias is not None:
                m.bias.data.zero_()
        for m in self._trainable(nn.BatchNorm2d):
            m.weight.data.fill_(1)
            m.bias.data.zero_()

    # -- checkpoints ------------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True, self_state=False):
        """Copy-by-name; never raises (the reference prints and continues)."""
        own = self_state if self_state else self.state_dict()
        for name, value in state_dict.items():
            if name not in own:
                print("Parameter {} is not in the model. ".format(name))
                continue
            try:
                own[name].copy_(value.data if hasattr(value, "data") else value)
            except Exception as exc:  # noqa: BLE001 -- mirror of the reference's lenient behaviour
                print("Parameter {} fails to load.".format(name))
                print("-----------------------------------------")
                print(exc)

    @contextmanager
    def set_activation_inplace(self):
        act = getattr(self, "act_fn", None)
        if act is not None and hasattr(act, "inplace"):
            act.inplace = True
            try:
                yield
            finally:
                act.inplace = False
        else:
            yield

    def total_parameters(self):
        total = sum(p.numel() for p in self.parameters())
        trainable = sum(p.numel() for p in self.parameters() if p.requires_grad)
        print("Total parameters : {}. Trainable parameters : {}".format(total, trainable))
        return total

    def forward(self, *x):
        raise NotImplementedError
