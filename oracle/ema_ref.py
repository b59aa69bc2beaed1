"""ORACLE (test infrastructure): restatement of ``torch_ema.ExponentialMovingAverage``.

The reference imports it (``nerf/utils.py:30``; used at :616, :1627, :1862, :1684-1695, :2056-2057, :2130-2132) from the PyPI
package ``torch-ema`` (``requirements.txt``: un-pinned, not vendored, not installed here).  This module restates the published
algorithm of torch-ema 0.3 — the surface the reference's Trainer touches: ``update`` / ``store`` / ``copy_to`` / ``restore`` /
``state_dict`` / ``load_state_dict`` — so that the reference's own ``Trainer`` can run in the drop-in test.  Parity for this
third-party arithmetic is therefore pinned to the published formula, not to the package (stated in DESIGN.md):

    decay_k = min(decay, (1 + k) / (10 + k))        k = number of updates so far, counting this one (use_num_updates=True)
    shadow  = shadow - (1 - decay_k) * (shadow - param)
"""
from __future__ import annotations

import torch


class ExponentialMovingAverage:
    def __init__(self, parameters, decay, use_num_updates=True):
        if decay < 0.0 or decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        parameters = list(parameters)
        self.shadow_params = [p.clone().detach() for p in parameters]
        self.collected_params = None
        self._params = parameters

    def _get(self, parameters):
        return self._params if parameters is None else list(parameters)

    def update(self, parameters=None):
        parameters = self._get(parameters)
        decay = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
        one_minus_decay = 1.0 - decay
        with torch.no_grad():
            for s_param, param in zip(self.shadow_params, parameters):
                tmp = s_param - param
                tmp.mul_(one_minus_decay)
                s_param.sub_(tmp)

    def copy_to(self, parameters=None):
        for s_param, param in zip(self.shadow_params, self._get(parameters)):
            param.data.copy_(s_param.data)

    def store(self, parameters=None):
        self.collected_params = [p.clone() for p in self._get(parameters)]

    def restore(self, parameters=None):
        if self.collected_params is None:
            raise RuntimeError("This ExponentialMovingAverage has no `store()`ed weights to `restore()`")
        for c_param, param in zip(self.collected_params, self._get(parameters)):
            param.data.copy_(c_param.data)

    def state_dict(self):
        return {"decay": self.decay, "num_updates": self.num_updates, "shadow_params": self.shadow_params,
                "collected_params": self.collected_params}

    def load_state_dict(self, state_dict):
        self.decay = state_dict["decay"]
        self.num_updates = state_dict["num_updates"]
        self.shadow_params = [p.to(s.device) for p, s in zip(state_dict["shadow_params"], self.shadow_params)]
        self.collected_params = state_dict["collected_params"]
