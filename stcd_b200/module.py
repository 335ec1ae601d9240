"""Shared plumbing of the drop-in ``nn.Module`` wrappers: plan caching and the eval-only contract.

A wrapper keeps the reference module's parameter names (so a reference ``state_dict`` loads) and
lowers itself to a libstcd_b200 plan per (device, H, W, chunk) on first use; anything that touches
the weights (``load_state_dict``, ``.to``, ``.cuda``, re-initialisation through
``networks.init_weights``) drops the packed copies.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from . import lowering as L


class PlannedModule(nn.Module):
    #: image pairs per pass through the layer stack (one plan workspace holds this many)
    default_chunk_pairs = 32

    def __init__(self):
        super().__init__()
        self._plans: Dict[tuple, object] = {}
        self.chunk_pairs = self.default_chunk_pairs

    # weights changed (load_state_dict / .to / re-init): packed copies are stale
    def _apply(self, fn, *a, **k):
        self._plans = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._plans = {}
        return super().load_state_dict(*a, **k)

    def invalidate_plans(self) -> None:
        self._plans = {}

    def lower(self, h: int, w: int) -> L.Program:  # pragma: no cover - abstract
        raise NotImplementedError

    def plan_for(self, x: torch.Tensor):
        from .plan import Plan
        if self.training:
            raise RuntimeError("stcd_b200 implements the eval-mode inference path; call .eval() first "
                               "(training stays with the reference, models/trainer.py)")
        if not x.is_cuda:
            raise RuntimeError("stcd_b200 has no CPU path: move the module and its inputs to a B200 (cuda) device")
        chunk = max(1, min(int(self.chunk_pairs), int(x.shape[0])))
        key = (x.device.index, int(x.shape[2]), int(x.shape[3]), chunk)
        plan = self._plans.get(key)
        if plan is None:
            plan = Plan(self.lower(key[1], key[2]), chunk, device=key[0])
            self._plans[key] = plan
        return plan
