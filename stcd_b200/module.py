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

    #: the reference loader's normalisation (CD_Dataset.MEAN / STD, data/dataset.py:171-172)
    IMAGENET_MEAN = (0.485, 0.456, 0.406)
    IMAGENET_STD = (0.229, 0.224, 0.225)

    @torch.no_grad()
    def forward_uint8(self, a: torch.Tensor, b: torch.Tensor, mean=None, std=None):
        """The same forward fed with the decoded uint8 HWC images ``[B, H, W, 3]`` the reference's loader starts
        from (data/dataset.py:196-203): ToTensor + Normalize run inside the input-pack kernel, bit-identical to
        ``self(normalize(a), normalize(b))`` at a quarter of the host-to-device bytes.  Returns what ``forward``
        returns."""
        norm = (tuple(mean or self.IMAGENET_MEAN), tuple(std or self.IMAGENET_STD))
        outs = self.plan_for(a, u8_norm=norm).forward(a, b)
        return self._wrap_outputs(outs)

    def _wrap_outputs(self, outs):
        return outs[0] if len(outs) == 1 else tuple(outs)

    def plan_for(self, x: torch.Tensor, u8_norm=None):
        from .plan import Plan
        if self.training:
            raise RuntimeError("stcd_b200 implements the eval-mode inference path; call .eval() first "
                               "(training stays with the reference, models/trainer.py)")
        if not x.is_cuda:
            raise RuntimeError("stcd_b200 has no CPU path: move the module and its inputs to a B200 (cuda) device")
        chunk = max(1, min(int(self.chunk_pairs), int(x.shape[0])))
        h, w = (int(x.shape[1]), int(x.shape[2])) if u8_norm is not None else (int(x.shape[2]), int(x.shape[3]))
        key = (x.device.index, h, w, chunk, u8_norm)
        plan = self._plans.get(key)
        if plan is None:
            prog = self.lower(h, w)
            if u8_norm is not None:
                for op in prog.ops:
                    if isinstance(op, L.InputPackSpec):
                        op.u8_norm = u8_norm
            plan = Plan(prog, chunk, device=key[0])
            self._plans[key] = plan
        return plan
