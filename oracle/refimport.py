"""Make the real reference importable from /root/reference (build container only).  TEST INFRASTRUCTURE.

The GPU box has no /root/reference: everything that runs there uses oracle/nets.py and the
committed fixtures instead.  Recipe: SURVEY.md App. F.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("STCD_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


class _Permissive:
    """Stands in for any symbol of an absent third-party package (timm, pretrainedmodels, ...)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Permissive()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Permissive()

    def __getitem__(self, k):
        return self.__dict__.setdefault("_items", {}).setdefault(k, _Permissive())

    def __setitem__(self, k, v):
        self.__dict__.setdefault("_items", {})[k] = v

    def __deepcopy__(self, memo):
        return _Permissive()

    def __iter__(self):
        return iter(())

    def __contains__(self, k):
        return False

    def __mro_entries__(self, bases):
        return (object,)

    def items(self):
        return []

    def keys(self):
        return []

    def copy(self):
        return {}


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        if name[:1].isupper():
            return type(name, (), {"__init__": lambda self, *a, **k: None})
        return _Permissive()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    PREFIXES = ("timm", "pretrainedmodels", "efficientnet_pytorch")

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.PREFIXES:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        import torch.nn as nn
        if module.__name__ == "timm.models.layers":
            class DropPath(nn.Module):           # identity in eval mode (ChangeFormer.py:11)
                def __init__(self, drop_prob=0.0):
                    super().__init__()

                def forward(self, x):
                    return x
            module.DropPath = DropPath
            module.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
            module.trunc_normal_ = nn.init.trunc_normal_
        if module.__name__ == "timm.models.registry":
            module.register_model = lambda f: f


_installed = False


def install() -> None:
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    sys.dont_write_bytecode = True          # /root/reference is read-only
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    have_real = []
    for pfx in _StubFinder.PREFIXES:
        try:
            importlib.import_module(pfx)
            have_real.append(pfx)
        except Exception:
            pass
    finder = _StubFinder()
    finder.PREFIXES = tuple(p for p in _StubFinder.PREFIXES if p not in have_real)
    sys.meta_path.insert(0, finder)
    if "gcn_lib" not in sys.modules:      # absent from the reference tree: the restated stand-in (parity unpinned)
        from oracle import gcn_lib_restated
        sys.modules["gcn_lib"] = gcn_lib_restated
    _installed = True


def ref_module(path: str):
    """e.g. ref_module('models.SiamUnet_diff')"""
    install()
    return importlib.import_module(path)


def disable_pretrained_download() -> None:
    """BIT builds its backbone with ``models.resnet18(pretrained=True, ...)`` (models/networks.py:234-235): a checkpoint
    download.  There is no network here and the harness re-draws every weight anyway (stcd_b200/synth.py), so the one
    line that fetches and loads the checkpoint is skipped; the module tree, names and forward are untouched."""
    r = ref_module("models.resnet")
    if getattr(r, "_stcd_patched", False):
        return
    orig = r._resnet
    r._resnet = lambda arch, block, layers, pretrained, progress, **kw: orig(arch, block, layers, False, progress, **kw)
    r._stcd_patched = True
    # IFNet: ``vgg16(pretrained=True)`` from torchvision (models/DSIFN.py:12): same treatment
    d = ref_module("models.DSIFN")
    import torchvision
    d.vgg16 = lambda pretrained=True: torchvision.models.vgg16(weights=None)


def segmentation_metric_class():
    """exec the source span of SegmentationMetric (train_stcd.py:515-593) without running the script."""
    import torch
    import torch.nn as nn
    src = open(os.path.join(REF_ROOT, "train_stcd.py"), encoding="utf-8").read().splitlines()
    start = next(i for i, l in enumerate(src) if l.startswith("class SegmentationMetric"))
    end = next(i for i in range(start + 1, len(src)) if src[i].startswith("class "))
    ns = {"torch": torch, "nn": nn}
    exec("\n".join(src[start:end]), ns)   # noqa: S102 - reference source, read-only tree
    return ns["SegmentationMetric"]
