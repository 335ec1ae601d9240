"""The C-ABI library loads and exports every symbol include/stcd_b200.h declares; without a GPU
every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import re

import pytest
import torch

from stcd_b200 import _lib


def _declared_symbols():
    text = _lib.HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(stcd_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    lib = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 14
    bound = {n for n, _, _ in _lib.SYMBOLS}
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/stcd_b200.h but not exported"
        assert name in bound, f"{name} has no ctypes signature in stcd_b200/_lib.py"
    assert lib.stcd_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header():
    # sizes the C side static_asserts / validates; a drifted ctypes mirror corrupts plans silently
    assert C.sizeof(_lib.Chunk) == 16 and C.sizeof(_lib.Tap) == 4
    assert C.sizeof(_lib.Phase) == 24


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_gpu_fails_loudly():
    lib = _lib.lib()
    assert lib.stcd_device_count() == 0
    h = C.c_void_p()
    rc = lib.stcd_plan_create(0, 4, C.byref(h))
    assert rc == 3 and not h.value          # STCD_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.stcd_last_error()
    cm = (C.c_int64 * 4)()
    buf = (C.c_uint8 * 16)()
    rc = lib.stcd_confusion_add_batch(buf, 3, 0.0, buf, 1, 1, 16, 2, cm, None, None)
    assert rc == 3
    from stcd_b200.metric import SegmentationMetric
    with pytest.raises(_lib.StcdError):
        SegmentationMetric(2)


def test_argument_validation_without_gpu():
    lib = _lib.lib()
    cm = (C.c_int64 * 4)()
    buf = (C.c_uint8 * 16)()
    assert lib.stcd_confusion_add_batch(None, 3, 0.0, buf, 1, 1, 16, 2, cm, None, None) == 1
    assert lib.stcd_confusion_add_batch(buf, 9, 0.0, buf, 1, 1, 16, 2, cm, None, None) == 1
    assert lib.stcd_confusion_add_batch(buf, 0, 0.0, buf, 1, 1, 16, 3, cm, None, None) == 1
    assert lib.stcd_forward(None, None, None, 1, None, 0, None) == 4      # STCD_ERR_STATE


# every registry family at a small legal size: the library's own descriptor checks run on the lowered programs without a
# GPU (a validation plan, device -1), so a lowering change that the C side rejects fails HERE, not on the GPU box
_VALIDATE = [
    ("SiamUnet_diff", (3, 2), 64, 64), ("SiamUnet_conc", (3, 2), 64, 64), ("SiamUnet_sub", (3, 2), 32, 48),
    ("SiamUnet_cross_conc", (3, 2), 48, 32), ("Unet", (3, 2), 32, 32), ("SNUNet_ECAM", (3, 2), 64, 96),
    ("SegCD", ("resnet34",), 64, 96), ("CDNet_model", (3,), 64, 96), ("DSIFN", (), 64, 96),
    ("ChangeFormerV6", (3, 2, False, 256), 256, 256), ("ChangeGNNV1", (3, 2, False, 256), 256, 256),
    ("SiamUnet_diff", (3, 2), 256, 256), ("SNUNet_ECAM", (3, 2), 256, 256),
]
_VALIDATE_SPLIT = [("SiamUnet_diff", (3, 2), 64, 64), ("SiamUnet_conc", (3, 2), 64, 64), ("SNUNet_ECAM", (3, 2), 64, 96),
                   ("SiamUnet_diff", (3, 2), 256, 256), ("SNUNet_ECAM", (3, 2), 256, 256)]


@pytest.mark.parametrize("name,args,h,w,precision", [v + ("bf16",) for v in _VALIDATE] + [v + ("tf32",) for v in _VALIDATE_SPLIT],
                         ids=[f"{n}-{h}x{w}" for n, _, h, w in _VALIDATE] + [f"{n}-{h}x{w}-tf32" for n, _, h, w in _VALIDATE_SPLIT])
def test_lowered_programs_pass_the_library_checks(name, args, h, w, precision):
    from stcd_b200.networks import CLASSES
    from stcd_b200.plan import Plan
    net = CLASSES[name](*args).eval()
    net.precision = precision
    prog = net.lower(h, w)
    for chunk in (1, 4):
        plan = Plan(prog, chunk, device=-1)
        assert plan.launches(chunk) == 0 or plan.launches(chunk) > 0      # recorded, never finalized
        lib = _lib.lib()
        assert lib.stcd_plan_finalize(plan._h) == 3                        # STCD_ERR_NO_DEVICE: nothing can run
        assert b"no CPU fallback" in lib.stcd_last_error()
        plan.close()
