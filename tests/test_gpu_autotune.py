"""The plan autotuner (stcd_plan_finalize: each conv op is timed with every admissible number of M sub-tiles per CTA pass and keeps
the fastest) must not change a single bit of the outputs: a sub-tile's MMAs and epilogue are the same sequence whatever the pass width."""
import os

import pytest
import torch

from stcd_b200 import siamunet, snunet, synth

pytestmark = pytest.mark.gpu


def _forward(make, name, pairs, h, w, chunk, autotune):
    old = os.environ.get("STCD_AUTOTUNE")
    os.environ["STCD_AUTOTUNE"] = "1" if autotune else "0"      # read when the plan is finalized (first forward)
    try:
        net = synth.prepare_(make().eval(), name).cuda()
        net.chunk_pairs = chunk
        x1, x2 = synth.image_pairs(pairs, h, w)
        y = net(x1.cuda(), x2.cuda())
        torch.cuda.synchronize()
        return [t.cpu() for t in (y if isinstance(y, (list, tuple)) else [y])]
    finally:
        if old is None:
            os.environ.pop("STCD_AUTOTUNE", None)
        else:
            os.environ["STCD_AUTOTUNE"] = old


@pytest.mark.parametrize("make,name,pairs,h,w,chunk", [
    (lambda: siamunet.SiamUnet_diff(3, 2), "SiamUnet_diff", 8, 128, 128, 8),
    (lambda: snunet.SNUNet_ECAM(3, 2), "SNUNet_ECAM", 8, 128, 128, 8),
])
def test_autotuned_plan_is_bit_identical(make, name, pairs, h, w, chunk):
    tuned = _forward(make, name, pairs, h, w, chunk, True)
    plain = _forward(make, name, pairs, h, w, chunk, False)
    assert len(tuned) == len(plain)
    for a, b in zip(tuned, plain):
        assert torch.equal(a, b), "the autotuner changed the logits"
