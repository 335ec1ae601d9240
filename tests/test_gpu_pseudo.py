"""Pseudo-label masks and the reliability score against the reference's own expressions
(train_stcd.py:104-123,176-196) evaluated with torch / the oracle metric on the CPU: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import metric as ometric
from stcd_b200 import pseudo

pytestmark = pytest.mark.gpu


def test_change_mask_is_the_reference_expression():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 1, 37, 53, generator=g) * 2
    x[0, 0, 0, :4] = torch.tensor([0.0, 1e-8, 1e-6, -0.0])            # sigmoid(x) > 0.5 is not x > 0 in fp32
    for thr in (0.5, 0.7):
        want = (torch.sigmoid(x) > thr).int()
        want[want == 1] = 255                                          # train_stcd.py:185
        got = pseudo.change_mask(x.cuda(), thr=thr).cpu()
        assert got.dtype == torch.uint8 and torch.equal(got, want[:, 0].to(torch.uint8))
    two = torch.randn(2, 2, 16, 24, generator=g)
    assert torch.equal(pseudo.change_mask(two.cuda(), kind="argmax", on_value=1).cpu(), two.argmax(1).to(torch.uint8))
    with pytest.raises(RuntimeError):
        pseudo.change_mask(x)


@pytest.mark.parametrize("cumulative", [True, False])
def test_reliability_matches_reference_loop(cumulative):
    g = torch.Generator().manual_seed(8)
    scorer = pseudo.ReliabilityScorer("cuda:0", cumulative=cumulative)
    cm = np.zeros((2, 2), np.int64)
    for _ in range(3):                                                 # three "images"
        base = torch.rand(1, 64, 64, generator=g) < 0.2
        masks = [((base ^ (torch.rand(1, 64, 64, generator=g) < p)).to(torch.uint8) * 255) for p in (0.10, 0.05, 0.0)]
        got = scorer.score([m.cuda() for m in masks])
        ious = []
        for m in masks[:-1]:
            if not cumulative:
                cm[:] = 0
            cm += ometric.confusion_matrix((m >= 1).numpy(), (masks[-1] >= 1).numpy())      # addBatch(preds[i], preds[-1])
            ious.append(ometric.scores(cm)["IoU"][1])
        assert got == pytest.approx(sum(ious) / len(ious), abs=0, rel=0)
