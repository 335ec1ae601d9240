"""K9 (confusion-matrix histogram + fused binarisation) against the oracle: bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import metric as ometric
from stcd_b200.metric import SegmentationMetric

pytestmark = pytest.mark.gpu


def test_matches_golden_from_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "segmentation_metric.npz"))
    m = SegmentationMetric(2, "cuda:0")
    label = torch.from_numpy(g["label"]).long()
    m.addBatch(torch.from_numpy(g["pred"]).int(), label)          # the reference's dtypes: int32 / int64
    m.addBatch(torch.from_numpy(g["pred2"]).int(), label)
    assert m.confusionMatrix.dtype == torch.float64
    assert np.array_equal(m.confusionMatrix.numpy(), g["cm"])
    np.testing.assert_array_equal(m.F1score().numpy(), g["f1"])
    np.testing.assert_array_equal(m.IntersectionOverUnion().numpy(), g["iou"])
    np.testing.assert_array_equal(m.OverallAccuracy().numpy(), g["oa"])
    np.testing.assert_array_equal(m.Precision().numpy(), g["precision"])
    np.testing.assert_array_equal(m.Recall().numpy(), g["recall"])
    np.testing.assert_allclose(m.Frequency_Weighted_Intersection_over_Union().numpy(), g["fwiou"], rtol=1e-15)


@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (1, 1, 7, 5), (3, 1, 33, 31), (2, 1, 64, 64), (4, 1, 256, 256)])
@pytest.mark.parametrize("pdt,ldt", [(torch.int32, torch.int64), (torch.uint8, torch.uint8), (torch.int64, torch.int32),
                                     (torch.bool, torch.int64)])
def test_add_batch_all_dtypes_and_ragged_sizes(shape, pdt, ldt):
    g = torch.Generator().manual_seed(sum(shape))
    for p in (0.0, 0.37, 1.0):
        pred = (torch.rand(shape, generator=g) < p)
        label = (torch.rand(shape, generator=g) < 0.2)
        m = SegmentationMetric(2)
        m.addBatch(pred.to(pdt).cuda(), label.to(ldt).cuda())
        want = ometric.confusion_matrix(pred.numpy(), label.numpy())
        assert np.array_equal(m.confusion_counts().cpu().numpy(), want)
        assert int(m.confusion_counts().sum()) == pred.numel()


def test_unaligned_views_take_the_scalar_path():
    g = torch.Generator().manual_seed(5)
    pred = (torch.rand(4099, generator=g) < 0.5).to(torch.uint8).cuda()
    label = (torch.rand(4099, generator=g) < 0.5).to(torch.uint8).cuda()
    m = SegmentationMetric(2)
    m.addBatch(pred[3:].clone()[1:], label[3:].clone()[1:])       # data_ptr not 16-byte aligned
    want = ometric.confusion_matrix(pred[4:].cpu().numpy(), label[4:].cpu().numpy())
    assert np.array_equal(m.confusion_counts().cpu().numpy(), want)


@pytest.mark.parametrize("kind,c", [("argmax", 2), ("sigmoid", 1), ("raw_ge", 1)])
def test_fused_binarise_matches_reference_expressions(kind, c):
    g = torch.Generator().manual_seed(11)
    logits = torch.randn(3, c, 48, 80, generator=g)
    # adversarial values: exact ties, the fp32 sigmoid plateau around 0, +-inf, thresholds
    special = torch.tensor([0.0, 1e-8, 5.9e-8, 6e-8, 1.2e-7, 1e-6, -1e-8, 0.5, 0.49999997, float("inf"), -float("inf")])
    logits.view(-1)[: special.numel()] = special
    if c == 2:
        logits[:, 1, :4] = logits[:, 0, :4]                        # argmax ties -> class 0
    label = (torch.rand(3, 48, 80, generator=g) < 0.1).long()
    pred_out = torch.empty(3, 48, 80, dtype=torch.uint8, device="cuda")
    m = SegmentationMetric(2)
    m.addLogits(logits.cuda(), label.cuda(), kind=kind, thr=0.5, pred_out=pred_out)
    want_pred = ometric.binarise(logits.numpy(), kind, 0.5)
    assert np.array_equal(pred_out.cpu().numpy(), want_pred)
    assert np.array_equal(m.confusion_counts().cpu().numpy(), ometric.confusion_matrix(want_pred, label.numpy()))
    m2 = SegmentationMetric(2)
    m2.addLogits(logits.cuda(), label.cuda(), kind=kind, thr=0.7)
    want2 = ometric.binarise(logits.numpy(), kind, 0.7)            # train_pse_cd.py:145 uses 0.7
    assert np.array_equal(m2.confusion_counts().cpu().numpy(), ometric.confusion_matrix(want2, label.numpy()))


def test_multiclass_and_accumulation_properties():
    g = torch.Generator().manual_seed(13)
    K = 7
    pred = torch.randint(0, K, (2, 1, 100, 77), generator=g)
    label = torch.randint(0, K, (2, 1, 100, 77), generator=g)
    m = SegmentationMetric(K)
    m.addBatch(pred.cuda(), label.cuda())
    assert np.array_equal(m.confusion_counts().cpu().numpy(), ometric.confusion_matrix(pred.numpy(), label.numpy(), K))
    # linearity: adding the batch in two halves == adding it at once; reset() clears
    m2 = SegmentationMetric(K)
    m2.addBatch(pred[:1].cuda(), label[:1].cuda())
    m2.addBatch(pred[1:].cuda(), label[1:].cuda())
    assert torch.equal(m.confusion_counts(), m2.confusion_counts())
    m2.reset()
    assert int(m2.confusion_counts().sum()) == 0


def test_full_size_checksum_c3():
    """Config C3's evaluator step: 16 x 1024^2 pixels; counts must sum to the pixel count and
    match numpy."""
    g = torch.Generator().manual_seed(17)
    logits = torch.randn(16, 1, 1024, 1024, generator=g)
    label = (torch.rand(16, 1024, 1024, generator=g) < 0.05).to(torch.uint8)
    m = SegmentationMetric(2)
    m.addLogits(logits.cuda(), label.cuda(), kind="sigmoid")
    cm = m.confusion_counts().cpu().numpy()
    assert cm.sum() == 16 * 1024 * 1024
    assert np.array_equal(cm, ometric.confusion_matrix(ometric.binarise(logits.numpy(), "sigmoid"), label.numpy()))
    with pytest.raises(AssertionError):
        m.addBatch(torch.zeros(2, 3, dtype=torch.int32).cuda(), torch.zeros(3, 2, dtype=torch.int64).cuda())


def test_out_of_range_classes_raise_like_the_reference_bincount():
    """train_stcd.py:572-579: bincount(...).reshape(numClass, numClass) fails when a class index is outside [0, numClass);
    the kernel skips such pixels, so every read of the matrix must raise instead of scoring a reduced pixel set."""
    pred = torch.zeros(2, 1, 16, 16, dtype=torch.int32)
    label = torch.zeros(2, 1, 16, 16, dtype=torch.int64)
    label[0, 0, 3, 4] = 255                                       # a raw {0, 255} mask passed as class indices
    with pytest.raises((RuntimeError, ValueError)):               # the reference itself
        ometric.confusion_matrix(pred.numpy(), label.numpy())
    m = SegmentationMetric(2)
    m.addBatch(pred.cuda(), label.cuda())
    with pytest.raises(ValueError, match="outside"):
        m.confusionMatrix
    with pytest.raises(ValueError):
        m.F1score()
    m.reset()
    pred[1, 0, 0, 0] = 7                                          # an out-of-range prediction
    label[0, 0, 3, 4] = 1
    m.addBatch(pred.cuda(), label.cuda())
    with pytest.raises(ValueError):
        m.IntersectionOverUnion()
    m.reset()
    label[0, 0, 3, 4] = -1
    m.addBatch(torch.zeros_like(pred).cuda(), label.cuda())
    with pytest.raises(ValueError):
        m.confusionMatrix
    # the documented way to feed raw masks still works and counts every pixel
    m.reset()
    raw = torch.zeros(2, 16, 16, dtype=torch.uint8)
    raw[0, 2, 2] = 255
    m.addBatch(raw.cuda(), raw.cuda(), raw_masks=True)
    assert m.confusionMatrix.sum() == raw.numel()


def test_gen_confusion_matrix_is_side_effect_free():
    g = torch.Generator().manual_seed(3)
    pred = (torch.rand(2, 1, 20, 20, generator=g) < 0.4).int()
    label = (torch.rand(2, 1, 20, 20, generator=g) < 0.3).long()
    m = SegmentationMetric(2)
    cm = m.genConfusionMatrix(pred.cuda(), label.cuda())
    assert np.array_equal(cm.numpy(), ometric.confusion_matrix(pred.numpy(), label.numpy()))
    assert int(m.confusion_counts().sum()) == 0


def test_confuse_matrix_meter_matches_restatement():
    """models/evaluator.py:99-122,150-167 (ConfuseMatrixMeter; parity unpinned: misc/metric_tool.py is absent upstream)."""
    from stcd_b200.metric import ConfuseMatrixMeter
    g = np.random.default_rng(9)
    meter = ConfuseMatrixMeter(n_class=2)
    tot_p, tot_l = [], []
    for i in range(3):
        pr = (g.random((2, 40, 56)) < 0.3).astype(np.int64)
        gt = (g.random((2, 40, 56)) < 0.2).astype(np.int64)
        if i == 1:
            gt[0, :4] = 255                                       # ignore label: masked by upstream's helper
        mf1 = meter.update_cm(pr=pr, gt=gt)                       # numpy in, as the reference calls it
        want = ometric.confuse_matrix_meter_scores(pr, gt)
        assert abs(mf1 - want["mf1"]) < 1e-12
        tot_p.append(pr)
        tot_l.append(gt)
    got = meter.get_scores()
    want = ometric.confuse_matrix_meter_scores(np.concatenate(tot_p), np.concatenate(tot_l))
    assert set(got) == set(want) == {"acc", "miou", "mf1", "iou_0", "iou_1", "F1_0", "F1_1", "precision_0", "precision_1",
                                     "recall_0", "recall_1"}
    for k in want:
        assert abs(got[k] - want[k]) < 1e-12, k
    # device tensors in: same numbers, no host round trip for the inputs
    meter.clear()
    meter.update_cm(pr=torch.from_numpy(tot_p[0]).cuda(), gt=torch.from_numpy(tot_l[0]).cuda())
    w0 = ometric.confuse_matrix_meter_scores(tot_p[0], tot_l[0])
    assert abs(meter.get_scores()["miou"] - w0["miou"]) < 1e-12
    with pytest.raises(ValueError):
        meter.update_cm(pr=np.full((1, 4, 4), 3), gt=np.zeros((1, 4, 4), np.int64))
