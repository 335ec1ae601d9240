"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

Run in the build container (the reference cannot travel to the GPU box):
    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden
Each fixture stores the recipe (model, seeds, shapes, gain) and the reference's fp32 logits;
weights and inputs are re-derived from the seeds by stcd_b200/synth.py, so fixtures stay small.
TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from oracle import refimport
from stcd_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (reference module path, class (= harness family in synth.GAINS), ctor args, batch, H, W)
CASES = {
    "siamunet_diff": ("models.SiamUnet_diff", "SiamUnet_diff", (3, 2), 2, 48, 32),
    "siamunet_conc": ("models.SiamUnet_conc", "SiamUnet_conc", (3, 2), 2, 32, 48),
    "siamunet_sub": ("models.SiamUnet_sub", "SiamUnet_sub", (3, 2), 2, 32, 48),
    "siamunet_crossconc": ("models.SiamUnet_crossconc", "SiamUnet_cross_conc", (3, 2), 2, 48, 32),
    "unet_ef": ("models.Unet", "Unet", (3, 2), 2, 32, 32),
    "snunet": ("models.SNUNet", "SNUNet_ECAM", (3, 2), 2, 32, 48),
    # smp.SegCD("resnet34", encoder_weights=None, classes=1): what train_stcd.py:637 instantiates (ResNet-34 row of C3)
    "segcd_r34": ("segmentation_models_pytorch", "SegCD", ("resnet34", 5, None), 2, 64, 96),
    # smp.SegCD("resnet50"): the encoder train_stcd.py:638 selects (Bottleneck blocks)
    "segcd_r50": ("segmentation_models_pytorch", "SegCD", ("resnet50", 5, None), 1, 64, 64),
    # ChangeGNNV1 runs only at its img_size (pos_embed is not resized); gcn_lib comes from oracle/gcn_lib_restated.py
    # (absent upstream dependency: this fixture pins everything EXCEPT the Grapher restatement itself)
    "changegnn_v1": ("models.ChangeVIG", "ChangeGNNV1", (3, 2, False, 256), 1, 256, 256),
    "changeformer_v6": ("models.ChangeFormer", "ChangeFormerV6", (3, 2, False, 256), 1, 256, 256),
    # CDNet34 = CDNet_model(in_channels, SEBasicBlock, [3, 4, 6, 3], num_classes) (models/DTCDSCN.py:316-320): the defaults
    "dtcdscn": ("models.DTCDSCN", "CDNet_model", (3,), 2, 64, 96),
    # BIT: BASE_Transformer(input_nc, output_nc, with_pos, resnet_stages_num, token_len, token_trans, enc_depth, dec_depth) =
    # registry key base_transformer_pos_s4_dd8 (models/networks.py:177-179); ResNet(3, 2) = base_resnet18 (:170-171).
    # The backbone's checkpoint download is skipped (refimport.disable_pretrained_download); weights are the harness' draw.
    "bit_dd8": ("models.networks", "BASE_Transformer", (3, 2, "learned", 4, 4, True, 1, 8), 2, 64, 96),
    "bit_resnet18": ("models.networks", "ResNet", (3, 2), 2, 64, 96),
    # IFNet = DSIFN(base, base) with ONE shared vgg16_base (models/networks.py:164-166); args = () -> built by _build_ifnet
    "ifnet": ("models.DSIFN", "DSIFN", (), 2, 64, 96),
    # ChangeGNNV2 (Cross_ConCat HFFM) and the "sub" Compare variant; gcn_lib from oracle/gcn_lib_restated.py as for changegnn_v1
    "changegnn_v2": ("models.ChangeVIG", "ChangeGNNV2", (3, 2, False, 256), 1, 256, 256),
    "changeformer_v1": ("models.ChangeFormer", "ChangeFormerV1", (), 1, 256, 256),
    "changeformer_v2": ("models.ChangeFormer", "ChangeFormerV2", (), 1, 256, 256),
    "changeformer_v3": ("models.ChangeFormer", "ChangeFormerV3", (), 1, 256, 256),
    "vig_v20": ("models.ChangeVIG", "VIG_V20_2", (3, 2, False, 256), 1, 256, 256),          # registry key "GNN"
    "changegnn_v2_sub": ("models.ChangeVIG", "ChangeGNNV2_Compare", (3, 2, False, 256, "MLP", 256, "sub"), 1, 256, 256),
}


def reference_net(case: str):
    mod, cls, args, *_ = CASES[case]
    if mod in ("models.networks", "models.DSIFN"):
        refimport.disable_pretrained_download()
    if cls == "DSIFN":
        m = refimport.ref_module(mod)
        base = m.vgg16_base()
        return synth.prepare_(m.DSIFN(base, base).eval(), cls)
    net = getattr(refimport.ref_module(mod), cls)(*args).eval()
    return synth.prepare_(net, cls)


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])            # optional: regenerate the named cases only
    for case, (mod, cls, args, b, h, w) in CASES.items():
        if only and case not in only:
            continue
        net = reference_net(case)
        x1, x2 = synth.image_pairs(b, h, w)
        with torch.no_grad():
            y = net(x1, x2)
        ys = y if isinstance(y, (list, tuple)) else [y]
        arrays = {f"out{i}": t.numpy() for i, t in enumerate(ys)}
        np.savez_compressed(os.path.join(OUT, f"{case}.npz"), gain=synth.GAINS[cls], batch=b, h=h, w=w,
                            weight_seed=synth.WEIGHT_SEED, data_seed=synth.DATA_SEED,
                            x1_sum=float(x1.double().sum()), n_out=len(ys), **arrays)
        print(case, [tuple(t.shape) for t in ys], "std", float(ys[-1].std()))
    if only and "segmentation_metric" not in only:
        return
    # evaluator: the reference's SegmentationMetric on seeded predictions/labels
    SM = refimport.segmentation_metric_class()
    g = torch.Generator().manual_seed(7)
    pred = (torch.rand(3, 1, 40, 56, generator=g) < 0.3).int()
    label = (torch.rand(3, 1, 40, 56, generator=g) < 0.1).long()
    m = SM(2)
    m.addBatch(pred, label)
    pred2 = (torch.rand(3, 1, 40, 56, generator=g) < 0.6).int()
    m.addBatch(pred2, label)
    np.savez_compressed(os.path.join(OUT, "segmentation_metric.npz"), pred=pred.numpy().astype(np.uint8), pred2=pred2.numpy().astype(np.uint8),
                        label=label.numpy().astype(np.uint8), cm=m.confusionMatrix.numpy(),
                        f1=m.F1score().numpy(), iou=m.IntersectionOverUnion().numpy(),
                        oa=m.OverallAccuracy().numpy(), precision=m.Precision().numpy(), recall=m.Recall().numpy(),
                        fwiou=m.Frequency_Weighted_Intersection_over_Union().numpy())
    print("segmentation_metric", m.confusionMatrix.tolist())


if __name__ == "__main__":
    main()
