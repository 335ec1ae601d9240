"""net_G(x1, x2) for FC-Siam-diff / FC-Siam-conc on the GPU against the oracle, the emulator and
the golden fixtures generated from the unmodified reference (north-star tolerances: logits within
2e-2 absolute on the bf16 path, binary change maps agreeing on >= 99.9 % of decided pixels)."""
import os

import numpy as np
import pytest
import torch

import parity
from oracle import emulate, nets
from oracle.make_golden import CASES
from stcd_b200 import siamunet, synth

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2
NETS = {"diff": (siamunet.SiamUnet_diff, synth.GAINS["SiamUnet_diff"]), "conc": (siamunet.SiamUnet_conc, synth.GAINS["SiamUnet_conc"]),
        "sub": (siamunet.SiamUnet_sub, synth.GAINS["SiamUnet_sub"]), "cross": (siamunet.SiamUnet_cross_conc, synth.GAINS["SiamUnet_cross_conc"]),
        "ef": (siamunet.Unet, synth.GAINS["Unet"])}


def _net(fusion):
    cls, gain = NETS[fusion]
    return synth.randomize_(cls(3, 2).eval(), gain=gain)


def _agreement(y, ref):
    margin = (ref[:, 1] - ref[:, 0]).abs()
    agree = (y[:, 1] > y[:, 0]) == (ref[:, 1] > ref[:, 0])
    return agree.float().mean().item(), agree[margin > BF16_TOL].float().mean().item()


@pytest.mark.parametrize("fusion", ["diff", "conc", "sub", "cross", "ef"])
def test_forward_matches_oracle_and_emulator(fusion):
    net = _net(fusion)
    x1, x2 = synth.image_pairs(5, 64, 96)
    with torch.no_grad():
        ref = nets.siamunet_forward(net.state_dict(), x1, x2, fusion)
    emu = emulate.run_program(net.lower(64, 96), x1, x2, chunk=4)[0]
    net = net.cuda()
    net.chunk_pairs = 4                      # 5 pairs -> one full chunk + a ragged one
    y = net(x1.cuda(), x2.cuda())
    if fusion in ("sub", "cross"):          # these two return [x11d] (SiamUnet_sub.py:177-180)
        assert isinstance(y, list) and len(y) == 1
        y = y[0]
    assert isinstance(y, torch.Tensor) and y.shape == ref.shape and y.dtype == torch.float32
    y = y.cpu()
    assert (y - emu).abs().max().item() < 1.5e-2, "kernel vs emulator (same rounding points; bf16 flips cascade)"
    assert (y - ref).abs().max().item() < BF16_TOL, "kernel vs fp32 oracle"
    # absolute AND relative criteria, all-pixel and decided-pixel agreement (tests/parity.py); config C1's family additionally
    # with a logit spread >= 0.25 so that the absolute bound cannot be met by small logits
    r = parity.check(f"siamunet_{fusion}:5x64x96", y, ref, "argmax", min_std=0.25 if fusion == "diff" else 0.15)
    assert r["agree_decided"] >= 0.999 and r["agree_all"] >= 0.995


_FUSION = {"siamunet_diff": "diff", "siamunet_conc": "conc", "siamunet_sub": "sub", "siamunet_crossconc": "cross", "unet_ef": "ef"}


@pytest.mark.parametrize("case", sorted(_FUSION))
def test_forward_matches_golden(case, golden_dir):
    g = np.load(os.path.join(golden_dir, f"{case}.npz"))
    fusion = _FUSION[case]
    net = synth.randomize_(NETS[fusion][0](3, 2).eval(), seed=int(g["weight_seed"]), gain=float(g["gain"])).cuda()
    assert float(g["gain"]) == NETS[fusion][1]
    x1, x2 = synth.image_pairs(int(g["batch"]), int(g["h"]), int(g["w"]), seed=int(g["data_seed"]))
    y = net(x1.cuda(), x2.cuda())
    y = (y[-1] if isinstance(y, list) else y).cpu()
    ref = torch.from_numpy(g["out0"])
    assert (y - ref).abs().max().item() < BF16_TOL
    assert _agreement(y, ref)[1] >= 0.999
    parity.check(f"golden:{case}", y, ref, "argmax", min_std=0.15, all_px=0.99)


def test_full_size_config_c1_and_host_path():
    """Config C1: SiamUnet_diff, 256x256, batch 8 — device path, host-buffer path, and the
    size-independent properties: determinism, batch-order equivariance, identical pair -> the
    |f1-f2| skips vanish so swapping T1 for a copy of T2 equals running (T2, T2)."""
    net = _net("diff").cuda()
    x1, x2 = synth.image_pairs(8, 256, 256)
    with torch.no_grad():
        ref = nets.siamunet_forward(net.cpu().state_dict(), x1[:2], x2[:2], "diff")
    net = net.cuda()
    y = net(x1.cuda(), x2.cuda())
    assert (y[:2].cpu() - ref).abs().max().item() < BF16_TOL
    parity.check("siamunet_diff:C1 8x256x256 (pairs 0-1 vs oracle)", y[:2], ref, "argmax")
    y2 = net(x1.cuda(), x2.cuda())
    assert torch.equal(y, y2), "forward must be deterministic"
    perm = torch.tensor([3, 1, 7, 0, 2, 6, 5, 4])
    yp = net(x1[perm].cuda(), x2[perm].cuda())
    assert torch.equal(yp, y[perm.cuda()]), "pairs are independent: permuting the batch permutes the logits"
    plan = net.plan_for(x1.cuda())
    outs = plan.forward_host(x1.pin_memory(), x2.pin_memory())
    assert torch.equal(outs[0], y.cpu()), "host-buffer path must equal the device path bit for bit"
    assert plan.launches(8) == 25


def test_reference_calling_conventions():
    net = _net("diff").cuda()
    x1, x2 = synth.image_pairs(2, 32, 32)
    y = net(x1.cuda(), x2.cuda())
    sd = net.state_dict()
    net2 = siamunet.SiamUnet_diff(3, 2)
    net2.load_state_dict(sd)
    y2 = net2.cuda().eval()(x1.cuda(), x2.cuda())
    assert torch.equal(y, y2)
    wrapped = torch.nn.DataParallel(net2, device_ids=[0])     # train_stcd.py:639 wraps the net like this
    y3 = wrapped(x1.cuda(), x2.cuda())
    assert torch.equal(y, y3) and wrapped.module is net2
    with pytest.raises(ValueError):
        net(x1.cuda(), x2[:1].cuda())
    with pytest.raises(TypeError):
        net(x1.cuda().half(), x2.cuda().half())
