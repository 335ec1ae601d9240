"""Host-side lowering (stcd_b200/lowering.py, siamunet.py) checked on the CPU: the emulator runs
the fused-op Program with the kernel's layouts and bf16 rounding points and must reproduce the
oracle within the bf16 tolerance the north star states (2e-2 absolute on the logits)."""
import numpy as np
import pytest
import torch

from oracle import emulate, nets
from stcd_b200 import lowering as L
from stcd_b200 import siamunet, synth

BF16_TOL = 2e-2


@pytest.mark.parametrize("fusion,cls,gain", [("diff", siamunet.SiamUnet_diff, synth.GAINS["SiamUnet_diff"]), ("conc", siamunet.SiamUnet_conc, synth.GAINS["SiamUnet_conc"])])
def test_siamunet_program_matches_oracle(fusion, cls, gain):
    net = synth.randomize_(cls(3, 2).eval(), gain=gain)
    x1, x2 = synth.image_pairs(3, 32, 48)
    with torch.no_grad():
        y = nets.siamunet_forward(net.state_dict(), x1, x2, fusion)
    prog = net.lower(32, 48)
    ye = emulate.run_program(prog, x1, x2, chunk=2)[0]      # 3 pairs in chunks of 2: ragged last chunk
    assert ye.shape == y.shape
    err = (ye - y).abs().max().item()
    assert err < BF16_TOL, err
    margin = (y[:, 1] - y[:, 0]).abs()
    agree = ((ye[:, 1] > ye[:, 0]) == (y[:, 1] > y[:, 0]))
    assert agree[margin > BF16_TOL].float().mean().item() >= 0.999
    assert agree.float().mean().item() >= 0.99
    assert 0.02 < (y[:, 1] > y[:, 0]).float().mean().item() < 0.98, "degenerate change map: parity would say nothing"


def test_snunet_program_matches_oracle():
    from stcd_b200 import snunet
    net = synth.prepare_(snunet.SNUNet_ECAM(3, 2).eval(), "SNUNet_ECAM")
    x1, x2 = synth.image_pairs(3, 32, 48)
    with torch.no_grad():
        y = nets.snunet_forward(net.state_dict(), x1, x2)
    prog = net.lower(32, 48)
    ye = emulate.run_program(prog, x1, x2, chunk=2)[0]
    assert ye.shape == y.shape
    assert (ye - y).abs().max().item() < BF16_TOL
    margin = (y[:, 1] - y[:, 0]).abs()
    agree = ((ye[:, 1] > ye[:, 0]) == (y[:, 1] > y[:, 0]))
    assert agree[margin > BF16_TOL].float().mean().item() >= 0.999
    assert 0.02 < (y[:, 1] > y[:, 0]).float().mean().item() < 0.98, "degenerate change map"
    # structure: 15 nested blocks x 2 convs + 10 up-convs + pack + fused ECAM head
    convs = [o for o in prog.ops if isinstance(o, L.ConvSpec)]
    assert len(convs) == 40 and len(prog.ops) == 42
    big = net.lower(256, 256)
    # 46.603 GMAC per pair at 256x256 (conv 43.92 + convT 2.68; + 1x1 head)
    assert abs(big.macs_per_pair() / 1e9 - 46.603) < 0.02
    assert max(len(o.srcs) for o in convs) <= L.MAX_SRC


def test_program_structure_and_macs():
    net = siamunet.SiamUnet_diff(3, 2).eval()
    prog = net.lower(256, 256)
    convs = [o for o in prog.ops if isinstance(o, L.ConvSpec)]
    assert len(convs) == 24 and len(prog.ops) == 25
    # 4.228 GMAC per pair for SiamUnet_diff at 256x256
    assert abs(prog.macs_per_pair() / 1e9 - 4.228) < 0.001
    up = [o for o in convs if o.name.startswith("upconv")]
    assert all(len(o.phases) == 4 and o.osy == 2 and o.osx == 2 for o in up)
    taps = sorted(ph.n_blocks * o.kc // 128 for o in up[:1] for ph in o.phases)
    assert taps == [1, 2, 2, 4]          # stride-2 ConvTranspose2d(k3): taps per output phase
    assert abs(siamunet.SiamUnet_conc(3, 2).eval().lower(256, 256).macs_per_pair() / 1e9 - 4.832) < 0.001


def test_fold_bn_matches_batchnorm():
    torch.manual_seed(0)
    bn = torch.nn.BatchNorm2d(8).eval()
    synth.randomize_(bn)
    bias = torch.randn(8)
    acc = torch.randn(2, 8, 5, 5)
    scale, shift = L.fold_bn(bias, {k: v for k, v in bn.state_dict().items()}, 8)
    want = bn(acc + bias[None, :, None, None])
    got = acc * torch.from_numpy(scale)[None, :, None, None] + torch.from_numpy(shift)[None, :, None, None]
    assert (want - got).abs().max().item() < 1e-5


def test_convT_phase_taps_equal_conv_transpose():
    torch.manual_seed(1)
    for (k, s, p, op) in [(3, 2, 1, 1), (2, 2, 0, 0), (4, 2, 1, 0)]:
        wt = torch.randn(5, 7, k, k)
        x = torch.randn(2, 5, 6, 4)
        want = torch.nn.functional.conv_transpose2d(x, wt, stride=s, padding=p, output_padding=op)
        got = torch.zeros_like(want)
        xp = torch.nn.functional.pad(x, (2, 2, 2, 2))
        for (oy, ox, taps) in L.convT_phase_taps(wt, s, p):
            acc = 0
            for (dy, dx, w) in taps:
                acc = acc + torch.einsum("nchw,oc->nohw", xp[:, :, 2 + dy: 2 + dy + 6, 2 + dx: 2 + dx + 4], w)
            got[:, :, oy::s, ox::s] = acc
        assert (want - got).abs().max().item() < 1e-4


def test_lowering_rejects_bad_shapes():
    net = siamunet.SiamUnet_diff(3, 2).eval()
    with pytest.raises(ValueError):
        net.lower(250, 256)
    with pytest.raises(RuntimeError):
        net.train()(torch.zeros(1, 3, 16, 16), torch.zeros(1, 3, 16, 16))
    with pytest.raises(RuntimeError):      # no CPU path
        net.eval()(torch.zeros(1, 3, 16, 16), torch.zeros(1, 3, 16, 16))
