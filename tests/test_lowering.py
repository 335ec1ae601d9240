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


@pytest.mark.parametrize("fusion,cls,gain", [("diff", siamunet.SiamUnet_diff, synth.GAINS["SiamUnet_diff"]), ("conc", siamunet.SiamUnet_conc, synth.GAINS["SiamUnet_conc"]),
                                             ("sub", siamunet.SiamUnet_sub, synth.GAINS["SiamUnet_sub"]),
                                             ("cross", siamunet.SiamUnet_cross_conc, synth.GAINS["SiamUnet_cross_conc"]),
                                             ("ef", siamunet.Unet, synth.GAINS["Unet"])])
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


def test_segcd_program_matches_oracle():
    """smp.SegCD(resnet34): space-to-depth stem, parity-class stride-2 convs, phase-decomposed nearest
    up-sampling + virtual concat, fused head -- all host-side lowering, checked through the emulator."""
    from stcd_b200 import segcd
    net = synth.prepare_(segcd.SegCD("resnet34").eval(), "SegCD")
    x1, x2 = synth.image_pairs(3, 64, 96)
    with torch.no_grad():
        y = nets.segcd_forward(net.state_dict(), x1, x2)
    prog = net.lower(64, 96)
    ye = emulate.run_program(prog, x1, x2, chunk=2)
    assert len(ye) == 3
    for a, b in zip(ye, y):
        assert a.shape == b.shape and (a - b).abs().max().item() < BF16_TOL
    change, change_e = y[2], ye[2]
    agree = (change_e > 0) == (change > 0)          # sigmoid(x) > 0.5 (train_stcd.py:477) up to the fp32 rounding at 0
    assert agree[change.abs() > BF16_TOL].float().mean().item() >= 0.999
    assert 0.02 < (change > 0).float().mean().item() < 0.98, "degenerate change map"
    convs = [o for o in prog.ops if isinstance(o, L.ConvSpec)]
    # 1 stem + 16 blocks x 2 + 3 downsamples + 5 decoder blocks x 2 = 46 convs; + pack, max-pool, head
    assert len(convs) == 46 and len(prog.ops) == 49
    dec1 = [o for o in convs if o.name.startswith("decoder") and o.name.endswith("conv1")]
    # nearest-x2 + conv as 4 output phases; the narrow ones (Cout 32, 16) fold all phases into GEMM N, Cout 64 folds the
    # two horizontal phases of each output-row parity (2 GEMM phases, N = 128: adjacent output pixels leave together)
    assert all((len(o.phases) == 4 and not o.fold_cs) or (len(o.phases) in (1, 2) and o.fold_cs) for o in dec1)
    assert [o.fold_cs for o in dec1] == [0, 0, 64, 32, 16] and [o.n_tile for o in dec1][2:] == [128, 128, 64]
    assert [len(o.phases) for o in dec1] == [4, 4, 2, 1, 1]
    big = net.lower(1024, 1024)
    # 500.397 GFLOP per pair at 1024x1024 (encoder 153.1 + decoder 96.6 + head 0.45 GMAC)
    assert abs(2 * big.macs_per_pair() / 1e9 - 500.397) < 0.5
    with pytest.raises(ValueError):
        net.lower(1000, 1024)
    with pytest.raises(NotImplementedError):
        segcd.SegCD("resnext50_32x4d")


def test_ffctlcd_program_matches_oracle():
    from stcd_b200 import segcd
    net = synth.prepare_(segcd.FFCTLCD("resnet18").eval(), "SegCD")
    x1, x2 = synth.image_pairs(2, 64, 64)
    with torch.no_grad():
        y = nets.ffctlcd_forward(net.state_dict(), x1, x2, layers=(2, 2, 2, 2))
    prog = net.lower(64, 64)
    ye = emulate.run_program(prog, x1, x2, chunk=2)
    for a, b in zip(ye, y):
        assert (a - b).abs().max().item() < BF16_TOL
    assert sum(isinstance(o, L.AbsDiffSpec) for o in prog.ops) == 5       # 4 skips + the bottleneck


def test_segcd_resnet50_program_matches_oracle():
    """The encoder the STCD script selects (train_stcd.py:638): Bottleneck blocks; the 1x1 conv ahead of a
    stride-2 3x3 runs per parity class of its space-to-depth input."""
    from stcd_b200 import segcd
    net = synth.prepare_(segcd.SegCD("resnet50").eval(), "SegCD")
    x1, x2 = synth.image_pairs(2, 64, 64)
    with torch.no_grad():
        y = nets.segcd_forward(net.state_dict(), x1, x2)
    prog = net.lower(64, 64)
    ye = emulate.run_program(prog, x1, x2, chunk=2)
    for a, b in zip(ye, y):
        assert (a - b).abs().max().item() < BF16_TOL
    convs = [o for o in prog.ops if isinstance(o, L.ConvSpec)]
    # stem + 16 blocks x 3 + 3 x 3 extra per-class conv1 launches + 4 downsamples + 10 decoder convs
    assert len(convs) == 1 + 48 + 9 + 4 + 10
    assert abs(2 * net.lower(1024, 1024).macs_per_pair() / 1e9 - 680.79) < 0.5


def test_dtcdscn_program_matches_oracle():
    """DTCDSCN (CDNet34): SE gates (two-pass), signed feature differences fused with the decoder addends, dilated centre block
    with statically dropped out-of-range taps, SCSE decoder, ConvTranspose2d phases -- checked through the emulator."""
    from stcd_b200 import dtcdscn
    net = synth.prepare_(dtcdscn.CDNet34(3, 2).eval(), "CDNet_model")
    x1, x2 = synth.image_pairs(3, 64, 96)
    with torch.no_grad():
        y = nets.dtcdscn_forward(net.state_dict(), x1, x2)
    prog = net.lower(64, 96)
    ye = emulate.run_program(prog, x1, x2, chunk=2)[0]
    assert ye.shape == y.shape == (3, 2, 64, 96) and (ye - y).abs().max().item() < BF16_TOL
    margin = (y[:, 1] - y[:, 0]).abs()
    agree = (ye[:, 1] > ye[:, 0]) == (y[:, 1] > y[:, 0])
    assert agree[margin > BF16_TOL].float().mean().item() >= 0.999
    assert 0.02 < (y[:, 1] > y[:, 0]).float().mean().item() < 0.98, "degenerate change map"
    convs = [o for o in prog.ops if isinstance(o, L.ConvSpec)]
    gates = [o for o in prog.ops if isinstance(o, L.ChannelGateSpec)]
    # stem + 16 blocks x 2 + 3 downsamples + 4 dilated + 4 decoder blocks x 3 + 3 head convs
    assert len(convs) == 1 + 32 + 3 + 4 + 12 + 3 and len(gates) == 16 + 4
    # e4 is 2x3 here: dilations 2, 4, 8 reach past it on one or both axes, only in-range taps stay
    dil = [o for o in convs if o.name.startswith("dblock_master")]
    assert [len(o.taps) // len(o.chunks) for o in dil] == [9, 3, 1, 1]
    big = net.lower(1024, 1024)
    dil = [o for o in big.ops if isinstance(o, L.ConvSpec) and o.name.startswith("dblock_master")]
    assert [o.kc for o in dil] == [64, 64, 32, 16] and [max(o.src_ey) for o in dil] == [2, 4, 8, 16]
    # the single-image branch is held as parameters only, like upstream (its forward is commented out, DTCDSCN.py:256-292)
    assert "decoder4.conv1.weight" in net.state_dict() and not any(o.name.startswith("decoder4.") for o in prog.ops)
    with pytest.raises(ValueError):
        net.lower(100, 96)


@pytest.mark.parametrize("variant", ["s4", "dd8", "dd8_dedim8", "resnet18"])
def test_bit_program_matches_oracle(variant):
    """BIT (BASE_Transformer) and its ResNet-18 baseline: backbone convs, nearest x2 + conv_pred as merged-tap phases, the token
    path as one op (checked semantically by the emulator from the PACKED weights), |x1 - x2|, bilinear x4, classifier."""
    from stcd_b200 import bit
    if variant == "resnet18":
        net, stages, cls = bit.ResNet(3, 2), 5, "ResNet"
    else:
        kw = {"s4": {}, "dd8": dict(enc_depth=1, dec_depth=8), "dd8_dedim8": dict(enc_depth=1, dec_depth=8, decoder_dim_head=8)}[variant]
        net, stages, cls = bit.BASE_Transformer(3, 2, with_pos="learned", resnet_stages_num=4, token_len=4, **kw), 4, "BASE_Transformer"
    net = synth.prepare_(net.eval(), cls)
    x1, x2 = synth.image_pairs(3, 64, 96)
    with torch.no_grad():
        y = nets.bit_forward(net.state_dict(), x1, x2, stages=stages)
    prog = net.lower(64, 96)
    ye = emulate.run_program(prog, x1, x2, chunk=2)[0]
    assert ye.shape == y.shape == (3, 2, 64, 96) and (ye - y).abs().max().item() < BF16_TOL
    margin = (y[:, 1] - y[:, 0]).abs()
    agree = (ye[:, 1] > ye[:, 0]) == (y[:, 1] > y[:, 0])
    assert agree[margin > BF16_TOL].float().mean().item() >= 0.999
    if variant in ("dd8", "resnet18"):          # the harness' head-bias offset is tuned for the dd8 key
        assert 0.02 < (y[:, 1] > y[:, 0]).float().mean().item() < 0.98, "degenerate change map"
    convs = [o for o in prog.ops if isinstance(o, L.ConvSpec)]
    bits = [o for o in prog.ops if isinstance(o, L.BitTransformerSpec)]
    # stem + 2 convs per BasicBlock + 1 downsample per strided/widening layer + conv_pred + 2 classifier convs
    assert len(convs) == 1 + 4 * (stages - 1) + (stages - 2) + 1 + 2
    if variant == "resnet18":
        assert not bits
    else:
        (b,) = bits
        assert (len(b.enc), len(b.dec), b.inner_dec) == {"s4": (1, 1, 512), "dd8": (1, 8, 512), "dd8_dedim8": (1, 8, 64)}[variant]
        assert b.enc.shape[1] == 2 * 32 + 3 * 512 * 32 + 32 * 512 + 32 + 2 * 32 + 64 * 32 + 64 + 32 * 64 + 32
    # the backbone's unused layer4 / fc are held as parameters, like upstream
    assert "resnet.fc.weight" in net.state_dict() and "resnet.layer4.1.conv2.weight" in net.state_dict()
    with pytest.raises(ValueError):
        net.lower(100, 96)


def test_dsifn_program_matches_oracle():
    """IFNet (DSIFN): shared VGG16 with max-pools fused into the conv epilogues, conv -> PReLU -> BN epilogues, channel attention
    over a virtual concat, spatial attention + BN, k2 s2 transposed convs as single-tap phases -- through the emulator."""
    from stcd_b200 import dsifn, networks
    net = synth.prepare_(networks.CLASSES["DSIFN"]().eval(), "DSIFN")
    x1, x2 = synth.image_pairs(3, 64, 96)
    with torch.no_grad():
        y = nets.dsifn_forward(net.state_dict(), x1, x2)
    prog = net.lower(64, 96)
    ye = emulate.run_program(prog, x1, x2, chunk=2)[0]
    assert ye.shape == y.shape == (3, 1, 64, 96) and (ye - y).abs().max().item() < BF16_TOL
    agree = (ye > 0) == (y > 0)                              # sigmoid(out) > 0.5
    assert agree[y.abs() > BF16_TOL].float().mean().item() >= 0.999
    assert 0.02 < (y > 0).float().mean().item() < 0.98, "degenerate change map"
    convs = [o for o in prog.ops if isinstance(o, L.ConvSpec)]
    # 13 VGG convs + 2 + 3 x 4 conv2d_bn + 4 transposed convs + the 1x1 head; 4 channel attentions, 5 spatial gates
    assert len(convs) == 13 + 14 + 4 + 1
    assert sum(isinstance(o, L.ChannelAttentionSpec) for o in prog.ops) == 4 and sum(isinstance(o, L.SpatialGateSpec) for o in prog.ops) == 5
    assert sum(o.out_pool is not None for o in convs) == 4, "the four max-pools ride in the producing conv's epilogue"
    # the deep-supervision side heads are parameters only (their outputs are discarded upstream, DSIFN.py:133,147,159,171)
    assert "o1_conv3.weight" in net.state_dict() and not any(o.name == "o1_conv3" for o in prog.ops)
    other = dsifn.DSIFN(dsifn.vgg16_base(), dsifn.vgg16_base())
    with pytest.raises(NotImplementedError):
        other.lower(64, 64)                                  # two different bases: not the registry's network
    with pytest.raises(ValueError):
        net.lower(72, 64)


def test_changegnn_program_matches_oracle():
    """Config C4's net: ViG Grapher blocks (graph op + grouped conv folded into a dense virtual-concat conv), GELU /
    PReLU-before-BN epilogues, bilinear resizes, ConvTranspose2d(k4, s2) phases -- checked through the emulator."""
    from stcd_b200 import changevig
    net = synth.prepare_(changevig.ChangeGNNV1().eval(), "ChangeGNNV1")
    x1, x2 = synth.image_pairs(1, 256, 256)
    with torch.no_grad():
        y = nets.changegnn_forward(net.state_dict(), x1, x2)
    prog = net.lower(256, 256)
    ye = emulate.run_program(prog, x1, x2, chunk=1)
    assert len(ye) == 5
    for a, b in zip(ye, y):
        assert a.shape == b.shape and (a - b).abs().max().item() < BF16_TOL
    margin = (y[-1][:, 1] - y[-1][:, 0]).abs()
    agree = (ye[-1][:, 1] > ye[-1][:, 0]) == (y[-1][:, 1] > y[-1][:, 0])
    assert agree[margin > BF16_TOL].float().mean().item() >= 0.999
    assert 0.02 < (y[-1][:, 1] > y[-1][:, 0]).float().mean().item() < 0.98, "degenerate change map"
    graphs = [o for o in prog.ops if isinstance(o, L.GraphConvSpec)]
    assert [(g.k, g.dilation, g.r) for g in graphs] == [(9, 1, 4)] * 2 + [(9, 1, 2)] * 2 + [(9, 2, 1)] * 4 + [(9, 3, 1)] * 4
    assert abs(2 * prog.macs_per_pair() / 1e9 - 283.07) < 0.5          # ~290 GFLOP per pair
    with pytest.raises(ValueError):
        net.lower(128, 128)            # pos_embed is not resized: the net only runs at img_size (ChangeVIG.py:87)


@pytest.mark.parametrize("mode", ["cross", "sub", "abs", "conc"])
def test_changegnn_v2_program_matches_oracle(mode):
    """ChangeGNNV2 / ChangeGNNV2_Compare: the V1 ViG encoder + HFFM (Cross_ConCat as block-diagonal convs | Sub | Abs | Conc split
    along K, residual bottleneck, Global_Local with its 1x1 / 3x3 / 7x7 depth-wise convs + 1x1 mix composed into one 7x7 conv)
    + VFFM, through the emulator."""
    from stcd_b200 import changevig
    cls = "ChangeGNNV2" if mode == "cross" else "ChangeGNNV2_Compare"
    net = changevig.ChangeGNNV2() if mode == "cross" else changevig.ChangeGNNV2_Compare(diff_mode=mode)
    net = synth.prepare_(net.eval(), cls)
    x1, x2 = synth.image_pairs(1, 256, 256)
    with torch.no_grad():
        y = nets.changegnn_v2_forward(net.state_dict(), x1, x2, mode)
    prog = net.lower(256, 256)
    ye = emulate.run_program(prog, x1, x2, chunk=1)
    assert len(ye) == len(y) == 1 and ye[0].shape == y[0].shape == (1, 2, 256, 256)
    assert (ye[0] - y[0]).abs().max().item() < BF16_TOL
    margin = (y[0][:, 1] - y[0][:, 0]).abs()
    agree = (ye[0][:, 1] > ye[0][:, 0]) == (y[0][:, 1] > y[0][:, 0])
    assert agree[margin > BF16_TOL].float().mean().item() >= 0.999
    assert 0.02 < (y[0][:, 1] > y[0][:, 0]).float().mean().item() < 0.98, "degenerate change map"
    assert sum(isinstance(o, L.GlobalLocalGateSpec) for o in prog.ops) == 4 and sum(isinstance(o, L.VffmSpec) for o in prog.ops) == 3
    if mode == "cross":
        # block-diagonal grouped conv: 1 / 1 / 2 / 2 channel blocks for 80 / 160 / 400 / 640 channels
        assert sum(".cross_conc.diff." in o.name for o in prog.ops if isinstance(o, L.ConvSpec)) == 6
    with pytest.raises(ValueError):
        net.lower(512, 512)


def test_vig_v20_program_matches_oracle():
    """VIG_V20_2 (registry key "GNN"): the ViG encoder under the prefix VIG_x2, conv_diff_V20 (Cross_ConCat), csam_V20 gates, AFF as a
    VFFM with an empty max branch, k2 s2 transposed convs -- through the emulator."""
    from stcd_b200 import changevig
    net = synth.prepare_(changevig.VIG_V20_2().eval(), "VIG_V20_2")
    x1, x2 = synth.image_pairs(1, 256, 256)
    with torch.no_grad():
        y = nets.vig_v20_forward(net.state_dict(), x1, x2)
    prog = net.lower(256, 256)
    ye = emulate.run_program(prog, x1, x2, chunk=1)
    assert len(ye) == len(y) == 1 and (ye[0] - y[0]).abs().max().item() < BF16_TOL
    margin = (y[0][:, 1] - y[0][:, 0]).abs()
    agree = (ye[0][:, 1] > ye[0][:, 0]) == (y[0][:, 1] > y[0][:, 0])
    assert agree[margin > BF16_TOL].float().mean().item() >= 0.999
    assert 0.02 < (y[0][:, 1] > y[0][:, 0]).float().mean().item() < 0.98, "degenerate change map"
    assert sum(isinstance(o, L.CsamGateSpec) for o in prog.ops) == 4 and sum(isinstance(o, L.VffmSpec) for o in prog.ops) == 3


def test_changeformer_program_matches_oracle():
    """Config C5's net: Linear layers as 1x1 convs, strided patch-embedding / spatial-reduction convs, LayerNorm,
    64-key attention and depth-wise conv ops -- checked through the emulator."""
    from stcd_b200 import changeformer
    net = synth.prepare_(changeformer.ChangeFormerV6().eval(), "ChangeFormerV6")
    x1, x2 = synth.image_pairs(1, 256, 256)
    with torch.no_grad():
        y = nets.changeformer_forward(net.state_dict(), x1, x2)
    prog = net.lower(256, 256)
    ye = emulate.run_program(prog, x1, x2, chunk=1)
    assert len(ye) == 5
    for a, b in zip(ye, y):
        assert a.shape == b.shape and (a - b).abs().max().item() < BF16_TOL
    margin = (y[-1][:, 1] - y[-1][:, 0]).abs()
    agree = (ye[-1][:, 1] > ye[-1][:, 0]) == (y[-1][:, 1] > y[-1][:, 0])
    assert agree[margin > BF16_TOL].float().mean().item() >= 0.999
    assert 0.02 < (y[-1][:, 1] > y[-1][:, 0]).float().mean().item() < 0.98, "degenerate change map"
    assert abs(2 * prog.macs_per_pair() / 1e9 - 277.459) < 0.01           # 277.459 GFLOP per pair
    assert sum(isinstance(o, L.AttentionSpec) for o in prog.ops) == 13 and sum(isinstance(o, L.LayerNormSpec) for o in prog.ops) == 44
    with pytest.raises(ValueError):
        net.lower(128, 128)


@pytest.mark.parametrize("version", ["1", "2", "3"])
def test_changeformer_v1_v2_program_matches_oracle(version):
    """ChangeFormerV1 / V2: Tenc (EncoderTransformer: 3x3 stride-2 patch embeds, depths 3-4-6-3, the never-called intra-patch blocks
    held as parameters), |fx1 - fx2| per scale, convprojection_base (V1) or TDec (V2) -- through the emulator."""
    from stcd_b200 import changeformer
    cls = f"ChangeFormerV{version}"
    net = synth.prepare_(getattr(changeformer, cls)().eval(), cls)
    x1, x2 = synth.image_pairs(1, 256, 256)
    with torch.no_grad():
        y = getattr(nets, f"changeformer_v{version}_forward")(net.state_dict(), x1, x2)
    prog = net.lower(256, 256)
    ye = emulate.run_program(prog, x1, x2, chunk=1)[0]
    assert ye.shape == y.shape == (1, 2, 256, 256) and (ye - y).abs().max().item() < BF16_TOL
    margin = (y[:, 1] - y[:, 0]).abs()
    agree = (ye[:, 1] > ye[:, 0]) == (y[:, 1] > y[:, 0])
    assert agree[margin > BF16_TOL].float().mean().item() >= 0.999
    assert 0.02 < (y[:, 1] > y[:, 0]).float().mean().item() < 0.98, "degenerate change map"
    assert "Tenc.patch_block2.0.attn.q.weight" in net.state_dict() and not any("patch_block" in o.name for o in prog.ops)
    assert sum(isinstance(o, L.AttentionSpec) for o in prog.ops) == 16 and sum(isinstance(o, L.AbsDiffSpec) for o in prog.ops) == 4
    if version == "3":             # PixelShuffle(4) head: one single-phase launch per output phase, all writing the fp32 output
        ps = [o for o in prog.ops if isinstance(o, L.ConvSpec) and ".pix_shuffle_conv." in o.name]
        assert len(ps) == 16 and all(o.osy == 4 and o.osx == 4 and len(o.phases) == 1 and o.out_ext == 0 for o in ps)
        assert sorted((o.phases[0].oy, o.phases[0].ox) for o in ps) == [(i, j) for i in range(4) for j in range(4)]
    with pytest.raises(ValueError):
        net.lower(512, 512)


def test_s2d_and_up2_tap_algebra():
    """The tap rewrites behind the SegCD lowering equal the reference ops they replace (fp32, no rounding)."""
    import torch.nn.functional as F
    torch.manual_seed(2)
    x = torch.randn(2, 5, 8, 12)
    w = torch.randn(7, 5, 3, 3)

    def s2d(t):        # [n, c, h, w] -> [n, 4c, h/2, w/2], channel (py*2+px)*c + ch
        return torch.cat([t[:, :, py::2, px::2] for py in range(2) for px in range(2)], 1)

    def run_taps(src, taps_per_seg, c, hg, wg):      # src [n, 4c or c, hg, wg] as class segments of c channels
        out = 0
        pad = F.pad(src, (2, 2, 2, 2))
        for k, taps in enumerate(taps_per_seg):
            for (dy, dx, wt) in taps:
                out = out + torch.einsum("nchw,oc->nohw", pad[:, k * c: (k + 1) * c, 2 + dy: 2 + dy + hg, 2 + dx: 2 + dx + wg], wt)
        return out

    # stride-2 3x3 conv == parity-class taps over the space-to-depth tensor
    want = F.conv2d(x, w, stride=2, padding=1)
    got = run_taps(s2d(x), L.s2d_conv_taps(w, pad=1), 5, 4, 6)
    assert (want - got).abs().max().item() < 1e-4
    # 1x1 stride-2 conv: only class (0, 0)
    w1 = torch.randn(7, 5, 1, 1)
    taps = L.s2d_conv_taps(w1, pad=0)
    assert [len(t) for t in taps] == [1, 0, 0, 0]
    assert (F.conv2d(x, w1, stride=2) - run_taps(s2d(x), taps, 5, 4, 6)).abs().max().item() < 1e-4
    # conv over nearest-upsampled low-res + full-res skip == 4 phases of (merged 2x2 taps, parity-class taps)
    lo = torch.randn(2, 3, 4, 6)
    wcat = torch.randn(7, 8, 3, 3)
    want = F.conv2d(torch.cat([F.interpolate(lo, scale_factor=2, mode="nearest"), x], 1), wcat, padding=1)
    got = torch.zeros_like(want)
    for a in range(2):
        for b in range(2):
            up = L.up2_conv_taps(wcat[:, :3], 1, a, b)
            assert len(up) == 4
            got[:, :, a::2, b::2] = run_taps(lo, [up], 3, 4, 6) + run_taps(s2d(x), L.s2d_conv_taps(wcat[:, 3:], 1, a, b), 5, 4, 6)
    assert (want - got).abs().max().item() < 1e-4
    # 7x7 stride-2 stem == 4x4 conv over space-to-depth input
    from stcd_b200.segcd import stem_s2d_taps
    img = torch.randn(2, 3, 16, 24)
    ws = torch.randn(6, 3, 7, 7)
    (_, _, taps), = stem_s2d_taps(ws)
    assert len(taps) == 16
    got = run_taps(s2d(img), [taps], 12, 8, 12)
    assert (F.conv2d(img, ws, stride=2, padding=3) - got).abs().max().item() < 1e-4


def test_program_structure_and_macs():
    net = siamunet.SiamUnet_diff(3, 2).eval()
    prog = net.lower(256, 256)
    convs = [o for o in prog.ops if isinstance(o, L.ConvSpec)]
    assert len(convs) == 24 and len(prog.ops) == 25
    # 4.228 GMAC per pair for SiamUnet_diff at 256x256
    assert abs(prog.macs_per_pair() / 1e9 - 4.228) < 0.001
    up = [o for o in convs if o.name.startswith("upconv")]
    assert all((len(o.phases) == 4 or o.fold_cs) and o.osy == 2 and o.osx == 2 for o in up)
    # upconv2 / upconv1: all phases folded into N (4*C <= 128); upconv3: the horizontal phases per output-row parity (2*C = 128)
    assert [o.fold_cs for o in up] == [0, 64, 32, 16] and [len(o.phases) for o in up] == [4, 2, 1, 1]
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


@pytest.mark.parametrize("family", ["siamunet_diff", "siamunet_conc", "snunet"])
def test_split_precision_program_meets_1e3(family):
    """precision = "tf32" (split-bf16 operands: lowering.Program.precision) through the emulator: logits within 1e-3 of the fp32
    oracle at a logit spread >= 0.25 -- the tolerance class the north star names for the tf32 path -- and the change maps
    agree on every decided pixel.  Plain tf32 rounding of the operands does NOT meet it."""
    from stcd_b200 import siamunet, snunet
    if family == "snunet":
        net = synth.prepare_(snunet.SNUNet_ECAM(3, 2).eval(), "SNUNet_ECAM")
        fwd = nets.snunet_forward
    else:
        fusion = family.split("_")[1]
        cls = {"diff": siamunet.SiamUnet_diff, "conc": siamunet.SiamUnet_conc}[fusion]
        net = synth.randomize_(cls(3, 2).eval(), gain=0.77)
        fwd = lambda sd, a, b: nets.siamunet_forward(sd, a, b, fusion)  # noqa: E731
    net.precision = "tf32"
    x1, x2 = synth.image_pairs(3, 32, 48)
    with torch.no_grad():
        ref = fwd(net.state_dict(), x1, x2)
    prog = net.lower(32, 48)
    assert prog.split and all(o.split for o in prog.ops if isinstance(o, (L.ConvSpec, L.InputPackSpec)))
    assert not any(o.xf_cs or o.fold_cs for o in prog.ops if isinstance(o, L.ConvSpec))
    y = emulate.run_program(prog, x1, x2, chunk=2)[0]
    err = (y - ref).abs().max().item()
    assert ref.std().item() >= 0.2 and err < 1e-3, (ref.std().item(), err)
    assert err < 2e-4, err                       # 16-bit operand mantissas: ~1e-4 of the logit spread
    agree = ((y[:, 1] > y[:, 0]) == (ref[:, 1] > ref[:, 0]))[(ref[:, 1] - ref[:, 0]).abs() > 1e-3]
    assert agree.all()
    net.precision = "bf16"                       # the same module lowers back to the bf16 program
    assert not net.lower(32, 48).split
    net.precision = "fp64"
    with pytest.raises(ValueError):
        net.lower(32, 48)


def test_precision_path_is_refused_where_it_is_not_implemented():
    from stcd_b200 import segcd
    net = segcd.SegCD("resnet34").eval()
    net.precision = "tf32"
    with pytest.raises(NotImplementedError):
        net.plan_precision
