"""The reference's model-selection API for the nets this library accelerates.

Mirror of ``define_G`` / ``init_net`` / ``init_weights`` (models/networks.py:138-215, 119-135,
85-116): same signatures, same registry keys (note ``SiamUnet_diff`` is registered as
``"SiamUnet_abs"``, :148-149), same duck-typed ``args`` (only ``net_G``, ``n_class``,
``embed_dim``, ``img_size`` are read), same re-initialisation (``normal``: Conv*/Linear* weights
~ N(0, gain), biases 0, BatchNorm2d weight ~ N(1, gain)), same ``NotImplementedError`` for an
unknown name.  The returned ``nn.Module`` has the reference's parameter names and
``forward(x1, x2)`` contract, backed by libstcd_b200 instead of torch.nn.functional.
"""
from __future__ import annotations

import torch
from torch.nn import init

from .siamunet import SiamUnet_conc, SiamUnet_cross_conc, SiamUnet_diff, SiamUnet_sub, Unet
from .changeformer import ChangeFormerV1, ChangeFormerV2, ChangeFormerV3, ChangeFormerV6
from .bit import BASE_Transformer, ResNet
from .changevig import ChangeGNNV1, ChangeGNNV2, ChangeGNNV2_Compare, VIG_V20_2
from .dsifn import DSIFN, vgg16_base
from .dtcdscn import CDNet34, CDNet_model
from .segcd import SegCD
from .snunet import SNUNet_ECAM

# registry keys of models/networks.py:144-214 that this library does NOT implement (yet): asking for one
# raises NotImplementedError like an unknown key does upstream, with the reason.
_REFERENCE_ONLY = (
    "ChangeFormerV4", "ChangeFormerV5",
)

_REGISTRY = {
    "Unet": lambda a: Unet(input_nbr=3, label_nbr=a.n_class),                      # networks.py:144-145 (FC-EF)
    "SiamUnet_sub": lambda a: SiamUnet_sub(input_nbr=3, label_nbr=a.n_class),      # networks.py:146-147
    "SiamUnet_cross_conc": lambda a: SiamUnet_cross_conc(input_nbr=3, label_nbr=a.n_class),   # networks.py:152-153
    "SiamUnet_abs": lambda a: SiamUnet_diff(input_nbr=3, label_nbr=a.n_class),     # networks.py:148-149
    "SiamUnet_conc": lambda a: SiamUnet_conc(input_nbr=3, label_nbr=a.n_class),    # networks.py:151-152
    "DTCDSCN": lambda a: CDNet34(in_channels=3, num_classes=a.n_class),            # networks.py:159-160
    # BIT, networks.py:170-182
    "base_resnet18": lambda a: ResNet(input_nc=3, output_nc=2, output_sigmoid=False),
    "base_transformer_pos_s4": lambda a: BASE_Transformer(input_nc=3, output_nc=2, token_len=4, resnet_stages_num=4, with_pos="learned"),
    "base_transformer_pos_s4_dd8": lambda a: BASE_Transformer(input_nc=3, output_nc=2, token_len=4, resnet_stages_num=4,
                                                              with_pos="learned", enc_depth=1, dec_depth=8),
    "base_transformer_pos_s4_dd8_dedim8": lambda a: BASE_Transformer(input_nc=3, output_nc=2, token_len=4, resnet_stages_num=4,
                                                                     with_pos="learned", enc_depth=1, dec_depth=8, decoder_dim_head=8),
    "IFNet": lambda a: _ifnet(),                                                   # networks.py:164-166
    "SNUNet": lambda a: SNUNet_ECAM(in_ch=3, out_ch=a.n_class),                    # networks.py:168-169
    "ChangeGNNV1": lambda a: ChangeGNNV1(embed_dim=a.embed_dim),                   # networks.py:199-200
    "ChangeFormerV1": lambda a: ChangeFormerV1(),                                  # networks.py:184-185
    "ChangeFormerV2": lambda a: ChangeFormerV2(),                                  # networks.py:186-187
    "ChangeFormerV3": lambda a: ChangeFormerV3(),                                  # networks.py:188-189
    "ChangeFormerV6": lambda a: ChangeFormerV6(embed_dim=a.embed_dim),             # networks.py:190-191
    # networks.py:201-208
    "ChangeGNNV2": lambda a: ChangeGNNV2(embed_dim=a.embed_dim, img_size=a.img_size),
    "ChangeGNNV2_sub": lambda a: ChangeGNNV2_Compare(embed_dim=a.embed_dim, img_size=a.img_size, diff_mode="sub"),
    "ChangeGNNV2_abs": lambda a: ChangeGNNV2_Compare(embed_dim=a.embed_dim, img_size=a.img_size, diff_mode="abs"),
    "ChangeGNNV2_conc": lambda a: ChangeGNNV2_Compare(embed_dim=a.embed_dim, img_size=a.img_size, diff_mode="conc"),
    "GNN": lambda a: VIG_V20_2(embed_dim=a.embed_dim),                             # networks.py:210-211
}


# class name (= the reference's) -> drop-in wrapper; synth.GAINS / bench.py / the tests key on these names
CLASSES = {"SiamUnet_diff": SiamUnet_diff, "SiamUnet_conc": SiamUnet_conc, "SiamUnet_sub": SiamUnet_sub,
           "SiamUnet_cross_conc": SiamUnet_cross_conc, "Unet": Unet, "SNUNet_ECAM": SNUNet_ECAM, "SegCD": SegCD,
           "ChangeGNNV1": ChangeGNNV1, "ChangeFormerV6": ChangeFormerV6,
           "BASE_Transformer": BASE_Transformer, "ResNet": ResNet, "ChangeGNNV2": ChangeGNNV2, "ChangeGNNV2_Compare": ChangeGNNV2_Compare, "VIG_V20_2": VIG_V20_2, "ChangeFormerV1": ChangeFormerV1, "ChangeFormerV2": ChangeFormerV2, "ChangeFormerV3": ChangeFormerV3,
           "CDNet_model": lambda in_channels=3, num_classes=2: CDNet34(in_channels, num_classes)}


def _ifnet() -> DSIFN:
    base_model = vgg16_base()
    return DSIFN(base_model, base_model)


CLASSES["DSIFN"] = _ifnet


def register(name: str, ctor) -> None:
    _REGISTRY[name] = ctor


def init_weights(net, init_type="normal", init_gain=0.02):
    """models/networks.py:85-116."""
    def init_func(m):
        classname = m.__class__.__name__
        if hasattr(m, "weight") and (classname.find("Conv") != -1 or classname.find("Linear") != -1):
            if init_type == "normal":
                init.normal_(m.weight.data, 0.0, init_gain)
            elif init_type == "xavier":
                init.xavier_normal_(m.weight.data, gain=init_gain)
            elif init_type == "kaiming":
                init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            elif init_type == "orthogonal":
                init.orthogonal_(m.weight.data, gain=init_gain)
            else:
                raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
            if hasattr(m, "bias") and m.bias is not None:
                init.constant_(m.bias.data, 0.0)
        elif classname.find("BatchNorm2d") != -1:
            init.normal_(m.weight.data, 1.0, init_gain)
            init.constant_(m.bias.data, 0.0)

    print("initialize network with %s" % init_type)
    net.apply(init_func)
    if hasattr(net, "invalidate_plans"):
        net.invalidate_plans()


def init_net(net, init_type="normal", init_gain=0.02, gpu_ids=[]):  # noqa: B006 - the reference's signature
    """models/networks.py:119-135."""
    if len(gpu_ids) > 0:
        assert torch.cuda.is_available()
        net.to(gpu_ids[0])
        if len(gpu_ids) > 1:
            net = torch.nn.DataParallel(net, gpu_ids)
    init_weights(net, init_type, init_gain=init_gain)
    return net


def define_G(args, init_type="normal", init_gain=0.02, gpu_ids=[]):  # noqa: B006
    """models/networks.py:138-215.  One extension beside the reference's fields (net_G, n_class, embed_dim, img_size):
    ``args.precision`` ("bf16" default | "tf32"), the north star's two tolerance classes (stcd_b200/module.py)."""
    name = args.net_G
    if name in _REGISTRY:
        net = _REGISTRY[name](args)
        precision = getattr(args, "precision", None)
        if precision is not None:
            net.precision = precision
            net.plan_precision          # raises here, not at the first forward, when the family has no such path
    elif name in _REFERENCE_ONLY:
        raise NotImplementedError("Generator model name [%s] is served by the reference only; stcd_b200 accelerates %s"
                                  % (name, sorted(_REGISTRY)))
    else:
        raise NotImplementedError("Generator model name [%s] is not recognized" % name)
    return init_net(net, init_type, init_gain, gpu_ids)
