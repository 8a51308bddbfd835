"""State-dict contract of the OTVM stage-4 eval model (the drop-in boundary).

The reference's ``eval.py:77-79`` loads ``weights/s4_OTVM.pth`` strictly, so the 785 key names and
shapes are part of the boundary.  This module enumerates them from the network topology alone
(torchvision ResNet-50 up to layer3 for the two STM encoders, ``models/trimap/STM.py:32-102``; the
GroupNorm/weight-standardised dilated ResNet-50 of FBA, ``models/alpha/FBA/resnet_GN_WS.py:90-134``
and ``models/alpha/FBA/models.py:208-269``; the FBA decoder ``:291-349`` and refinement module
``:395-416``) without importing the reference.

Every entry carries a *role* so the fixture generator (``otvm_b200/fixtures.py``) can draw
deterministic, name-keyed values for it.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import NamedTuple, Tuple


class Entry(NamedTuple):
    shape: Tuple[int, ...]
    role: str          # conv_w | conv_b | norm_w | norm_b | bn_mean | bn_var | bn_count | const
    dtype: str = "float32"


def _conv(spec, name, cout, cin, k, bias):
    spec[name + ".weight"] = Entry((cout, cin, k, k), "conv_w")
    if bias:
        spec[name + ".bias"] = Entry((cout,), "conv_b")


def _gn(spec, name, c):
    spec[name + ".weight"] = Entry((c,), "norm_w")
    spec[name + ".bias"] = Entry((c,), "norm_b")


def _bn(spec, name, c):
    _gn(spec, name, c)
    spec[name + ".running_mean"] = Entry((c,), "bn_mean")
    spec[name + ".running_var"] = Entry((c,), "bn_var")
    spec[name + ".num_batches_tracked"] = Entry((), "bn_count", "int64")


RESNET50_LAYERS = ((64, 3), (128, 4), (256, 6), (512, 3))   # (planes, blocks)


def _bottleneck_stack(spec, prefix, layer_names, norm):
    """ResNet-50 bottleneck stacks; ``norm`` is ``_bn`` (torchvision) or ``_gn`` (FBA GN+WS)."""
    inplanes = 64
    for lname, (planes, blocks) in zip(layer_names, RESNET50_LAYERS):
        for b in range(blocks):
            p = f"{prefix}.{lname}.{b}"
            _conv(spec, p + ".conv1", planes, inplanes, 1, False); norm(spec, p + ".bn1", planes)
            _conv(spec, p + ".conv2", planes, planes, 3, False);   norm(spec, p + ".bn2", planes)
            _conv(spec, p + ".conv3", planes * 4, planes, 1, False); norm(spec, p + ".bn3", planes * 4)
            if b == 0:
                _conv(spec, p + ".downsample.0", planes * 4, inplanes, 1, False)
                norm(spec, p + ".downsample.1", planes * 4)
            inplanes = planes * 4


def _resblock(spec, p, c):
    _conv(spec, p + ".conv1", c, c, 3, True)
    _conv(spec, p + ".conv2", c, c, 3, True)


def state_spec(hdim: int = 16) -> "OrderedDict[str, Entry]":
    """(name -> Entry) for ``EvalModel(stage=4)`` wrapping ``FullModel_eval(stage=4, hdim=16)``."""
    s: "OrderedDict[str, Entry]" = OrderedDict()
    s["IMG_MEAN"] = Entry((1, 1, 3, 1, 1), "const")
    s["IMG_STD"] = Entry((1, 1, 3, 1, 1), "const")

    # ---- alpha network (FBA) -------------------------------------------------------------
    _conv(s, "NET.encoder.conv1", 64, 11, 7, False); _gn(s, "NET.encoder.bn1", 64)
    _bottleneck_stack(s, "NET.encoder", ("layer1", "layer2", "layer3", "layer4"), _gn)
    for i in range(4):
        _conv(s, f"NET.decoder.ppm.{i}.1", 256, 2048, 1, True); _gn(s, f"NET.decoder.ppm.{i}.2", 256)
    _conv(s, "NET.decoder.conv_up1.0", 256, 3072, 3, True); _gn(s, "NET.decoder.conv_up1.1", 256)
    _conv(s, "NET.decoder.conv_up1.3", 256, 256, 3, True);  _gn(s, "NET.decoder.conv_up1.4", 256)
    _conv(s, "NET.decoder.conv_up2.0", 256, 512, 3, True);  _gn(s, "NET.decoder.conv_up2.1", 256)
    _conv(s, "NET.decoder.conv_up3.0", 64, 320, 3, True);   _gn(s, "NET.decoder.conv_up3.1", 64)
    _conv(s, "NET.decoder.conv_up4.0", 32, 72, 3, True)
    _conv(s, "NET.decoder.conv_up4.2", 16, 32, 3, True)
    _conv(s, "NET.decoder.conv_up4.4", 7, 16, 1, True)
    _conv(s, "NET.refine.conv1.0", 64, 73, 3, True); _gn(s, "NET.refine.conv1.1", 64)
    for l in ("layer1", "layer2"):
        _conv(s, f"NET.refine.{l}.conv1", 64, 64, 3, False); _gn(s, f"NET.refine.{l}.bn1", 64)
        _conv(s, f"NET.refine.{l}.conv2", 64, 64, 3, False); _gn(s, f"NET.refine.{l}.bn2", 64)
    _conv(s, "NET.refine.pred.0", 32, 64, 3, True)
    _conv(s, "NET.refine.pred.2", 16, 32, 3, True)
    _conv(s, "NET.refine.pred.4", 10, 16, 1, True)
    s["LAPLOSS.KERNEL"] = Entry((5, 5), "const")

    # ---- trimap propagation network (STM) ------------------------------------------------
    s["trimap.IMG_MEAN"] = Entry((1, 1, 3, 1, 1), "const")
    s["trimap.IMG_STD"] = Entry((1, 1, 3, 1, 1), "const")
    em = "trimap.model.Encoder_M"
    s[em + ".mean"] = Entry((1, 3, 1, 1), "const"); s[em + ".std"] = Entry((1, 3, 1, 1), "const")
    _conv(s, em + ".conv1_m", 64, 1, 7, False)
    _conv(s, em + ".conv1_o", 64, 1, 7, False)
    _conv(s, em + ".conv1_a", 64, 1, 7, False)
    _conv(s, em + ".conv1_h", 64, hdim, 7, False)
    _conv(s, em + ".conv1", 64, 3, 7, False); _bn(s, em + ".bn1", 64)
    _bottleneck_stack(s, em, ("res2", "res3", "res4"), _bn)
    eq = "trimap.model.Encoder_Q"
    s[eq + ".mean"] = Entry((1, 3, 1, 1), "const"); s[eq + ".std"] = Entry((1, 3, 1, 1), "const")
    _conv(s, eq + ".conv1", 64, 3, 7, False); _bn(s, eq + ".bn1", 64)
    _bottleneck_stack(s, eq, ("res2", "res3", "res4"), _bn)
    for kv in ("KV_M_r4", "KV_Q_r4"):
        _conv(s, f"trimap.model.{kv}.Key", 128, 1024, 3, True)
        _conv(s, f"trimap.model.{kv}.Value", 512, 1024, 3, True)
    d = "trimap.model.Decoder"
    _conv(s, d + ".convFM", 256, 1024, 3, True); _resblock(s, d + ".ResMM", 256)
    for rf, cin in (("RF3", 512), ("RF2", 256)):
        _conv(s, f"{d}.{rf}.convFS", 256, cin, 3, True)
        _resblock(s, f"{d}.{rf}.ResFS", 256); _resblock(s, f"{d}.{rf}.ResMM", 256)
    _conv(s, d + ".pred", 3, 256, 3, True)
    s["trimap.LOSS.weight"] = Entry((3,), "const")
    return s


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
