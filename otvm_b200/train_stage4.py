"""The stage-4 joint training step (BASELINE configs[4]; reference ``train.py:349-375``, ``models/alpha/model.py:189-312``,
``models/trimap/model.py:133-154``) with the fused ``Memory.read`` (forward + recompute backward, :mod:`otvm_b200.train`)
on the training path, under ``DistributedDataParallel``.

Scope.  The kernels of this repo are inference kernels; what a training step needs besides the read -- the backward of ~180
convolutions, GroupNorm, the losses -- is PyTorch autograd over cuDNN / cuBLAS, exactly as in the reference.  This module
is therefore a *PyTorch* restatement of the reference's training graph (functional ops over the reference's own 785
parameter names, so an ``s3_OTVM.pth`` / ``s4_OTVM.pth`` checkpoint loads strictly), written from the behaviour of the
reference, with ONE operator replaced: ``Memory.forward`` -> :func:`otvm_b200.train.memory_read`.  It exists so that the
fused read can be trained end to end and measured in the step it was built for; it is pinned against the reference's own
training step (losses, outputs, gradients) by ``tests/golden/train_step_s4.npz`` (``oracle/make_golden_train.py``).

Modes follow ``train.py:311-316``: the trimap network's BatchNorm runs in eval mode ("STM disables BN during training"),
the FBA network has GroupNorm only, there is no dropout -- so the forward arithmetic is the inference arithmetic.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import train as T
from .models import _grow

# ------------------------------------------------------------------------------------------------------------------
# layers (sd: name -> tensor; parameters of the module below)
# ------------------------------------------------------------------------------------------------------------------

def _conv(sd, name, x, stride=1, padding=0, dilation=1):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride, padding, dilation)


def _ws_conv(sd, name, x, stride=1, padding=0, dilation=1):
    """weight-standardised convolution, FBA/layers_WS.py:13-23"""
    w = sd[name + ".weight"]
    w = w - w.mean(dim=1, keepdim=True).mean(dim=2, keepdim=True).mean(dim=3, keepdim=True)
    std = torch.sqrt(torch.var(w.flatten(1), dim=1) + 1e-12).view(-1, 1, 1, 1) + 1e-5
    return F.conv2d(x, w / std, sd.get(name + ".bias"), stride, padding, dilation)


def _bn(sd, name, x):
    """BatchNorm2d in eval mode (train.py:312-316)"""
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"], sd[name + ".bias"],
                        False, 0.0, 1e-5)


def _gn(sd, name, x):
    return F.group_norm(x, 32, sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def _up(x, scale=None, size=None):
    return F.interpolate(x, size=size, scale_factor=scale, mode="bilinear", align_corners=False)


def _pad16(x):
    h, w = x.shape[-2:]
    nh, nw = h + (16 - h % 16) % 16, w + (16 - w % 16) % 16
    lh, lw = (nh - h) // 2, (nw - w) // 2
    pad = (lw, nw - w - lw, lh, nh - h - lh)
    return (F.pad(x, pad) if sum(pad) else x), pad


def _crop(x, pad):
    lw, uw, lh, uh = pad
    return x[:, :, lh:x.shape[2] - uh if uh else None, lw:x.shape[3] - uw if uw else None] if sum(pad) else x


# ---- STM (models/trimap/STM.py) -------------------------------------------------------------------------------------

def _tv_block(sd, p, x, stride):
    o = F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x)))
    o = F.relu(_bn(sd, p + ".bn2", _conv(sd, p + ".conv2", o, stride=stride, padding=1)))
    o = _bn(sd, p + ".bn3", _conv(sd, p + ".conv3", o))
    if (p + ".downsample.0.weight") in sd:
        x = _bn(sd, p + ".downsample.1", _conv(sd, p + ".downsample.0", x, stride=stride))
    return F.relu(o + x)


def _tv_layers(sd, p, x):
    outs = []
    for lname, blocks, stride in (("res2", 3, 1), ("res3", 4, 2), ("res4", 6, 2)):
        for b in range(blocks):
            x = _tv_block(sd, f"{p}.{lname}.{b}", x, stride if b == 0 else 1)
        outs.append(x)
    return outs


def _encoder_q(sd, frame, p="trimap.model.Encoder_Q"):
    f = (frame - sd[p + ".mean"]) / sd[p + ".std"]
    x = F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", f, stride=2, padding=3)))
    r2, r3, r4 = _tv_layers(sd, p, F.max_pool2d(x, 3, 2, 1))
    return r4, r3, r2


def _encoder_m(sd, frame, masks, p="trimap.model.Encoder_M"):
    """STM.py:56-74: masks = (trimap 3 | alpha 1 | hidden 16); unknown, fg, alpha and hidden get their own 7x7 stems"""
    f = (frame - sd[p + ".mean"]) / sd[p + ".std"]
    x = (_conv(sd, p + ".conv1", f, stride=2, padding=3) + _conv(sd, p + ".conv1_m", masks[:, 1:2], stride=2, padding=3)
         + _conv(sd, p + ".conv1_o", masks[:, 2:3], stride=2, padding=3)
         + _conv(sd, p + ".conv1_a", masks[:, 3:4], stride=2, padding=3)
         + _conv(sd, p + ".conv1_h", masks[:, 4:], stride=2, padding=3))
    x = F.relu(_bn(sd, p + ".bn1", x))
    return _tv_layers(sd, p, F.max_pool2d(x, 3, 2, 1))[2]


def _resblock(sd, p, x):
    r = _conv(sd, p + ".conv1", F.relu(x), padding=1)
    return x + _conv(sd, p + ".conv2", F.relu(r), padding=1)


def _refine(sd, p, f, pm):
    s = _resblock(sd, p + ".ResFS", _conv(sd, p + ".convFS", f, padding=1))
    return _resblock(sd, p + ".ResMM", s + _up(pm, scale=2))


def _stm_memorize(sd, frame, masks):
    frame, _ = _pad16(frame)
    masks, _ = _pad16(masks)
    r4 = _encoder_m(sd, frame, masks)
    p = "trimap.model.KV_M_r4"
    return _conv(sd, p + ".Key", r4, padding=1).unsqueeze(2), _conv(sd, p + ".Value", r4, padding=1).unsqueeze(2)


def _stm_segment(sd, frame, keys, values, read_fn):
    frame, pad = _pad16(frame)
    r4, r3, r2 = _encoder_q(sd, frame)
    p = "trimap.model.KV_Q_r4"
    m4 = read_fn(keys, values, _conv(sd, p + ".Key", r4, padding=1), _conv(sd, p + ".Value", r4, padding=1))
    d = "trimap.model.Decoder"
    m = _resblock(sd, d + ".ResMM", _conv(sd, d + ".convFM", m4, padding=1))
    m = _refine(sd, d + ".RF3", r3, m)
    m = _refine(sd, d + ".RF2", r2, m)
    return _crop(_up(_conv(sd, d + ".pred", F.relu(m), padding=1), scale=4), pad)


# ---- FBA (models/alpha/FBA) ---------------------------------------------------------------------------------------

def _gn_block(sd, p, x, stride, dilation):
    o = F.relu(_gn(sd, p + ".bn1", _ws_conv(sd, p + ".conv1", x)))
    o = F.relu(_gn(sd, p + ".bn2", _ws_conv(sd, p + ".conv2", o, stride=stride, padding=dilation, dilation=dilation)))
    o = _gn(sd, p + ".bn3", _ws_conv(sd, p + ".conv3", o))
    if (p + ".downsample.0.weight") in sd:
        x = _gn(sd, p + ".downsample.1", _ws_conv(sd, p + ".downsample.0", x, stride=stride))
    return F.relu(o + x)


def _fusion(alpha, img, Fg, Bg):
    """fba_fusion, FBA/models.py:279-288"""
    Fn = alpha * img + (1 - alpha ** 2) * Fg - alpha * (1 - alpha) * Bg
    Bn = (1 - alpha) * img + (2 * alpha - alpha ** 2) * Bg - alpha * (1 - alpha) * Fn
    Fn, Bn = torch.clamp(Fn, 0, 1), torch.clamp(Bn, 0, 1)
    alpha = (alpha * 0.1 + torch.sum((img - Bn) * (Fn - Bn), 1, keepdim=True)) / \
            (torch.sum((Fn - Bn) * (Fn - Bn), 1, keepdim=True) + 0.1)
    return torch.cat([torch.clamp(alpha, 0, 1), Fn, Bn], 1)


def _head(raw7, img):
    return _fusion(torch.clamp(raw7[:, 0:1], 0, 1), img, torch.sigmoid(raw7[:, 1:4]), torch.sigmoid(raw7[:, 4:7]))


def _matting(sd, x11, img, two_chan):
    """MattingModule.forward, FBA/models.py:32-45 (refinement=True) -> output [B,7], hid [B,16], refine_output [B,7],
    refine_trimap logits [B,3]"""
    e = "NET.encoder"
    outs = [x11]
    x = F.relu(_gn(sd, e + ".bn1", _ws_conv(sd, e + ".conv1", x11, stride=2, padding=3)))
    outs.append(x)
    x = F.max_pool2d(x, 3, 2, 1)
    for lname, blocks, stride, d0, d in (("layer1", 3, 1, 1, 1), ("layer2", 4, 2, 1, 1), ("layer3", 6, 1, 1, 2),
                                         ("layer4", 3, 1, 2, 4)):
        for b in range(blocks):
            x = _gn_block(sd, f"{e}.{lname}.{b}", x, stride if b == 0 else 1, d0 if b == 0 else d)
        outs.append(x)
    p = "NET.decoder"
    conv5 = outs[-1]
    ppm = [conv5]
    for i, s in enumerate((1, 2, 3, 6)):
        y = F.adaptive_avg_pool2d(conv5, s)
        y = F.leaky_relu(_gn(sd, f"{p}.ppm.{i}.2", _ws_conv(sd, f"{p}.ppm.{i}.1", y)), 0.01)
        ppm.append(_up(y, size=tuple(conv5.shape[2:])))
    x = torch.cat(ppm, 1)
    x = F.leaky_relu(_gn(sd, p + ".conv_up1.1", _ws_conv(sd, p + ".conv_up1.0", x, padding=1)), 0.01)
    x = F.leaky_relu(_gn(sd, p + ".conv_up1.4", _ws_conv(sd, p + ".conv_up1.3", x, padding=1)), 0.01)
    x = torch.cat((_up(x, scale=2), outs[-4]), 1)
    x = F.leaky_relu(_gn(sd, p + ".conv_up2.1", _ws_conv(sd, p + ".conv_up2.0", x, padding=1)), 0.01)
    x = torch.cat((_up(x, scale=2), outs[-5]), 1)
    x = F.leaky_relu(_gn(sd, p + ".conv_up3.1", _ws_conv(sd, p + ".conv_up3.0", x, padding=1)), 0.01)
    x_dec = torch.cat((_up(x, scale=2), outs[-6][:, :3], img), 1)
    h = F.leaky_relu(_conv(sd, p + ".conv_up4.0", torch.cat((x_dec, two_chan), 1), padding=1), 0.01)
    h = F.leaky_relu(_conv(sd, p + ".conv_up4.2", h, padding=1), 0.01)
    output = _head(_conv(sd, p + ".conv_up4.4", h), img)
    r = "NET.refine"
    x = torch.cat((x_dec, two_chan, output[:, :1]), 1)
    x = F.leaky_relu(_gn(sd, r + ".conv1.1", _ws_conv(sd, r + ".conv1.0", x, padding=1)), 0.01)
    for l in ("layer1", "layer2"):
        o = F.relu(_gn(sd, f"{r}.{l}.bn1", _ws_conv(sd, f"{r}.{l}.conv1", x, padding=1)))
        x = F.relu(_gn(sd, f"{r}.{l}.bn2", _ws_conv(sd, f"{r}.{l}.conv2", o, padding=1)) + x)
    x = F.leaky_relu(_conv(sd, r + ".pred.0", x, padding=1), 0.01)
    hid = F.leaky_relu(_conv(sd, r + ".pred.2", x, padding=1), 0.01)
    raw = _conv(sd, r + ".pred.4", hid)
    return output, hid, _head(raw[:, :7], img), raw[:, -3:]


# ---- trimap encoding (models/alpha/model.py:40-53, utils/utils.py:12-39): no gradient through the distance channels ----

def _edt_clicks(t2: torch.Tensor) -> torch.Tensor:
    """[N,2,H,W] {0,1} masks (bg, fg) -> [N,6,H,W] Gaussians of the exact Euclidean distance (host, like the reference)"""
    from scipy import ndimage
    N, _, H, W = t2.shape
    m = t2.detach().cpu().numpy()
    out = np.zeros((N, 6, H, W), np.float32)
    for n in range(N):
        for k in range(2):
            if m[n, k].any():
                d2 = ndimage.distance_transform_edt(m[n, k] == 0).astype(np.float32) ** 2
                for j, s in enumerate((0.02, 0.08, 0.16)):
                    out[n, 3 * k + j] = np.exp(-d2 / (2 * (s * 320) ** 2))
    return torch.from_numpy(out).to(t2.device)


def _trimap8(tri3):
    """tri3 [N,3,H,W] soft (bg, unknown, fg) -> 8 channels: 6 distance Gaussians (of the argmax classes, constants) + the
    soft bg and fg channels (differentiable)"""
    cls = tri3.max(dim=1)[1]
    t2 = torch.stack([(cls == 0).float(), (cls == 2).float()], dim=1)
    return torch.cat([_edt_clicks(t2), tri3[:, 0:1], tri3[:, 2:3]], dim=1)


# ---- losses (utils/loss_func.py) ----------------------------------------------------------------------------------------

def _l1(x, y):
    return torch.mean(torch.abs(x - y))


def _grad_xy(img):
    dy = F.pad(img[:, :, 1:, :] - img[:, :, :-1, :], (0, 0, 0, 1))
    dx = F.pad(img[:, :, :, 1:] - img[:, :, :, :-1], (0, 1, 0, 0))
    return dx, dy


def _l1_grad(pred, gt, eps=1.001e-5):
    px, py = _grad_xy(pred)
    gx, gy = _grad_xy(gt)
    return _l1(torch.sqrt(px ** 2 + py ** 2 + eps), torch.sqrt(gx ** 2 + gy ** 2 + eps))


def _exclusion(a, b, level=3, eps=1.001e-5):
    lx, ly = [], []
    for _ in range(level):
        ax, ay = _grad_xy(a)
        bx, by = _grad_xy(b)
        kx = 2.0 * torch.mean(torch.abs(ax)) / (torch.mean(torch.abs(bx)) + eps)
        ky = 2.0 * torch.mean(torch.abs(ay)) / (torch.mean(torch.abs(by)) + eps)
        sa_x, sa_y = torch.sigmoid(ax) * 2 - 1, torch.sigmoid(ay) * 2 - 1
        sb_x, sb_y = torch.sigmoid(bx * kx) * 2 - 1, torch.sigmoid(by * ky) * 2 - 1
        lx.append((torch.mean(sa_x ** 2 * sb_x ** 2, dim=(1, 2, 3)) + eps) ** 0.25)
        ly.append((torch.mean(sa_y ** 2 * sb_y ** 2, dim=(1, 2, 3)) + eps) ** 0.25)
        a, b = F.avg_pool2d(a, 2, 2), F.avg_pool2d(b, 2, 2)
    return torch.mean(sum(lx) / float(level)) + torch.mean(sum(ly) / float(level))


def _lap_pyramid(img, kernel, levels=5):
    C = img.shape[1]
    k = kernel.repeat(C, 1, 1, 1)
    gauss = lambda x, kk: F.conv2d(F.pad(x, (2, 2, 2, 2), mode="reflect"), kk, groups=C)
    cur, pyr = img, []
    for _ in range(levels):
        down = gauss(cur, k)[:, :, ::2, ::2]
        up = torch.zeros(down.shape[0], C, down.shape[2] * 2, down.shape[3] * 2, device=img.device, dtype=img.dtype)
        up[:, :, ::2, ::2] = down                           # zeros interleaved (pyrUp), then 4 x Gaussian
        pyr.append(cur - gauss(up, 4 * k))
        cur = down
    return pyr


def _lap_loss(img, tgt, kernel):
    """LapLoss.forward with normalize=True (inputs here are multiples of 32: no padding branch)"""
    assert img.shape[-2] % 32 == 0 and img.shape[-1] % 32 == 0, "training crops are multiples of 32 (config.py:27)"
    loss = sum((2 ** l) * torch.sum(torch.abs(a - b)) for l, (a, b) in
               enumerate(zip(_lap_pyramid(img, kernel), _lap_pyramid(tgt, kernel))))
    return loss / tgt.numel()


def _fba_loss(preds, trimasks, gts, fgs, bgs, imgs, lap_kernel):
    """FullModel.fba_single_image_loss (models/alpha/model.py:107-187) with normalize=True -> (L_alpha_comp, L_lap, L_grad)"""
    S = preds.shape[1]
    L1s, Lls, Lgs, al, Fs, Bs = [], [], [], [], [], []
    for c in range(S):
        gt, tm, img, a = gts[:, c], trimasks[:, c].bool(), imgs[:, c], preds[:, c, :1]
        cF = torch.where((tm & (gt > 0)).repeat(1, 3, 1, 1), preds[:, c, 1:4], fgs[:, c])
        cB = torch.where(tm.repeat(1, 3, 1, 1), preds[:, c, 4:], bgs[:, c])
        L_a1 = _l1(a, gt)
        L_ac = _l1(cF * gt + cB * (1. - gt), img)
        L_FBc = _l1(fgs[:, c] * a + bgs[:, c] * (1. - a), img)
        L_FB1 = _l1(cF, fgs[:, c]) + _l1(cB, bgs[:, c])
        L1s.append(L_a1 + L_ac + 0.25 * (L_FBc + L_FB1))
        Lgs.append(_l1_grad(a, gt) + 0.25 * _exclusion(cF, cB))
        Lls.append(_lap_loss(a, gt, lap_kernel) + 0.25 * (_lap_loss(cF, fgs[:, c], lap_kernel) + _lap_loss(cB, bgs[:, c], lap_kernel)))
        al.append(a); Fs.append(cF); Bs.append(cB)
    L1, Ll, Lg = sum(L1s) / S, sum(Lls) / S, sum(Lgs) / S
    if S > 1:
        al, Fs, Bs = torch.stack(al, 1), torch.stack(Fs, 1), torch.stack(Bs, 1)
        Lg = Lg + F.mse_loss(al[:, 1:] - al[:, :-1], gts[:, 1:] - gts[:, :-1]) + 0.25 * (
            F.mse_loss(Fs[:, 1:] - Fs[:, :-1], fgs[:, 1:] - fgs[:, :-1]) + F.mse_loss(Bs[:, 1:] - Bs[:, :-1], bgs[:, 1:] - bgs[:, :-1]))
    return L1, Ll, Lg


# ------------------------------------------------------------------------------------------------------------------
class Stage4Model(nn.Module):
    """``FullModel`` of ``models/alpha/model.py:10`` at stage 4 wrapping ``FullModel`` of ``models/trimap/model.py:15``: same
    785 ``state_dict`` keys, same ``forward(a, fg, bg, ignore_region=None, tri=None)`` and the same first four outputs
    (L_alpha_comp, L_lap, L_grad, L_trimap) followed by the refined alphas and the predicted trimaps."""

    def __init__(self, read_fn: Optional[Callable] = None):
        super().__init__()
        _grow(self, "")
        self.trimap = nn.Module()
        _grow(self.trimap, "trimap.")
        for p in self.parameters():
            p.requires_grad_(True)
        self.read_fn = read_fn or T.memory_read
        self.IMG_SCALE = 1.0 / 255

    def _sd(self) -> Dict[str, torch.Tensor]:
        sd = dict(self.named_parameters())
        sd.update(dict(self.named_buffers()))
        return sd

    def forward(self, a, fg, bg, ignore_region=None, tri=None):
        sd = self._sd()
        B, S, _, H, W = a.shape
        with torch.no_grad():                                   # preprocess, models/alpha/model.py:55-65
            gts = a
            fgs, bgs = fg.flip([2]) * self.IMG_SCALE, bg.flip([2]) * self.IMG_SCALE
            imgs = fgs * gts + bgs * (1. - gts)
            cls = tri.float().max(dim=2)[1].unsqueeze(2).float() * 0.5
            trimasks = ((cls > 0) & (cls < 1)).float()
            imgs_n = (imgs - sd["IMG_MEAN"]) / sd["IMG_STD"]
        tri_prop = [tri[:, 0]] + [None] * (S - 1)               # propagated trimaps (input of the alpha network)
        tri_ref = [tri[:, 0]] + [None] * (S - 1)                # refined trimaps (input of the memory)
        out_dec, out_ref, logit_prop, logit_ref = [], [], [], []
        keys = vals = None
        for t in range(S):                                      # :205-246
            tri8 = _trimap8(tri_prop[t])
            output, hid, refine_output, refine_logit = _matting(sd, torch.cat([imgs_n[:, t], tri8], 1), imgs[:, t], tri8[:, -2:])
            out_dec.append(output); out_ref.append(refine_output); logit_ref.append(refine_logit)
            if t > 0:
                tri_ref[t] = F.softmax(refine_logit, dim=1)
            if t < S - 1:                                       # trimap single step, models/trimap/model.py:133-154
                k4, v4 = _stm_memorize(sd, imgs[:, t], torch.cat([tri_ref[t], refine_output[:, :1], hid], 1))
                keys = k4 if keys is None else torch.cat([keys, k4], 2)
                vals = v4 if vals is None else torch.cat([vals, v4], 2)
                logit = _stm_segment(sd, imgs[:, t + 1], keys, vals, self.read_fn)
                logit_prop.append(logit)
                tri_prop[t + 1] = F.softmax(logit, dim=1)
        out_dec, out_ref = torch.stack(out_dec, 1), torch.stack(out_ref, 1)
        la = _fba_loss(out_dec, trimasks, gts, fgs, bgs, imgs, sd["LAPLOSS.KERNEL"])
        lb = _fba_loss(out_ref, trimasks, gts, fgs, bgs, imgs, sd["LAPLOSS.KERNEL"])
        gt_cls = torch.argmax(tri, dim=2)
        l_tri = F.cross_entropy(torch.stack(logit_prop, 1).reshape(-1, 3, H, W), gt_cls[:, 1:].reshape(-1, H, W)) + \
            F.cross_entropy(torch.stack(logit_ref, 1).reshape(-1, 3, H, W), gt_cls.reshape(-1, H, W))
        return [la[0] + lb[0], la[1] + lb[1], la[2] + lb[2], l_tri, out_ref[:, :, :1], torch.stack(tri_ref, 1)]


def step(model, opt, sample, autocast_dtype=torch.bfloat16):
    """train.py:349-375: forward, loss = sum of the four means, zero_grad, backward (DDP all-reduces), optimiser step"""
    a, fg, bg, tri = sample
    with torch.autocast(device_type=a.device.type, dtype=autocast_dtype, enabled=autocast_dtype is not None):
        out = model(a, fg, bg, ignore_region=None, tri=tri)
    losses = [o.float().mean() for o in out[:4]]
    loss = sum(losses)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    # ONE small tensor carries all scalars of the step: the reference all-reduces five scalars one by one and adds a
    # barrier every step (train.py:377-388); a caller that logs needs a single all_reduce of this tensor, if any
    return torch.stack([loss.detach()] + [l.detach() for l in losses])
