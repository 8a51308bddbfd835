"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; never imported by ``otvm_b200``.

A CPU restatement (plain PyTorch fp32 functional ops + numpy/scipy, NCHW, no custom kernels) of the
per-frame inference hot path of Hongje/OTVM, written from the behaviour of the reference, each function
citing the reference ``file:line`` it follows (paths relative to the reference root).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.

Parity status: the reference has NO golden vectors or tests for this path, so the
oracle is pinned by running the UNMODIFIED reference in the build container on the seeded weights/frames of
``otvm_b200/fixtures.py`` and committing its outputs under ``tests/golden/`` (``oracle/make_golden.py`` is the
generating script).  ``tests/test_oracle_golden.py`` checks this file against those vectors.

Third-party arithmetic on the path that is not under the reference tree: ``torch`` conv / GEMM / softmax /
group_norm / interpolate (reference pins "pytorch 1.8.2", README.md:11-14; 2.11.0 here), the torchvision
ResNet-50 topology (``models/trimap/STM.py:43,79``) and ``cv2.distanceTransform(DIST_L2, 0)``
(``utils/utils.py:21``), which is the exact Euclidean distance transform; it is restated here with
``scipy.ndimage.distance_transform_edt`` (checked equal to cv2 in ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from scipy import ndimage

# --------------------------------------------------------------------------------------------------
# primitive layers
# --------------------------------------------------------------------------------------------------

def conv2d(sd, name, x, stride=1, padding=0, dilation=1):
    """nn.Conv2d forward with the named weight (+bias when present)."""
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride, padding, dilation)


def ws_weight(w):
    """Weight standardisation, models/alpha/FBA/layers_WS.py:15-21: per output channel subtract the
    mean over (cin,kh,kw), divide by sqrt(unbiased var + 1e-12) + 1e-5."""
    m = w.mean(dim=1, keepdim=True).mean(dim=2, keepdim=True).mean(dim=3, keepdim=True)   # same order as :16-17
    w = w - m
    std = torch.sqrt(torch.var(w.flatten(1), dim=1) + 1e-12).view(-1, 1, 1, 1) + 1e-5
    return w / std


def ws_conv2d(sd, name, x, stride=1, padding=0, dilation=1):
    """models/alpha/FBA/layers_WS.py:13-23."""
    return F.conv2d(x, ws_weight(sd[name + ".weight"]), sd.get(name + ".bias"), stride, padding, dilation)


def batchnorm_eval(sd, name, x, eps=1e-5):
    """nn.BatchNorm2d in eval mode (STM ResNets; eval.py runs model.eval())."""
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                        sd[name + ".weight"], sd[name + ".bias"], False, 0.0, eps)


def groupnorm32(sd, name, x, eps=1e-5):
    """layers_WS.BatchNorm2d == nn.GroupNorm(32, C), models/alpha/FBA/layers_WS.py:26-27."""
    return F.group_norm(x, 32, sd[name + ".weight"], sd[name + ".bias"], eps)


def up_bilinear(x, scale=None, size=None):
    return F.interpolate(x, size=size, scale_factor=scale, mode="bilinear", align_corners=False)


# --------------------------------------------------------------------------------------------------
# STM (models/trimap/STM.py)
# --------------------------------------------------------------------------------------------------

def _tv_bottleneck(sd, p, x, stride):
    """torchvision Bottleneck (v1.5: stride on the 3x3), eval-mode BN."""
    o = F.relu(batchnorm_eval(sd, p + ".bn1", conv2d(sd, p + ".conv1", x)))
    o = F.relu(batchnorm_eval(sd, p + ".bn2", conv2d(sd, p + ".conv2", o, stride=stride, padding=1)))
    o = batchnorm_eval(sd, p + ".bn3", conv2d(sd, p + ".conv3", o))
    if (p + ".downsample.0.weight") in sd:
        x = batchnorm_eval(sd, p + ".downsample.1", conv2d(sd, p + ".downsample.0", x, stride=stride))
    return F.relu(o + x)


def _tv_layers(sd, p, x):
    """resnet.layer1..layer3 registered as res2/res3/res4, STM.py:49-51."""
    outs = []
    for lname, blocks, stride in (("res2", 3, 1), ("res3", 4, 2), ("res4", 6, 2)):
        for b in range(blocks):
            x = _tv_bottleneck(sd, f"{p}.{lname}.{b}", x, stride if b == 0 else 1)
        outs.append(x)
    return outs  # r2, r3, r4


def encoder_q(sd, frame, p="trimap.model.Encoder_Q"):
    """Encoder_Q.forward, STM.py:92-102.  frame: [N,3,H,W] RGB in [0,1]."""
    f = (frame - sd[p + ".mean"]) / sd[p + ".std"]
    x = conv2d(sd, p + ".conv1", f, stride=2, padding=3)
    c1 = F.relu(batchnorm_eval(sd, p + ".bn1", x))
    x = F.max_pool2d(c1, 3, 2, 1)
    r2, r3, r4 = _tv_layers(sd, p, x)
    return r4, r3, r2


def encoder_m(sd, frame, in_m, in_o, in_a, in_h, p="trimap.model.Encoder_M"):
    """Encoder_M.forward, STM.py:56-74: five 7x7 s2 stems summed before bn1."""
    f = (frame - sd[p + ".mean"]) / sd[p + ".std"]
    x = (conv2d(sd, p + ".conv1_m", in_m.unsqueeze(1), stride=2, padding=3)
         + conv2d(sd, p + ".conv1_o", in_o.unsqueeze(1), stride=2, padding=3)
         + conv2d(sd, p + ".conv1_a", in_a.unsqueeze(1), stride=2, padding=3)
         + conv2d(sd, p + ".conv1_h", in_h, stride=2, padding=3))
    x = conv2d(sd, p + ".conv1", f, stride=2, padding=3) + x
    c1 = F.relu(batchnorm_eval(sd, p + ".bn1", x))
    x = F.max_pool2d(c1, 3, 2, 1)
    r2, r3, r4 = _tv_layers(sd, p, x)
    return r4


def key_value(sd, p, r4):
    """KeyValue.forward, STM.py:173-174."""
    return conv2d(sd, p + ".Key", r4, padding=1), conv2d(sd, p + ".Value", r4, padding=1)


def memory_read(m_in, m_out, q_in, q_out):
    """Memory.forward, STM.py:144-163.
    m_in [B,De,T,h,w] keys, m_out [B,Do,T,h,w] values, q_in [B,De,h,w], q_out [B,Do,h,w] -> [B,2*Do,h,w]."""
    B, De, T, h, w = m_in.shape
    Do = m_out.shape[1]
    mi = m_in.reshape(B, De, T * h * w).transpose(1, 2)          # [B, THW, De]
    qi = q_in.reshape(B, De, h * w)                              # [B, De, HW]
    p = torch.bmm(mi, qi) / math.sqrt(De)                        # [B, THW, HW]
    p = F.softmax(p, dim=1)                                      # over the THW memory locations
    mem = torch.bmm(m_out.reshape(B, Do, T * h * w), p).view(B, Do, h, w)
    return torch.cat([mem, q_out], dim=1)


def _resblock(sd, p, x):
    """ResBlock.forward, STM.py:23-30 (indim == outdim, no downsample in the decoder)."""
    r = conv2d(sd, p + ".conv1", F.relu(x), padding=1)
    r = conv2d(sd, p + ".conv2", F.relu(r), padding=1)
    return x + r


def _refine(sd, p, f, pm):
    """Refine.forward, STM.py:113-117."""
    s = _resblock(sd, p + ".ResFS", conv2d(sd, p + ".convFS", f, padding=1))
    m = s + up_bilinear(pm, scale=2)
    return _resblock(sd, p + ".ResMM", m)


def stm_decoder(sd, m4in, r3, r2, p="trimap.model.Decoder"):
    """Decoder.forward, STM.py:129-137."""
    m4 = _resblock(sd, p + ".ResMM", conv2d(sd, p + ".convFM", m4in, padding=1))
    m3 = _refine(sd, p + ".RF3", r3, m4)
    m2 = _refine(sd, p + ".RF2", r2, m3)
    p2 = conv2d(sd, p + ".pred", F.relu(m2), padding=1)
    return up_bilinear(p2, scale=4)


def pad_to(x, d, value=0.0):
    """pad_divide_by, helpers.py:24-40 / models/alpha/common.py:6-27 (centred padding)."""
    h, w = x.shape[-2:]
    nh = h + (d - h % d) % d
    nw = w + (d - w % d) % d
    lh, lw = (nh - h) // 2, (nw - w) // 2
    pad = (lw, nw - w - lw, lh, nh - h - lh)
    if sum(pad) > 0:
        x = F.pad(x, pad, value=value)
    return x, pad


def crop(x, pad):
    lw, uw, lh, uh = pad
    if lh + uh > 0:
        x = x[:, :, lh:x.shape[2] - uh, :]
    if lw + uw > 0:
        x = x[:, :, :, lw:x.shape[3] - uw]
    return x


def stm_segment(sd, frame, keys, values):
    """STM.segment, STM.py:239-257.  keys/values: [1,C,T,h,w] (already squeezed)."""
    frame, pad = pad_to(frame, 16)
    r4, r3, r2 = encoder_q(sd, frame)
    k4, v4 = key_value(sd, "trimap.model.KV_Q_r4", r4)
    m4 = memory_read(keys, values, k4, v4)
    return crop(stm_decoder(sd, m4, r3, r2), pad), dict(k4=k4, v4=v4, m4=m4, r4=r4)


def stm_memorize(sd, frame, masks):
    """STM.memorize, STM.py:201-228 with masks = cat(tri3, alpha1, hid16), models/trimap/model.py:231.
    Returns k4 [1,128,1,h,w], v4 [1,512,1,h,w]."""
    frame, _ = pad_to(frame, 16)
    masks, _ = pad_to(masks, 16)
    r4 = encoder_m(sd, frame, masks[:, 1], masks[:, 2], masks[:, 3], masks[:, 4:])
    k4, v4 = key_value(sd, "trimap.model.KV_M_r4", r4)
    return k4.unsqueeze(2), v4.unsqueeze(2)


# --------------------------------------------------------------------------------------------------
# FBA (models/alpha/FBA)
# --------------------------------------------------------------------------------------------------

def _gn_bottleneck(sd, p, x, stride, dilation):
    """resnet_GN_WS.Bottleneck.forward :69-88 after ResnetDilated._nostride_dilate (FBA/models.py:236-249)."""
    o = F.relu(groupnorm32(sd, p + ".bn1", ws_conv2d(sd, p + ".conv1", x)))
    o = F.relu(groupnorm32(sd, p + ".bn2",
                           ws_conv2d(sd, p + ".conv2", o, stride=stride, padding=dilation, dilation=dilation)))
    o = groupnorm32(sd, p + ".bn3", ws_conv2d(sd, p + ".conv3", o))
    if (p + ".downsample.0.weight") in sd:
        x = groupnorm32(sd, p + ".downsample.1", ws_conv2d(sd, p + ".downsample.0", x, stride=stride))
    return F.relu(o + x)


# (layer, blocks, stride of block 0, dilation of block 0's 3x3, dilation of later blocks)
_FBA_LAYERS = (("layer1", 3, 1, 1, 1), ("layer2", 4, 2, 1, 1), ("layer3", 6, 1, 1, 2), ("layer4", 3, 1, 2, 4))


def fba_encoder(sd, x, p="NET.encoder"):
    """ResnetDilated.forward, FBA/models.py:251-269 (dilate_scale=8)."""
    outs = [x]
    x = F.relu(groupnorm32(sd, p + ".bn1", ws_conv2d(sd, p + ".conv1", x, stride=2, padding=3)))
    outs.append(x)
    x = F.max_pool2d(x, 3, 2, 1)
    for lname, blocks, stride, d0, d in _FBA_LAYERS:
        for b in range(blocks):
            x = _gn_bottleneck(sd, f"{p}.{lname}.{b}", x, stride if b == 0 else 1, d0 if b == 0 else d)
        outs.append(x)
    return outs


def fba_fusion(alpha, img, Fg, Bg):
    """fba_fusion, FBA/models.py:279-288.  NB :281 reads the F already updated by :280 (before its clamp)."""
    Fn = alpha * img + (1 - alpha ** 2) * Fg - alpha * (1 - alpha) * Bg
    Bn = (1 - alpha) * img + (2 * alpha - alpha ** 2) * Bg - alpha * (1 - alpha) * Fn
    Fn = torch.clamp(Fn, 0, 1)
    Bn = torch.clamp(Bn, 0, 1)
    la = 0.1
    alpha = (alpha * la + torch.sum((img - Bn) * (Fn - Bn), 1, keepdim=True)) / \
            (torch.sum((Fn - Bn) * (Fn - Bn), 1, keepdim=True) + la)
    return torch.clamp(alpha, 0, 1), Fn, Bn


def _head(out7, img):
    alpha = torch.clamp(out7[:, 0:1], 0, 1)
    Fg = torch.sigmoid(out7[:, 1:4])
    Bg = torch.sigmoid(out7[:, 4:7])
    return torch.cat(fba_fusion(alpha, img, Fg, Bg), 1)


def fba_decoder(sd, conv_out, img, two_chan, p="NET.decoder"):
    """fba_decoder.forward, FBA/models.py:351-392."""
    conv5 = conv_out[-1]
    hw = conv5.shape[2:]
    ppm = [conv5]
    for i, s in enumerate((1, 2, 3, 6)):
        y = F.adaptive_avg_pool2d(conv5, s)
        y = F.leaky_relu(groupnorm32(sd, f"{p}.ppm.{i}.2", ws_conv2d(sd, f"{p}.ppm.{i}.1", y)), 0.01)
        ppm.append(up_bilinear(y, size=tuple(hw)))
    x = torch.cat(ppm, 1)
    x = F.leaky_relu(groupnorm32(sd, p + ".conv_up1.1", ws_conv2d(sd, p + ".conv_up1.0", x, padding=1)), 0.01)
    x = F.leaky_relu(groupnorm32(sd, p + ".conv_up1.4", ws_conv2d(sd, p + ".conv_up1.3", x, padding=1)), 0.01)
    x = torch.cat((up_bilinear(x, scale=2), conv_out[-4]), 1)
    x = F.leaky_relu(groupnorm32(sd, p + ".conv_up2.1", ws_conv2d(sd, p + ".conv_up2.0", x, padding=1)), 0.01)
    x = torch.cat((up_bilinear(x, scale=2), conv_out[-5]), 1)
    x = F.leaky_relu(groupnorm32(sd, p + ".conv_up3.1", ws_conv2d(sd, p + ".conv_up3.0", x, padding=1)), 0.01)
    x = torch.cat((up_bilinear(x, scale=2), conv_out[-6][:, :3], img), 1)          # 70 ch
    x2 = torch.cat((x, two_chan), 1)                                               # 72 ch
    h = F.leaky_relu(conv2d(sd, p + ".conv_up4.0", x2, padding=1), 0.01)           # plain nn.Conv2d
    hid = F.leaky_relu(conv2d(sd, p + ".conv_up4.2", h, padding=1), 0.01)
    raw = conv2d(sd, p + ".conv_up4.4", hid)
    return hid, _head(raw, img), x, raw


def _gn_basicblock(sd, p, x):
    """resnet_GN_WS.BasicBlock.forward :32-48."""
    o = F.relu(groupnorm32(sd, p + ".bn1", ws_conv2d(sd, p + ".conv1", x, padding=1)))
    o = groupnorm32(sd, p + ".bn2", ws_conv2d(sd, p + ".conv2", o, padding=1))
    return F.relu(o + x)


def fba_refine(sd, x_dec, img, two_chan, pred_alpha, p="NET.refine"):
    """RefinementModule.forward, FBA/models.py:417-435."""
    x = torch.cat((x_dec, two_chan, pred_alpha), 1)                                # 73 ch
    x = F.leaky_relu(groupnorm32(sd, p + ".conv1.1", ws_conv2d(sd, p + ".conv1.0", x, padding=1)), 0.01)
    x = _gn_basicblock(sd, p + ".layer1", x)
    x = _gn_basicblock(sd, p + ".layer2", x)
    x = F.leaky_relu(conv2d(sd, p + ".pred.0", x, padding=1), 0.01)
    hid = F.leaky_relu(conv2d(sd, p + ".pred.2", x, padding=1), 0.01)
    raw = conv2d(sd, p + ".pred.4", hid)
    return hid, _head(raw[:, :7], img), raw[:, -3:], raw


def matting_forward(sd, x11, img, two_chan):
    """MattingModule.forward, FBA/models.py:32-45 (refinement=True)."""
    conv_out = fba_encoder(sd, x11)
    hid_d, output, x_dec, raw_d = fba_decoder(sd, conv_out, img, two_chan)
    hid, refine_output, refine_trimap, raw_r = fba_refine(sd, x_dec, img, two_chan, output[:, :1])
    return dict(output=output, hid=hid, refine_output=refine_output, refine_trimap=refine_trimap,
                raw_decoder=raw_d, raw_refine=raw_r, conv5=conv_out[-1])


# --------------------------------------------------------------------------------------------------
# trimap encoding with the host EDT (utils/utils.py:12-39, models/alpha/model.py:40-53)
# --------------------------------------------------------------------------------------------------

def edt_sq(mask_nonzero: np.ndarray) -> np.ndarray:
    """Exact squared Euclidean distance (int64) from every pixel to the nearest ZERO pixel of the input,
    i.e. what cv2.distanceTransform(src, DIST_L2, DIST_MASK_PRECISE) computes before its sqrt."""
    _, idx = ndimage.distance_transform_edt(mask_nonzero, return_indices=True)
    yy, xx = np.indices(mask_nonzero.shape)
    return (idx[0] - yy).astype(np.int64) ** 2 + (idx[1] - xx).astype(np.int64) ** 2


def trimap_transform(trimap2):
    """utils/utils.py:25-39.  trimap2: [N,2,H,W] in {0,1} (bg mask, fg mask) -> [N,6,H,W]."""
    N, _, H, W = trimap2.shape
    dev = trimap2.device
    trimap2 = trimap2.cpu()                 # the distance transform runs on the HOST, like utils/utils.py:12-23 (cv2)
    clicks = torch.zeros(N, 6, H, W)
    L = 320
    for n in range(N):
        for k in range(2):
            tk = trimap2[n, k]
            if int((tk != 0).sum()) > 0:
                src = ((1.0 - tk).numpy() * 255).astype(np.uint8)
                d = torch.from_numpy(np.sqrt(edt_sq(src != 0).astype(np.float32)))
                dm = -d ** 2
                for j, s in enumerate((0.02, 0.08, 0.16)):
                    clicks[n, 3 * k + j] = torch.exp(dm / (2 * ((s * L) ** 2)))
    return clicks.to(dev)


def make_trimap8(tri3):
    """FullModel.make_trimap (TRIMAP_CHANNEL == 8), models/alpha/model.py:40-53.  tri3 [N,3,H,W] soft
    (bg, unknown, fg) -> [N,8,H,W] = 6 distance channels + soft bg + soft fg."""
    cls = tri3.max(dim=1)[1]
    t2 = torch.stack([(cls == 0).float(), (cls == 2).float()], dim=1)
    return torch.cat([trimap_transform(t2), tri3[:, 0:1], tri3[:, 2:3]], dim=1)


def trimap_from_alpha(alpha, radius):
    """EvalModel.make_trimap_gt with trimap3=None, models/alpha/model.py:342-362 (EPS = 0).
    alpha [N,1,H,W] -> one-hot [N,3,H,W] (bg, unknown, fg)."""
    alpha = torch.where(alpha < 0, torch.zeros_like(alpha), alpha)
    alpha = torch.where(alpha > 1, torch.ones_like(alpha), alpha)
    unk = ((alpha > 0) & (alpha < 1)).float()
    unk = F.max_pool2d(unk, kernel_size=2 * radius + 1, stride=1, padding=radius)
    cls = torch.where(unk > 0.5, torch.ones_like(alpha), 2 * alpha).long()
    return F.one_hot(cls.squeeze(1), 3).permute(0, 3, 1, 2).float()


# --------------------------------------------------------------------------------------------------
# EvalModel.forward (models/alpha/model.py:391-512): the per-frame state machine
# --------------------------------------------------------------------------------------------------

class OracleEvalModel:
    """Same call contract as the reference ``EvalModel`` at stage 4, including the one-trimap inputs ``tri=`` (a BGR
    0..255 trimap image) and ``tri_gt=`` (a one-hot / soft 3-channel trimap) that seed frame 0, models/alpha/model.py:395-401."""

    IMG_SCALE = 1.0 / 255

    def __init__(self, state_dict, dilate_kernel=12, device="cpu"):
        """``device='cuda'`` runs the same PyTorch ops through cuDNN / cuBLAS -- the reference's own GPU path (fp32 eager,
        host distance transform); used by bench.py as the "PyTorch eager on the same GPU" baseline."""
        self.sd = {k: (v.detach().float() if v.is_floating_point() else v).to(device) for k, v in state_dict.items()}
        self.radius = dilate_kernel
        self.memories = None
        self.trace = {}

    @torch.no_grad()
    def __call__(self, a, fg, bg, tri=None, tri_gt=None, first_frame=False, last_frame=False,
                 memorize=False, max_memory_num=2, large_input=False):
        sd = self.sd
        # preprocess_gt :380-389 (the trimap_transform at :377 is dead work: tris_gt is never read)
        gts = a
        fgs = fg.flip([2]) * self.IMG_SCALE
        bgs = bg.flip([2]) * self.IMG_SCALE
        scaled_imgs = fgs * gts + bgs * (1.0 - gts)
        tri3_gt = torch.stack([trimap_from_alpha(gts[b], self.radius) for b in range(gts.shape[0])])
        if tri is not None:                                # :395-396 user trimap image, BGR 0..255 -> (bg, un, fg) in [0,1]
            tri = tri.flip([2]) * self.IMG_SCALE
        elif tri_gt is not None:                           # :397-399 make_trimap_gt(None, trimap3=tri): argmax -> one-hot
            tri = tri_gt
            tri3_gt = F.one_hot(tri_gt.max(dim=2)[1], 3).permute(0, 1, 4, 2, 3).float()
        else:                                              # :400-401
            tri = tri3_gt
        img = scaled_imgs.squeeze(0)                       # [1,3,H,W]
        tri_ = tri.squeeze(0)                              # [1,3,H,W]
        img, pad = pad_to(img, 32)                         # :408
        if sum(pad) > 0:                                   # :409-410 pad bg with 1, others with 0
            tri_ = torch.cat((F.pad(tri_[:, :1], pad, value=1.0), F.pad(tri_[:, 1:], pad, value=0.0)), 1)
        mean, std = sd["IMG_MEAN"].squeeze(0), sd["IMG_STD"].squeeze(0)
        img_n = (img - mean) / std                         # :414
        tr = self.trace = {}
        if first_frame:                                    # :424-430
            self.memories = {"key": None, "val": None}
            preds_trimap = tri_
        else:                                              # :431-443
            logit, aux = stm_segment(sd, img, self.memories["key"].squeeze(1), self.memories["val"].squeeze(1))
            tr.update(seg_logit=logit, **aux)
            preds_trimap = F.softmax(logit, dim=1)
        tri8 = make_trimap8(preds_trimap)                  # :416 / :442
        x11 = torch.cat([img_n, tri8], dim=1)              # :445
        net = matting_forward(sd, x11, img, tri8[:, -2:])  # :449
        tr.update(tri8=tri8, **net)
        preds_alpha = net["refine_output"][:, :1]          # :452
        preds_trimap = F.softmax(net["refine_trimap"], dim=1)   # :460
        if not last_frame:                                 # :461-493
            masks = torch.cat([preds_trimap, preds_alpha, net["hid"]], dim=1)     # trimap/model.py:231
            k4, v4 = stm_memorize(sd, img, masks)
            new = {"key": k4.unsqueeze(1), "val": v4.unsqueeze(1)}                # [1,1,C,1,h,w]
            tr.update(mem_k=k4, mem_v=v4)
            self._update_bank(new, first_frame, memorize, max_memory_num)
        preds_trimap = crop(preds_trimap, pad).unsqueeze(0)
        preds_alpha = crop(preds_alpha, pad).unsqueeze(0)
        return scaled_imgs, preds_trimap, tri3_gt, preds_alpha, gts

    def _update_bank(self, new, first_frame, memorize, max_memory_num):
        """Memory-bank policy, models/alpha/model.py:472-493."""
        m = self.memories
        if max_memory_num == 0:
            if first_frame:
                self.memories = new
        elif max_memory_num == 1:
            self.memories = new
        else:
            if first_frame:
                self.memories = new
            else:
                for k in ("key", "val"):
                    if memorize or m[k].size(3) == 1:
                        m[k] = torch.cat([m[k], new[k]], dim=3)
                    else:
                        m[k] = torch.cat([m[k][:, :, :, :-1], new[k]], dim=3)
            m = self.memories
            if m["key"].size(3) > max_memory_num:
                for k in ("key", "val"):
                    m[k] = torch.cat([m[k][:, :, :, :1], m[k][:, :, :, 2:]], dim=3)
