"""Thin tensor-level wrappers over the C ABI (``include/otvm_b200.h``).

Activations are NHWC torch tensors ``[N, H, W, C]`` that may be channel slices of a wider buffer
(``stride(2)`` is the per-pixel leading dimension).  Every wrapper launches on torch's current CUDA
stream and raises :class:`otvm_b200._lib.OtvmError` on failure.  No wrapper computes anything in PyTorch.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib, split
from ._lib import ACT_LEAKY, ACT_NONE, ACT_RELU, BF16, F32, ConvParams, ReadBwdParams, ReadParams, check  # noqa: F401


def _dt(t) -> int:
    """dtype word of the C ABI for an activation tensor: OTVM_F32, OTVM_BF16, or OTVM_SPLIT_DTYPE(planes, stride) when
    the bf16 tensor is the plane-0 view of a split tensor (:mod:`otvm_b200.split`)"""
    if t.dtype == torch.float32:
        return F32
    assert t.dtype == torch.bfloat16, t.dtype
    a = split.arena_of(t)
    return a.dtype_word if a is not None else BF16


DRY = False            # planning pass (Engine._measure): wrappers allocate / return without calling the library


class Profiler:
    """Optional per-op CUDA-event timing (bench.py): op family -> [events], flops, bytes."""

    def __init__(self):
        self.spans = {}

    def span(self, key, flops=0.0, nbytes=0.0):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.spans.setdefault(key, []).append((e0, e1, flops, nbytes))
        return e0, e1

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for k, v in self.spans.items():
            out[k] = dict(calls=len(v), ms=sum(a.elapsed_time(b) for a, b, _, _ in v),
                          flops=sum(f for _, _, f, _ in v), bytes=sum(n for _, _, _, n in v))
        return out


GN_FUSE = os.environ.get("OTVM_GN_FUSE", "1") != "0"      # GroupNorm inside the convolution kernel where possible
PROFILER: "Profiler | None" = None
PROFILE_SHAPES = False          # dev: one profiler key per conv shape


def device_error_flags(clear: bool = True) -> int:
    """sticky device-side error flags (synchronises): bit 0 = a GroupNorm-fused convolution's grid barrier timed out"""
    return int(_lib.load().otvm_device_error_flags(int(clear)))


def launch_count() -> int:
    return int(_lib.load().otvm_launch_count())


def _timed(key, fn, flops=0.0, nbytes=0.0):
    if PROFILER is None:
        return fn()
    e0, e1 = PROFILER.span(key, flops, nbytes)
    e0.record()
    r = fn()
    e1.record()
    return r


def _planes(t):
    return max(1, _dt(t) & 0xff) if t.dtype == torch.bfloat16 else 1


def _nbytes(*ts):
    return float(sum(t.shape.numel() * t.element_size() * _planes(t) for t in ts if t is not None))


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ld(t: torch.Tensor) -> int:
    """per-pixel leading dimension of an NHWC tensor (possibly a channel slice)"""
    assert t.dim() == 4 and t.stride(3) == 1, (t.shape, t.stride())
    ld = t.stride(2)
    assert t.shape[1] == 1 or t.stride(1) == t.shape[2] * ld
    assert t.shape[0] == 1 or t.stride(0) == t.shape[1] * t.shape[2] * ld
    return ld


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def conv2d(x, w, bias, out, *, stride=1, pad=0, dil=1, act=ACT_NONE, relu_in=False, res=None, out_relu=None,
           gn_stats=None, gn_stats_zeroed=False, out_strides=None, cin=None, workspace=None, gn_fuse=None, gn_raw_out=None,
           gn_group_ch=0, query_fuse=False, groups=1):
    """x [N,H,W,Cin(view)], w [Cout,KH,KW,Cin] packed, out NHWC view (or any buffer with ``out_strides`` =
    (pixel_stride, channel_stride) in elements, used for the channel-major value bank).

    ``gn_fuse=(gamma, beta, eps)``: GroupNorm(32) + affine (+res, +act) inside the convolution kernel when the
    library can (one co-resident wave, see include/otvm_b200.h); the call then returns ``True``.  When it cannot,
    nothing of the normalisation is done, the raw convolution (+ statistics) is written to ``gn_raw_out`` and
    ``False`` is returned (the caller follows with :func:`gn_apply`).  Without ``gn_fuse`` the function returns
    ``out``.

    Channel slices of a wide normalised layer: pass the ``[a:b]`` views of the packed weights (dim 1 of split weights),
    gamma / beta / out / res, ``gn_group_ch`` = full Cout / 32 and a statistics slot of its own per slice
    (``otvm_conv_params.gn_group_ch``).  ``query_fuse=True`` only answers whether the fused path would be taken.

    ``groups=G``: ``w`` stacks G filter banks along its Cout axis ([G*Cout][KH][KW][Cin]), ``bias`` is [G*Cout], and image
    n of the batch (N == G) uses bank n (``otvm_conv_params.groups``)."""
    if DRY:
        return out if gn_fuse is None else True
    lib = _lib.load()
    N, H, W, Cx = x.shape
    wplanes, w_plane_stride = 1, 0
    if w.dim() == 5:                       # split weights [planes][Cout][KH][KW][Cin] (or a [:, a:b] channel slice)
        wplanes, w_plane_stride, w = w.shape[0], w.stride(0), w[0]
    Cout, KH, KW, Cin = w.shape
    assert Cout % groups == 0 and (groups == 1 or N == groups), (groups, N, w.shape)
    Cout //= groups
    assert (cin or Cx) == Cin, (x.shape, w.shape)
    p = ConvParams()
    p.inp, p.in_ld = x.data_ptr(), _ld(x)
    p.N, p.H, p.W, p.Cin = N, H, W, Cin
    p.weight, p.bias = w.data_ptr(), (bias.data_ptr() if bias is not None else None)
    p.Cout, p.KH, p.KW, p.stride, p.pad, p.dil = Cout, KH, KW, stride, pad, dil
    p.out = out.data_ptr()
    if out_strides is None:
        p.out_ps, p.out_cs = _ld(out), 1
    else:
        p.out_ps, p.out_cs = out_strides
    p.res, p.res_ld = (res.data_ptr(), _ld(res)) if res is not None else (None, 0)
    p.out_relu, p.out_relu_ld = (out_relu.data_ptr(), _ld(out_relu)) if out_relu is not None else (None, 0)
    p.act, p.relu_in, p.dtype = act, int(relu_in), _dt(x)
    p.w_plane_stride = w_plane_stride or w.numel()
    p.gn_group_ch = gn_group_ch
    p.groups = groups
    assert wplanes == max(1, p.dtype & 0xff), "weights and activations must have the same number of planes"
    p.out_f32 = int(out.dtype == torch.float32 and x.dtype != torch.float32)
    p.gn_stats = gn_stats.data_ptr() if gn_stats is not None else None
    p.gn_stats_zeroed = int(gn_stats_zeroed)
    fused = None
    if gn_fuse is not None:
        gamma, beta, eps = gn_fuse
        p.gn_gamma, p.gn_beta, p.gn_eps = gamma.data_ptr(), beta.data_ptr(), eps
        if workspace is not None:
            p.workspace, p.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
        fused = GN_FUSE and bool(lib.otvm_conv2d_can_fuse_gn(C.byref(p)))
        if query_fuse:
            return fused
        if not fused:                     # plain convolution + statistics; residual / activation belong to gn_apply
            p.gn_gamma = p.gn_beta = None
            p.res, p.res_ld, p.act = None, 0, ACT_NONE
            p.out, p.out_ps, p.out_cs = gn_raw_out.data_ptr(), _ld(gn_raw_out), 1
    if workspace is not None:
        p.workspace, p.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    assert w.dtype == x.dtype and (bias is None or bias.dtype == torch.float32)
    if PROFILER is None:
        check(lib.otvm_conv2d(C.byref(p), _stream()), "otvm_conv2d")
        return out if fused is None else fused
    Ho = (H + 2 * pad - dil * (KH - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (KW - 1) - 1) // stride + 1
    key = "conv_tcgen05" if lib.otvm_conv2d_uses_tensor_cores(C.byref(p)) else "conv_ffma"
    if PROFILE_SHAPES:
        key += f" Cin={Cin} Cout={Cout} k={KH} s={stride} d={dil} {H}x{W}"
    es = x.element_size() * _planes(x)
    nb = (N * H * W * Cin + Cout * KH * KW * Cin) * es + N * Ho * Wo * Cout * out.element_size() * _planes(out)
    _timed(key, lambda: check(lib.otvm_conv2d(C.byref(p), _stream()), "otvm_conv2d"),
           2.0 * N * Ho * Wo * Cout * KH * KW * Cin, float(nb))
    return out if fused is None else fused


def zero_(t):
    """cudaMemsetAsync of a contiguous tensor on the current stream (GroupNorm statistics arena)"""
    if DRY:
        return t
    assert t.is_contiguous()
    check(_lib.load().otvm_zero_async(_p(t), t.numel() * t.element_size(), _stream()), "otvm_zero_async")
    return t


def gn_stats(x, stats):
    if DRY:
        return None
    lib = _lib.load()
    N, H, W, Cc = x.shape
    _timed("gn_stats", lambda: check(lib.otvm_gn_stats(_p(x), _ld(x), N, H * W, Cc, _dt(x), _p(stats),
                                                       _stream()), "otvm_gn_stats"), nbytes=_nbytes(x))


def gn_apply(x, stats, gamma, beta, out, *, act=ACT_NONE, res=None, eps=1e-5):
    if DRY:
        return out
    lib = _lib.load()
    N, H, W, Cc = x.shape
    _timed("gn_apply", lambda: check(
        lib.otvm_gn_apply(_p(x), _ld(x), N, H * W, Cc, _dt(x), _p(stats), _p(gamma), _p(beta), eps,
                          _p(res), _ld(res) if res is not None else 0, act, _p(out), _ld(out), _stream()),
        "otvm_gn_apply"), nbytes=_nbytes(x, out, res))
    return out


def upsample(x, out, *, add=None, out_relu=None, out_nchw_f32=False):
    """bilinear, align_corners=False.  ``out`` is NHWC [N,Ho,Wo,C] with the dtype of ``x``; when C is not a
    multiple of 4 (the 3 STM logits) ``x`` and ``out`` are fp32 and ``out`` may instead be NCHW planes
    [C,Ho,Wo] (``out_nchw_f32``)."""
    if DRY:
        return out
    lib = _lib.load()
    N, Hi, Wi, Cc = x.shape
    if out_nchw_f32:
        Ho, Wo = out.shape[-2:]
        old = 0
    else:
        Ho, Wo = out.shape[1:3]
        old = _ld(out)
    if Cc % 4 or out_nchw_f32:
        assert out.dtype == torch.float32 and add is None
    _timed("upsample", lambda: check(
        lib.otvm_upsample_bilinear(_p(x), _ld(x), N, Hi, Wi, Cc, Ho, Wo, _p(add),
                                   _ld(add) if add is not None else 0, _p(out), old, _p(out_relu),
                                   _ld(out_relu) if out_relu is not None else 0, _dt(x),
                                   int(out_nchw_f32), _stream()), "otvm_upsample_bilinear"),
        nbytes=_nbytes(x, out, add, out_relu))
    return out


def maxpool3x3s2(x, out):
    if DRY:
        return out
    lib = _lib.load()
    N, H, W, Cc = x.shape
    _timed("maxpool", lambda: check(lib.otvm_maxpool3x3s2(_p(x), _ld(x), N, H, W, Cc, _p(out), _ld(out),
                                                          _dt(x), _stream()), "otvm_maxpool3x3s2"),
           nbytes=_nbytes(x, out))
    return out


def ppm_pool(x, out, scratch):
    if DRY:
        return out
    lib = _lib.load()
    N, H, W, Cc = x.shape
    _timed("ppm_pool", lambda: check(lib.otvm_ppm_pool(_p(x), _ld(x), N, H, W, Cc, _p(out), _p(scratch),
                                                       _dt(x), _stream()), "otvm_ppm_pool"), nbytes=_nbytes(x))
    return out


def memory_read_workspace(M, HW, De, Do, dtype=None):
    """bytes of fp32 scratch for :func:`memory_read` (an upper bound over every element format)"""
    if DRY:
        return 16
    return int(_lib.load().otvm_memory_read_workspace(M, HW, De, Do, BF16))


def memory_read_backward(keys, vals, ldv, query, out, dout, lse, dkeys, dvals, dquery, M):
    """gradients of :func:`memory_read` (fp32): keys [M,De] / vals [Do,ldv] / query NHWC as in the forward (fp32 or bf16),
    out / dout [1,h,w,>=Do] fp32 views, lse [HW] from the forward; dkeys [M,De], dvals [Do,*], dquery [HW,De] overwritten"""
    lib = _lib.load()
    p = ReadBwdParams()
    p.keys, p.vals, p.ldv = keys.data_ptr(), vals.data_ptr(), ldv
    p.query, p.q_ld = query.data_ptr(), _ld(query)
    p.out, p.out_ld = out.data_ptr(), _ld(out)
    p.dout, p.dout_ld = dout.data_ptr(), _ld(dout)
    p.lse = lse.data_ptr()
    p.dkeys, p.dvals, p.dldv, p.dquery = dkeys.data_ptr(), dvals.data_ptr(), dvals.stride(0), dquery.data_ptr()
    p.M, p.HW, p.De, p.Do = M, query.shape[1] * query.shape[2], keys.shape[-1], vals.shape[0]
    p.dtype = _dt(keys)
    assert out.dtype == dout.dtype == lse.dtype == dkeys.dtype == torch.float32
    check(lib.otvm_memory_read_backward(C.byref(p), _stream()), "otvm_memory_read_backward")


def memory_read(keys, vals, ldv, query, out, M, workspace, *, force_simt=False, lse=None):
    """keys [>=M, De] rows; vals [Do, ldv] channel-major; query NHWC view with De channels; out NHWC view with
    Do channels (a slice of the 2*Do decoder input)."""
    if DRY:
        return out
    lib = _lib.load()
    p = ReadParams()
    De, Do = keys.shape[-1], vals.shape[0]
    HW = query.shape[1] * query.shape[2]
    p.keys, p.vals, p.ldv = keys.data_ptr(), vals.data_ptr(), ldv
    p.query, p.q_ld = query.data_ptr(), _ld(query)
    p.out, p.out_ld = out.data_ptr(), _ld(out)
    p.M, p.HW, p.De, p.Do = M, HW, De, Do
    p.dtype = _dt(keys)
    p.workspace, p.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    p.force_simt = int(force_simt)
    p.lse = lse.data_ptr() if lse is not None else None
    es = keys.element_size() * _planes(keys)
    _timed("memory_read", lambda: check(lib.otvm_memory_read(C.byref(p), _stream()), "otvm_memory_read"),
           2.0 * M * HW * (De + Do), float((De + Do) * M * es + De * HW * es + Do * HW * es))
    return out


def preprocess(a, fg, bg, H, W, Hp, Wp, pad_top, pad_left, radius, mean_std, img, scaled_img, tri3, imgn, scratch):
    if DRY:
        return None
    lib = _lib.load()
    ms = (C.c_float * 6)(*mean_std)
    _timed("glue", lambda: check(
        lib.otvm_preprocess(_p(a), _p(fg), _p(bg), H, W, Hp, Wp, pad_top, pad_left, radius, ms, _p(img),
                            _p(scaled_img), _p(tri3), _p(imgn), _ld(imgn), _dt(imgn), _p(scratch),
                            _stream()), "otvm_preprocess"))


def trimap_encode(tri, tri_ld, is_logit, img, Hp, Wp, mean_std, x11, cat_dst, extras, d2, scratch, seeds):
    if DRY:
        return None
    lib = _lib.load()
    ms = (C.c_float * 6)(*mean_std)
    _timed("trimap_encode_edt", lambda: check(
        lib.otvm_trimap_encode(_p(tri), tri_ld, int(is_logit), _p(img), Hp, Wp, ms, _p(x11), _ld(x11),
                               _p(cat_dst), _ld(cat_dst) if cat_dst is not None else 0, _dt(x11),
                               _p(extras), _p(d2), _p(scratch), _p(seeds), _stream()), "otvm_trimap_encode"))


def edt_sq(seed, d2, scratch):
    if DRY:
        return d2
    lib = _lib.load()
    H, W = seed.shape
    check(lib.otvm_edt_sq(_p(seed), H, W, _p(d2), _p(scratch), _stream()), "otvm_edt_sq")
    return d2


def fba_head(raw, raw_ld, dtype, extras, P, out7, alpha_dst=None, alpha_ld=0):
    if DRY:
        return None
    lib = _lib.load()
    # element format of alpha_dst (and of `raw` unless it is fp32)
    fmt = _dt(alpha_dst) if alpha_dst is not None else (_dt(raw) if raw.dtype != torch.float32 else
                                                       (F32 if dtype == torch.float32 else BF16))
    _timed("glue", lambda: check(
        lib.otvm_fba_head(_p(raw), raw_ld, fmt, int(raw.dtype == torch.float32), _p(extras), P, _p(out7),
                          _p(alpha_dst), alpha_ld, _stream()), "otvm_fba_head"))


def head_conv_fba(x, w, bias, raw, extras, P, out7, alpha_dst=None, alpha_ld=0):
    """1x1 head convolution (x [1,H,W,16] -> raw [..][8 or 12] fp32, w [Cout,16] fp32) + clamp / sigmoid / fba_fusion"""
    if DRY:
        return None
    lib = _lib.load()
    _timed("glue", lambda: check(
        lib.otvm_head_conv_fba(_p(x), _ld(x), _dt(x), _p(w), _p(bias), w.shape[0], _p(raw), raw.shape[-1], _p(extras), P,
                               _p(out7), _p(alpha_dst), alpha_ld, _stream()), "otvm_head_conv_fba"))


def frame_outputs(raw10, raw_ld, fused, hid, extras, Hp, Wp, H, W, pad_top, pad_left, mean_std, mem_in,
                  alpha_out, trimap_out):
    if DRY:
        return None
    lib = _lib.load()
    ms = (C.c_float * 6)(*mean_std)
    _timed("glue", lambda: check(
        lib.otvm_frame_outputs(_p(raw10), raw_ld, _p(fused), _p(hid), _ld(hid), _p(extras), Hp, Wp, H, W,
                               pad_top, pad_left, ms, _p(mem_in), _ld(mem_in) if mem_in is not None else 0,
                               _dt(hid), _p(alpha_out), _p(trimap_out), _stream()), "otvm_frame_outputs"))


def unpack_frame_u8(fg_u8, bg_u8, a, fg, bg):
    """fg_u8 [H,W,3|4] / bg_u8 [H,W,3] uint8 (cv2.imread layout) -> a [H,W], fg / bg [3,H,W] fp32 (EvalModel inputs)"""
    H, W, c = fg_u8.shape
    check(_lib.load().otvm_unpack_frame_u8(_p(fg_u8), c, _p(bg_u8), H, W, _p(a), _p(fg), _p(bg), _stream()),
          "otvm_unpack_frame_u8")


def alpha_to_u8(alpha, out):
    check(_lib.load().otvm_alpha_to_u8(_p(alpha), alpha.numel(), _p(out), _stream()), "otvm_alpha_to_u8")
    return out


def nchw_to_nhwc(x, out):
    if DRY:
        return out
    lib = _lib.load()
    N, Cc, H, W = x.shape
    check(lib.otvm_nchw_to_nhwc(_p(x), N, Cc, H * W, _p(out), _ld(out), _dt(out), _stream()), "nchw_to_nhwc")
    return out


def nhwc_to_nchw(x, out):
    if DRY:
        return out
    lib = _lib.load()
    N, H, W, Cc = x.shape
    check(lib.otvm_nhwc_to_nchw(_p(x), _ld(x), N, Cc, H * W, _p(out), _dt(x), _stream()), "nhwc_to_nchw")
    return out
