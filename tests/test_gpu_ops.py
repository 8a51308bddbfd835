"""Kernel-level parity (through the C ABI) against plain PyTorch fp32 ops / the oracle, on the GPU.

Tolerances are scale-relative max errors (tests/util.py:rel_err): 1e-4 for the strict fp32 kernels (pure
re-association noise) and for the split-bf16 formats ("x2" = two bf16 planes, "x3" = three: the tensor cores multiply
the planes pairwise, otvm_b200/split.py), 1e-2 for plain bf16 storage with fp32 accumulation -- always compared
against the SAME op evaluated in fp32 on the operands as stored; integer outputs (EDT, trimap classes) are bit-exact.
"""
import math
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import ROOT, rel_err

sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu
DEV = "cuda"
X2, X3 = "x2", "x3"                       # split-bf16 element formats (2 / 3 planes)
TOL = {torch.float32: 1e-4, torch.bfloat16: 1e-2, X2: 1e-4, X3: 1e-4}
DTYPES = [torch.float32, torch.bfloat16, X2, X3]
TC_DTYPES = [torch.bfloat16, X2, X3]      # formats the tcgen05 kernels take
PLANES = {X2: 2, X3: 3}
_ARENAS = {}


def _ops():
    from otvm_b200 import ops
    return ops


def _arena(fmt):
    from otvm_b200.split import SplitArena
    if fmt not in _ARENAS:
        _ARENAS[fmt] = SplitArena(PLANES[fmt], 640 << 20, DEV)
    return _ARENAS[fmt]


@pytest.fixture(autouse=True)
def _reset_arenas():
    for a in _ARENAS.values():
        a.off = 0
    yield


def zeros(shape, dtype):
    """zero-filled device tensor of an element format (split formats: the plane-0 view of an arena tensor)"""
    if dtype in PLANES:
        return _arena(dtype).alloc(tuple(shape), zero=True)
    return torch.zeros(*shape, dtype=dtype, device=DEV)


def put(dst, x):
    """dst[...] = x (fp32, any device) in dst's element format"""
    from otvm_b200.split import arena_of
    a = arena_of(dst) if dst.dtype == torch.bfloat16 else None
    if a is not None and a.planes > 1:
        a.write(dst, x.to(DEV))
    else:
        dst.copy_(x.to(DEV))


def nhwc(x, dtype, ld=None):
    """NCHW fp32 cpu -> NHWC device tensor (optionally a channel slice of a wider buffer)."""
    N, C, H, W = x.shape
    ld = ld or C
    buf = zeros((N, H, W, ld), dtype)
    put(buf[..., :C], x.permute(0, 2, 3, 1))
    return buf[..., :C]


def flt(x):
    from otvm_b200.split import to_float
    return to_float(x)


def nchw(x):
    return flt(x).permute(0, 3, 1, 2).cpu()


def rnd(dtype, x):
    """x as stored in the element format"""
    if dtype in PLANES:
        from otvm_b200.split import split_planes
        return split_planes(x, PLANES[dtype]).float().sum(0)
    return x.to(dtype).float()


def wpack(w, dtype):
    """[Cout,Cin,KH,KW] fp32 -> packed [Cout,KH,KW,Cin] filter bank in the element format"""
    p = w.permute(0, 2, 3, 1).contiguous()
    if dtype in PLANES:
        from otvm_b200.split import split_planes
        return split_planes(p, PLANES[dtype]).to(DEV)
    return p.to(DEV, dtype)


CONV_CASES = [
    # Cin, Cout, k, stride, pad, dil, H, W
    (64, 64, 1, 1, 0, 1, 16, 16),
    (64, 256, 3, 1, 1, 1, 16, 24),
    (128, 128, 3, 2, 1, 1, 16, 16),
    (256, 512, 1, 2, 0, 1, 16, 16),
    (256, 256, 3, 1, 2, 2, 16, 16),
    (512, 512, 3, 1, 4, 4, 8, 8),
    (3, 64, 7, 2, 3, 1, 32, 32),
    (22, 64, 7, 2, 3, 1, 32, 40),
    (80, 32, 3, 1, 1, 1, 16, 16),
    (16, 7, 1, 1, 0, 1, 8, 8),
    (256, 3, 3, 1, 1, 1, 8, 8),
    (2048, 256, 1, 1, 0, 1, 2, 2),
    (1024, 128, 3, 1, 1, 1, 8, 8),
    (320, 64, 3, 1, 1, 1, 20, 12),
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d(case, dtype):
    ops = _ops()
    Cin, Cout, k, s, p, d, H, W = case
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    x = torch.randn(1, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
    b = torch.randn(Cout, generator=g)
    want = F.conv2d(rnd(dtype, x), rnd(dtype, w), b, s, p, d)
    res = torch.randn_like(want)
    want_res = F.leaky_relu(want + rnd(dtype, res), 0.01)
    xd = nhwc(x, dtype, ld=Cin + 8 if Cin % 4 == 0 else None)
    wd = wpack(w, dtype)
    bd = b.to(DEV)
    Ho, Wo = want.shape[2:]
    out = zeros((1, Ho, Wo, Cout + 4), dtype)[..., :Cout] if Cout % 4 == 0 else \
        zeros((1, Ho, Wo, Cout), dtype)
    ops.conv2d(xd, wd, bd, out, stride=s, pad=p, dil=d)
    assert rel_err(nchw(out), want) < TOL[dtype]
    # fused epilogue: residual + LeakyReLU + ReLU'd second output
    out2 = zeros((1, Ho, Wo, Cout), dtype)
    outr = zeros((1, Ho, Wo, Cout), dtype)
    ops.conv2d(xd, wd, bd, out2, stride=s, pad=p, dil=d, res=nhwc(res, dtype), act=ops.ACT_LEAKY, out_relu=outr)
    assert rel_err(nchw(out2), want_res) < TOL[dtype]
    assert rel_err(nchw(outr), F.relu(want_res)) < TOL[dtype]
    # relu on the input (STM ResBlock)
    out3 = zeros((1, Ho, Wo, Cout), dtype)
    ops.conv2d(xd, wd, bd, out3, stride=s, pad=p, dil=d, relu_in=True, act=ops.ACT_RELU)
    assert rel_err(nchw(out3), F.relu(F.conv2d(F.relu(rnd(dtype, x)), rnd(dtype, w), b, s, p, d))) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
def test_conv2d_channel_major_out_and_f32_head(dtype):
    """value-bank store (out[c*ldv + p]) and the fp32 head output used for the 7/10-channel predictions"""
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 64, 8, 8, generator=g); w = torch.randn(32, 64, 3, 3, generator=g) / 24; b = torch.randn(32, generator=g)
    want = F.conv2d(rnd(dtype, x), rnd(dtype, w), b, 1, 1)
    wd = wpack(w, dtype)
    bank = zeros((32, 3 * 64), dtype)
    ops.conv2d(nhwc(x, dtype), wd, b.to(DEV), bank[:, 64:], pad=1, out_strides=(1, bank.shape[1]))
    assert rel_err(flt(bank[:, 64:128]).cpu().view(32, 8, 8), want[0]) < TOL[dtype]
    assert float(flt(bank[:, :64]).abs().max()) == 0 and float(flt(bank[:, 128:]).abs().max()) == 0
    o32 = torch.zeros(1, 8, 8, 36, dtype=torch.float32, device=DEV)
    ops.conv2d(nhwc(x, dtype), wd, b.to(DEV), o32[..., :32], pad=1)
    assert rel_err(nchw(o32[..., :32]), want) < (1e-4 if dtype == torch.float32 else 2e-3)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C,H,W", [(64, 32, 32), (256, 16, 8), (2048, 8, 8), (256, 1, 1), (256, 3, 3), (128, 5, 7)])
def test_groupnorm(C, H, W, dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(1, C, H, W, generator=g) * 3 + 1.5
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    res = torch.randn(1, C, H, W, generator=g)
    want = F.relu(F.group_norm(rnd(dtype, x), 32, gamma, beta, 1e-5) + rnd(dtype, res))
    xd = nhwc(x, dtype)
    stats = torch.zeros(64, dtype=torch.float64, device=DEV)
    ops.gn_stats(xd, stats)
    xr = rnd(dtype, x).double().view(32, -1)
    assert rel_err(stats.cpu().view(32, 2)[:, 0], xr.sum(1)) < 1e-5
    assert rel_err(stats.cpu().view(32, 2)[:, 1], (xr * xr).sum(1)) < 1e-5
    out = zeros(tuple(xd.shape), dtype)
    ops.gn_apply(xd, stats, gamma.to(DEV), beta.to(DEV), out, act=ops.ACT_RELU, res=nhwc(res, dtype))
    assert rel_err(nchw(out), want) < TOL[dtype]
    # statistics fused into the conv epilogue must agree with the stand-alone pass
    w = torch.randn(C, 64, 1, 1, generator=g) / 8
    xin = torch.randn(1, 64, H, W, generator=g)
    raw = zeros((1, H, W, C), dtype)
    st2 = torch.zeros(64, dtype=torch.float64, device=DEV)
    ops.conv2d(nhwc(xin, dtype), wpack(w, dtype), None, raw, gn_stats=st2)
    ops.gn_stats(raw, stats)
    assert rel_err(st2.cpu(), stats.cpu()) < 1e-5


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C,Hi,Wi,Ho,Wo", [(256, 8, 8, 16, 16), (64, 16, 12, 32, 24), (256, 1, 1, 8, 8),
                                           (256, 2, 2, 8, 12), (256, 3, 3, 8, 8), (256, 6, 6, 16, 16)])
def test_upsample_bilinear(C, Hi, Wi, Ho, Wo, dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(Hi * Wo)
    x = torch.randn(1, C, Hi, Wi, generator=g); add = torch.randn(1, C, Ho, Wo, generator=g)
    want = rnd(dtype, add) + F.interpolate(rnd(dtype, x), size=(Ho, Wo), mode="bilinear", align_corners=False)
    out = zeros((1, Ho, Wo, C + 64), dtype)[..., 64:]
    outr = zeros((1, Ho, Wo, C), dtype)
    ops.upsample(nhwc(x, dtype), out, add=nhwc(add, dtype), out_relu=outr)
    assert rel_err(nchw(out), want) < TOL[dtype]
    assert rel_err(nchw(outr), F.relu(want)) < TOL[dtype]


def test_upsample_logits_x4_fp32():
    ops = _ops()
    x = torch.randn(1, 3, 16, 20)
    want = F.interpolate(x, scale_factor=4, mode="bilinear", align_corners=False)
    xd = torch.zeros(1, 16, 20, 4, device=DEV); xd[..., :3] = x.permute(0, 2, 3, 1).to(DEV)
    out = torch.zeros(1, 64, 80, 4, device=DEV)
    ops.upsample(xd[..., :3], out[..., :3])
    assert rel_err(nchw(out[..., :3]), want) < 1e-6
    planes = torch.zeros(3, 64, 80, device=DEV)
    ops.upsample(xd[..., :3], planes, out_nchw_f32=True)
    assert rel_err(planes.cpu(), want[0]) < 1e-6


@pytest.mark.parametrize("dtype", DTYPES)
def test_maxpool_and_ppm(dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 64, 18, 22, generator=g)
    out = zeros((1, 9, 11, 64), dtype)
    ops.maxpool3x3s2(nhwc(x, dtype, ld=80), out)
    assert rel_err(nchw(out), F.max_pool2d(rnd(dtype, x), 3, 2, 1)) == 0
    for H, W in ((8, 8), (16, 12), (64, 64), (5, 9)):
        f = torch.randn(1, 128, H, W, generator=g)
        pooled = zeros((50, 128), dtype)
        ops.ppm_pool(nhwc(f, dtype), pooled, torch.zeros(H * 12 * 128, device=DEV))
        off = 0
        for s in (1, 2, 3, 6):
            want = F.adaptive_avg_pool2d(rnd(dtype, f), s)[0].reshape(128, s * s).t()
            assert rel_err(flt(pooled[off:off + s * s]).cpu(), want) < TOL[dtype], (H, W, s)
            off += s * s


def test_edt_bit_exact():
    import otvm_oracle as O
    ops = _ops()
    r = np.random.RandomState(0)
    for (H, W), p in (((37, 53), 0.02), ((64, 64), 0.3), ((5, 90), 0.5), ((128, 96), 0.0005), ((33, 31), 0.0),
                      ((512, 512), 0.00002), ((300, 200), 0.2), ((257, 130), 0.001)):
        seed = (r.uniform(size=(H, W)) < p)
        if p > 0:
            seed[r.randint(H), r.randint(W)] = True
        d2 = torch.zeros(H, W, dtype=torch.int32, device=DEV)
        ops.edt_sq(torch.from_numpy(seed.astype(np.uint8)).to(DEV), d2, torch.zeros(H, W, dtype=torch.int32, device=DEV))
        got = d2.cpu().numpy()
        if seed.any():
            assert np.array_equal(got.astype(np.int64), O.edt_sq(~seed))
        else:
            assert (got == 0x3fffffff).all()


MEM_CASES = [(1, 4, 4, 1.0), (3, 6, 5, 1.0), (2, 8, 8, 6.0), (5, 7, 9, 0.3), (8, 16, 16, 2.0), (3, 12, 20, 3.0)]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("simt", [True, False])
@pytest.mark.parametrize("case", MEM_CASES)
def test_memory_read(case, simt, dtype):
    """Memory.forward (STM.py:144-163) on random banks, incl. sharp (scale 6) and flat softmaxes, ragged sizes"""
    import otvm_oracle as O
    ops = _ops()
    T, h, w, sc = case
    r = np.random.RandomState(T * 100 + h)
    t = lambda *s: torch.from_numpy(r.standard_normal(s).astype(np.float32))
    m_in, m_out = t(1, 128, T, h, w) * sc, t(1, 512, T, h, w)
    q_in, q_out = t(1, 128, h, w) * sc, t(1, 512, h, w)
    want = O.memory_read(rnd(dtype, m_in), rnd(dtype, m_out), rnd(dtype, q_in), rnd(dtype, q_out))
    hw, cap = h * w, T + 2
    keys = zeros((cap * hw, 128), dtype)
    vals = zeros((512, cap * hw), dtype)
    put(keys[:T * hw], m_in[0].permute(1, 2, 3, 0).reshape(T * hw, 128))
    put(vals[:, :T * hw], m_out[0].reshape(512, T * hw))
    q = nhwc(q_in, dtype)
    out = zeros((1, h, w, 1024), dtype)
    put(out[..., 512:], q_out.permute(0, 2, 3, 1))
    ws = torch.zeros(ops.memory_read_workspace(cap * hw, hw, 128, 512) // 4 + 1, device=DEV)
    ops.memory_read(keys, vals, vals.shape[1], q, out[..., :512], T * hw, ws, force_simt=simt)
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-4       # bf16: P is rounded to bf16 before P.V on tensor cores
    assert rel_err(nchw(out), want) < tol


def test_memory_read_golden_vectors():
    """the reference's own Memory.forward outputs (tests/golden/memory_read.npz)"""
    from util import golden
    ops = _ops()
    g = golden("memory_read")
    ci = 0
    while f"c{ci}_out" in g:
        T, h, w = (int(v) for v in g[f"c{ci}_shape"])
        sc = float(g[f"c{ci}_scale"])
        r = np.random.RandomState(1000 + ci)
        t = lambda *s: torch.from_numpy(r.standard_normal(s).astype(np.float32))
        m_in, m_out = t(1, 128, T, h, w) * sc, t(1, 512, T, h, w)
        q_in, q_out = t(1, 128, h, w) * sc, t(1, 512, h, w)
        hw = h * w
        keys = m_in[0].permute(1, 2, 3, 0).reshape(T * hw, 128).contiguous().to(DEV)
        vals = m_out[0].reshape(512, T * hw).contiguous().to(DEV)
        out = torch.zeros(1, h, w, 512, device=DEV)
        ws = torch.zeros(ops.memory_read_workspace(T * hw, hw, 128, 512) // 4 + 1, device=DEV)
        ops.memory_read(keys, vals, T * hw, nhwc(q_in, torch.float32), out, T * hw, ws)
        assert rel_err(nchw(out)[0], g[f"c{ci}_out"][:512]) < 1e-4, ci
        ci += 1
    assert ci == 5


def test_memory_read_baseline_sizes():
    """the fused tcgen05 read at BASELINE.json's full sizes -- 512^2 with T=8 (cfg 2) and T=16 (north-star target), 1024^2
    with T=16 (cfg 3: HW=4096, THW=65536, a 1 GB affinity for the unfused form) -- against softmax(K Q^T / sqrt(128)) V in
    fp32 on the same bf16-rounded operands (STM.py:153-158); north-star tolerance for bf16: 1e-2 (measured 2.8e-3 .. 4e-3,
    scripts/bench_read.py)"""
    ops = _ops()
    torch.manual_seed(0)
    for hw_side, T in [(32, 8), (32, 16), (64, 16)]:
        HW = hw_side * hw_side
        M = T * HW
        k = torch.randn(M, 128, device=DEV).bfloat16()
        v = torch.randn(512, M, device=DEV).bfloat16()
        q = (torch.randn(1, hw_side, hw_side, 128, device=DEV) * 1.5).bfloat16()
        out = torch.zeros(1, hw_side, hw_side, 1024, device=DEV, dtype=torch.bfloat16)
        ws = torch.zeros(ops.memory_read_workspace(M, HW, 128, 512) // 4 + 1, device=DEV)
        p = torch.softmax((k.float() @ q.view(HW, 128).float().t()) / math.sqrt(128), dim=0)       # [M, HW]
        want = (v.float() @ p).t()                                                                   # [HW, 512]
        del p
        ops.memory_read(k, v, M, q, out[..., :512], M, ws)
        torch.cuda.synchronize()
        got = out[0, ..., :512].float().view(HW, 512)
        assert float((got - want).abs().max() / want.abs().max()) < 1e-2, (hw_side, T)
        assert float(out[..., 512:].float().abs().max()) == 0.0          # the query-value half is not this kernel's to write


TC_CASES = [
    # Cin, Cout, k, pad, dil, H, W   (stride 1: the tcgen05 implicit-GEMM path)
    (3072, 256, 3, 1, 1, 32, 32),
    (64, 64, 3, 1, 1, 96, 128),
    (96, 64, 3, 1, 1, 64, 64),
    (96, 32, 3, 1, 1, 40, 72),
    (32, 16, 3, 1, 1, 64, 64),
    (16, 10, 1, 0, 1, 64, 64),
    (2048, 512, 1, 0, 1, 16, 16),
    (512, 2048, 1, 0, 1, 16, 16),
    (512, 512, 3, 4, 4, 32, 32),
    (1024, 512, 3, 1, 1, 32, 32),
    (256, 256, 3, 1, 1, 20, 36),
    (128, 128, 3, 1, 1, 9, 13),
]


@pytest.fixture(params=[1, 0], ids=["halo", "per_tap"])
def conv_halo(request):
    """3x3 stride-1 convs: one shared-memory input patch for all 9 taps (default) vs one TMA box per tap"""
    import ctypes
    from otvm_b200 import _lib
    lib = _lib.load()
    lib.otvm_debug_set_conv_halo.argtypes = [ctypes.c_int]
    lib.otvm_debug_set_conv_halo(request.param)
    yield request.param
    lib.otvm_debug_set_conv_halo(-1)


@pytest.mark.parametrize("dtype", TC_DTYPES)
@pytest.mark.parametrize("case", TC_CASES)
def test_conv2d_tcgen05(case, conv_halo, dtype):
    """stride-1 convs must run on the tcgen05 kernel and match fp32 math on the operands as stored (bf16, or split
    bf16 with 3 / 6 plane products per K step)"""
    if conv_halo == 0 and case[2] != 3:
        pytest.skip("per-tap mode only differs for 3x3 convolutions")
    ops = _ops()
    tol, tol_head, tol_st = (1e-2, 2e-3, 2e-3) if dtype == torch.bfloat16 else (1e-4, 1e-4, 1e-4)
    Cin, Cout, k, p, d, H, W = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(1, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
    b = torch.randn(Cout, generator=g)
    want = F.conv2d(rnd(dtype, x), rnd(dtype, w), b, 1, p, d)
    res = torch.randn_like(want)
    want2 = F.relu(want + rnd(dtype, res))
    xd = nhwc(x, dtype, ld=Cin + 8)
    wd = wpack(w, dtype)
    Ho, Wo = want.shape[2:]
    head = Cout % 8 != 0
    out = torch.zeros(1, Ho, Wo, Cout, device=DEV) if head else zeros((1, Ho, Wo, Cout + 8), dtype)[..., :Cout]
    outr = None if head else zeros((1, Ho, Wo, Cout), dtype)
    stats = torch.zeros(64, dtype=torch.float64, device=DEV) if Cout % 32 == 0 else None
    prof = ops.Profiler()
    ops.PROFILER = prof
    try:
        ops.conv2d(xd, wd, b.to(DEV), out, pad=p, dil=d, gn_stats=stats)
        if not head:
            out2 = zeros((1, Ho, Wo, Cout), dtype)
            ops.conv2d(xd, wd, b.to(DEV), out2, pad=p, dil=d, res=nhwc(res, dtype), act=ops.ACT_RELU, out_relu=outr)
    finally:
        ops.PROFILER = None
    assert set(prof.summary()) == {"conv_tcgen05"}
    assert rel_err(nchw(out), want) < (tol_head if head else tol)
    if not head:
        assert rel_err(nchw(out2), want2) < tol
        assert rel_err(nchw(outr), want2) < tol
    if stats is not None:
        q = rnd(dtype, want).double()[0].reshape(32, -1)
        assert rel_err(stats.cpu().view(32, 2)[:, 0], q.sum(1)) < tol_st
        assert rel_err(stats.cpu().view(32, 2)[:, 1], (q * q).sum(1)) < tol_st
    # with a workspace, small grids take the split-K route (fp32 partial tiles + fused finish kernel)
    ws = torch.zeros(8 << 20, device=DEV)
    out3 = torch.zeros_like(out.contiguous()) if head else zeros((1, Ho, Wo, Cout), dtype)
    stats3 = torch.zeros(64, dtype=torch.float64, device=DEV) if stats is not None else None
    ops.conv2d(xd, wd, b.to(DEV), out3, pad=p, dil=d, gn_stats=stats3, workspace=ws)
    assert rel_err(nchw(out3), want) < (tol_head if head else tol)
    if stats is not None:
        assert rel_err(stats3.cpu(), stats.cpu()) < tol_st
    if not head:
        out4 = zeros((1, Ho, Wo, Cout), dtype); outr4 = zeros((1, Ho, Wo, Cout), dtype)
        ops.conv2d(xd, wd, b.to(DEV), out4, pad=p, dil=d, res=nhwc(res, dtype), act=ops.ACT_RELU, out_relu=outr4, workspace=ws)
        assert rel_err(nchw(out4), want2) < tol and rel_err(nchw(outr4), want2) < tol


PERSIST_CASES = [
    # Cin, Cout, k, pad, dil, H, W, gn   (3x3 stride-1, Cout in {16, 32, 64}: the persistent patch-mode kernel)
    (64, 64, 3, 1, 1, 96, 128, True),        # one tile per CTA
    (64, 64, 3, 1, 1, 296, 304, True),       # ~5 tiles per CTA, ragged rows
    (96, 64, 3, 1, 1, 200, 168, True),       # three K-chunks per tile
    (96, 32, 3, 1, 1, 250, 180, False),      # ragged columns
    (64, 32, 3, 1, 1, 160, 160, False),
    (32, 16, 3, 1, 1, 320, 320, False),
    (64, 64, 3, 2, 2, 128, 128, False),      # dilation 2
    (128, 32, 3, 1, 1, 64, 72, False),       # two 64-channel chunks
    (64, 64, 3, 1, 1, 512, 512, True),       # the refinement-module layer itself (2048 tiles, auto mode)
]


@pytest.mark.parametrize("dtype", TC_DTYPES)
@pytest.mark.parametrize("case", PERSIST_CASES)
def test_conv2d_persistent(case, dtype):
    """persistent patch-mode tcgen05 kernel (resident filter bank, tile loop, double-buffered TMEM accumulators and
    staging tiles) == fp32 math on the bf16-rounded operands, incl. the GroupNorm statistics of the stored values"""
    import ctypes
    from otvm_b200 import _lib
    ops = _ops()
    lib = _lib.load()
    lib.otvm_debug_set_conv_persist.argtypes = [ctypes.c_int]
    lib.otvm_debug_conv_persist_launches.restype = ctypes.c_longlong
    tol, tol_st = (1e-2, 2e-3) if dtype == torch.bfloat16 else (1e-4, 1e-4)
    Cin, Cout, k, p, d, H, W, gn = case
    g = torch.Generator().manual_seed(sum(case[:7]))
    x = torch.randn(1, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
    b = torch.randn(Cout, generator=g)
    want = F.conv2d(rnd(dtype, x), rnd(dtype, w), b, 1, p, d)
    if not gn:
        want = F.leaky_relu(want, 0.01)
    xd = nhwc(x, dtype, ld=Cin + 8)
    wd = wpack(w, dtype)
    auto = H * W >= 512 * 512
    lib.otvm_debug_set_conv_persist(-1 if auto else 1)
    try:
        for trial in range(2):
            out = zeros((1, H, W, Cout + 8), dtype)[..., :Cout]
            stats = torch.zeros(64, dtype=torch.float64, device=DEV) if gn else None
            n0 = lib.otvm_debug_conv_persist_launches()
            ops.conv2d(xd, wd, b.to(DEV), out, pad=p, dil=d, gn_stats=stats, act=ops.ACT_NONE if gn else ops.ACT_LEAKY)
            torch.cuda.synchronize()
            # (split operands: the filter bank + patch ring of the 3-chunk layers / of three planes exceed one CTA; those
            # fall back to the one-tile kernel -- results are checked either way)
            if dtype == torch.bfloat16 or (dtype == X2 and Cin <= 96 and d == 1):
                assert lib.otvm_debug_conv_persist_launches() == n0 + 1, "persistent kernel not selected"
            assert rel_err(nchw(out), want) < tol
            if gn:
                q = rnd(dtype, want).double()[0].reshape(32, -1)
                assert rel_err(stats.cpu().view(32, 2)[:, 0], q.sum(1)) < tol_st
                assert rel_err(stats.cpu().view(32, 2)[:, 1], (q * q).sum(1)) < tol_st
    finally:
        lib.otvm_debug_set_conv_persist(-1)
    # same layer through the one-tile-per-CTA kernel: both kernels round identically (fp32 accumulate, one bf16 rounding)
    lib.otvm_debug_set_conv_persist(0)
    try:
        out0 = zeros((1, H, W, Cout), dtype)
        ops.conv2d(xd, wd, b.to(DEV), out0, pad=p, dil=d, act=ops.ACT_NONE if gn else ops.ACT_LEAKY)
    finally:
        lib.otvm_debug_set_conv_persist(-1)
    assert rel_err(nchw(out0), nchw(out)) < tol_st


FUSED_GN_CASES = [
    # Cin, Cout, k, pad, dil, H, W, with_res, act
    (64, 64, 1, 0, 1, 32, 32, False, "relu"),
    (64, 256, 1, 0, 1, 32, 32, True, "relu"),
    (256, 256, 3, 2, 2, 24, 40, False, "relu"),
    (512, 2048, 1, 0, 1, 16, 16, True, "relu"),
    (256, 256, 3, 1, 1, 32, 32, False, "leaky"),
    (128, 128, 3, 1, 1, 64, 64, False, "none"),
    (64, 64, 3, 1, 1, 64, 64, True, "relu"),
]


@pytest.mark.parametrize("dtype", TC_DTYPES)
@pytest.mark.parametrize("case", FUSED_GN_CASES)
def test_conv2d_fused_groupnorm(case, dtype):
    """conv -> GroupNorm(32) -> (+res) -> act in ONE kernel (statistics, grid barrier, normalise from tensor memory)
    against F.conv2d + F.group_norm in fp32 on the bf16-rounded operands (layers_WS.py:26-27, resnet_GN_WS.py:69-88)"""
    ops = _ops()
    Cin, Cout, k, p, d, H, W, with_res, act = case
    g = torch.Generator().manual_seed(sum(case[:7]))
    x = torch.randn(1, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
    b = torch.randn(Cout, generator=g)
    gamma, beta = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g)
    y = F.group_norm(F.conv2d(rnd(dtype, x), rnd(dtype, w), b, 1, p, d), 32, gamma, beta, 1e-5)
    res = torch.randn_like(y)
    if with_res:
        y = y + rnd(dtype, res)
    y = {"relu": F.relu, "leaky": lambda t: F.leaky_relu(t, 0.01), "none": lambda t: t}[act](y)
    actc = {"relu": ops.ACT_RELU, "leaky": ops.ACT_LEAKY, "none": ops.ACT_NONE}[act]
    for trial in range(2):                         # second call on a re-zeroed slot: the barrier counter is reusable
        arena = torch.zeros(72, dtype=torch.float64, device=DEV)
        out = zeros((1, H, W, Cout + 8), dtype)[..., :Cout]
        raw = zeros((1, H, W, Cout), dtype)
        fused = ops.conv2d(nhwc(x, dtype), wpack(w, dtype), b.to(DEV), out, pad=p, dil=d,
                           gn_stats=arena, gn_stats_zeroed=True, gn_fuse=(gamma.to(DEV), beta.to(DEV), 1e-5), gn_raw_out=raw,
                           res=nhwc(res, dtype) if with_res else None, act=actc)
        assert fused is True, "these shapes are single-wave tcgen05 grids"
        torch.cuda.synchronize()
        assert rel_err(nchw(out), y) < (1e-2 if dtype == torch.bfloat16 else 1e-4)


GROUPED_CASES = [
    # Cin, Cout, k, stride, H, W, with_res      (STM encoder layers at 128 x 160 and 512 x 512 frames)
    (64, 64, 1, 1, 32, 40, False), (64, 64, 3, 1, 32, 40, False), (64, 256, 1, 1, 32, 40, True), (256, 128, 1, 1, 32, 40, False),
    (128, 128, 3, 2, 32, 40, False), (256, 512, 1, 2, 32, 40, False), (1024, 256, 1, 1, 8, 10, False),
    (256, 256, 3, 1, 32, 32, False), (256, 1024, 1, 1, 32, 32, True), (64, 64, 3, 1, 128, 128, False),
]


@pytest.mark.parametrize("dtype", TC_DTYPES)
@pytest.mark.parametrize("case", GROUPED_CASES)
def test_conv2d_grouped_pair(case, dtype):
    """otvm_conv_params.groups = 2: image n of a batch of two is convolved with filter bank n (+ its bias, its residual)
    in ONE launch == two separate convolutions (the layers of Encoder_Q / Encoder_M, STM.py:33-102)"""
    ops = _ops()
    Cin, Cout, k, stride, H, W, with_res = case
    g = torch.Generator().manual_seed(sum(case[:6]))
    x = torch.randn(2, Cin, H, W, generator=g)
    w = torch.randn(2, Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
    b = torch.randn(2, Cout, generator=g)
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    res = torch.randn(2, Cout, Ho, Wo, generator=g)
    want = torch.cat([F.conv2d(rnd(dtype, x[i:i + 1]), rnd(dtype, w[i]), b[i], stride, pad) for i in range(2)])
    if with_res:
        want = want + rnd(dtype, res)
    want = F.relu(want)
    wd = [wpack(w[i], dtype) for i in range(2)]
    wpair = torch.cat(wd, dim=1 if wd[0].dim() == 5 else 0).contiguous()            # banks stacked along Cout
    xd = nhwc(x, dtype)
    out = zeros((2, Ho, Wo, Cout), dtype)
    ops.conv2d(xd, wpair, b.reshape(-1).to(DEV), out, stride=stride, pad=pad, act=ops.ACT_RELU,
               res=nhwc(res, dtype) if with_res else None, groups=2)
    torch.cuda.synchronize()
    assert rel_err(nchw(out), want) < TOL[dtype]
    # and bit-identical to the two single launches when the tile shape is the same (it is for these grids' second image)
    single = zeros((2, Ho, Wo, Cout), dtype)
    rd = nhwc(res, dtype) if with_res else None
    for i in range(2):
        ops.conv2d(xd[i:i + 1], wd[i], b[i].to(DEV), single[i:i + 1], stride=stride, pad=pad, act=ops.ACT_RELU,
                   res=rd[i:i + 1] if with_res else None)
    torch.cuda.synchronize()
    assert rel_err(nchw(out), nchw(single)) < (4e-3 if dtype == torch.bfloat16 else 1e-5)


CLUSTER_CASES = [
    # Cin, Cout, k, N, H, W, with_res, relu2     (short grids with 16..47 K iterations: STM res4 / decoder layers at 1/16)
    (256, 256, 3, 1, 32, 32, False, False), (256, 256, 3, 2, 32, 32, False, False), (1024, 256, 1, 2, 32, 32, False, False),
    (256, 256, 3, 1, 32, 32, True, True), (1024, 1024, 1, 1, 16, 16, True, False), (1024, 256, 1, 1, 24, 40, False, True),
]


@pytest.mark.parametrize("dtype", TC_DTYPES)
@pytest.mark.parametrize("case", CLUSTER_CASES)
def test_conv2d_cluster_split_k(case, dtype):
    """cluster split-K: the K slices of a tile run as a thread-block cluster, ranks >= 1 park fp32 partial tiles in shared
    memory and rank 0 adds them through distributed shared memory inside its normal epilogue (bias, residual, activation,
    second ReLU output) == fp32 convolution on the stored operands"""
    import ctypes
    from otvm_b200 import _lib
    ops = _ops()
    lib = _lib.load()
    lib.otvm_debug_conv_cluster_launches.restype = ctypes.c_longlong
    Cin, Cout, k, N, H, W, with_res, relu2 = case
    g = torch.Generator().manual_seed(sum(case[:6]))
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(N, Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
    b = torch.randn(N, Cout, generator=g)
    res = torch.randn(N, Cout, H, W, generator=g)
    want = torch.cat([F.conv2d(rnd(dtype, x[i:i + 1]), rnd(dtype, w[i]), b[i], 1, k // 2) for i in range(N)])
    if with_res:
        want = want + rnd(dtype, res)
    wd = [wpack(w[i], dtype) for i in range(N)]
    wp = torch.cat(wd, dim=1 if wd[0].dim() == 5 else 0).contiguous() if N > 1 else wd[0]
    out, outr = zeros((N, H, W, Cout), dtype), zeros((N, H, W, Cout), dtype)
    n0 = lib.otvm_debug_conv_cluster_launches()
    ops.conv2d(nhwc(x, dtype), wp, b.reshape(-1).to(DEV), out, pad=k // 2, res=nhwc(res, dtype) if with_res else None,
               out_relu=outr if relu2 else None, groups=N)
    torch.cuda.synchronize()
    assert lib.otvm_debug_conv_cluster_launches() == n0 + 1, "the cluster split-K path was not taken"
    assert rel_err(nchw(out), want) < TOL[dtype]
    if relu2:
        assert rel_err(nchw(outr), F.relu(want)) < TOL[dtype]


@pytest.mark.parametrize("dtype", TC_DTYPES)
@pytest.mark.parametrize("nslice", [2, 4])
def test_conv2d_fused_groupnorm_channel_slices(nslice, dtype):
    """a wide normalised layer as channel slices (otvm_conv_params.gn_group_ch): every slice holds whole GroupNorm groups,
    has its own statistics slot and takes the fused path; together they equal GroupNorm(32) of the whole layer -- at the
    FBA layer4 shape (512 -> 2048 at 64 x 64: 512 CTAs unsliced, which cannot be one co-resident wave)"""
    ops = _ops()
    Cin, Cout, H, W = 512, 2048, 64, 64
    g = torch.Generator().manual_seed(nslice)
    x = torch.randn(1, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 1, 1, generator=g) / math.sqrt(Cin)
    gamma, beta = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g)
    res = torch.randn(1, Cout, H, W, generator=g)
    y = F.relu(F.group_norm(F.conv2d(rnd(dtype, x), rnd(dtype, w)), 32, gamma, beta, 1e-5) + rnd(dtype, res))
    xd, wd, rd = nhwc(x, dtype), wpack(w, dtype), nhwc(res, dtype)
    out, raw = zeros((1, H, W, Cout), dtype), zeros((1, H, W, Cout), dtype)
    gd, bd = gamma.to(DEV), beta.to(DEV)
    whole = ops.conv2d(xd, wd, None, out, gn_stats=torch.zeros(72, dtype=torch.float64, device=DEV), gn_stats_zeroed=True,
                       gn_fuse=(gd, bd, 1e-5), gn_raw_out=raw, res=rd, act=ops.ACT_RELU, query_fuse=True)
    assert whole is False, "the unsliced layer is more than one co-resident wave"
    step = Cout // nslice
    wv0 = wd[:, :step] if wd.dim() == 5 else wd[:step]
    if not ops.conv2d(xd, wv0, None, out[..., :step], gn_stats=torch.zeros(72, dtype=torch.float64, device=DEV),
                      gn_stats_zeroed=True, gn_fuse=(gd[:step], bd[:step], 1e-5), gn_raw_out=raw[..., :step], res=rd[..., :step],
                      act=ops.ACT_RELU, gn_group_ch=Cout // 32, query_fuse=True):
        assert dtype == X3 and nslice == 2, "only three planes of 1024 channels exceed one co-resident wave"
        pytest.skip("three planes: 256 CTAs of this footprint are not co-resident (the engine takes 4 slices)")
    for i in range(nslice):
        a, c = i * step, (i + 1) * step
        wv = wd[:, a:c] if wd.dim() == 5 else wd[a:c]
        fused = ops.conv2d(xd, wv, None, out[..., a:c], gn_stats=torch.zeros(72, dtype=torch.float64, device=DEV),
                           gn_stats_zeroed=True, gn_fuse=(gd[a:c], bd[a:c], 1e-5), gn_raw_out=raw[..., a:c], res=rd[..., a:c],
                           act=ops.ACT_RELU, gn_group_ch=Cout // 32)
        assert fused is True
    torch.cuda.synchronize()
    assert rel_err(nchw(out), y) < (1e-2 if dtype == torch.bfloat16 else 1e-4)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("Cout,ld", [(7, 8), (10, 12)])
def test_head_conv_fba(Cout, ld, dtype):
    """the 1x1 head convolution + clamp / sigmoid / fba_fusion in one pointwise kernel against F.conv2d + the oracle's
    fusion (FBA/models.py:279-288, 347, 383-390, 415) on the stored activations"""
    import otvm_oracle as O
    ops = _ops()
    H, W = 24, 40
    g = torch.Generator().manual_seed(Cout)
    x = torch.randn(1, 16, H, W, generator=g)
    w = torch.randn(Cout, 16, 1, 1, generator=g) / 3
    b = torch.randn(Cout, generator=g) / 3
    img = torch.rand(1, 3, H, W, generator=g)
    raw_want = F.conv2d(rnd(dtype, x), w, b)
    fused_want = O._head(raw_want[:, :7], img)                                   # [1,7,H,W]: alpha, F, B
    P = H * W
    extras = torch.zeros(P, 8, device=DEV); extras[:, :3] = img[0].permute(1, 2, 0).reshape(P, 3).to(DEV)
    raw = torch.full((1, H, W, ld), 7.0, device=DEV)
    out7 = torch.zeros(P, 8, device=DEV)
    cat = zeros((1, H, W, 80), dtype)
    ops.head_conv_fba(nhwc(x, dtype), w.view(Cout, 16).to(DEV), b.to(DEV), raw, extras, P, out7, cat[..., 72:73], cat.stride(2))
    torch.cuda.synchronize()
    tol = 1e-2 if dtype == torch.bfloat16 else 1e-5
    assert rel_err(raw[0, ..., :Cout].permute(2, 0, 1).cpu(), raw_want[0]) < 1e-5
    assert float(raw[..., Cout:].abs().max()) == 0.0                             # the pad columns of a quad are zeroed
    assert rel_err(out7[:, :7].t().reshape(7, H, W).cpu(), fused_want[0]) < 1e-5
    assert rel_err(nchw(cat[..., 72:73])[0, 0], fused_want[0, 0]) < tol


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("s", [1, 2, 3, 6])
def test_conv1x1_on_ppm_cells(s, dtype):
    """the pyramid-pooling 1x1 convs run on s*s pixels (small-M kernel) with fused GroupNorm statistics"""
    ops = _ops()
    g = torch.Generator().manual_seed(s)
    x = torch.randn(1, 2048, s, s, generator=g); w = torch.randn(256, 2048, 1, 1, generator=g) / 45; b = torch.randn(256, generator=g)
    want = F.conv2d(rnd(dtype, x), rnd(dtype, w), b)
    out = zeros((1, s, s, 256), dtype)
    stats = torch.zeros(64, dtype=torch.float64, device=DEV)
    ops.conv2d(nhwc(x, dtype), wpack(w, dtype), b.to(DEV), out, gn_stats=stats)
    assert rel_err(nchw(out), want) < TOL[dtype]
    q = rnd(dtype, nchw(out))[0].double().reshape(32, -1)
    assert rel_err(stats.cpu().view(32, 2)[:, 0], q.sum(1)) < 1e-5
    assert rel_err(stats.cpu().view(32, 2)[:, 1], (q * q).sum(1)) < 1e-5


def test_fused_groupnorm_next_to_foreign_work():
    """the grid-synchronising GroupNorm-fused convolution while ANOTHER stream keeps the SMs busy (cuBLAS GEMMs the
    library knows nothing about): its CTAs become resident late, the bounded barrier must simply wait (no trap, no
    error flag) and the results must equal the quiet run"""
    ops = _ops()
    dtype = X2
    g = torch.Generator().manual_seed(3)
    Cin, Cout, H, W = 64, 256, 128, 128                       # 256 CTAs: two per SM
    x = torch.randn(1, Cin, H, W, generator=g); w = torch.randn(Cout, Cin, 1, 1, generator=g) / 8
    gamma, beta = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g)
    xd, wd = nhwc(x, dtype), wpack(w, dtype)
    y = F.relu(F.group_norm(F.conv2d(rnd(dtype, x), rnd(dtype, w)), 32, gamma, beta, 1e-5))

    def run():
        arena = torch.zeros(72, dtype=torch.float64, device=DEV)
        out = zeros((1, H, W, Cout), dtype); raw = zeros((1, H, W, Cout), dtype)
        fused = ops.conv2d(xd, wd, None, out, gn_stats=arena, gn_stats_zeroed=True,
                           gn_fuse=(gamma.to(DEV), beta.to(DEV), 1e-5), gn_raw_out=raw, act=ops.ACT_RELU)
        assert fused is True
        return out

    assert ops.device_error_flags() == 0
    side = torch.cuda.Stream()
    a = torch.randn(8192, 8192, device=DEV, dtype=torch.bfloat16)
    with torch.cuda.stream(side):
        for _ in range(40):
            a @ a                                             # ~1 ms each, every SM
    outs = [run() for _ in range(24)]
    torch.cuda.synchronize()
    assert ops.device_error_flags() == 0, "grid barrier of the fused GroupNorm timed out"
    for o in outs[::7]:
        assert rel_err(nchw(o), y) < 1e-4
