"""dev: the full-resolution 3x3 layers of a 512x512 frame through the persistent patch-mode kernel vs the one-tile-per-CTA
kernel (20 launches of each in one CUDA graph, us per launch)."""
import ctypes, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops, _lib
lib = _lib.load()
lib.otvm_debug_set_conv_persist.argtypes = [ctypes.c_int]
N = 20


def bench(ci, co, H, gn, mode):
    lib.otvm_debug_set_conv_persist(mode)
    g = torch.Generator(device="cuda").manual_seed(ci * 1000 + co + H)      # the SAME problem for both kernels (round 1
    x = torch.randn(1, H, H, ci, device="cuda", generator=g).bfloat16()     # drew fresh inputs per call, so its max|diff|
    w = (torch.randn(co, 3, 3, ci, device="cuda", generator=g) / math.sqrt(ci * 9)).bfloat16()   # column meant nothing)
    b = torch.zeros(co, device="cuda")
    out = torch.empty(1, H, H, co, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(72, dtype=torch.float64, device="cuda") if gn else None
    def body():
        for _ in range(N): ops.conv2d(x, w, b, out, pad=1, gn_stats=stats, gn_stats_zeroed=True, act=ops.ACT_NONE if gn else ops.ACT_LEAKY)
    body(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): body()
    for _ in range(2): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / 5 / N, out.float().cpu()


for ci, co, H, gn in [(64, 64, 512, True), (96, 64, 512, True), (96, 32, 512, False), (64, 32, 512, False), (32, 16, 512, False),
                      (64, 64, 256, True), (64, 64, 1024, True)]:
    t0, o0 = bench(ci, co, H, gn, 0)
    t1, o1 = bench(ci, co, H, gn, 1)
    fl = 2.0 * H * H * co * 9 * ci
    by = H * H * (ci + co) * 2
    print(f"{ci}->{co} k3 {H}^2 gn={int(gn)}: one-tile {t0:7.2f} us   persistent {t1:7.2f} us  ({fl / t1 / 1e6:.0f} TFLOP/s, {by / t1 / 1e3:.0f} GB/s)"
          f"   max|diff| {float((o0 - o1).abs().max()):.3g}", flush=True)
lib.otvm_debug_set_conv_persist(-1)
