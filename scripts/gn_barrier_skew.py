"""dev: where the ~5.5 us between 'accumulator ready' and 'first normalised chunk' of a GroupNorm-fused convolution go:
per-CTA arrival spread at the grid barrier (global time of the accumulator-ready stamp) vs the barrier mechanics."""
import ctypes, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops, _lib
from otvm_b200.split import SplitArena, split_planes
lib = _lib.load(); lib.otvm_debug_set_conv_timestamps.argtypes = [ctypes.c_void_p]
ar = SplitArena(2, 1 << 30, "cuda")
ws = torch.empty(16 << 20, device="cuda")
for Cin, Cout, k, d, H, W in [(64, 256, 1, 1, 128, 128), (256, 1024, 1, 1, 64, 64), (256, 256, 3, 2, 64, 64), (1024, 256, 1, 1, 64, 64)]:
    x = ar.alloc((1, H, W, Cin)); ar.write(x, torch.randn(1, H, W, Cin, device="cuda"))
    w = split_planes(torch.randn(Cout, k, k, Cin, device="cuda") / math.sqrt(Cin * k * k), 2)
    out = ar.alloc((1, H, W, Cout)); raw = ar.alloc((1, H, W, Cout)); res = ar.alloc((1, H, W, Cout)); ar.write(res, torch.randn(1, H, W, Cout, device="cuda"))
    g = torch.ones(Cout, device="cuda"); b = torch.zeros(Cout, device="cuda")
    st = torch.zeros(72, dtype=torch.float64, device="cuda")
    def run():
        st.zero_()
        return ops.conv2d(x, w, None, out, pad=d * (k // 2), dil=d, workspace=ws, act=ops.ACT_RELU, res=res, gn_stats=st,
                          gn_stats_zeroed=True, gn_fuse=(g, b, 1e-5), gn_raw_out=raw)
    for _ in range(3): run()
    dbg = torch.zeros(4096, 64, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    lib.otvm_debug_set_conv_timestamps(ctypes.c_void_p(dbg.data_ptr())); run(); torch.cuda.synchronize()
    lib.otvm_debug_set_conv_timestamps(None)
    t = dbg[dbg[:, 0] > 0].cpu().double()
    t0 = t[:, 10].min()
    start = (t[:, 10] - t0) / 1e3                                   # us, global
    acc = start + (t[:, 5] - t[:, 0]) / 1965.0                      # accumulator ready, global us (1965 MHz)
    c0 = start + (t[:, 8] - t[:, 0]) / 1965.0                       # first chunk of pass 2
    end = (t[:, 12] - t0) / 1e3
    q = lambda v: f"min {v.min():5.1f} med {v.median():5.1f} max {v.max():5.1f}"
    last = int(torch.argmax(acc))                                   # the CTA whose accumulator was ready last
    rel = lambda i: (t[last, i] - t[last, 5]) / 1965.0
    print(f"   last CTA, us after its accumulator: statistics pass {rel(13):.2f}, column sums + atomics {rel(14):.2f}, release-increment issued "
          f"{rel(15):.2f}, barrier passed {rel(16):.2f} ({int(t[last, 18])} closed polls), scale/shift ready {rel(17):.2f}, first chunk {rel(8):.2f}")
    med = lambda i: ((t[:, i] - t[:, 5]) / 1965.0).median()
    print(f"   median CTA: statistics pass {med(13):.2f}, atomics {med(14):.2f}, increment {med(15):.2f}, barrier passed {med(16):.2f}, "
          f"scale/shift {med(17):.2f}, first chunk {med(8):.2f}")
    print(f"Cin={Cin} Cout={Cout} k={k} {H}x{W} ctas={len(t)}: CTA start {q(start)} | accumulator ready {q(acc)} | pass-2 chunk 0 {q(c0)} | "
          f"end {q(end)}  => barrier released {c0.min() - acc.max():.1f} us after the LAST accumulator")
