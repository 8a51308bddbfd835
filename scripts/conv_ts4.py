"""dev: where a CTA of the one-tile conv kernel spends its time, split operands (bf16x2), for the short-K layers that are
bound by their epilogue: per-CTA clock64 stamps (median over CTAs, cycles) + the kernel's span from %globaltimer.
   python scripts/conv_ts4.py [bf16x2]"""
import ctypes, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops, _lib
from otvm_b200.split import SplitArena, split_planes
lib = _lib.load()
lib.otvm_debug_set_conv_timestamps.argtypes = [ctypes.c_void_p]
fmt = next((a for a in sys.argv[1:] if a.startswith("bf16")), "bf16x2")
planes = {"bf16": 1, "bf16x2": 2, "bf16x3": 3}[fmt]
ar = SplitArena(planes, 1 << 30, "cuda")
def act(shape, rand=True):
    t = ar.alloc(shape)
    if rand: ar.write(t, torch.randn(shape, device="cuda"))
    return t
ws = torch.empty(16 << 20, device="cuda")
shapes = [(64, 256, 1, 1, 128, 128), (128, 512, 1, 1, 64, 64), (256, 1024, 1, 1, 32, 32), (256, 1024, 1, 1, 64, 64),
          (1024, 256, 1, 1, 32, 32), (256, 256, 3, 1, 32, 32), (128, 128, 3, 1, 64, 64), (256, 64, 1, 1, 128, 128)]
for Cin, Cout, k, d, H, W in shapes:
    x = act((1, H, W, Cin))
    w = torch.randn(Cout, k, k, Cin, device="cuda") / math.sqrt(Cin * k * k)
    w = split_planes(w, planes) if planes > 1 else w.bfloat16()
    out = act((1, H, W, Cout), rand=False); raw = act((1, H, W, Cout), rand=False); res = act((1, H, W, Cout))
    b = torch.zeros(Cout, device="cuda"); g = torch.ones(Cout, device="cuda")
    for name in ("plain", "res", "gnfuse+res"):
        st = torch.zeros(72, dtype=torch.float64, device="cuda")
        def run():
            if name == "plain": ops.conv2d(x, w, b, out, pad=d * (k // 2), dil=d, workspace=ws, act=ops.ACT_RELU)
            elif name == "res": ops.conv2d(x, w, b, out, pad=d * (k // 2), dil=d, workspace=ws, act=ops.ACT_RELU, res=res)
            else:
                st.zero_()
                return ops.conv2d(x, w, None, out, pad=d * (k // 2), dil=d, workspace=ws, act=ops.ACT_RELU, res=res, gn_stats=st,
                                  gn_stats_zeroed=True, gn_fuse=(g, b, 1e-5), gn_raw_out=raw)
        for _ in range(2): fused = run()
        dbg = torch.zeros(4096, 64, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        lib.otvm_debug_set_conv_timestamps(ctypes.c_void_p(dbg.data_ptr()))
        run()
        torch.cuda.synchronize()
        lib.otvm_debug_set_conv_timestamps(None)
        t = dbg[dbg[:, 0] > 0].cpu()
        rel = (t[:, :10] - t[:, :1]).float()
        med = lambda i: int(rel[:, i].median())
        span = (t[:, 12].max() - t[:, 10].min()).item() / 1e3
        cta = (t[:, 12] - t[:, 10]).float().median().item() / 1e3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        print(f"Cin={Cin} Cout={Cout} k={k} {H}x{W} {name}{'' if name != 'gnfuse+res' else ' fused=' + str(fused)}: ctas={len(t)} "
              f"pdl={med(1)} full0={med(3)} mma_issued={med(4)} accum={med(5)} chunk0={med(8)} chunks_done={med(9)} epi_done={med(6)} "
              f"end={med(7)} | cta {cta:.1f} us, kernel span {span:.1f} us, loop {e0.elapsed_time(e1) * 50:.1f} us/launch")
