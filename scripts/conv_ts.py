import ctypes, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops, _lib
lib = _lib.load()
lib.otvm_debug_set_conv_timestamps.argtypes = [ctypes.c_void_p]
for Cin, Cout, k, d, H, W in [(64, 256, 1, 1, 128, 128), (256, 256, 3, 1, 128, 128), (64, 64, 3, 1, 512, 512)]:
    x = torch.randn(1, H, W, Cin, device="cuda").bfloat16(); w = (torch.randn(Cout, k, k, Cin, device="cuda") / math.sqrt(Cin * k * k)).bfloat16()
    out = torch.empty(1, H, W, Cout, device="cuda", dtype=torch.bfloat16); b = torch.zeros(Cout, device="cuda")
    for _ in range(2): ops.conv2d(x, w, b, out, pad=d * (k // 2), dil=d)
    dbg = torch.zeros(4096, 16, dtype=torch.int64, device="cuda")
    lib.otvm_debug_set_conv_timestamps(ctypes.c_void_p(dbg.data_ptr()))
    ops.conv2d(x, w, b, out, pad=d * (k // 2), dil=d)
    torch.cuda.synchronize()
    lib.otvm_debug_set_conv_timestamps(None)
    t = dbg[dbg[:, 0] > 0].cpu()
    t0 = t[:, 0].min()
    rel = (t - t[:, :1]).float()
    print(f"Cin={Cin} Cout={Cout} k={k} {H}x{W}: ctas={len(t)} start spread={int((t[:,0]-t0).max())} cyc; per-CTA median cycles since CTA start:")
    for i, name in enumerate(["start", "setup done", "first TMA issued", "first stage full", "all MMA issued", "accum ready", "epilogue done", "cta end", "first tmem ld done", "chunk loop done", "after staging barrier"]):
        print(f"   {name:18s} med={rel[:, i].median():9.0f} max={rel[:, i].max():9.0f}")
    print("   kernel span (max end - min start):", int(t[:, 7].max() - t0))
