import ctypes, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops, _lib
lib = _lib.load()
lib.otvm_debug_set_conv_timestamps.argtypes = [ctypes.c_void_p]
for Cin, Cout, k, d, H, W in [(2048, 128, 1, 1, 8, 16), (2048, 16, 1, 1, 8, 16)]:
    x = torch.randn(1, H, W, Cin, device="cuda").bfloat16(); w = (torch.randn(Cout, k, k, Cin, device="cuda") / math.sqrt(Cin * k * k)).bfloat16()
    out = torch.empty(1, H, W, Cout, device="cuda", dtype=torch.bfloat16); b = torch.zeros(Cout, device="cuda")
    for _ in range(2): ops.conv2d(x, w, b, out, pad=d * (k // 2), dil=d)
    dbg = torch.zeros(4096, 64, dtype=torch.int64, device="cuda")
    lib.otvm_debug_set_conv_timestamps(ctypes.c_void_p(dbg.data_ptr()))
    ops.conv2d(x, w, b, out, pad=d * (k // 2), dil=d)
    torch.cuda.synchronize()
    lib.otvm_debug_set_conv_timestamps(None)
    t = dbg[dbg[:, 0] > 0].cpu()
    rel = (t - t[:, :1]).float()
    names = ["start", "setup done", "first TMA issued", "first stage full", "all MMA issued", "accum ready", "epilogue done", "cta end", "first tmem ld done", "chunk loop done", "after staging barrier"]
    print(f"Cin={Cin} Cout={Cout} k={k} {H}x{W}: ctas={len(t)}: " + " | ".join(f"{n}={rel[:, i].median():.0f}" for i, n in enumerate(names) if i not in (0, 7)))
    print("   mma  full-wait passed at:", [int(v) for v in rel[0, 16:32]])
    print("   tma  issued at          :", [int(v) for v in rel[0, 32:48]])
