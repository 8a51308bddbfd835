"""dev: per-tile phase stamps of the persistent patch-mode conv kernel (CTA 0 and a middle CTA)"""
import ctypes, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops, _lib
lib = _lib.load()
lib.otvm_debug_set_conv_timestamps.argtypes = [ctypes.c_void_p]
lib.otvm_debug_set_conv_persist.argtypes = [ctypes.c_int]
lib.otvm_debug_set_conv_persist(1)
names = ["prod issued", "mma: acc free", "mma: last patch in", "mma: committed", "epi: acc full", "epi: drained", "epi: store issued"]
for ci, co, H, gn in [(64, 64, 512, True), (64, 64, 512, False), (96, 64, 512, True), (32, 16, 512, False)]:
    x = torch.randn(1, H, H, ci, device="cuda").bfloat16()
    w = (torch.randn(co, 3, 3, ci, device="cuda") / math.sqrt(ci * 9)).bfloat16(); b = torch.zeros(co, device="cuda")
    out = torch.empty(1, H, H, co, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(72, dtype=torch.float64, device="cuda") if gn else None
    run = lambda: ops.conv2d(x, w, b, out, pad=1, gn_stats=stats, gn_stats_zeroed=True, act=ops.ACT_NONE if gn else ops.ACT_LEAKY)
    for _ in range(3): run()
    dbg = torch.zeros(148, 256, dtype=torch.int64, device="cuda")
    lib.otvm_debug_set_conv_timestamps(ctypes.c_void_p(dbg.data_ptr()))
    run(); torch.cuda.synchronize()
    lib.otvm_debug_set_conv_timestamps(None)
    t = dbg.cpu().view(148, 32, 8)
    print(f"== {ci}->{co} k3 {H}^2 gn={int(gn)}")
    for cta in (0, 77):
        t0 = int(t[cta, 0, 0])
        print(f"  CTA {cta}: tile | " + " | ".join(names))
        for i in range(15):
            if t[cta, i, 0] == 0: break
            print(f"    {i:2d} | " + " | ".join(f"{int(t[cta, i, k]) - t0:7d}" for k in range(7)))
lib.otvm_debug_set_conv_persist(-1)
