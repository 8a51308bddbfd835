"""dev tool: fused Memory.read at BASELINE sizes: tcgen05 vs FFMA kernels, checked against torch on the GPU"""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops
torch.manual_seed(0)
for (hw_side, T) in [(32, 8), (32, 16), (64, 16)]:
    HW = hw_side * hw_side; M = T * HW
    k = torch.randn(M, 128, device="cuda").bfloat16(); v = torch.randn(512, M, device="cuda").bfloat16()
    q = (torch.randn(1, hw_side, hw_side, 128, device="cuda") * 1.5).bfloat16()
    out = torch.zeros(1, hw_side, hw_side, 1024, device="cuda", dtype=torch.bfloat16)
    ws = torch.zeros(ops.memory_read_workspace(M, HW, 128, 512, torch.bfloat16) // 4 + 1, device="cuda")
    p = torch.softmax((k.float() @ q.view(HW, 128).float().t()) / math.sqrt(128), dim=0)       # [M, HW]
    want = (v.float() @ p).t()                                                                   # [HW, 512]
    for simt in (False, True):
        for _ in range(3): ops.memory_read(k, v, M, q, out[..., :512], M, ws, force_simt=simt)
        n = 20 if not simt else 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): ops.memory_read(k, v, M, q, out[..., :512], M, ws, force_simt=simt)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        err = float((out[0, ..., :512].float().view(HW, 512) - want).abs().max() / want.abs().max())
        fl = 2.0 * M * HW * 640; by = (640 * M + 128 * HW + 512 * HW) * 2
        print(f"HW={HW} T={T} {'ffma' if simt else 'tcgen05'}: {us:8.1f} us  {fl/us/1e6:8.1f} TFLOP/s  {by/us/1e3:7.1f} GB/s  rel_err={err:.2e}", flush=True)
