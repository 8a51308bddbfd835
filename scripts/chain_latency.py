"""dev: dependent-chain latency per kernel inside a CUDA graph (what a frame's critical path pays per layer).

A chain of NCH convolutions x -> y -> x -> ... of one shape (Cin == Cout, or alternating pair shapes) is captured in a
graph and replayed; us/kernel = replay time / NCH.  Compared with the kernel's own in-CTA span (scripts/conv_ts.py)
this separates launch-to-launch dependency cost from time inside the kernel."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops, _lib
lib = _lib.load()
NCH = 40


def chain(shapes, H, W, pdl, reps=20):
    """shapes: list of (Cin, Cout, k, dil) applied cyclically; consecutive shapes must chain (Cout_i == Cin_{i+1})"""
    lib.otvm_set_pdl(1 if pdl else 0)
    # every buffer exists BEFORE the capture (a tensor created inside body() would put its fill kernel into the graph)
    bufs = {(C, tag): torch.randn(1, H, W, C, device="cuda").bfloat16()
            for C in {c for sh in shapes for c in sh[:2]} for tag in (0, 1, 2)}
    def t(C, tag):
        return bufs[(C, tag)]
    wts = [((torch.randn(co, k, k, ci, device="cuda") / math.sqrt(ci * k * k)).bfloat16(), torch.zeros(co, device="cuda"))
           for ci, co, k, d in shapes]
    work = torch.empty(16 << 20, dtype=torch.float32, device="cuda")
    def body():
        x = t(shapes[0][0], 0)
        for i in range(NCH):
            ci, co, k, d = shapes[i % len(shapes)]
            y = t(co, 1 + i % 2)
            ops.conv2d(x, wts[i % len(shapes)][0], wts[i % len(shapes)][1], y, pad=d * (k // 2), dil=d, workspace=work)
            x = y
    body(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    for _ in range(3): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps / NCH


CASES = [
    ("256->256 k1 32^2 (16 CTAs)", [(256, 256, 1, 1)], 32, 32),
    ("256->1024->256 k1 32^2", [(256, 1024, 1, 1), (1024, 256, 1, 1)], 32, 32),
    ("256->256 k3 32^2", [(256, 256, 3, 1)], 32, 32),
    ("64->256->64 k1 128^2", [(64, 256, 1, 1), (256, 64, 1, 1)], 128, 128),
    ("128->512->128 k1 64^2", [(128, 512, 1, 1), (512, 128, 1, 1)], 64, 64),
    ("128->128 k3 64^2", [(128, 128, 3, 1)], 64, 64),
    ("64->64 k3 128^2", [(64, 64, 3, 1)], 128, 128),
    ("256->256 k3 64^2", [(256, 256, 3, 1)], 64, 64),
    ("64->64 k3 512^2", [(64, 64, 3, 1)], 512, 512),
]
if __name__ == "__main__":
    for name, shapes, H, W in CASES:
        on, off = chain(shapes, H, W, True), chain(shapes, H, W, False)
        print(f"{name:32s} us/kernel in graph chain: PDL {on:6.2f}   no-PDL {off:6.2f}", flush=True)
    lib.otvm_set_pdl(1)
