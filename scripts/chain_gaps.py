"""dev: where the time between dependent kernels of a graph goes: %globaltimer stamps of every CTA of each conv in a
dependent chain (entry / after griddepcontrol.wait / exit), plus a chain of trivial pointwise kernels as the floor."""
import ctypes, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otvm_b200 import ops, _lib
lib = _lib.load()
lib.otvm_debug_set_conv_timestamps.argtypes = [ctypes.c_void_p]
NCH = 10


def run(name, ci, co, k, d, H, W, pdl):
    lib.otvm_set_pdl(1 if pdl else 0)
    xs = [torch.randn(1, H, W, ci, device="cuda").bfloat16(), torch.randn(1, H, W, co, device="cuda").bfloat16()]
    assert ci == co
    w = (torch.randn(co, k, k, ci, device="cuda") / math.sqrt(ci * k * k)).bfloat16(); b = torch.zeros(co, device="cuda")
    work = torch.empty(16 << 20, dtype=torch.float32, device="cuda")
    dbg = torch.zeros(NCH, 4096, 64, dtype=torch.int64, device="cuda")
    def body(stamp):
        for i in range(NCH):
            lib.otvm_debug_set_conv_timestamps(ctypes.c_void_p(dbg[i].data_ptr()) if stamp else None)
            ops.conv2d(xs[i % 2], w, b, xs[1 - i % 2], pad=d * (k // 2), dil=d, workspace=work)
        lib.otvm_debug_set_conv_timestamps(None)
    body(False); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body(True)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    t = dbg.cpu()
    t0 = None
    rows = []
    for i in range(NCH):
        m = t[i][t[i][:, 10] > 0]
        if len(m) == 0: continue
        if t0 is None: t0 = int(m[:, 10].min())
        rows.append((len(m), int(m[:, 10].min()) - t0, int(m[:, 10].max()) - t0, int(m[:, 11].min()) - t0, int(m[:, 11].max()) - t0,
                     int(m[:, 12].min()) - t0, int(m[:, 12].max()) - t0))
    print(f"{name} pdl={pdl}: per kernel (ns from first entry): ctas | entry min..max | wait-done min..max | exit min..max")
    for i, r in enumerate(rows):
        prev_end = rows[i - 1][6] if i else 0
        print(f"   k{i}: {r[0]:4d} | {r[1]:7d}..{r[2]:7d} | {r[3]:7d}..{r[4]:7d} | {r[5]:7d}..{r[6]:7d} | wait-done - prev exit = {r[3] - prev_end:6d}  body = {r[6] - r[3]:6d}")


def trivial_chain(n=40, reps=20):
    x = torch.randn(1, 16, 16, 64, device="cuda").bfloat16(); y = torch.empty(1, 8, 8, 64, device="cuda", dtype=torch.bfloat16)
    def body():
        for _ in range(n): ops.maxpool3x3s2(x, y)
    body(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): body()
    for _ in range(3): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps / n


if __name__ == "__main__":
    for pdl in (1, 0):
        lib.otvm_set_pdl(pdl)
        print(f"trivial maxpool chain pdl={pdl}: {trivial_chain():.2f} us/kernel")
    for pdl in (1, 0):
        run("256->256 k1 32^2", 256, 256, 1, 1, 32, 32, pdl)
    run("64->64 k3 128^2", 64, 64, 3, 1, 128, 128, 1)
    run("256->256 k3 64^2", 256, 256, 3, 1, 64, 64, 1)
    lib.otvm_set_pdl(1)
