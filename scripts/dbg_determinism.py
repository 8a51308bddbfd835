"""dev: run the same 6-frame clip several times per configuration in ONE process; report run-to-run differences"""
import os, sys, torch, ctypes
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from frames_util import build_model
from util import rel_err
from otvm_b200 import _lib, ops
from otvm_b200.fixtures import make_frame
lib = _lib.load(); lib.otvm_debug_set_conv_halo.argtypes = [ctypes.c_int]
WATCH = ["q_key", "m4in", "seg_logits", "x11", "cat1", "raw7", "raw10", "hid", "alpha_out"]
def run(n=6, **cfg):
    os.environ["OTVM_OVERLAP"] = cfg.get("overlap", "0"); os.environ["OTVM_CUDA_GRAPHS"] = cfg.get("graphs", "0")
    lib.otvm_set_pdl(cfg.get("pdl", 0)); ops.GN_FUSE = cfg.get("fuse", False); lib.otvm_debug_set_conv_halo(cfg.get("halo", 0))
    simt = cfg.get("simt_read", False)
    orig = ops.memory_read
    if simt: ops.memory_read = lambda *a, **k: orig(*a, **dict(k, force_simt=True))
    model, _ = build_model("tempered", "bf16")
    res = []
    for i in range(n):
        a, fg, bg = make_frame(0, i, 128, 160)
        out = model(a.cuda(), fg.cuda(), bg.cuda(), first_frame=(i == 0), last_frame=(i == n - 1), memorize=(i % 3 != 2), max_memory_num=4)
        torch.cuda.synchronize()
        b = model.engine.plan(128, 160).bufs
        res.append({k: b[k].float().clone().cpu() for k in WATCH if k in b})
    ops.memory_read = orig
    return res
def first_diff(A, B):
    for i, (x, y) in enumerate(zip(A, B)):
        for k in WATCH:
            if k in x and k in y and not torch.equal(x[k], y[k]):
                return f"frame {i} buffer {k} max|d|={float((x[k]-y[k]).abs().max()):.2e}"
    return "identical"
for name, cfg in [("baseline(all off)", {}), ("halo", dict(halo=-1)), ("simt_read", dict(simt_read=True)), ("pdl", dict(pdl=1)),
                  ("fuse", dict(fuse=True)), ("graphs", dict(graphs="1")), ("overlap", dict(overlap="1")),
                  ("all on", dict(halo=-1, pdl=1, fuse=True, graphs="1", overlap="1"))]:
    runs = [run(**cfg) for _ in range(4)]
    print(f"{name:20s}:", " | ".join(first_diff(runs[0], r) for r in runs[1:]), flush=True)
