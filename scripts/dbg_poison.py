import os, sys, torch
os.environ["OTVM_DEBUG_POISON"] = "1"; os.environ["OTVM_CUDA_GRAPHS"] = "0"; os.environ["OTVM_OVERLAP"] = sys.argv[1] if len(sys.argv) > 1 else "0"
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from frames_util import build_model
from otvm_b200 import ops
from otvm_b200.fixtures import make_frame
ops.GN_FUSE = False
model, _ = build_model("tempered", "bf16")
for i in range(4):
    a, fg, bg = make_frame(0, i, 128, 160)
    out = model(a.cuda(), fg.cuda(), bg.cuda(), first_frame=(i == 0), last_frame=False, memorize=(i % 3 != 2), max_memory_num=4)
    torch.cuda.synchronize()
    pl = model.engine.plan(128, 160)
    bad = [k for k, v in pl.bufs.items() if v.is_floating_point() and not torch.isfinite(v.float()).all()]
    print("frame", i, "alpha finite:", bool(torch.isfinite(out[3]).all()), "buffers with NaN:", bad[:40], flush=True)
bank = model.engine.bank(pl)
print("bank finite", bool(torch.isfinite(bank.keys.float()).all()), bool(torch.isfinite(bank.vals.float()).all()))
