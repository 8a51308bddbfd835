import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from frames_util import run_clip
for kind, prec, H, W in [("tempered", "fp32", 128, 128), ("default", "fp32", 128, 128), ("tempered", "fp32", 120, 152), ("tempered", "bf16", 128, 128)]:
    try:
        rows = run_clip(kind, prec, H, W, 3, max_mem=2 if H == 120 else 8)
        for i, e in enumerate(rows):
            print(kind, prec, H, W, "frame", i, " ".join(f"{k}={v:.1e}" for k, v in e.items()), flush=True)
    except Exception as ex:
        import traceback; traceback.print_exc()

rows = run_clip("tempered", "bf16", 256, 256, 3, with_mean=True)
for i, e in enumerate(rows):
    print("bf16 256 frame", i, " ".join(f"{k}={v[0]:.1e}/{v[1]:.1e}" for k, v in e.items()), flush=True)
