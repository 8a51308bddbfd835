"""dev: per-stage scale-relative max errors (engine vs oracle, teacher-forced) for precision modes / fixtures / sizes.
    python scripts/frame_errors.py kind size nframes precision [precision...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from frames_util import run_clip
kind, size, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
for prec in sys.argv[4:]:
    rows = run_clip(kind, prec, size, size, n)
    for i, e in enumerate(rows):
        print(f"{kind} {size} {prec} frame {i}: " + " ".join(f"{k}={v:.1e}" for k, v in e.items()), flush=True)
