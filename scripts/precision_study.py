"""CPU study: which stored precision does each part of the frame need?  (dev tool, uses the oracle as the model)

Every convolution of the oracle is re-run with its INPUT activations and its weights rounded to an emulated storage
format (fp32 math, like a tensor-core MMA with fp32 accumulation):
  1 = bf16, 2 = bf16 hi+lo (two planes, ~16 mantissa bits), 3 = three planes (~24 bits), 0 = exact fp32.
A policy maps layer-name prefixes to a format; the error of every traced tensor is reported against the exact oracle
on the same frame (teacher-forced bank).

    python scripts/precision_study.py [size] [fixture]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import otvm_oracle as O  # noqa: E402
from otvm_b200.fixtures import make_frame, make_state_dict  # noqa: E402
from util import rel_err, mean_err  # noqa: E402

_conv2d, _ws_conv2d = O.conv2d, O.ws_conv2d
POLICY = []          # [(prefix, planes)], first match wins
WCACHE = {}


def planes(x, n):
    if n == 0:
        return x
    out = torch.zeros_like(x)
    r = x
    for _ in range(n):
        h = r.to(torch.bfloat16).float()
        out = out + h
        r = r - h
    return out


def fmt(name):
    """(activation planes, weight planes) of the first matching policy entry"""
    for e in POLICY:
        if name.startswith(e[0]):
            return (e[1], e[2] if len(e) > 2 else e[1])
    return (0, 0)


def q_conv2d(sd, name, x, stride=1, padding=0, dilation=1):
    n, nw = fmt(name)
    w = planes(sd[name + ".weight"], nw)
    return O.F.conv2d(planes(x, n), w, sd.get(name + ".bias"), stride, padding, dilation)


def q_ws_conv2d(sd, name, x, stride=1, padding=0, dilation=1):
    n, nw = fmt(name)
    key = (name, nw)
    if key not in WCACHE:
        WCACHE[key] = planes(O.ws_weight(sd[name + ".weight"]), nw)
    return O.F.conv2d(planes(x, n), WCACHE[key], sd.get(name + ".bias"), stride, padding, dilation)


def run(policy, sd, frames, exact):
    """exact: list of (ref outputs, trace, memories) from the exact run; returns per-frame error dicts"""
    POLICY[:] = policy
    O.conv2d, O.ws_conv2d = q_conv2d, q_ws_conv2d
    try:
        m = O.OracleEvalModel(sd, dilate_kernel=12)
        rows = []
        for i, (a, fg, bg) in enumerate(frames):
            out = m(a, fg, bg, first_frame=(i == 0), last_frame=False, memorize=True, max_memory_num=8)
            ref, tr, mem = exact[i]
            e = {}
            for k in ("seg_logit", "m4", "k4", "conv5", "raw_decoder", "output", "raw_refine", "hid", "refine_output",
                      "mem_k", "mem_v"):
                if k in tr and k in m.trace:
                    e[k] = (rel_err(m.trace[k], tr[k]), mean_err(m.trace[k], tr[k]))
            e["alpha"] = (rel_err(out[3], ref[3]), mean_err(out[3], ref[3]))
            e["trimap"] = (rel_err(out[1], ref[1]), mean_err(out[1], ref[1]))
            rows.append(e)
            m.memories = {k: v.clone() for k, v in mem.items()}          # teacher forcing
        return rows
    finally:
        O.conv2d, O.ws_conv2d = _conv2d, _ws_conv2d


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    kind = sys.argv[2] if len(sys.argv) > 2 else "tempered"
    torch.set_num_threads(os.cpu_count())
    sd = make_state_dict(kind)
    frames = [make_frame(0, i, size, size) for i in range(2)]
    m = O.OracleEvalModel(sd, dilate_kernel=12)
    exact = []
    for i, (a, fg, bg) in enumerate(frames):
        out = m(a, fg, bg, first_frame=(i == 0), last_frame=False, memorize=True, max_memory_num=8)
        exact.append((out, dict(m.trace), {k: v.clone() for k, v in m.memories.items()}))
    policies = {
        "act x2, weights bf16": [("", 2, 1)],
        "act bf16, weights x2": [("", 1, 2)],
        "all bf16": [("", 1)],
        "all x2": [("", 2)],
        "all x3": [("", 3)],
        "STM bf16, FBA x2": [("trimap.", 1), ("", 2)],
        "STM bf16, FBA enc bf16, dec+refine x2": [("trimap.", 1), ("NET.encoder", 1), ("", 2)],
        "STM bf16, FBA enc x2, dec+refine bf16": [("trimap.", 1), ("NET.encoder", 2), ("", 1)],
        "STM x2, FBA x2, refine bf16": [("NET.refine", 1), ("", 2)],
        "only heads exact": [("NET.decoder.conv_up4", 0), ("NET.refine.pred", 0), ("", 1)],
    }
    only = os.environ.get("ONLY")
    for name, pol in policies.items():
        if only and name not in only.split(";"):
            continue
        rows = run(pol, sd, frames, exact)
        print(f"== {name}")
        for i, e in enumerate(rows):
            print(f"  frame {i}: " + "  ".join(f"{k}={v[0]:.1e}/{v[1]:.1e}" for k, v in e.items()))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
