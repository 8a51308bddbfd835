"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_err(got, want):
    """max|got-want| / max|want| — the scale-relative error every tolerance in tests/ is quoted in."""
    got = torch.as_tensor(np.asarray(got) if not torch.is_tensor(got) else got).double().cpu()
    want = torch.as_tensor(np.asarray(want) if not torch.is_tensor(want) else want).double().cpu()
    assert got.shape == want.shape, (got.shape, want.shape)
    return float((got - want).abs().max() / max(float(want.abs().max()), 1e-12))


def frac_off(got, want, tol):
    """fraction of elements whose scale-relative error exceeds tol."""
    got = torch.as_tensor(got).double().cpu()
    want = torch.as_tensor(want).double().cpu()
    scale = max(float(want.abs().max()), 1e-12)
    return float(((got - want).abs() > tol * scale).double().mean())


def mean_err(got, want):
    """mean|got-want| / max|want| — robust companion of rel_err for bf16 paths (a few pixels flip trimap class)."""
    got = torch.as_tensor(np.asarray(got) if not torch.is_tensor(got) else got).double().cpu()
    want = torch.as_tensor(np.asarray(want) if not torch.is_tensor(want) else want).double().cpu()
    assert got.shape == want.shape, (got.shape, want.shape)
    return float((got - want).abs().mean() / max(float(want.abs().max()), 1e-12))
