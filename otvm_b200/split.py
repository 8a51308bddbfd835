"""Split-bf16 tensors: device memory for activations stored as P bf16 planes (``include/otvm_b200.h``, "dtype").

A value is the sum of its planes: plane 0 = bf16(v), plane 1 = bf16(v - plane 0), plane 2 = bf16(remainder) -- 8 / 16 /
24 significant bits for P = 1 / 2 / 3.  The tensor cores multiply planes pairwise (``csrc/conv_tc.cu``,
``csrc/memory_read_tc.cu``), which is how the tcgen05 path reproduces the reference's fp32 forward
(``models/trimap/STM.py``, ``models/alpha/FBA/models.py`` are fp32 end to end) within 1e-2 / 1e-3.

All split tensors that meet in one kernel call must share ONE plane stride, so they are carved from an arena
``[P][capacity]``: a tensor handed around in Python is the plane-0 view (an ordinary bf16 ``torch.Tensor``; channel
slices ``t[..., a:b]`` stay valid for every plane), the C ABI finds plane k at ``data_ptr + k * plane_stride``.
PyTorch is used for memory only; :meth:`SplitArena.read` / :meth:`SplitArena.write` convert at the boundary (tests,
stand-alone STM entry points).
"""
from __future__ import annotations

import math
import weakref
from typing import Optional

import torch

BF16, BF16X2, BF16X3 = 1, 2, 3
_ALIGN = 1024          # bytes: TMA base alignment (16) and whole swizzle atoms

_ARENAS: "weakref.WeakSet[SplitArena]" = weakref.WeakSet()     # live arenas (an arena dies with its engine plan)


class SplitArena:
    """``planes`` x ``plane_bytes`` bump allocator.  ``device='meta'`` measures only (planning pass)."""

    def __init__(self, planes: int, plane_bytes: int, device):
        assert planes in (1, 2, 3)
        self.planes = planes
        self.device = torch.device(device)
        self.plane_bytes = (max(int(plane_bytes), 4096) + 4095) // 4096 * 4096
        self.off = 0                                   # bytes used in every plane
        self.meta = self.device.type == "meta"
        self.buf = torch.empty(planes, self.plane_bytes // 2, dtype=torch.bfloat16, device=self.device)
        if not self.meta:
            self.base = self.buf.data_ptr()
            _ARENAS.add(self)

    # -- C ABI ------------------------------------------------------------------------------------------
    @property
    def dtype_word(self) -> int:
        """OTVM_SPLIT_DTYPE(planes, plane_bytes); a one-plane arena is plain OTVM_BF16"""
        if self.planes == 1:
            return BF16
        return (BF16X3 if self.planes == 3 else BF16X2) | ((self.plane_bytes // 4096) << 8)

    # -- allocation -------------------------------------------------------------------------------------
    def alloc(self, shape, zero: bool = False) -> torch.Tensor:
        n = int(math.prod(shape))
        nbytes = (n * 2 + _ALIGN - 1) // _ALIGN * _ALIGN
        if self.meta:
            self.off += nbytes
            return torch.empty(tuple(shape), dtype=torch.bfloat16, device="meta")
        if self.off + nbytes > self.plane_bytes:
            raise MemoryError(f"split arena exhausted: {self.off + nbytes} > {self.plane_bytes} bytes per plane")
        e0 = self.off // 2
        self.off += nbytes
        if zero:
            self.buf[:, e0:e0 + n].zero_()
        return self.buf[0, e0:e0 + n].view(tuple(shape))

    # -- boundary conversions (PyTorch, not on the hot path) ----------------------------------------------
    def planes_of(self, t: torch.Tensor) -> torch.Tensor:
        """[planes, *t.shape] strided view of every plane of the plane-0 view ``t``"""
        off = t.storage_offset()
        assert t.untyped_storage().data_ptr() == self.buf.untyped_storage().data_ptr(), "tensor is not from this arena"
        return self.buf.as_strided((self.planes, *t.shape), (self.plane_bytes // 2, *t.stride()), off)

    def read(self, t: torch.Tensor) -> torch.Tensor:
        """fp32 value of a split tensor (sum of its planes)"""
        return self.planes_of(t).float().sum(dim=0)

    def write(self, t: torch.Tensor, value: torch.Tensor):
        """store an fp32 tensor (broadcastable to ``t``) as planes"""
        r = value.to(device=t.device, dtype=torch.float32).expand(t.shape).clone()
        pl = self.planes_of(t)
        for k in range(self.planes):
            h = r.to(torch.bfloat16)
            pl[k].copy_(h)
            r -= h.float()
        return t


def arena_of(t: torch.Tensor) -> Optional[SplitArena]:
    """the arena a bf16 tensor was carved from (None: an ordinary bf16 tensor = one plane)"""
    if t.dtype != torch.bfloat16 or not _ARENAS:
        return None
    p = t.data_ptr()
    for a in list(_ARENAS):
        if a.base <= p < a.base + a.plane_bytes:
            return a
    return None


def to_float(t: torch.Tensor) -> torch.Tensor:
    """fp32 value of any activation tensor (fp32, bf16 or split)"""
    a = arena_of(t)
    return a.read(t) if a is not None and a.planes > 1 else t.float()


def split_planes(x: torch.Tensor, planes: int) -> torch.Tensor:
    """[planes, *x.shape] bf16 planes of an fp32 tensor (weights are packed with this)"""
    r = x.float().clone()
    out = []
    for _ in range(planes):
        h = r.to(torch.bfloat16)
        out.append(h)
        r = r - h.float()
    return torch.stack(out, dim=0)
