"""Host-side mirror of the reference's model interface for the inference path (drop-in boundary).

Same class names, constructor arguments, ``forward`` keyword arguments, return tuple and ``state_dict``
keys as the reference, so ``eval.py`` runs unchanged on top of it:

* ``FullModel_eval``  <- models/trimap/model.py:173-281 (``forward(..., memorize=, segment=, memories=, hid=)``)
* ``EvalModel``       <- models/alpha/model.py:314-512  (``forward(a, fg, bg, tri, tri_gt, first_frame,
  last_frame, memorize, max_memory_num, large_input)`` -> ``(scaled_imgs, preds_trimap, tri_gt, preds_alpha,
  scaled_gts)``)

The modules only *hold* the parameters (785 keys, strict ``load_state_dict`` like eval.py:79); all arithmetic
is done by :class:`otvm_b200.engine.Engine` through the sm_100a C ABI.  There is no PyTorch fallback: on a
machine without the compiled library or without a CUDA device ``forward`` raises.
"""
from __future__ import annotations

import os

import torch
from torch import nn

from .engine import PRECISION_MODES, Engine
from .spec import state_spec

_PARAM_ROLES = ("conv_w", "conv_b", "norm_w", "norm_b")
_TORCH_DT = {"float32": torch.float32, "int64": torch.int64}
# precision modes (engine.PRECISION_MODES): "bf16x2" (alias "fast", the default) and "bf16x3" ("strict") run every
# contraction on the tcgen05 tensor cores with split-bf16 operands and meet the fp32 reference to 1e-2 / 1e-3; "bf16" is
# plain bf16 storage (fastest, but it misses the reference by up to 0.7 on alpha: DESIGN.md section 4); "fp32" = FFMA.
PRECISIONS = PRECISION_MODES
DEFAULT_PRECISION = "bf16x2"


class _Node(nn.Module):
    """Anonymous container so that dotted state_dict keys match the reference module tree."""


def _grow(root: nn.Module, prefix: str):
    """Register every spec entry under ``prefix`` on ``root`` (prefix stripped)."""
    from .fixtures import _const
    for name, e in state_spec().items():
        if not name.startswith(prefix) or (prefix == "" and name.startswith("trimap.")):
            continue
        parts = name[len(prefix):].split(".")
        m = root
        for p in parts[:-1]:
            if not hasattr(m, p):
                m.add_module(p, _Node())
            m = getattr(m, p)
        if e.role == "const":
            t = torch.from_numpy(_const(name, e.shape)).clone()
        elif e.role == "bn_var" or e.role == "norm_w":
            t = torch.ones(e.shape, dtype=_TORCH_DT[e.dtype])
        else:
            t = torch.zeros(e.shape, dtype=_TORCH_DT[e.dtype])
        if e.role in _PARAM_ROLES:
            m.register_parameter(parts[-1], nn.Parameter(t, requires_grad=False))
        else:
            m.register_buffer(parts[-1], t)


class FullModel_eval(nn.Module):
    """Trimap propagation wrapper (reference models/trimap/model.py:173).  Holds ``model.*`` (STM)."""

    def __init__(self, dilate_kernel=None, eps=0, ignore_label=255, stage=4, hdim=16):
        super().__init__()
        if stage != 4:
            raise NotImplementedError("otvm_b200 implements the stage-4 inference path only")
        self.DILATION_KERNEL, self.EPS, self.stage, self.hdim = dilate_kernel, eps, stage, hdim
        self.num_object = 1
        self.memory_update = False
        _grow(self, "trimap.")

    # -- the wrapper runs on its own too (reference models/trimap/model.py:247-264) ---------------------------
    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._engine = None
        return r

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def set_precision(self, precision: str):
        assert precision in PRECISIONS
        self.precision, self._engine = precision, None
        return self

    @property
    def engine(self) -> Engine:
        if getattr(self, "_engine", None) is None:
            dev = self.IMG_MEAN.device
            if dev.type != "cuda":
                raise RuntimeError("otvm_b200 runs on a CUDA device only (no CPU fallback): call model.cuda()")
            sd = {"trimap." + k: v for k, v in self.state_dict().items()}
            precision = getattr(self, "precision", None) or os.environ.get("OTVM_PRECISION", DEFAULT_PRECISION)
            dtype, planes = PRECISIONS[precision]
            self._engine = Engine(sd, dtype, dev, int(os.environ.get("OTVM_BANK_CAPACITY", "16")), fba=False, planes=planes)
        return self._engine

    @torch.no_grad()
    def forward(self, a, fg, bg, tri=None, first_frame=False, og_shape=None, memorize=False, segment=False,
                memories=None, hid=None, save_memory=False, max_memory_num=2, memory_update=False, memorize_gt=False):
        """``memorize=True``: ``bg`` = frame (RGB in [0,1]) [1,3,H,W], ``tri`` [1,3,H,W], ``a`` = alpha [1,1,H,W],
        ``hid`` [1,16,H,W] -> ``{'key': [1,1,128,1,h,w], 'val': [1,1,512,1,h,w]}`` (``_forward_memorize`` :227-239).
        ``segment=True``: ``fg`` = query frame [1,3,H,W], ``memories`` = dict of key/val banks [1,1,C,T,h,w] -> trimap
        logits [1,3,H,W] (``_forward_segment`` :241-245).  The stage-1 propagation loop (neither flag) is not part of the
        stage-4 inference path."""
        if memorize:
            if bg.shape[0] != 1:
                raise NotImplementedError("the reference memorises one frame at a time (b=1)")
            masks = torch.cat([tri, a, hid], dim=1) if self.hdim > 0 else tri
            key, val = self.engine.stm_memorize(bg, masks)
            return {"key": key, "val": val}
        if segment:
            if fg.shape[0] != 1:
                raise NotImplementedError("the reference segments one frame at a time (b=1)")
            return self.engine.stm_segment(fg, memories["key"], memories["val"])
        raise NotImplementedError("FullModel_eval without memorize= / segment= is the stage-1 propagation loop "
                                  "(models/trimap/model.py:174-225); otvm_b200 covers the stage-4 inference path")


class EvalModel(nn.Module):
    """Alpha + trimap eval model (reference models/alpha/model.py:314)."""

    def __init__(self, dilate_kernel=None, eps=0, trimap=None, stage=4, precision=None):
        super().__init__()
        if stage != 4 or trimap is None:
            raise NotImplementedError("otvm_b200 implements the stage-4 inference path only")
        self.stage, self.refinement = stage, True
        self.DILATION_KERNEL, self.EPS = dilate_kernel, eps
        self.IMG_SCALE = 1.0 / 255
        self.TRIMAP_CHANNEL = 8
        self.memory_update = False
        _grow(self, "")
        self.trimap = trimap
        self.precision = precision or os.environ.get("OTVM_PRECISION", DEFAULT_PRECISION)
        assert self.precision in PRECISIONS, self.precision
        self.bank_capacity = int(os.environ.get("OTVM_BANK_CAPACITY", "16"))
        self._engine = None
        self._last_plan = None

    @property
    def memories(self):
        """``{'key': [1,1,128,T,h,w], 'val': [1,1,512,T,h,w]}`` fp32 copies of the bank in the reference's layout and
        order (models/alpha/model.py:425-429,472-493), or ``None`` entries before the first frame.  The engine owns
        the bank (pre-allocated slots, in-place replacement, possibly split-bf16 storage and a deferred memorize
        pass): reading this property first runs a pending pass, so what is returned is complete."""
        pl = self._last_plan
        if self._engine is None or pl is None or pl.bank is None or pl.bank.T == 0:
            return {"key": None, "val": None}
        self._engine.flush(pl)
        h, w = pl.Hp // 16, pl.Wp // 16
        b = pl.bank
        return {"key": b.key_tensor().view(1, 1, -1, b.T, h, w), "val": b.val_tensor().view(1, 1, -1, b.T, h, w)}

    # -- parameter changes invalidate the packed weights -------------------------------------------------
    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._engine = None
        return r

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def set_precision(self, precision: str):
        assert precision in PRECISIONS
        self.precision, self._engine = precision, None
        return self

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            dev = self.IMG_MEAN.device
            if dev.type != "cuda":
                raise RuntimeError("otvm_b200 runs on a CUDA device only (no CPU fallback): call model.cuda()")
            dtype, planes = PRECISIONS[self.precision]
            self._engine = Engine(self.state_dict(), dtype, dev, self.bank_capacity, planes=planes)
        return self._engine

    @torch.no_grad()
    def forward(self, a, fg, bg, tri=None, tri_gt=None, first_frame=False, last_frame=False,
                memorize=False, max_memory_num=2, large_input=False):
        if a.shape[0] != 1 or a.shape[1] != 1:
            raise NotImplementedError("eval.py feeds one frame at a time (batch 1, sample length 1)")
        if self.DILATION_KERNEL is None:
            # the reference draws a random radius in [0, 25] per call when none is given (models/alpha/model.py:352-354);
            # eval.py always passes one (eval.py:67-72), and a random trimap width has no parity meaning
            raise NotImplementedError("dilate_kernel=None (random trimap width) is a training-time option; pass the radius")
        if self.EPS != 0:
            raise NotImplementedError("eps != 0 (models/alpha/model.py:345-346) is not implemented; eval.py uses EPS = 0")
        eng = self.engine
        H, W = a.shape[-2:]
        dev = self.IMG_MEAN.device
        # the engine copies these into its static input buffers (H2D straight from pinned host memory is fine)
        a_ = a.float().reshape(H, W)
        fg_ = fg.float().reshape(3, H, W)
        bg_ = bg.float().reshape(3, H, W)
        user_tri = None
        if first_frame and (tri is not None or tri_gt is not None):
            # models/alpha/model.py:395-401: a user trimap (BGR, 0..255) or a GT one-hot trimap seeds frame 0
            user_tri = (tri.to(dev).float().flip([2]) * self.IMG_SCALE) if tri is not None else tri_gt.to(dev).float()
            user_tri = user_tri.reshape(3, H, W)
        if max_memory_num > eng.bank_capacity:
            raise ValueError(f"max_memory_num={max_memory_num} exceeds the bank capacity {eng.bank_capacity} "
                             "(set OTVM_BANK_CAPACITY)")
        scaled, trimap, tri3, alpha = eng.frame(a_, fg_, bg_, first_frame=first_frame, last_frame=last_frame,
                                                memorize=memorize, max_memory_num=max_memory_num,
                                                radius=self.DILATION_KERNEL, user_tri=user_tri)
        self.memory_update = memorize
        pl = self._last_plan = eng.plan(H, W)
        t3 = tri3.view(pl.Hp, pl.Wp, 4)[pl.pad_top:pl.pad_top + H, pl.pad_left:pl.pad_left + W, :3]
        tri_gt_out = t3.permute(2, 0, 1).reshape(1, 1, 3, H, W)
        if tri_gt is not None:                                 # make_trimap_gt(None, trimap3=tri) (:363-366)
            cls = tri_gt.to(dev).float().reshape(3, H, W).max(dim=0)[1]
            tri_gt_out = torch.nn.functional.one_hot(cls, 3).permute(2, 0, 1).float().reshape(1, 1, 3, H, W)
        return (scaled.view(1, 1, 3, H, W), trimap.view(1, 1, 3, H, W), tri_gt_out,
                alpha.view(1, 1, 1, H, W), a)
