"""Per-frame OTVM inference engine over the sm_100a C ABI.

Host-side mirror of the reference dataflow (paths relative to the reference root):

* ``segment``  – STM.segment, models/trimap/STM.py:239-257 (Encoder_Q :92-102, KeyValue :173-174,
  Memory.read :144-163, Decoder :129-137)
* ``matting``  – MattingModule.forward, models/alpha/FBA/models.py:32-45 (ResnetDilated :251-269,
  fba_decoder :351-392, RefinementModule :417-435)
* ``memorize`` – STM.memorize, models/trimap/STM.py:201-228 (Encoder_M :56-74)
* ``MemoryBank`` – the eviction policy of EvalModel.forward, models/alpha/model.py:472-493

Everything numerical runs in the hand-written CUDA kernels of ``csrc/``; PyTorch only owns device memory and
the stream.  Weights are prepared ONCE per ``load_state_dict`` (they are constant in eval): eval-mode
BatchNorm is folded into the preceding convolution, weight standardisation (layers_WS.py:15-21) is applied
to the weights themselves, the five 7x7 stems of Encoder_M (STM.py:63,67) become one 22-channel stem, and
all filters are re-laid out [Cout][KH][KW][Cin] (K-major) in the activation dtype.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import ops
from .ops import ACT_LEAKY, ACT_NONE, ACT_RELU
from .split import SplitArena, split_planes, to_float

# precision modes: name -> (activation dtype, bf16 planes per value).  "bf16x2" / "bf16x3" store every activation and
# weight as 2 / 3 bf16 planes and multiply them pairwise on the tensor cores (split.py): fp32-grade results, which the
# fp32 reference needs, at 3 / 6 tcgen05 products per K step.  "fp32" is the FFMA path.
PRECISION_MODES = {"fp32": (torch.float32, 0), "bf16": (torch.bfloat16, 1), "bf16x2": (torch.bfloat16, 2),
                   "bf16x3": (torch.bfloat16, 3), "fast": (torch.bfloat16, 2), "strict": (torch.bfloat16, 3)}

DE, DO = 128, 512          # key / value channels (STM.py:184-185)
BN_EPS = 1e-5


def _ws(w: torch.Tensor) -> torch.Tensor:
    """layers_WS.py:15-21 in fp32 (same association order as the reference)."""
    m = w.mean(dim=1, keepdim=True).mean(dim=2, keepdim=True).mean(dim=3, keepdim=True)
    w = w - m
    std = torch.sqrt(torch.var(w.view(w.size(0), -1), dim=1) + 1e-12).view(-1, 1, 1, 1) + 1e-5
    return w / std


class PackedWeights:
    """state_dict -> kernel-ready tensors on ``device`` in ``dtype``."""

    def __init__(self, sd: Dict[str, torch.Tensor], dtype: torch.dtype, device, fba: bool = True, planes: int = 1):
        """``fba=False`` packs the trimap-propagation network only (a stand-alone ``FullModel_eval``).
        ``planes`` > 1: bf16 weights are stored split, [planes][Cout][KH][KW][Cin]."""
        self.dtype, self.device, self.planes = dtype, device, planes
        self.conv: Dict[str, tuple] = {}
        self.norm: Dict[str, tuple] = {}
        # stem channel padding: the tcgen05 path needs Cin % 16 == 0 (one 32-byte swizzle span per pixel)
        bf = dtype == torch.bfloat16
        self.cin_img, self.cin_mem = (16, 32) if bf else (4, 24)
        f = lambda k: sd[k].detach().to("cpu", torch.float32)

        def put(name, w, b, cin_pad=None):
            cout, cin, kh, kw = w.shape
            if cin_pad and cin_pad > cin:
                w = torch.cat([w, w.new_zeros(cout, cin_pad - cin, kh, kw)], dim=1)
            wp = w.permute(0, 2, 3, 1).contiguous()
            wp = split_planes(wp, planes).to(device) if bf and planes > 1 else wp.to(device=device, dtype=dtype)
            bp = b.contiguous().to(device=device, dtype=torch.float32) if b is not None else None
            self.conv[name] = (wp, bp)

        def bn_fold(w, bn, b=None):
            s = f(bn + ".weight") / torch.sqrt(f(bn + ".running_var") + BN_EPS)
            bias = f(bn + ".bias") - f(bn + ".running_mean") * s
            if b is not None:
                bias = bias + b * s
            return w * s.view(-1, 1, 1, 1), bias

        # ---- STM encoders: torchvision ResNet-50 (conv1..layer3) with folded BN ------------------------
        for enc in ("trimap.model.Encoder_Q", "trimap.model.Encoder_M"):
            if enc.endswith("_M"):          # order must match frame_outputs' 22-channel memorize input
                w = torch.cat([f(enc + ".conv1.weight"), f(enc + ".conv1_m.weight"), f(enc + ".conv1_o.weight"),
                               f(enc + ".conv1_a.weight"), f(enc + ".conv1_h.weight")], dim=1)
                put(enc + ".stem", *bn_fold(w, enc + ".bn1"), cin_pad=self.cin_mem)
            else:
                put(enc + ".stem", *bn_fold(f(enc + ".conv1.weight"), enc + ".bn1"), cin_pad=self.cin_img)
            for lname, blocks in (("res2", 3), ("res3", 4), ("res4", 6)):
                for b in range(blocks):
                    p = f"{enc}.{lname}.{b}"
                    for c in ("1", "2", "3"):
                        put(f"{p}.conv{c}", *bn_fold(f(f"{p}.conv{c}.weight"), f"{p}.bn{c}"))
                    if b == 0:
                        put(p + ".downsample", *bn_fold(f(p + ".downsample.0.weight"), p + ".downsample.1"))
        # The two encoders have the same stages with different weights: store every res2..res4 layer as ONE pair tensor
        # (bank 0 = Encoder_Q, bank 1 = Encoder_M) for the grouped launches of Engine._tv_encoder; the per-encoder
        # entries become views of it (no copy).
        self.pair: Dict[str, tuple] = {}
        if bf:
            q, m = "trimap.model.Encoder_Q", "trimap.model.Encoder_M"
            for name in [n for n in self.conv if n.startswith(q + ".res")]:
                sfx = name[len(q):]
                (wq, bq), (wm, bm) = self.conv[q + sfx], self.conv[m + sfx]
                bp = torch.stack([bq, bm])
                if planes > 1:
                    wp = torch.stack([wq, wm], dim=1)                     # [planes][2][Cout][KH][KW][Cin]
                    self.conv[q + sfx], self.conv[m + sfx] = (wp[:, 0], bp[0]), (wp[:, 1], bp[1])
                    self.pair[sfx] = (wp.view(planes, 2 * wq.shape[1], *wq.shape[2:]), bp.view(-1))
                else:
                    wp = torch.stack([wq, wm], dim=0)                     # [2][Cout][KH][KW][Cin]
                    self.conv[q + sfx], self.conv[m + sfx] = (wp[0], bp[0]), (wp[1], bp[1])
                    self.pair[sfx] = (wp.view(2 * wq.shape[0], *wq.shape[1:]), bp.view(-1))
        for kv in ("trimap.model.KV_M_r4", "trimap.model.KV_Q_r4"):
            put(kv + ".Key", f(kv + ".Key.weight"), f(kv + ".Key.bias"))
            put(kv + ".Value", f(kv + ".Value.weight"), f(kv + ".Value.bias"))
        d = "trimap.model.Decoder"
        plain = [d + ".convFM", d + ".ResMM.conv1", d + ".ResMM.conv2", d + ".pred"]
        for rf in ("RF3", "RF2"):
            plain += [f"{d}.{rf}.convFS"] + [f"{d}.{rf}.{rb}.conv{c}" for rb in ("ResFS", "ResMM") for c in "12"]
        if fba:
            plain += ["NET.decoder.conv_up4.2", "NET.decoder.conv_up4.4", "NET.refine.pred.0", "NET.refine.pred.2",
                      "NET.refine.pred.4"]
        for name in plain:
            put(name, f(name + ".weight"), f(name + ".bias"))
        # the two 1x1 heads run in a pointwise kernel with fp32 weights (ops.head_conv_fba)
        self.head: Dict[str, tuple] = {}
        if fba:
            for name in ("NET.decoder.conv_up4.4", "NET.refine.pred.4"):
                w = f(name + ".weight")
                self.head[name] = (w.reshape(w.shape[0], w.shape[1]).contiguous().to(device),
                                   f(name + ".bias").contiguous().to(device))
        ms = lambda m, s: [float(v) for v in f(m).flatten()] + [float(v) for v in f(s).flatten()]
        self.ms_q = ms("trimap.model.Encoder_Q.mean", "trimap.model.Encoder_Q.std")
        self.ms_m = ms("trimap.model.Encoder_M.mean", "trimap.model.Encoder_M.std")
        if not fba:
            return
        put("NET.decoder.conv_up4.0", f("NET.decoder.conv_up4.0.weight"), f("NET.decoder.conv_up4.0.bias"), cin_pad=96)

        # ---- FBA: weight-standardised convs + GroupNorm affine -----------------------------------------
        def ws_put(name, cin_pad=None):
            b = f(name + ".bias") if (name + ".bias") in sd else None
            put(name, _ws(f(name + ".weight")), b, cin_pad=cin_pad)

        def gn_put(name):
            self.norm[name] = (f(name + ".weight").to(device), f(name + ".bias").to(device))

        ws_put("NET.encoder.conv1", cin_pad=16); gn_put("NET.encoder.bn1")
        for lname, blocks in (("layer1", 3), ("layer2", 4), ("layer3", 6), ("layer4", 3)):
            for b in range(blocks):
                p = f"NET.encoder.{lname}.{b}"
                for c in ("1", "2", "3"):
                    ws_put(f"{p}.conv{c}"); gn_put(f"{p}.bn{c}")
                if b == 0:
                    ws_put(p + ".downsample.0"); gn_put(p + ".downsample.1")
        for i in range(4):
            ws_put(f"NET.decoder.ppm.{i}.1"); gn_put(f"NET.decoder.ppm.{i}.2")
        for c, n in (("conv_up1.0", "conv_up1.1"), ("conv_up1.3", "conv_up1.4"), ("conv_up2.0", "conv_up2.1"),
                     ("conv_up3.0", "conv_up3.1")):
            ws_put("NET.decoder." + c); gn_put("NET.decoder." + n)
        ws_put("NET.refine.conv1.0", cin_pad=96); gn_put("NET.refine.conv1.1")
        for l in ("layer1", "layer2"):
            for c in ("1", "2"):
                ws_put(f"NET.refine.{l}.conv{c}"); gn_put(f"NET.refine.{l}.bn{c}")

        self.ms_alpha = ms("IMG_MEAN", "IMG_STD")


class FramePlan:
    """Device buffers for one padded frame size (allocated once, reused every frame)."""

    def __init__(self, H, W, dtype, device, multiple=32, arena=None):
        """``arena``: every buffer of the plan's activation dtype is carved from it (one plane stride for all split
        tensors of the plan, split.py); ``device='meta'`` + a meta arena = planning pass that only adds up sizes."""
        self.H, self.W = H, W
        self.arena, self.bank, self.stm_banks = arena, None, {}
        self.Hp, self.Wp = H + (multiple - H % multiple) % multiple, W + (multiple - W % multiple) % multiple
        self.pad_top, self.pad_left = (self.Hp - H) // 2, (self.Wp - W) // 2     # models/alpha/common.py:17-19
        self.dtype, self.device = dtype, torch.device(device)
        self.bufs: Dict[str, torch.Tensor] = {}
        # GroupNorm statistics arena: one [32][2] fp64 slot per normalised convolution, zeroed ONCE per frame
        # (one memset node instead of one in front of each of the 66 convolutions)
        self.gn_arena = torch.zeros(320, 72, dtype=torch.float64, device=device)   # [32][2] sums + barrier counter + pad
        self.gn_slots: Dict[str, int] = {}
        self.pending: Optional[int] = None          # bank slot whose memorize pass is deferred to the next frame

    def gn_slot(self, name):
        i = self.gn_slots.setdefault(name, len(self.gn_slots))
        return self.gn_arena[i]

    def buf(self, name, shape, dtype=None, zero=False):
        t = self.bufs.get(name)
        if t is None:
            dt = dtype or self.dtype
            if dt == torch.bfloat16 and self.arena is not None:
                t = self.arena.alloc(shape, zero=zero)
            else:
                t = (torch.zeros if zero else torch.empty)(shape, dtype=dt, device=self.device)
            if not zero and t.is_floating_point() and os.environ.get("OTVM_DEBUG_POISON") == "1" and not t.is_meta:
                t.fill_(float("nan"))               # dev: a kernel reading bytes nobody wrote shows up as NaN
            self.bufs[name] = t
        assert tuple(t.shape) == tuple(shape), (name, t.shape, shape)
        return t


class MemoryBank:
    """Pre-allocated key/value bank with in-place slot replacement.

    Reference policy (models/alpha/model.py:472-493): slot 0 (first frame) is kept for ever; a ``memorize``
    frame appends, any other frame overwrites the most recent slot; above ``max_memory_num`` the second-oldest
    slot is dropped.  Softmax over the memory axis is order-invariant, so instead of re-concatenating tensors
    every frame the new key/value are written straight into the slot that would be dropped.
    Layout: keys [cap*HW, De] rows (location-major), values [Do, cap*HW] channel-major (K-major operand of the
    P.V contraction).
    """

    def __init__(self, hw, cap, dtype, device, arena=None):
        self.hw, self.cap, self.arena = hw, cap, arena
        if arena is not None and dtype == torch.bfloat16:
            self.keys = arena.alloc((cap * hw, DE), zero=True)
            self.vals = arena.alloc((DO, cap * hw), zero=True)
        else:
            self.keys = torch.zeros(cap * hw, DE, dtype=dtype, device=device)
            self.vals = torch.zeros(DO, cap * hw, dtype=dtype, device=device)
        self.order = []                 # logical (oldest..newest) -> physical slot

    @property
    def T(self):
        return len(self.order)

    def reset(self):
        self.order = []

    def next_slot(self, first_frame, memorize, max_memory_num):
        """Returns (slot to write, new logical order) or (None, order) when the frame is not stored."""
        o = list(self.order)
        if max_memory_num == 0:
            return (0, [0]) if first_frame else (None, o)
        if max_memory_num == 1 or first_frame:
            return 0, [0]
        if memorize or len(o) == 1:
            if len(o) + 1 > max_memory_num:            # append then drop logical slot 1 -> reuse its storage
                s = o[1]
                return s, [o[0]] + o[2:] + [s]
            s = len(o)
            assert s < self.cap, "memory bank capacity exceeded"
            return s, o + [s]
        return o[-1], o                                # overwrite the most recent slot

    def key_slot(self, s):
        return self.keys[s * self.hw:(s + 1) * self.hw]

    def store(self, key, val):
        """overwrite the bank with fp32 tensors key [T*hw, De], val [Do, T*hw] (boundary / tests; PyTorch)"""
        n = key.shape[0]
        for dst, src in ((self.keys[:n], key), (self.vals[:, :n], val)):
            if self.arena is not None and dst.dtype == torch.bfloat16:
                self.arena.write(dst, src)
            else:
                dst.copy_(src)

    def key_tensor(self, T=None):
        """fp32 [1,1,De,T,h*w] copy of the held keys in LOGICAL order (reference layout, models/alpha/model.py:472)"""
        order = self.order if T is None else list(range(T))
        k = to_float(self.keys).view(self.cap, self.hw, DE)[order]
        return k.permute(2, 0, 1).reshape(1, 1, DE, len(order), self.hw)

    def val_tensor(self, T=None):
        order = self.order if T is None else list(range(T))
        v = to_float(self.vals).view(DO, self.cap, self.hw)[:, order]
        return v.reshape(1, 1, DO, len(order), self.hw)

    def val_slot_ptr_offset(self, s):
        return s * self.hw


class _Fork:
    """``with eng.fork("tag") as f: ...`` issues the body on a side stream forked from the current one (also inside a
    CUDA-graph capture, where it becomes a parallel branch of the graph); ``f.join()`` makes the current stream wait
    for it.  Independent chains of small-grid kernels (downsample branch of a bottleneck, the four pyramid-pooling
    branches, the skip-connection halves of the STM decoder, Key / Value projections) fill SMs the main chain
    leaves idle.  Every branch gets its own split-K workspace; buffers are keyed by layer name, so branches never
    share outputs."""

    def __init__(self, eng, tag):
        self.eng, self.tag, self.stream = eng, f"{eng._ws_tag}.{tag}", None

    def __enter__(self):
        eng = self.eng
        if eng.fork_enabled and ops.PROFILER is None and not ops.DRY:
            self.stream = eng.streams.get(self.tag)
            if self.stream is None:
                self.stream = eng.streams[self.tag] = torch.cuda.Stream(device=eng.device)
            self.stream.wait_stream(torch.cuda.current_stream())
            self.ctx = torch.cuda.stream(self.stream)
            self.ctx.__enter__()
            self.prev_tag, eng._ws_tag = eng._ws_tag, self.tag
            eng._open_forks += 1
        return self

    def __exit__(self, *exc):
        if self.stream is not None:
            self.eng._ws_tag = self.prev_tag
            self.ctx.__exit__(*exc)
        return False

    def join(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
            self.eng._open_forks -= 1


class Engine:
    def __init__(self, state_dict, dtype=torch.float32, device="cuda", bank_capacity=16, fba=True, planes=None):
        self.dtype, self.device = dtype, torch.device(device)
        self.planes = (planes or 1) if dtype == torch.bfloat16 else 0
        self.w = PackedWeights(state_dict, dtype, self.device, fba=fba, planes=max(self.planes, 1))
        self.workspaces: Dict[str, torch.Tensor] = {}      # split-K scratch per stream tag (shared by all plans)
        self.bank_capacity = bank_capacity
        self.plans: Dict[tuple, FramePlan] = {}
        self.max_plans = max(1, int(os.environ.get("OTVM_MAX_PLANS", "3")))
        self.copy_stream_enabled = os.environ.get("OTVM_COPY_STREAM", "1") != "0"
        self.copy_stream = None
        self.use_graphs = os.environ.get("OTVM_CUDA_GRAPHS", "1") != "0"
        # deferred memorize: frame t's Encoder_M / KV_M pass (STM.py:201-228) only feeds frame t+1's Memory.read, so
        # it is issued at the START of frame t+1 on a side stream, concurrently with Encoder_Q / KV_Q of t+1 (both
        # are chains of small-grid convolutions that leave most SMs idle on their own), and joined before the read
        self.defer_memorize = os.environ.get("OTVM_OVERLAP", "1") != "0"
        self.fork_enabled = os.environ.get("OTVM_FORK", "1") != "0"     # parallel graph branches inside a frame
        self.streams: Dict[str, "torch.cuda.Stream"] = {}
        self._ws_tag = "main"
        self._open_forks = 0                       # branches issued and not yet joined
        self.tv_pair = os.environ.get("OTVM_TV_PAIR", "1") != "0"   # grouped Encoder_Q + Encoder_M launches (_tv_encoder)
        self._gn_slices: Dict[tuple, int] = {}     # (conv, input shape) -> channel slices of _ws_gn_sliced (0: none)
        self.graphs: Dict[tuple, "torch.cuda.CUDAGraph"] = {}
        self.warm, self.seen = set(), set()
        self.graph_launches: Dict[tuple, int] = {}
        self.replayed_launches = 0                 # kernels executed through graph replays (bench.py gpu_launches)

    # ------------------------------------------------------------------------------------------------
    def _new_plan(self, H, W, multiple, dry_fn) -> FramePlan:
        """bf16 modes: a planning pass (meta tensors, ops.DRY) walks the frame once to add up every activation buffer,
        then ONE arena [planes][bytes] is allocated for the plan (all split tensors share its plane stride)."""
        arena = None
        if self.dtype == torch.bfloat16:
            probe = FramePlan(H, W, self.dtype, "meta", multiple, arena=SplitArena(self.planes, 0, "meta"))
            ops.DRY = True
            try:
                dry_fn(probe)
            finally:
                ops.DRY = False
            arena = SplitArena(self.planes, probe.arena.off, self.device)
        return FramePlan(H, W, self.dtype, self.device, multiple, arena=arena)

    def plan(self, H, W) -> FramePlan:
        k = (H, W)
        if k not in self.plans:
            self._evict(k)
            self.plans[k] = self._new_plan(H, W, 32, self._dry_frame)
        elif next(reversed(self.plans)) != k:
            self.plans[k] = self.plans.pop(k)              # most recently used last
        return self.plans[k]

    def _evict(self, incoming):
        """at most ``max_plans`` frame sizes stay resident (OTVM_MAX_PLANS, default 3): a plan owns every activation
        buffer, its memory bank and its CUDA graphs (GBs at 1024^2), and a dataset walks through many resolutions.
        The least recently used plan goes first; its memories go with it (the reference drops them per clip too)."""
        while len(self.plans) >= self.max_plans:
            old = next(iter(self.plans))
            pl = self.plans.pop(old)
            for g in [g for g in self.graphs if (g[0], g[1]) == (pl.H, pl.W)]:
                del self.graphs[g]
                self.graph_launches.pop(g, None)
            self.warm.discard((pl.H, pl.W))
            self.seen = {g for g in self.seen if (g[0], g[1]) != (pl.H, pl.W)}
            del pl

    def bank(self, pl: FramePlan) -> MemoryBank:
        if pl.bank is None:
            pl.bank = MemoryBank((pl.Hp // 16) * (pl.Wp // 16), self.bank_capacity, self.dtype, pl.device, pl.arena)
        return pl.bank

    def _dry_frame(self, pl: FramePlan):
        """every buffer a frame can touch: first frame (with a user trimap) and a steady-state frame with a memorize pass"""
        H, W = pl.H, pl.W
        bank = self.bank(pl)
        for n, shp in (("in_a", (H, W)), ("in_fg", (3, H, W)), ("in_bg", (3, H, W))):
            pl.buf(n, shp, torch.float32)
        for first in (True, False):
            tri = torch.empty(3, H, W, device="meta") if first else None
            self._frame_body(pl, bank, first_frame=first, slot=0, radius=1, user_tri=tri, pending=None if first else 0)

    def workspace(self, pl) -> torch.Tensor:
        """64 MB fp32 split-K scratch of the current stream tag (concurrent branches never share one)"""
        if pl.device.type == "meta":
            return torch.empty(16, dtype=torch.float32, device="meta")
        ws = self.workspaces.get(self._ws_tag)
        if ws is None:
            ws = self.workspaces[self._ws_tag] = torch.empty(16 << 20, dtype=torch.float32, device=self.device)
        return ws

    def fork(self, tag):
        return _Fork(self, tag)

    # ---- building blocks ---------------------------------------------------------------------------
    def _conv(self, pl, name, x, out_name=None, *, out=None, cout_f32=False, stride=1, pad=0, dil=1, **kw):
        w, b = self.w.conv[name]
        N, H, W, _ = x.shape
        kh, cout = w.shape[-3], w.shape[-4]
        Ho = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
        Wo = (W + 2 * pad - dil * (kh - 1) - 1) // stride + 1
        if out is None:
            out = pl.buf(out_name or name, (N, Ho, Wo, cout))
        ws = self.workspace(pl)
        return ops.conv2d(x, w, b, out, stride=stride, pad=pad, dil=dil, workspace=ws, **kw)

    # ---- STM encoders (torchvision ResNet-50 stem..layer3, BN folded) -------------------------------------
    # Encoder_Q (current frame) and Encoder_M (previous frame + its trimap / alpha / hidden maps) are the same stages with
    # different weights.  Every activation is ONE [2, H, W, C] buffer (image 0 = Q, image 1 = M): an encoder running
    # alone works on its image (``idx`` 0 / 1), and when both are due -- every steady-state frame, because frame t-1's
    # memorize pass is deferred to the start of frame t -- each layer is ONE grouped launch over both images
    # (``idx`` None, otvm_conv_params.groups = 2) instead of two small-grid launches competing on two streams.
    _TV = ("trimap.model.Encoder_Q", "trimap.model.Encoder_M")

    def _tv_buf(self, pl, key, shape, idx):
        t = pl.buf("tv." + key, (2,) + tuple(shape[1:]))
        return t if idx is None else t[idx:idx + 1]

    def _tv_conv(self, pl, idx, sfx, x, *, stride=1, pad=0, act=ACT_NONE, res=None):
        if idx is None:
            (w, b), groups = self.w.pair[sfx], 2
        else:
            (w, b), groups = self.w.conv[self._TV[idx] + sfx], 1
        N, H, W, _ = x.shape
        kh, cout = w.shape[-3], w.shape[-4] // groups
        Ho = (H + 2 * pad - (kh - 1) - 1) // stride + 1
        Wo = (W + 2 * pad - (kh - 1) - 1) // stride + 1
        out = self._tv_buf(pl, sfx, (1, Ho, Wo, cout), idx)
        return ops.conv2d(x, w, b, out, stride=stride, pad=pad, act=act, res=res, workspace=self.workspace(pl), groups=groups)

    def _tv_bottleneck(self, pl, idx, p, x, stride):
        """torchvision Bottleneck with BN folded; ReLUs and the residual add live in the conv epilogues."""
        idn, f = x, None
        if (self._TV[0] + p + ".downsample") in self.w.conv:
            with self.fork("ds") as f:
                idn = self._tv_conv(pl, idx, p + ".downsample", x, stride=stride)
        t1 = self._tv_conv(pl, idx, p + ".conv1", x, act=ACT_RELU)
        t2 = self._tv_conv(pl, idx, p + ".conv2", t1, stride=stride, pad=1, act=ACT_RELU)
        if f is not None:
            f.join()
        return self._tv_conv(pl, idx, p + ".conv3", t2, res=idn, act=ACT_RELU)

    def _tv_encoder(self, pl, idx, x_q=None, x_m=None):
        """``idx`` 0: Encoder_Q on ``x_q``; 1: Encoder_M on ``x_m``; None: both (grouped).  Returns r2, r3, r4 with the
        batch dimension of the mode ([1, ...] or [2, ...])."""
        xs = (x_q, x_m)
        N, H, W, _ = (x_q if x_q is not None else x_m).shape
        c1 = self._tv_buf(pl, "stem", (1, (H + 1) // 2, (W + 1) // 2, 64), None)
        f = None
        for i in ((0, 1) if idx is None else (idx,)):
            w, b = self.w.conv[self._TV[i] + ".stem"]
            if idx is None and i == 1:                     # the second stem beside the first
                with self.fork("stem") as f:
                    ops.conv2d(xs[i], w, b, c1[i:i + 1], stride=2, pad=3, act=ACT_RELU, workspace=self.workspace(pl))
            else:
                ops.conv2d(xs[i], w, b, c1[i:i + 1], stride=2, pad=3, act=ACT_RELU, workspace=self.workspace(pl))
        if f is not None:
            f.join()
        c1 = c1 if idx is None else c1[idx:idx + 1]
        N, H, W, C = c1.shape
        x = ops.maxpool3x3s2(c1, self._tv_buf(pl, "pool", (1, (H + 1) // 2, (W + 1) // 2, C), idx))
        feats = []
        for lname, blocks, stride in (("res2", 3, 1), ("res3", 4, 2), ("res4", 6, 2)):
            for b in range(blocks):
                x = self._tv_bottleneck(pl, idx, f".{lname}.{b}", x, stride if b == 0 else 1)
            feats.append(x)
        return feats       # r2, r3, r4

    def _gn(self, pl, name, x, *, act, res=None, out=None, stats=None):
        g, b = self.w.norm[name]
        if stats is None:
            stats = pl.buf("gn_stats", (64,), torch.float64)
            ops.gn_stats(x, stats)
        return ops.gn_apply(x, stats, g, b, out if out is not None else x, act=act, res=res)

    def _ws_gn(self, pl, conv, norm, x, *, act, res=None, out=None, stride=1, pad=0, dil=1, raw_name=None):
        """WS-conv -> GroupNorm(32) -> activation (+residual).  GN statistics are accumulated by the conv
        epilogue; the normalise/affine/activation pass runs in place unless ``out`` is given."""
        stats = pl.gn_slot(conv)                    # zeroed by _frame_body (one memset per frame)
        g, b = self.w.norm[norm]
        w = self.w.conv[conv][0]
        N, H, W, _ = x.shape
        kh, cout = w.shape[-3], w.shape[-4]
        Ho = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
        Wo = (W + 2 * pad - dil * (kh - 1) - 1) // stride + 1
        raw = pl.buf(raw_name or conv, (N, Ho, Wo, cout))
        dst = out if out is not None else raw
        # one kernel when the grid is a single co-resident wave: statistics, grid barrier, normalise from TMEM.
        # Only while nothing else is in flight on another stream: two grid-synchronising kernels sharing the SMs
        # could each hold slots the other needs to become fully resident.
        alone = self._ws_tag == "main" and self._open_forks == 0
        if alone and cout >= 128 and not ops.DRY and self._ws_gn_sliced(pl, conv, g, b, x, dst, raw, act, res, stride, pad, dil):
            return dst
        if alone and self._conv(pl, conv, x, out=dst, stride=stride, pad=pad, dil=dil, gn_stats=stats,
                                gn_stats_zeroed=True, gn_fuse=(g, b, 1e-5), gn_raw_out=raw, act=act, res=res):
            return dst
        if not alone:
            self._conv(pl, conv, x, out=raw, stride=stride, pad=pad, dil=dil, gn_stats=stats, gn_stats_zeroed=True)
        return self._gn(pl, norm, raw, act=act, res=res, out=out, stats=stats)

    def _ws_gn_sliced(self, pl, conv, g, b, x, dst, raw, act, res, stride, pad, dil) -> bool:
        """A wide normalised layer whose grid is more than one co-resident wave (FBA layer4: 512 / 1024 -> 2048 at 1/8
        resolution, 512 CTAs) as 2 or 4 channel slices, each a launch of its own with its own statistics slot and grid
        barrier: GroupNorm groups are 64 consecutive channels, so slices hold whole groups and every slice takes the
        in-kernel GroupNorm path (no raw tensor round trip, no gn_apply pass).  False: not applicable, nothing launched."""
        w, bias = self.w.conv[conv]
        if w.dim() != 5 or os.environ.get("OTVM_GN_SLICES", "1") == "0":
            return False
        cout = w.shape[1]
        ws = self.workspace(pl)

        def run(a, c, i, query=False):
            return ops.conv2d(x, w[:, a:c], bias[a:c] if bias is not None else None, dst[..., a:c], stride=stride, pad=pad,
                              dil=dil, workspace=ws, gn_stats=pl.gn_slot(f"{conv}#slice{i}"), gn_stats_zeroed=True,
                              gn_fuse=(g[a:c], b[a:c], 1e-5), gn_raw_out=raw[..., a:c], act=act,
                              res=res[..., a:c] if res is not None else None, gn_group_ch=cout // 32, query_fuse=query)

        key = (conv, tuple(x.shape))
        ns = self._gn_slices.get(key)
        if ns is None:
            ns = 0
            if not ops.conv2d(x, w, bias, dst, stride=stride, pad=pad, dil=dil, workspace=ws, gn_stats=pl.gn_slot(conv),
                              gn_stats_zeroed=True, gn_fuse=(g, b, 1e-5), gn_raw_out=raw, act=act, res=res, query_fuse=True):
                for n in (2, 4, 8):
                    if cout % n == 0 and (cout // n) % max(cout // 32, 8) == 0 and run(0, cout // n, 0, query=True):
                        ns = n
                        break
            self._gn_slices[key] = ns
        if ns < 2:
            return False
        step = cout // ns
        for i in range(ns):
            ok = run(i * step, (i + 1) * step, i)
            assert ok, "a channel slice lost the fused GroupNorm path it was planned with"
        return True

    def _gn_bottleneck(self, pl, p, x, stride, dil, out=None):
        # (no parallel downsample branch here: the GroupNorm-fused kernels synchronise their whole grid and must never
        # share the GPU with another grid-synchronising kernel, see _ws_gn)
        t = self._ws_gn(pl, p + ".conv1", p + ".bn1", x, act=ACT_RELU)
        t = self._ws_gn(pl, p + ".conv2", p + ".bn2", t, act=ACT_RELU, stride=stride, pad=dil, dil=dil)
        if (p + ".downsample.0") in self.w.conv:
            idn = self._ws_gn(pl, p + ".downsample.0", p + ".downsample.1", x, act=ACT_NONE, stride=stride)
        else:
            idn = x
        return self._ws_gn(pl, p + ".conv3", p + ".bn3", t, act=ACT_RELU, res=idn, out=out)

    def _resblock(self, pl, p, x, x_relu, *, out_name, final_relu=False, out_relu_name=None):
        """STM ResBlock (STM.py:23-30): x + conv2(relu(conv1(relu(x)))).  ``x_relu`` is relu(x), written by
        the producer of ``x``; returns (y, relu(y) or None)."""
        r = self._conv(pl, p + ".conv1", x_relu, pad=1, act=ACT_RELU)
        N, H, W, C = x.shape
        y = pl.buf(out_name, (N, H, W, C))
        yr = pl.buf(out_relu_name, (N, H, W, C)) if out_relu_name else None
        self._conv(pl, p + ".conv2", r, out=y, pad=1, res=x, act=ACT_RELU if final_relu else ACT_NONE, out_relu=yr)
        return y, yr

    # ---- STM ---------------------------------------------------------------------------------------
    def segment(self, pl: FramePlan, bank: MemoryBank, join=None, feats=None):
        """Propagated trimap logits [Hp*Wp][4] fp32 for the current frame (imgn must be ready).  ``feats``: r2, r3, r4 of
        Encoder_Q when the caller ran the grouped encoder pass."""
        imgn = pl.bufs["imgn"]                      # [1,Hp,Wp,cin_img]: 3 normalised channels + zeros
        r2, r3, r4 = feats if feats is not None else self._tv_encoder(pl, 0, x_q=imgn)
        N, h, w, _ = r4.shape
        m4in = pl.buf("m4in", (1, h, w, 2 * DO))
        # parallel branches: the skip-connection halves of the decoder (STM.py:113-114 only need r3 / r2) and the
        # Value projection run beside Key -> Memory.read -> convFM -> ResMM
        skips = {}
        for rf, f in (("RF3", r3), ("RF2", r2)):
            with self.fork(rf) as fk:
                skips[rf] = (self._stm_skip(pl, rf, f), fk)
        with self.fork("val") as fv:
            self._conv(pl, "trimap.model.KV_Q_r4.Value", r4, out=m4in[..., DO:], pad=1)
        qk = self._conv(pl, "trimap.model.KV_Q_r4.Key", r4, "q_key", pad=1)
        M = bank.T * bank.hw
        ws_bytes = ops.memory_read_workspace(bank.cap * bank.hw, h * w, DE, DO, self.dtype)
        ws = pl.buf(f"read_ws.{bank.cap}", (ws_bytes // 4,), torch.float32)
        if join is not None:
            torch.cuda.current_stream().wait_stream(join)      # deferred memorize of the previous frame has landed
        # strict mode (three planes): the fused tcgen05 read multiplies two planes per operand (2^-16), and with the
        # reference-like ill-conditioned attention of random weights (logits of several hundred) that is 6e-5 on the read
        # and 1.4e-3 on alpha after the softmax / FBA amplification -- the fp32 FFMA read keeps strict mode inside 1e-3
        ops.memory_read(bank.keys, bank.vals, bank.vals.shape[1], qk, m4in[..., :DO], M, ws, force_simt=self.planes == 3)
        fv.join()
        return self._stm_decoder(pl, m4in, skips)

    def _stm_skip(self, pl, rf, f):
        """Refine.forward's skip half, STM.py:113-114: ResFS(convFS(f))"""
        d = "trimap.model.Decoder"
        N, H, W, _ = f.shape
        s0 = pl.buf(f"dec.{rf}.s0", (1, H, W, 256)); s0r = pl.buf(f"dec.{rf}.s0r", (1, H, W, 256))
        self._conv(pl, f"{d}.{rf}.convFS", f, out=s0, pad=1, out_relu=s0r)
        s, _ = self._resblock(pl, f"{d}.{rf}.ResFS", s0, s0r, out_name=f"dec.{rf}.s")
        return s

    def _stm_decoder(self, pl, m4in, skips):
        d = "trimap.model.Decoder"
        N, h, w, _ = m4in.shape
        x0 = pl.buf("dec.x0", (1, h, w, 256)); x0r = pl.buf("dec.x0r", (1, h, w, 256))
        self._conv(pl, d + ".convFM", m4in, out=x0, pad=1, out_relu=x0r)
        m, _ = self._resblock(pl, d + ".ResMM", x0, x0r, out_name="dec.m4")
        for rf in ("RF3", "RF2"):
            s, fk = skips[rf]
            fk.join()
            N, H, W, _ = s.shape
            mmr = pl.buf(f"dec.{rf}.mmr", (1, H, W, 256))
            mm = ops.upsample(m, pl.buf(f"dec.{rf}.mm", (1, H, W, 256)), add=s, out_relu=mmr)     # STM.py:115
            # the last block's output is only read through F.relu (STM.py:134) -> fold it into the epilogue
            m, _ = self._resblock(pl, f"{d}.{rf}.ResMM", mm, mmr, out_name=f"dec.{rf}.m", final_relu=rf == "RF2")
        N, H, W, _ = m.shape
        p2 = pl.buf("dec.p2", (1, H, W, 4), torch.float32, zero=True)
        self._conv(pl, d + ".pred", m, out=p2[..., :3], pad=1)                        # relu(m2) applied above
        logits = pl.buf("seg_logits", (1, pl.Hp, pl.Wp, 4), torch.float32, zero=True)
        ops.upsample(p2[..., :3], logits[..., :3])                                    # STM.py:136
        return logits

    def memorize(self, pl: FramePlan, bank: MemoryBank, slot: int, r4=None):
        """Encode (frame, trimap, alpha, hidden) and write key/value straight into bank slot ``slot``.  ``r4``: Encoder_M's
        last feature map when the caller ran the grouped encoder pass."""
        mem_in = pl.bufs["mem_in"]                  # [1,Hp,Wp,cin_mem]: 22 channels + zeros
        if r4 is None:
            _, _, r4 = self._tv_encoder(pl, 1, x_m=mem_in)
        N, h, w, _ = r4.shape
        kdst = bank.key_slot(slot).view(1, h, w, DE)
        vdst = bank.vals[:, slot * bank.hw:]
        with self.fork("val") as fv:
            self._conv(pl, "trimap.model.KV_M_r4.Value", r4, out=vdst, pad=1, out_strides=(1, bank.vals.shape[1]))
        self._conv(pl, "trimap.model.KV_M_r4.Key", r4, out=kdst, pad=1)
        fv.join()

    def _stage_inputs(self, pl: FramePlan, a, fg, bg):
        """bring a / fg / bg into the plan's static input buffers (what the CUDA graphs read).

        Host tensors (eval.py hands the loader's CPU tensors over) go through a COPY STREAM into one of two staging
        sets: the H2D of frame t+1 (7 fp32 planes, 7.3 MB at 512^2) then overlaps the kernels of frame t instead of
        sitting in front of them on the compute stream, and only a device-to-device copy (a few us) stays in stream
        order.  A staging set is reused once the D2D copy that last read it has been issued two frames earlier."""
        H, W = pl.H, pl.W
        f32 = torch.float32
        dst = [pl.buf("in_a", (H, W), f32), pl.buf("in_fg", (3, H, W), f32), pl.buf("in_bg", (3, H, W), f32)]
        src = [a, fg, bg]
        if all(t.device.type == "cuda" for t in src) or not self.copy_stream_enabled:
            for d, t in zip(dst, src):
                d.copy_(t, non_blocking=True)
            return
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(device=self.device)
        k = pl.stage_idx = (getattr(pl, "stage_idx", 1) + 1) % 2
        stage = [pl.buf(f"stage{k}.{n}", tuple(d.shape), f32) for n, d in zip("afb", dst)]
        cur = torch.cuda.current_stream()
        if not hasattr(pl, "stage_done"):
            pl.stage_done = [None, None]
        with torch.cuda.stream(self.copy_stream):
            if pl.stage_done[k] is not None:
                self.copy_stream.wait_event(pl.stage_done[k])       # the D2D copy of frame t-2 has read this set
            for d, t in zip(stage, src):
                d.copy_(t, non_blocking=True)                       # pinned source: truly asynchronous
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        cur.wait_event(ready)
        for d, t in zip(dst, stage):
            d.copy_(t, non_blocking=True)
        pl.stage_done[k] = torch.cuda.Event()
        pl.stage_done[k].record(cur)

    def flush(self, pl: FramePlan):
        """Run a deferred memorize pass now (anything that looks at the bank outside ``frame`` calls this)."""
        if pl.pending is not None:
            self.memorize(pl, self.bank(pl), pl.pending)
            pl.pending = None

    # ---- stand-alone STM entry points (FullModel_eval.forward(memorize=) / (segment=)) -----------------
    def _stm_plan(self, H, W) -> FramePlan:
        """buffers for STM.memorize / STM.segment called on their own: pad to 16 (STM.py:204,241), not 32"""
        k = ("stm", H, W)
        if k not in self.plans:
            self._evict(k)
            self.plans[k] = self._new_plan(H, W, 16, self._dry_stm)
        return self.plans[k]

    def _stm_bank(self, pl: FramePlan) -> MemoryBank:
        """bank of a stand-alone STM plan (the caller passes the memories in; capacity = bank_capacity frames)"""
        if pl.bank is None:
            pl.bank = MemoryBank((pl.Hp // 16) * (pl.Wp // 16), self.bank_capacity, self.dtype, pl.device, pl.arena)
        return pl.bank

    def _dry_stm(self, pl: FramePlan):
        pl.buf("mem_in", (1, pl.Hp, pl.Wp, self.w.cin_mem), zero=True)
        pl.buf("imgn", (1, pl.Hp, pl.Wp, self.w.cin_img), zero=True)
        bank = self._stm_bank(pl)
        self.memorize(pl, bank, 0)
        self.segment(pl, bank)

    def stm_memorize(self, frame, masks):
        """STM.memorize (STM.py:201-228): frame [1,3,H,W] RGB in [0,1], masks [1,20,H,W] = trimap 3 | alpha 1 | hidden 16
        (models/trimap/model.py:231).  Returns key [1,1,128,1,h,w], value [1,1,512,1,h,w] fp32."""
        H, W = frame.shape[-2:]
        pl = self._stm_plan(H, W)
        mean = torch.tensor(self.w.ms_m[:3], device=self.device).view(1, 3, 1, 1)
        std = torch.tensor(self.w.ms_m[3:], device=self.device).view(1, 3, 1, 1)
        lw, lh = (pl.Wp - W) // 2, (pl.Hp - H) // 2
        pad = (lw, pl.Wp - W - lw, lh, pl.Hp - H - lh)
        f = (torch.nn.functional.pad(frame.float(), pad) - mean) / std           # padding happens BEFORE Encoder_M normalises
        m = torch.nn.functional.pad(masks.float(), pad)
        x = torch.cat([f, m[:, 1:]], dim=1).contiguous()                         # rgb 3 | unknown | fg | alpha | hidden 16
        mem_in = pl.buf("mem_in", (1, pl.Hp, pl.Wp, self.w.cin_mem), zero=True)
        ops.nchw_to_nhwc(x, mem_in[..., :22])
        bank = self._stm_bank(pl)
        self.memorize(pl, bank, 0)
        h, w = pl.Hp // 16, pl.Wp // 16
        key = bank.key_tensor(1).view(1, 1, DE, 1, h, w).contiguous()
        val = bank.val_tensor(1).view(1, 1, DO, 1, h, w).contiguous()
        return key, val

    def stm_segment(self, frame, keys, values):
        """STM.segment (STM.py:239-257): frame [1,3,H,W] RGB in [0,1], keys [1,1,128,T,h,w], values [1,1,512,T,h,w].
        Returns the trimap logits [1,3,H,W] fp32 (padding cropped)."""
        H, W = frame.shape[-2:]
        pl = self._stm_plan(H, W)
        mean = torch.tensor(self.w.ms_q[:3], device=self.device).view(1, 3, 1, 1)
        std = torch.tensor(self.w.ms_q[3:], device=self.device).view(1, 3, 1, 1)
        lw, lh = (pl.Wp - W) // 2, (pl.Hp - H) // 2
        f = (torch.nn.functional.pad(frame.float(), (lw, pl.Wp - W - lw, lh, pl.Hp - H - lh)) - mean) / std
        imgn = pl.buf("imgn", (1, pl.Hp, pl.Wp, self.w.cin_img), zero=True)
        ops.nchw_to_nhwc(f.contiguous(), imgn[..., :3])
        T = keys.shape[3]
        hw = (pl.Hp // 16) * (pl.Wp // 16)
        assert keys.shape[-2:] == (pl.Hp // 16, pl.Wp // 16), "memory and query frame sizes differ"
        bank = self._stm_bank(pl)
        if T > bank.cap:
            raise ValueError(f"{T} memory frames exceed the bank capacity {bank.cap} (set OTVM_BANK_CAPACITY)")
        # (values arrive [Do, T, h*w] contiguous; the bank's channel rows are cap*hw long, frames 0..T-1 first)
        bank.store(keys[0, 0].permute(1, 2, 3, 0).reshape(T * hw, DE).to(self.device),
                   values[0, 0].reshape(DO, T * hw).to(self.device))
        bank.order = list(range(T))
        logits = self.segment(pl, bank)
        out = logits[0, lh:lh + H, lw:lw + W, :3].permute(2, 0, 1).unsqueeze(0).contiguous()
        return out

    # ---- FBA ---------------------------------------------------------------------------------------
    def matting(self, pl: FramePlan):
        """x11 / cat4[64:72] / extras must be ready.  Produces raw heads, fused outputs and ``hid``."""
        Hp, Wp = pl.Hp, pl.Wp
        e = "NET.encoder"
        x11 = pl.bufs["x11"]
        cat3 = pl.buf("cat3", (1, Hp // 2, Wp // 2, 320))       # up(conv_up2) | conv_out[-5]
        cat2 = pl.buf("cat2", (1, Hp // 4, Wp // 4, 512))       # up(conv_up1) | conv_out[-4]
        cat1 = pl.buf("cat1", (1, Hp // 8, Wp // 8, 3072))      # conv5 | ppm x4
        c1 = self._ws_gn(pl, e + ".conv1", e + ".bn1", x11, act=ACT_RELU, stride=2, pad=3, out=cat3[..., 256:],
                         raw_name="enc.stem.raw")
        x = ops.maxpool3x3s2(c1, pl.buf("enc.pool", (1, Hp // 4, Wp // 4, 64)))
        cfg = (("layer1", 3, 1, 1, 1, cat2[..., 256:]), ("layer2", 4, 2, 1, 1, None),
               ("layer3", 6, 1, 1, 2, None), ("layer4", 3, 1, 2, 4, cat1[..., :2048]))
        for lname, blocks, stride, d0, dd, last_out in cfg:
            for b in range(blocks):
                x = self._gn_bottleneck(pl, f"{e}.{lname}.{b}", x, stride if b == 0 else 1, d0 if b == 0 else dd,
                                        out=last_out if b == blocks - 1 else None)
        # pyramid pooling (FBA/models.py:357-362)
        conv5 = x
        h8, w8 = conv5.shape[1:3]
        pooled = pl.buf("ppm.pooled", (50, 2048))
        ops.ppm_pool(conv5, pooled, pl.buf("ppm.rows", (h8 * 12 * 2048,), torch.float32))
        off = 0
        forks = []
        for i, s in enumerate((1, 2, 3, 6)):                # four independent branches
            # (a 1x1 convolution does not care how its pixels are arranged: the 3x3 and 6x6 grids go in as one row of 9 / 36
            # pixels, wide enough for the tcgen05 kernel's 8-pixel tile rows; 1 and 4 cells stay on the small-M FFMA kernel)
            shape = (1, 1, s * s, 2048) if s * s >= 8 else (1, s, s, 2048)
            cells = pooled[off:off + s * s].view(shape); off += s * s
            with self.fork(f"ppm{i}") as fk:
                y = self._ws_gn(pl, f"NET.decoder.ppm.{i}.1", f"NET.decoder.ppm.{i}.2", cells, act=ACT_LEAKY)
                ops.upsample(y.view(1, s, s, 256), cat1[..., 2048 + 256 * i: 2304 + 256 * i])
            forks.append(fk)
        for fk in forks:
            fk.join()
        dn = "NET.decoder"
        x = self._ws_gn(pl, dn + ".conv_up1.0", dn + ".conv_up1.1", cat1, act=ACT_LEAKY, pad=1)
        x = self._ws_gn(pl, dn + ".conv_up1.3", dn + ".conv_up1.4", x, act=ACT_LEAKY, pad=1)
        ops.upsample(x, cat2[..., :256])
        x = self._ws_gn(pl, dn + ".conv_up2.0", dn + ".conv_up2.1", cat2, act=ACT_LEAKY, pad=1)
        ops.upsample(x, cat3[..., :256])
        x = self._ws_gn(pl, dn + ".conv_up3.0", dn + ".conv_up3.1", cat3, act=ACT_LEAKY, pad=1)
        cat4 = pl.bufs["cat4"]                                  # up(conv_up3) | imgn | img | two_chan | alpha | 0
        ops.upsample(x, cat4[..., :64])
        x = self._conv(pl, dn + ".conv_up4.0", cat4, pad=1, act=ACT_LEAKY)
        hid_d = self._conv(pl, dn + ".conv_up4.2", x, pad=1, act=ACT_LEAKY)
        raw7 = pl.buf("raw7", (1, Hp, Wp, 8), torch.float32, zero=True)
        P = Hp * Wp
        extras = pl.bufs["extras"]
        out7 = pl.buf("out7", (P, 8), torch.float32)
        ops.head_conv_fba(hid_d, *self.w.head[dn + ".conv_up4.4"], raw7, extras, P, out7, cat4[..., 72:73], cat4.stride(2))
        # refinement (FBA/models.py:417-435)
        r = "NET.refine"
        x = self._ws_gn(pl, r + ".conv1.0", r + ".conv1.1", cat4, act=ACT_LEAKY, pad=1)
        for l in ("layer1", "layer2"):
            t = self._ws_gn(pl, f"{r}.{l}.conv1", f"{r}.{l}.bn1", x, act=ACT_RELU, pad=1)
            x = self._ws_gn(pl, f"{r}.{l}.conv2", f"{r}.{l}.bn2", t, act=ACT_RELU, pad=1, res=x)
        x = self._conv(pl, r + ".pred.0", x, pad=1, act=ACT_LEAKY)
        hid = self._conv(pl, r + ".pred.2", x, "hid", pad=1, act=ACT_LEAKY)
        raw10 = pl.buf("raw10", (1, Hp, Wp, 12), torch.float32, zero=True)
        fused = pl.buf("fused", (P, 8), torch.float32)
        ops.head_conv_fba(hid, *self.w.head[r + ".pred.4"], raw10, extras, P, fused)
        return dict(raw7=raw7, out7=out7, raw10=raw10, fused=fused, hid=hid, conv5=conv5)

    # ---- one frame (EvalModel.forward with tri=None, tri_gt=None) ------------------------------------
    def _frame_body(self, pl, bank, *, first_frame, slot, radius, user_tri=None, pending=None):
        """Issue every kernel of one frame on the current stream (inputs already in the plan's static buffers).
        ``pending``: bank slot of the previous frame's deferred memorize pass, run first on the side stream."""
        H, W, Hp, Wp, P = pl.H, pl.W, pl.Hp, pl.Wp, pl.Hp * pl.Wp
        f32 = torch.float32
        join = None
        # both encoders are due (steady state): one grouped pass after the preprocessing below instead of two streams
        pair = pending is not None and not first_frame and self.tv_pair and bool(self.w.pair)
        if pending is not None and not pair:
            if ops.PROFILER is None and not ops.DRY:
                if "memorize" not in self.streams:
                    self.streams["memorize"] = torch.cuda.Stream(device=self.device)
                join = self.streams["memorize"]
                join.wait_stream(torch.cuda.current_stream())          # fork (also inside a graph capture)
                with torch.cuda.stream(join):
                    self._ws_tag = "memorize"
                    try:
                        self.memorize(pl, bank, pending)
                    finally:
                        self._ws_tag = "main"
            else:
                self.memorize(pl, bank, pending)                         # instrumented pass: one stream, no overlap
        ops.zero_(pl.gn_arena)
        a, fg, bg = pl.bufs["in_a"], pl.bufs["in_fg"], pl.bufs["in_bg"]
        img = pl.buf("img", (P, 4), f32)
        scaled = pl.buf("scaled_img", (3, H, W), f32)
        tri3 = pl.buf("tri3", (P, 4), f32)
        imgn = pl.buf("imgn", (1, Hp, Wp, self.w.cin_img), zero=True)[..., :4]
        ops.preprocess(a, fg, bg, H, W, Hp, Wp, pl.pad_top, pl.pad_left, radius, self.w.ms_q, img, scaled, tri3, imgn,
                       pl.buf("pre_scratch", (2 * H * W,), torch.uint8))
        x11 = pl.buf("x11", (1, Hp, Wp, 16))
        cat4 = pl.buf("cat4", (1, Hp, Wp, 96), zero=True)
        extras = pl.buf("extras", (P, 8), f32)
        d2 = pl.buf("d2", (2, P), torch.int32)
        enc_args = (img, Hp, Wp, self.w.ms_alpha, x11, cat4[..., 64:72], extras, d2,
                    pl.buf("edt_scratch", (2, P), torch.int32), pl.buf("seeds", (2, P), torch.uint8))
        if first_frame:
            tri_first = tri3
            if user_tri is not None:         # user / GT trimap for frame 0 (:395-401); padding = bg (:409-410)
                tri_first = pl.buf("tri_user", (Hp, Wp, 4), f32)
                tri_first.zero_(); tri_first[..., 0] = 1.0
                tri_first[pl.pad_top:pl.pad_top + H, pl.pad_left:pl.pad_left + W, :3] = user_tri.permute(1, 2, 0)
            ops.trimap_encode(tri_first, 4, False, *enc_args)                   # preds_trimap = tri_ (:429)
        else:
            feats = None
            if pair:
                r2, r3, r4 = self._tv_encoder(pl, None, x_q=pl.bufs["imgn"], x_m=pl.bufs["mem_in"])
                feats = (r2[0:1], r3[0:1], r4[0:1])
                if ops.PROFILER is None and not ops.DRY:     # KV_M projections beside the query path, joined before the read
                    if "memorize" not in self.streams:
                        self.streams["memorize"] = torch.cuda.Stream(device=self.device)
                    join = self.streams["memorize"]
                    join.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(join):
                        self._ws_tag = "memorize"
                        try:
                            self.memorize(pl, bank, pending, r4=r4[1:2])
                        finally:
                            self._ws_tag = "main"
                else:
                    self.memorize(pl, bank, pending, r4=r4[1:2])
            logits = self.segment(pl, bank, join, feats=feats)
            join = None
            ops.trimap_encode(logits, 4, True, *enc_args)                       # softmax + make_trimap (:440-442)
        net = self.matting(pl)
        alpha = pl.buf("alpha_out", (H, W), f32)
        trimap = pl.buf("trimap_out", (3, H, W), f32)
        mem_in = pl.buf("mem_in", (1, Hp, Wp, self.w.cin_mem), zero=True)[..., :24] if slot is not None else None
        ops.frame_outputs(net["raw10"], 12, net["fused"], net["hid"], extras, Hp, Wp, H, W, pl.pad_top, pl.pad_left,
                          self.w.ms_m, mem_in, alpha, trimap)
        if join is not None:                          # (first frames never carry a pending pass; kept for safety)
            torch.cuda.current_stream().wait_stream(join)
        if slot is not None and not self.defer_memorize:
            self.memorize(pl, bank, slot)

    def frame(self, a, fg, bg, *, first_frame, last_frame, memorize, max_memory_num, radius, user_tri=None):
        """a [H,W] fp32, fg/bg [3,H,W] fp32 BGR 0..255 (device or pinned host).  Returns the plan's output
        buffers: scaled_img [3,H,W], trimap [3,H,W], tri3 [Hp*Wp,4] one-hot (padded), alpha [H,W].

        Steady-state frames are replayed from a CUDA graph keyed by (bank size, destination slot): the frame is
        ~180 kernel launches whose host-side issue cost would otherwise bound the frame rate."""
        H, W = a.shape[-2:]
        pl = self.plan(H, W)
        bank = self.bank(pl)
        self._stage_inputs(pl, a, fg, bg)
        if first_frame:
            bank.reset()
            pl.pending = None                        # the reference drops the old memories too (:425-429)
        slot, order = (None, bank.order)
        if not last_frame:
            slot, order = bank.next_slot(first_frame, memorize, max_memory_num)
        pending, pl.pending = pl.pending, None
        key = (H, W, bank.T, pending, slot if not self.defer_memorize else slot is not None, radius)
        body = lambda: self._frame_body(pl, bank, first_frame=first_frame, slot=slot, radius=radius, user_tri=user_tri,
                                        pending=pending)
        steady = slot is not None and bank.T >= max(2, max_memory_num)      # bank full: keys repeat from now on
        if not self.use_graphs or first_frame or ops.PROFILER is not None:
            body()
        elif key in self.graphs:
            self.graphs[key].replay()
            self.replayed_launches += self.graph_launches[key]
        elif (H, W) not in self.warm or not (steady or key in self.seen):
            body()                                   # eager: allocates buffers / sets func attributes; growing bank
            self.warm.add((H, W)); self.seen.add(key)
        else:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(g):
                body()
            self.graphs[key] = g
            self.graph_launches[key] = ops.launch_count() - n0      # kernels recorded in this graph
            g.replay()
            self.replayed_launches += self.graph_launches[key]
        if slot is not None:
            bank.order = order
            if self.defer_memorize:
                pl.pending = slot
        b = pl.bufs
        return b["scaled_img"], b["trimap_out"], b["tri3"], b["alpha_out"]
