"""Training-side pieces of SURVEY.md section 8(f) rank 2: the fused ``Memory.read`` under autograd, and the STM read block
it sits in, trainable under ``DistributedDataParallel`` (NCCL gradient all-reduce).

* :class:`FusedMemoryRead` / :func:`memory_read` -- ``Memory.forward`` (reference ``models/trimap/STM.py:144-163``) as a
  ``torch.autograd.Function``: the forward is the fused kernel (``otvm_memory_read``, which also saves the per-query
  log-sum-exp), the backward ``otvm_memory_read_backward`` recomputes 32 x 32 tiles of the affinity, so nothing of size
  THW x HW is stored for the backward either (the reference keeps the full softmax matrix alive: 1 GB at 1024^2 / T=16).
* :class:`STMReadBlock` -- the slice of the trimap-propagation network around the read with the reference's parameter
  names and shapes: ``KV_M_r4`` / ``KV_Q_r4`` (``KeyValue``, STM.py:166-174), ``Memory`` and the decoder's ``convFM`` +
  ``pred`` (STM.py:120-137) on 1/16-resolution features.  :func:`train_step` runs one optimisation step on it the way
  ``train.py:349-375`` does (bf16 autocast, backward, optimiser step); under DDP the gradients of its parameters
  are all-reduced over NCCL (14.2 M parameters, 56.7 MB of fp32 gradients per step).

What is NOT here: the full stage-4 model (the two ResNet-50 encoders, the FBA network and the losses of
``models/alpha/model.py:189-312``).  Their backward passes are PyTorch / cuDNN territory and are out of this repo's hot
path; this module provides the one operator of that step that this repo owns, with its gradient, and shows it training
under DDP.  PyTorch (autograd, ``nn.Conv2d``, DDP, the optimiser) is used as the training framework here, like the
reference uses it.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import ops

DE, DO = 128, 512


class FusedMemoryRead(torch.autograd.Function):
    """mem[B, Do, h, w] = sum_m softmax_m(K.Q / sqrt(De)) V  for keys [B,De,T,h,w], values [B,Do,T,h,w], query [B,De,h,w]"""

    @staticmethod
    def forward(ctx, m_in, m_out, q_in):
        assert m_in.is_cuda, "the fused read is a CUDA kernel (no CPU fallback)"
        B, De, T, h, w = m_in.shape
        Do = m_out.shape[1]
        M, HW = T * h * w, h * w
        out_dtype = q_in.dtype
        # kernel layouts: keys [M, De] location-major rows, values [Do, M] channel-major, query / output NHWC (fp32)
        keys = m_in.detach().float().permute(0, 2, 3, 4, 1).reshape(B, M, De).contiguous()
        vals = m_out.detach().float().reshape(B, Do, M).contiguous()
        q = q_in.detach().float().permute(0, 2, 3, 1).contiguous()
        out = torch.empty(B, h, w, Do, device=q.device, dtype=torch.float32)
        lse = torch.empty(B, HW, device=q.device, dtype=torch.float32)
        ws = torch.empty(ops.memory_read_workspace(M, HW, De, Do) // 4 + 1, device=q.device, dtype=torch.float32)
        for b in range(B):
            ops.memory_read(keys[b], vals[b], M, q[b:b + 1], out[b:b + 1], M, ws, lse=lse[b])
        ctx.save_for_backward(keys, vals, q, out, lse)
        ctx.shape = (B, De, Do, T, h, w, m_in.dtype, m_out.dtype, q_in.dtype)
        return out.permute(0, 3, 1, 2).to(out_dtype)

    @staticmethod
    def backward(ctx, grad):
        keys, vals, q, out, lse = ctx.saved_tensors
        B, De, Do, T, h, w, dt_k, dt_v, dt_q = ctx.shape
        M, HW = T * h * w, h * w
        dout = grad.detach().float().permute(0, 2, 3, 1).contiguous()
        dk = torch.empty(B, M, De, device=q.device, dtype=torch.float32)
        dv = torch.empty(B, Do, M, device=q.device, dtype=torch.float32)
        dq = torch.empty(B, HW, De, device=q.device, dtype=torch.float32)
        for b in range(B):
            ops.memory_read_backward(keys[b], vals[b], M, q[b:b + 1], out[b:b + 1], dout[b:b + 1], lse[b], dk[b], dv[b],
                                     dq[b], M)
        d_m_in = dk.view(B, T, h, w, De).permute(0, 4, 1, 2, 3).to(dt_k)
        d_m_out = dv.view(B, Do, T, h, w).to(dt_v)
        d_q_in = dq.view(B, h, w, De).permute(0, 3, 1, 2).to(dt_q)
        return d_m_in, d_m_out, d_q_in


def memory_read(m_in, m_out, q_in, q_out):
    """``Memory.forward`` (STM.py:144-163): ``cat([read(m_in, m_out, q_in), q_out], dim=1)`` with the fused kernels"""
    return torch.cat([FusedMemoryRead.apply(m_in, m_out, q_in), q_out], dim=1)


class KeyValue(nn.Module):
    """STM.py:166-174"""

    def __init__(self, indim=1024, keydim=DE, valdim=DO):
        super().__init__()
        self.Key = nn.Conv2d(indim, keydim, kernel_size=3, padding=1, stride=1)
        self.Value = nn.Conv2d(indim, valdim, kernel_size=3, padding=1, stride=1)

    def forward(self, x):
        return self.Key(x), self.Value(x)


class STMReadBlock(nn.Module):
    """r4 features of T memory frames + r4 features of the query frame -> 3-class trimap logits at 1/16 resolution.
    Parameter names follow the reference (``KV_M_r4.Key.weight``, ``KV_Q_r4.Value.bias``, ``Decoder.convFM.weight`` ...),
    so slices of a reference ``state_dict`` load into it.  ``read_fn`` is the read operator (default: the fused kernels;
    tests inject the composite PyTorch formulation to check gradients and to run the DDP plumbing on CPU)."""

    def __init__(self, read_fn: Optional[Callable] = None, mdim=256):
        super().__init__()
        self.KV_M_r4 = KeyValue()
        self.KV_Q_r4 = KeyValue()
        self.Decoder = nn.Module()
        self.Decoder.convFM = nn.Conv2d(2 * DO, mdim, kernel_size=3, padding=1, stride=1)
        self.Decoder.pred = nn.Conv2d(mdim, 3, kernel_size=3, padding=1, stride=1)
        self.read_fn = read_fn or memory_read

    def forward(self, r4_mem, r4_query):
        """r4_mem [B, T, 1024, h, w], r4_query [B, 1024, h, w] -> logits [B, 3, h, w]"""
        B, T = r4_mem.shape[:2]
        k, v = self.KV_M_r4(r4_mem.flatten(0, 1))
        k = k.view(B, T, *k.shape[1:]).transpose(1, 2)           # [B, De, T, h, w]
        v = v.view(B, T, *v.shape[1:]).transpose(1, 2)
        qk, qv = self.KV_Q_r4(r4_query)
        m4 = self.read_fn(k, v, qk, qv)                          # [B, 1024, h, w]
        return self.Decoder.pred(F.relu(self.Decoder.convFM(m4)))


def synthetic_batch(batch, T, h, w, seed, device):
    """features / labels of the shapes one stage-4 sample produces at 1/16 resolution (320x320 crops -> 20x20, config.py:27)"""
    g = torch.Generator().manual_seed(seed)
    r4m = torch.randn(batch, T, 1024, h, w, generator=g).to(device) * 0.5
    r4q = torch.randn(batch, 1024, h, w, generator=g).to(device) * 0.5
    lab = torch.randint(0, 3, (batch, h, w), generator=g).to(device)
    return r4m, r4q, lab


def train_step(model, opt, batch, autocast_dtype=torch.bfloat16):
    """one optimisation step in the shape of train.py:349-375: autocast forward, loss, backward (DDP all-reduces the
    gradients inside), optimiser step.  Returns the loss (a device tensor: no host sync here)."""
    r4m, r4q, lab = batch
    opt.zero_grad(set_to_none=True)
    with torch.autocast(device_type=r4q.device.type, dtype=autocast_dtype, enabled=autocast_dtype is not None):
        logits = model(r4m, r4q)
    loss = F.cross_entropy(logits.float(), lab)
    loss.backward()
    opt.step()
    return loss.detach()


def allreduce_bytes(model) -> int:
    """bytes of gradients DDP all-reduces per step (fp32 gradients of every trainable parameter)"""
    return sum(p.numel() * 4 for p in model.parameters() if p.requires_grad)
