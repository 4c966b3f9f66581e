"""Single-GPU attention entry points with the reference's names and signatures
(burst_attn/flash_triton.py:1013-1160: ``flash_attn_func``, ``flash_attn_kvpacked_func``,
``flash_attn_qkvpacked_func``; layout [batch, seqlen, nheads, headdim], ``causal`` bottom-right aligned like the
Triton kernel's ``seqlen_k - seqlen_q`` offset).  The reference keeps a vanilla Triton FlashAttention copy there
that its ring op never calls; here the three wrappers run the same sm_100a tile kernels as the ring (one local
"round", no communication), reading Q / K / V straight out of the packed tensor through strided views (TMA takes
the strides; nothing is unpacked or copied on the way in).

``bias`` (the Triton copy's additive attention bias, lao.py:102-105,155-173): the per-KEY form -- shape
``(batch | 1, nheads, 1, seqlen_k)``, the reference's "vector" bias, e.g. an ALiBi row or a key-padding mask of
-inf -- runs in the tile kernels (forward: one extra K = 16 step on the tensor core per score tile; backward: a
per-thread scalar in the exponent's FMA).  The "matrix" form ``(.., seqlen_q, seqlen_k)`` is not supported (at the
sequence lengths this path is built for it does not fit memory) and raises.  As in the reference no gradient
flows into the bias.
"""
from __future__ import annotations

import math

import torch

from .burst_attn_interface import _bwd_round, _fwd_round, _fwd_round_needs_state, _pad_head_dim, _unpad
from .chunk_ops import get_ops

__all__ = ["flash_attn_func", "flash_attn_kvpacked_func", "flash_attn_qkvpacked_func"]


def _key_bias(bias, q, k):
    """(batch | 1, nheads, 1, seqlen_k) -> fp32 [B|1 (stride 0), H, Sk] view for the chunk operators."""
    if bias is None:
        return None
    B, H, Sk = q.shape[0], q.shape[2], k.shape[1]
    if bias.dim() != 4 or bias.shape[2] != 1 or bias.shape[3] != Sk or bias.shape[1] != H or bias.shape[0] not in (1, B):
        raise NotImplementedError(f"only a per-key bias of shape (batch | 1, {H}, 1, {Sk}) is supported by the sm_100a "
                                  f"tile kernels, got {tuple(bias.shape)}")
    b3 = bias.detach().to(torch.float32).reshape(bias.shape[0], H, Sk).contiguous()
    return b3.expand(B, H, Sk)


def _local_forward(q, k, v, causal, softmax_scale, bias=None):
    ops = get_ops()
    scale = softmax_scale or 1.0 / math.sqrt(q.shape[-1])
    (qp, kp, vp), D = _pad_head_dim(ops, [q, k, v])
    B, Sq, H = qp.shape[0], qp.shape[1], qp.shape[2]
    out = torch.empty(qp.shape, dtype=qp.dtype, device=qp.device)
    lse = torch.empty((B, H, Sq), dtype=torch.float32, device=qp.device)
    o_acc = torch.empty(qp.shape, dtype=torch.float32, device=qp.device) if _fwd_round_needs_state(kp, 1) else None
    _fwd_round(ops, qp, kp, vp, o_acc, lse, out, scale, causal, kp.shape[1] - Sq, True, True, 1, bias)
    return out, lse, scale, (qp, kp, vp), D


def _local_backward(do, qp, kp, vp, out, lse, causal, scale, bias=None, deterministic=False):
    ops = get_ops()
    (g,), _ = _pad_head_dim(ops, [do])
    g, out = g.contiguous(), out.contiguous()
    B, Sq, H = qp.shape[0], qp.shape[1], qp.shape[2]
    delta = torch.empty((B, H, Sq), dtype=torch.float32, device=qp.device)
    ops.delta(out, g, delta, 1)
    f32 = dict(dtype=torch.float32, device=qp.device)
    dq, dk, dv = torch.zeros(qp.shape, **f32), torch.zeros(kp.shape, **f32), torch.zeros(vp.shape, **f32)
    _bwd_round(ops, g, qp, kp, vp, delta, lse, dq, dk, dv, scale, causal, kp.shape[1] - Sq, 1, deterministic, bias)
    return dq, dk, dv


def _cast(src, like, D):
    dst = torch.empty(src.shape, dtype=like.dtype, device=src.device)
    get_ops().cast(src, dst, 1)
    return _unpad(dst, D)


def _check(bias, *ts):
    for t in ts:
        assert t.stride(-1) == 1, "the head_dim axis must be contiguous"


class FlashAttnFunc(torch.autograd.Function):
    """q: (batch, seqlen_q, nheads, headdim); k, v: (batch, seqlen_k, nheads, headdim)  (reference :1122-1168)."""

    @staticmethod
    def forward(ctx, q, k, v, bias=None, causal=False, softmax_scale=None):
        _check(bias, q, k, v)
        ctx.bias = _key_bias(bias, q, k)
        out, lse, ctx.softmax_scale, saved, ctx.head_dim = _local_forward(q, k, v, causal, softmax_scale, ctx.bias)
        ctx.save_for_backward(*saved, out, lse)
        ctx.causal = causal
        return _unpad(out, ctx.head_dim)

    @staticmethod
    def backward(ctx, do):
        qp, kp, vp, out, lse = ctx.saved_tensors
        dq, dk, dv = _local_backward(do, qp, kp, vp, out, lse, ctx.causal, ctx.softmax_scale, ctx.bias)
        return _cast(dq, qp, ctx.head_dim), _cast(dk, kp, ctx.head_dim), _cast(dv, vp, ctx.head_dim), None, None, None


class FlashAttnKVPackedFunc(torch.autograd.Function):
    """q: (batch, seqlen_q, nheads, headdim); kv: (batch, seqlen_k, 2, nheads, headdim)  (reference :1073-1119)."""

    @staticmethod
    def forward(ctx, q, kv, bias=None, causal=False, softmax_scale=None):
        _check(bias, q, kv)
        ctx.bias = _key_bias(bias, q, kv[:, :, 0])
        out, lse, ctx.softmax_scale, saved, ctx.head_dim = _local_forward(q, kv[:, :, 0], kv[:, :, 1], causal,
                                                                          softmax_scale, ctx.bias)
        ctx.save_for_backward(*saved, out, lse)
        ctx.causal = causal
        return _unpad(out, ctx.head_dim)

    @staticmethod
    def backward(ctx, do):
        qp, kp, vp, out, lse = ctx.saved_tensors
        dq, dk, dv = _local_backward(do, qp, kp, vp, out, lse, ctx.causal, ctx.softmax_scale, ctx.bias)
        dkv = torch.stack([_cast(dk, kp, ctx.head_dim), _cast(dv, vp, ctx.head_dim)], dim=2)
        return _cast(dq, qp, ctx.head_dim), dkv, None, None, None


class FlashAttnQKVPackedFunc(torch.autograd.Function):
    """qkv: (batch, seqlen, 3, nheads, headdim)  (reference :1021-1070)."""

    @staticmethod
    def forward(ctx, qkv, bias=None, causal=False, softmax_scale=None):
        _check(bias, qkv)
        ctx.bias = _key_bias(bias, qkv[:, :, 0], qkv[:, :, 1])
        out, lse, ctx.softmax_scale, saved, ctx.head_dim = _local_forward(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], causal,
                                                                          softmax_scale, ctx.bias)
        ctx.save_for_backward(*saved, out, lse)
        ctx.causal = causal
        return _unpad(out, ctx.head_dim)

    @staticmethod
    def backward(ctx, do):
        qp, kp, vp, out, lse = ctx.saved_tensors
        dq, dk, dv = _local_backward(do, qp, kp, vp, out, lse, ctx.causal, ctx.softmax_scale, ctx.bias)
        dqkv = torch.stack([_cast(t, qp, ctx.head_dim) for t in (dq, dk, dv)], dim=2)
        return dqkv, None, None, None


flash_attn_func = FlashAttnFunc.apply
flash_attn_kvpacked_func = FlashAttnKVPackedFunc.apply
flash_attn_qkvpacked_func = FlashAttnQKVPackedFunc.apply
