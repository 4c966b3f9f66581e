"""Helpers shared by the -m gpu parity tests (all call through the C-ABI)."""
import ctypes

import torch

from burst_attn import native as nat
from burst_attn.chunk_ops import NativeOps
from oracle import attention_oracle as orc

# bf16/fp16 parity tolerances against the fp64 oracle.  fp16: the reference's own
# test/checker.py:6-10 (rtol=1e-3, atol=1e-2).  bf16 is never tested by the
# reference; we state rtol=1.6e-2, atol=2e-2 (8 mantissa bits) as the hard cap.
TOL = {torch.float16: dict(rtol=1e-3, atol=1e-2), torch.bfloat16: dict(rtol=1.6e-2, atol=2e-2)}


def selftest(mode, a, b, out_dtype=torch.float32):
    if mode == 2:
        out = torch.zeros(8192, device=a.device, dtype=torch.int16)
    else:
        out = torch.zeros(a.shape[0], 128, device=a.device, dtype=torch.float32)  # modes 4/5: a is [256,128]
    rc = nat.selftest_lib().ba_selftest(mode, a.data_ptr(), b.data_ptr(), out.data_ptr(), nat.dtype_code(a.dtype),
                               nat.stream_ptr(a.device))
    nat.check_selftest(rc, "ba_selftest")
    torch.cuda.synchronize()
    return out


def fwd_chunks(q, k_chunks, v_chunks, scale, causal=False, offsets=None, seq_dim=1):
    """Run ba_fwd_chunk over a list of K/V chunks with carried state; returns (out, lse)."""
    ops = NativeOps()
    B, S, H = q.shape[0], q.shape[seq_dim], q.shape[3 - seq_dim]
    n = len(k_chunks)
    out = torch.empty_like(q)
    lse = torch.empty(B, H, S, device=q.device, dtype=torch.float32)
    o_acc = torch.empty(q.shape, device=q.device, dtype=torch.float32) if n > 1 else None
    for c in range(n):
        off = 0 if offsets is None else offsets[c]
        ops.fwd_chunk(q, k_chunks[c], v_chunks[c], o_acc, lse, out, scale, causal, off, c == 0, c == n - 1, seq_dim)
    torch.cuda.synchronize()
    return out, lse


def err_stats(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    d = (got - ref).abs()
    return dict(max_abs=float(d.max()), mean_abs=float(d.mean()), ref_max=float(ref.abs().max()),
                nan=int(torch.isnan(got).sum()), argmax=tuple(int(x) for x in torch.unravel_index(d.argmax(), d.shape)))
