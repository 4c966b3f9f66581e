"""Parity of ba_bwd_delta / ba_bwd_chunk / cast / accumulate (through the C-ABI)
against the CPU oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from burst_attn.chunk_ops import NativeOps  # noqa: E402
from gpu_util import TOL  # noqa: E402
from oracle import attention_oracle as orc  # noqa: E402


def _mk(B, S, H, dtype, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(B, S, H, 128, device="cuda", generator=g) * scale).to(dtype)


def run_bwd(q, k, v, do, causal=False, off=0):
    """dense fwd via oracle for (o, lse) consistency, then native delta + bwd chunk."""
    ops = NativeOps()
    scale = 128 ** -0.5
    mode = "none" if not causal else ("causal" if off == 0 else "causal_strict")
    o_ref, lse_ref = orc.chunk_forward(q.cpu(), k.cpu(), v.cpu(), None, None, scale, mode)
    lse_safe = torch.where(torch.isinf(lse_ref), torch.zeros_like(lse_ref), lse_ref)
    o_dev = o_ref.to(q.dtype).cuda()
    lse_dev = lse_ref.float().cuda().contiguous()
    B, Sq, H, _ = q.shape
    delta = torch.empty(B, H, Sq, device="cuda", dtype=torch.float32)
    ops.delta(o_dev, do, delta, 1)
    dq = torch.zeros(q.shape, device="cuda", dtype=torch.float32)
    dk = torch.zeros(k.shape, device="cuda", dtype=torch.float32)
    dv = torch.zeros(v.shape, device="cuda", dtype=torch.float32)
    ops.bwd_chunk(do, q, k, v, delta, lse_dev, dq, dk, dv, scale, causal, off, 1)
    torch.cuda.synchronize()
    delta_ref = orc.compute_delta(o_dev.cpu(), do.cpu())
    rdq, rdk, rdv = orc.chunk_backward(do.cpu(), q.cpu(), k.cpu(), v.cpu(), delta_ref, lse_safe, scale, mode)
    if mode == "causal_strict":  # rows that saw nothing: lse=-inf -> p must be 0 (oracle used lse=0 there)
        dead = torch.isinf(lse_ref)
        assert dead.any()
        rdq, rdk, rdv = orc.chunk_backward(
            do.cpu(), q.cpu(), k.cpu(), v.cpu(), delta_ref,
            torch.where(dead, torch.full_like(lse_ref, 1e30), lse_ref), scale, mode)
    return (delta, dq, dk, dv), (delta_ref, rdq, rdk, rdv)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,Sq,Sk,H", [(1, 128, 128, 1), (2, 256, 384, 2), (1, 200, 333, 2), (1, 512, 128, 3)])
def test_bwd_noncausal(dtype, B, Sq, Sk, H):
    q, do = _mk(B, Sq, H, dtype, 1), _mk(B, Sq, H, dtype, 2)
    k, v = _mk(B, Sk, H, dtype, 3), _mk(B, Sk, H, dtype, 4)
    got, ref = run_bwd(q, k, v, do)
    torch.testing.assert_close(got[0].double().cpu(), ref[0], rtol=1e-2, atol=2e-2)
    for g, r in zip(got[1:], ref[1:]):
        torch.testing.assert_close(g.double().cpu(), r, **TOL[dtype])


@pytest.mark.parametrize("S", [128, 384, 300])
@pytest.mark.parametrize("off", [0, -1])
def test_bwd_causal(S, off):
    dtype = torch.bfloat16
    q, do, k, v = (_mk(1, S, 2, dtype, s) for s in (5, 6, 7, 8))
    got, ref = run_bwd(q, k, v, do, causal=True, off=off)
    for g, r in zip(got[1:], ref[1:]):
        torch.testing.assert_close(g.double().cpu(), r, **TOL[dtype])


def test_bwd_accumulates_into_existing_values():
    dtype = torch.bfloat16
    q, do, k, v = (_mk(1, 256, 2, dtype, s) for s in (9, 10, 11, 12))
    ops = NativeOps()
    scale = 128 ** -0.5
    o, lse = orc.dense_attention(q.cpu(), k.cpu(), v.cpu(), scale)
    delta = orc.compute_delta(o, do.cpu()).float().cuda()
    lse = lse.float().cuda()
    acc = [torch.full(t.shape, 0.5, device="cuda", dtype=torch.float32) for t in (q, k, v)]
    ops.bwd_chunk(do, q, k, v, delta, lse, acc[0], acc[1], acc[2], scale, False, 0, 1)
    ops.bwd_chunk(do, q, k, v, delta, lse, acc[0], acc[1], acc[2], scale, False, 0, 1)
    torch.cuda.synchronize()
    rdq, rdk, rdv = orc.chunk_backward(do.cpu(), q.cpu(), k.cpu(), v.cpu(), delta.cpu(), lse.cpu(), scale)
    for g, r in zip(acc, (rdq, rdk, rdv)):
        torch.testing.assert_close(g.double().cpu(), 0.5 + 2 * r, rtol=2e-2, atol=4e-2)


def test_cast_and_accumulate():
    ops = NativeOps()
    src = torch.randn(2, 100, 3, 128, device="cuda")
    dst = torch.empty(2, 100, 3, 128, device="cuda", dtype=torch.bfloat16)
    ops.cast(src, dst, 1)
    assert torch.equal(dst, src.to(torch.bfloat16))
    acc = torch.randn(2, 100, 3, 128, device="cuda")
    exp = acc + src
    ops.accumulate(src, acc, 1)
    assert torch.equal(acc, exp)
    # strided half views
    half = torch.zeros(2, 50, 3, 128, device="cuda", dtype=torch.float16)
    ops.cast(src.narrow(1, 50, 50), half, 1)
    assert torch.equal(half, src[:, 50:].to(torch.float16))


@pytest.mark.parametrize("causal", [False, True])
def test_bwd_deterministic_is_bitwise_reproducible(causal):
    """deterministic=True (reference: forwarded to flash-attn, burst_attn_interface.py:318-320):
    dQ is reduced in key-block order -> identical bits run to run, same values as the oracle."""
    dtype = torch.bfloat16
    B, S, H = 1, 2048, 4
    q, do, k, v = (_mk(B, S, H, dtype, s) for s in (21, 22, 23, 24))
    ops = NativeOps()
    scale = 128 ** -0.5
    o, lse = orc.dense_attention(q.cpu(), k.cpu(), v.cpu(), scale, causal)
    delta = orc.compute_delta(o, do.cpu()).float().cuda()
    lse = lse.float().cuda()
    outs = []
    for _ in range(3):
        acc = [torch.zeros(t.shape, device="cuda", dtype=torch.float32) for t in (q, k, v)]
        ops.bwd_chunk(do, q, k, v, delta, lse, acc[0], acc[1], acc[2], scale, causal, 0, 1, deterministic=True)
        torch.cuda.synchronize()
        outs.append(acc)
    for a in outs[1:]:
        for x, y in zip(a, outs[0]):
            assert torch.equal(x, y)
    _, _, rdq, rdk, rdv = orc.dense_attention_bwd(q.cpu(), k.cpu(), v.cpu(), do.cpu(), scale, causal)
    for g, r in zip(outs[0], (rdq, rdk, rdv)):
        torch.testing.assert_close(g.double().cpu(), r, **TOL[dtype])


def test_bwd_large_property_dv_linear_in_do():
    """Size-independent property at a size the oracle cannot finish: dV is linear in dO (delta, lse fixed
    per call), and dK/dQ/dV of one head match torch SDPA's autograd on the GPU."""
    dtype = torch.bfloat16
    B, S, H = 1, 4096, 2
    q, k, v, do = (_mk(B, S, H, dtype, s) for s in (31, 32, 33, 34))
    qq, kk, vv = (t.float().permute(0, 2, 1, 3).clone().requires_grad_() for t in (q, k, v))
    o_ref = torch.nn.functional.scaled_dot_product_attention(qq, kk, vv)
    g = torch.autograd.grad(o_ref, (qq, kk, vv), do.float().permute(0, 2, 1, 3))
    ops = NativeOps()
    scale = 128 ** -0.5
    out = torch.empty_like(q)
    lse = torch.empty(B, H, S, device="cuda")
    ops.fwd_chunk(q, k, v, None, lse, out, scale, False, 0, True, True, 1)
    delta = torch.empty(B, H, S, device="cuda")
    ops.delta(out, do, delta, 1)
    acc = [torch.zeros(t.shape, device="cuda") for t in (q, k, v)]
    ops.bwd_chunk(do, q, k, v, delta, lse, acc[0], acc[1], acc[2], scale, False, 0, 1)
    torch.cuda.synchronize()
    for got, ref in zip(acc, g):
        torch.testing.assert_close(got, ref.permute(0, 2, 1, 3), rtol=2e-2, atol=2e-2)
