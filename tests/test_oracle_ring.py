"""The single-process ring-schedule simulators must reproduce dense attention
and its autograd gradients in all three shard layouts."""
import pytest
import torch

from oracle import attention_oracle as orc

CASES = [("none", "contiguous", False), ("zigzag", "zigzag", True), ("striped", "striped", True)]


@pytest.mark.parametrize("mode,layout,causal", CASES)
@pytest.mark.parametrize("W", [1, 2, 4])
def test_ring_fwd_bwd_equals_dense_autograd(mode, layout, causal, W):
    torch.manual_seed(1)
    B, S, H, D = 2, 16 * W, 3, 8
    q, k, v, do = (torch.randn(B, S, H, D, dtype=torch.float64) for _ in range(4))
    scale = D ** -0.5
    qr, kr, vr = (t.clone().requires_grad_() for t in (q, k, v))
    o_ref, lse_ref = orc.dense_attention(qr, kr, vr, scale, causal)
    dq_ref, dk_ref, dv_ref = torch.autograd.grad(o_ref, (qr, kr, vr), do)
    # analytic dense bwd agrees with autograd
    _, _, dq_a, dk_a, dv_a = orc.dense_attention_bwd(q, k, v, do, scale, causal)
    for a, b in ((dq_a, dq_ref), (dk_a, dk_ref), (dv_a, dv_ref)):
        torch.testing.assert_close(a, b, rtol=1e-10, atol=1e-10)

    sh = lambda t, dim=1: [orc.shard(t, r, W, layout, dim) for r in range(W)]
    qs, ks, vs, dos = sh(q), sh(k), sh(v), sh(do)
    os_, lses = orc.ring_forward(qs, ks, vs, scale, mode)
    torch.testing.assert_close(orc.unshard(os_, layout), o_ref.detach(), rtol=1e-10, atol=1e-10)
    torch.testing.assert_close(orc.unshard(lses, layout, dim=2), lse_ref.detach(), rtol=1e-10, atol=1e-10)
    dqs, dks, dvs = orc.ring_backward(qs, ks, vs, os_, lses, dos, scale, mode)
    torch.testing.assert_close(orc.unshard(dqs, layout), dq_ref, rtol=1e-9, atol=1e-9)
    torch.testing.assert_close(orc.unshard(dks, layout), dk_ref, rtol=1e-9, atol=1e-9)
    torch.testing.assert_close(orc.unshard(dvs, layout), dv_ref, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("layout", ["contiguous", "zigzag", "striped"])
def test_shard_roundtrip(layout):
    t = torch.arange(2 * 24 * 3).reshape(2, 24, 3).double()
    parts = [orc.shard(t, r, 4, layout) for r in range(4)]
    torch.testing.assert_close(orc.unshard(parts, layout), t)


def test_flops_formula():
    # benchmarks/benchmark.py:17-20; SURVEY 8(d): C2 fwd 7.037e13
    assert abs(orc.attention_flops(1, 65536, 32, 128) - 7.037e13) / 7.037e13 < 1e-3
    assert orc.attention_flops(1, 1024, 2, 64, causal=True, mode="fwd_bwd") == 3.5 * 4 * 1024 * 1024 * 2 * 64 / 2
