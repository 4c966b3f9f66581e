"""BASELINE.json configs[0] (C1: the reference's own CPU-runnable case, B=1, S=4096, H=8, D=64) through the
public API on one GPU.  head_dim 64 has its own tile (kernels templated on the head dim: one SW128 box per operand
tile, 4 k-steps for QK^T / S^T / dP^T, N = 64 for PV / dV / dK / dQ); other head dims <= 128 are zero-padded up to
the next tile width (burst_attn_interface.py ``_pad_head_dim``).  Results must match dense attention."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from burst_attn import burst_attn_func, burst_attn_func_striped  # noqa: E402
from gpu_util import TOL  # noqa: E402
from oracle import attention_oracle as orc  # noqa: E402


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_c1_head_dim_64(causal, dtype):
    torch.manual_seed(0)
    b, s, n, d = 1, 4096, 8, 64
    q, k, v, do = (torch.randn(b, s, n, d, device="cuda", dtype=dtype) for _ in range(4))
    qq, kk, vv = (t.clone().requires_grad_() for t in (q, k, v))
    o = burst_attn_func(qq, kk, vv, None, "cuda", causal)
    assert o.shape == q.shape and o.dtype == dtype and o.is_contiguous()
    dq, dk, dv = torch.autograd.grad(o, (qq, kk, vv), do)
    assert dq.shape == q.shape and dk.shape == k.shape and dv.shape == v.shape
    o_ref, _, dq_ref, dk_ref, dv_ref = orc.dense_attention_bwd(q.cpu(), k.cpu(), v.cpu(), do.cpu(), None, causal)
    torch.testing.assert_close(o.double().cpu(), o_ref, **TOL[dtype])
    for g, r in ((dv, dv_ref), (dk, dk_ref), (dq, dq_ref)):
        torch.testing.assert_close(g.double().cpu(), r, **TOL[torch.bfloat16] if dtype == torch.bfloat16 else
                                   dict(rtol=1e-3, atol=1e-2))


def test_normal_layout_head_dim_64_and_striped_32():
    torch.manual_seed(1)
    q, k, v, do = (torch.randn(1, 8, 512, 64, device="cuda", dtype=torch.bfloat16) for _ in range(4))
    qq, kk, vv = (t.clone().requires_grad_() for t in (q, k, v))
    o = burst_attn_func(qq, kk, vv, None, None, False)  # C1's own layout [B, H, S, D]
    dq, dk, dv = torch.autograd.grad(o, (qq, kk, vv), do)
    p = lambda t: t.permute(0, 2, 1, 3).cpu()  # noqa: E731
    o_ref, _, dq_ref, dk_ref, dv_ref = orc.dense_attention_bwd(p(q), p(k), p(v), p(do))
    torch.testing.assert_close(p(o).double(), o_ref, **TOL[torch.bfloat16])
    for g, r in ((dv, dv_ref), (dk, dk_ref), (dq, dq_ref)):
        torch.testing.assert_close(p(g).double(), r, **TOL[torch.bfloat16])
    q, k, v, do = (torch.randn(2, 300, 4, 32, device="cuda", dtype=torch.float16) for _ in range(4))
    qq, kk, vv = (t.clone().requires_grad_() for t in (q, k, v))
    o = burst_attn_func_striped(qq, kk, vv, None, "cuda", True)
    dq, dk, dv = torch.autograd.grad(o, (qq, kk, vv), do)
    o_ref, _, dq_ref, dk_ref, dv_ref = orc.dense_attention_bwd(q.cpu(), k.cpu(), v.cpu(), do.cpu(), None, True)
    torch.testing.assert_close(o.double().cpu(), o_ref, **TOL[torch.float16])
    for g, r in ((dv, dv_ref), (dk, dk_ref), (dq, dq_ref)):
        torch.testing.assert_close(g.double().cpu(), r, rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("D", [64, 128])
def test_native_tile_is_used_for_64_and_128(D):
    """No padding copy for the two native tile widths: the kernels see the user's tensors (same data_ptr)."""
    from burst_attn import chunk_ops
    from burst_attn.burst_attn_interface import _pad_head_dim
    q = torch.randn(1, 256, 2, D, device="cuda", dtype=torch.bfloat16)
    (qp,), d = _pad_head_dim(chunk_ops.get_ops(), [q])
    assert qp.data_ptr() == q.data_ptr() and d == D
    (q96,), d = _pad_head_dim(chunk_ops.get_ops(), [q[..., :40].contiguous()])
    assert q96.shape[-1] == 64 and d == 40


@pytest.mark.parametrize("causal", [False, True])
def test_head_dim_64_kernels_directly_with_carried_state_and_offsets(causal):
    """ba_fwd_chunk / ba_bwd_chunk at D = 64: chained K/V chunks (carried state), ragged sizes, a causal offset."""
    from burst_attn.chunk_ops import NativeOps
    dtype = torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(3)
    mk = lambda S: torch.randn(2, S, 3, 64, device="cuda", generator=g).to(dtype)  # noqa: E731
    Sq, Sc = 300, 200
    q, do = mk(Sq), mk(Sq)
    ks, vs = [mk(Sc), mk(Sc)], [mk(Sc), mk(Sc)]
    scale = 64 ** -0.5
    ops = NativeOps()
    out = torch.empty_like(q)
    lse = torch.empty(2, 3, Sq, device="cuda", dtype=torch.float32)
    o_acc = torch.empty(q.shape, device="cuda", dtype=torch.float32)
    offs = [100, -100] if causal else [0, 0]
    o_ref = l_ref = None
    for c in range(2):
        ops.fwd_chunk(q, ks[c], vs[c], o_acc, lse, out, scale, causal, offs[c], c == 0, c == 1, 1)
        o_ref, l_ref = orc.chunk_forward(q.cpu(), ks[c].cpu(), vs[c].cpu(), o_ref, l_ref, scale,
                                         ("causal_offset", offs[c]) if causal else "none")
    torch.cuda.synchronize()
    torch.testing.assert_close(out.double().cpu(), o_ref, **TOL[dtype])
    torch.testing.assert_close(lse.double().cpu(), l_ref, rtol=1e-3, atol=2e-3)
    delta = torch.empty_like(lse)
    ops.delta(out, do, delta, 1)
    for c in range(2):
        acc = [torch.zeros(t.shape, device="cuda", dtype=torch.float32) for t in (q, ks[c], vs[c])]
        ops.bwd_chunk(do, q, ks[c], vs[c], delta, lse, acc[0], acc[1], acc[2], scale, causal, offs[c], 1)
        torch.cuda.synchronize()
        ref = orc.chunk_backward(do.cpu(), q.cpu(), ks[c].cpu(), vs[c].cpu(), orc.compute_delta(out.cpu(), do.cpu()),
                                 l_ref, scale, ("causal_offset", offs[c]) if causal else "none")
        for a, r in zip(acc, ref):
            torch.testing.assert_close(a.double().cpu(), r, **TOL[dtype])
