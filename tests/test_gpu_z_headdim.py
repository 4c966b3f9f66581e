"""BASELINE.json configs[0] (C1: the reference's own CPU-runnable case, B=1, S=4096, H=8, D=64) through the
public API on one GPU.  head_dim 64 runs on the 128-wide tile by zero padding (burst_attn_interface.py
``_pad_head_dim``); results must match dense attention on the unpadded tensors."""
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
