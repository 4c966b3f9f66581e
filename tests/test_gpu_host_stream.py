"""Host-resident operands (burst_attn/host_stream.py): burst_attn_func called with pinned CPU tensors on one rank
streams Q/K/V/dO up and O/dQ/dK/dV down under the L2-blocked sub-launches.  Same results as the device call."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from burst_attn import burst_attn_func  # noqa: E402
from gpu_util import TOL  # noqa: E402
from oracle import attention_oracle as orc  # noqa: E402


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("S,blk", [(1024, 256), (1536 + 200, 512), (384, 32768)])
def test_host_resident_call_matches_oracle_and_device_call(monkeypatch, causal, S, blk):
    monkeypatch.setenv("BA_L2_BLOCK", str(blk))
    torch.manual_seed(0)
    dtype = torch.bfloat16
    q, k, v, do = (torch.randn(1, S, 3, 128).to(dtype).pin_memory() for _ in range(4))
    qq, kk, vv = (t.clone().pin_memory().requires_grad_() for t in (q, k, v))
    o = burst_attn_func(qq, kk, vv, None, "cuda", causal)
    dq, dk, dv = torch.autograd.grad(o, (qq, kk, vv), do)
    torch.cuda.synchronize()
    assert all(t.device.type == "cpu" and t.dtype == dtype for t in (o, dq, dk, dv))
    o_ref, _, dq_ref, dk_ref, dv_ref = orc.dense_attention_bwd(q, k, v, do, None, causal)
    for got, ref in ((o, o_ref), (dq, dq_ref), (dk, dk_ref), (dv, dv_ref)):
        torch.testing.assert_close(got.double(), ref, **TOL[dtype])
    # and the plain device call on the same inputs
    qd, kd, vd = (t.cuda().requires_grad_() for t in (q, k, v))
    od = burst_attn_func(qd, kd, vd, None, "cuda", causal)
    gd = torch.autograd.grad(od, (qd, kd, vd), do.cuda())
    for got, ref in zip((o, dq, dk, dv), (od, *gd)):
        torch.testing.assert_close(got.float(), ref.float().cpu(), rtol=2e-2, atol=2e-2)


def test_cpu_tensors_without_pinning_fail_loudly():
    q = torch.randn(1, 128, 1, 128, dtype=torch.bfloat16)
    with pytest.raises(AssertionError):
        burst_attn_func(q, q, q, None, "cuda", False)
