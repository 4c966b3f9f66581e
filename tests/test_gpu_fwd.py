"""Parity of ba_fwd_chunk (through the C-ABI) against the CPU oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import TOL, fwd_chunks  # noqa: E402
from oracle import attention_oracle as orc  # noqa: E402


def _mk(B, S, H, D, dtype, seed=0, seq_dim=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    shape = (B, S, H, D) if seq_dim == 1 else (B, H, S, D)
    return torch.randn(shape, device="cuda", generator=g).to(dtype)


def _to_bshd(t, seq_dim):
    return t if seq_dim == 1 else t.permute(0, 2, 1, 3)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,Sq,Sk,H", [(1, 256, 256, 2), (2, 128, 384, 3), (1, 200, 333, 2), (1, 1, 77, 1),
                                       (1, 640, 512, 1)])
def test_single_chunk_noncausal(dtype, B, Sq, Sk, H):
    q, k, v = _mk(B, Sq, H, 128, dtype, 1), _mk(B, Sk, H, 128, dtype, 2), _mk(B, Sk, H, 128, dtype, 3)
    out, lse = fwd_chunks(q, [k], [v], 128 ** -0.5)
    o_ref, lse_ref = orc.dense_attention(q.cpu(), k.cpu(), v.cpu())
    torch.testing.assert_close(out.double().cpu(), o_ref, **TOL[dtype])
    torch.testing.assert_close(lse.double().cpu(), lse_ref, rtol=1e-3, atol=2e-3)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("S", [128, 256, 512, 300, 1000])
def test_single_chunk_causal(dtype, S):
    q, k, v = (_mk(2, S, 2, 128, dtype, s) for s in (4, 5, 6))
    out, lse = fwd_chunks(q, [k], [v], 128 ** -0.5, causal=True)
    o_ref, lse_ref = orc.dense_attention(q.cpu(), k.cpu(), v.cpu(), causal=True)
    torch.testing.assert_close(out.double().cpu(), o_ref, **TOL[dtype])
    torch.testing.assert_close(lse.double().cpu(), lse_ref, rtol=1e-3, atol=2e-3)


@pytest.mark.parametrize("nchunks", [2, 4])
@pytest.mark.parametrize("Sq,Sc", [(256, 256), (384, 128), (250, 200)])
def test_carried_state_chain_equals_dense(nchunks, Sq, Sc):
    dtype = torch.bfloat16
    q = _mk(1, Sq, 2, 128, dtype, 7)
    ks = [_mk(1, Sc, 2, 128, dtype, 10 + c) for c in range(nchunks)]
    vs = [_mk(1, Sc, 2, 128, dtype, 20 + c) for c in range(nchunks)]
    out, lse = fwd_chunks(q, ks, vs, 128 ** -0.5)
    o_ref, lse_ref = orc.dense_attention(q.cpu(), torch.cat(ks, 1).cpu(), torch.cat(vs, 1).cpu())
    torch.testing.assert_close(out.double().cpu(), o_ref, **TOL[dtype])
    torch.testing.assert_close(lse.double().cpu(), lse_ref, rtol=1e-3, atol=2e-3)


def test_chain_with_scale_jumps_forces_rescale():
    """Later chunks have much larger logits -> the lazy O rescale path must fire."""
    dtype = torch.bfloat16
    q = _mk(1, 256, 1, 128, dtype, 30)
    ks = [_mk(1, 256, 1, 128, dtype, 31) * 0.1, _mk(1, 256, 1, 128, dtype, 32) * 3.0, _mk(1, 256, 1, 128, dtype, 33)]
    vs = [_mk(1, 256, 1, 128, dtype, 34 + c) for c in range(3)]
    out, lse = fwd_chunks(q, ks, vs, 128 ** -0.5)
    o_ref, lse_ref = orc.dense_attention(q.cpu(), torch.cat(ks, 1).cpu(), torch.cat(vs, 1).cpu())
    torch.testing.assert_close(out.double().cpu(), o_ref, **TOL[dtype])
    torch.testing.assert_close(lse.double().cpu(), lse_ref, rtol=1e-3, atol=2e-3)


def test_strict_causal_offset_matches_oracle_chunk():
    """causal_offset=-1 (striped 'causal_shift'): row 0 sees nothing and keeps its state."""
    dtype = torch.bfloat16
    S = 384
    q = _mk(1, S, 2, 128, dtype, 40)
    ks = [_mk(1, S, 2, 128, dtype, 41 + c) for c in range(2)]
    vs = [_mk(1, S, 2, 128, dtype, 43 + c) for c in range(2)]
    out, lse = fwd_chunks(q, ks, vs, 128 ** -0.5, causal=True, offsets=[0, -1])
    o, l = orc.chunk_forward(q.cpu(), ks[0].cpu(), vs[0].cpu(), None, None, 128 ** -0.5, "causal")
    o, l = orc.chunk_forward(q.cpu(), ks[1].cpu(), vs[1].cpu(), o, l, 128 ** -0.5, "causal_strict")
    torch.testing.assert_close(out.double().cpu(), o, **TOL[dtype])
    torch.testing.assert_close(lse.double().cpu(), l, rtol=1e-3, atol=2e-3)


def test_normal_layout_bhsd_and_half_views():
    dtype = torch.bfloat16
    q, k, v = (_mk(2, 512, 3, 128, dtype, s, seq_dim=2) for s in (50, 51, 52))
    out, lse = fwd_chunks(q, [k], [v], 128 ** -0.5, seq_dim=2)
    o_ref, lse_ref = orc.dense_attention(*(_to_bshd(t, 2).cpu() for t in (q, k, v)))
    torch.testing.assert_close(_to_bshd(out, 2).double().cpu(), o_ref, **TOL[dtype])
    # half-sequence views: second half of q against first half of k/v, no copies
    qh, kh, vh = q.narrow(2, 256, 256), k.narrow(2, 0, 256), v.narrow(2, 0, 256)
    out2, _ = fwd_chunks(qh, [kh], [vh], 128 ** -0.5, seq_dim=2)
    o_ref2, _ = orc.dense_attention(*(_to_bshd(t, 2).cpu() for t in (qh, kh, vh)))
    torch.testing.assert_close(_to_bshd(out2, 2).double().cpu(), o_ref2, **TOL[dtype])


def test_linearity_in_v_at_large_size():
    """Size-independent property at a size the oracle cannot finish: O is linear in V."""
    dtype = torch.bfloat16
    S = 8192
    q, k = _mk(1, S, 4, 128, dtype, 60), _mk(1, S, 4, 128, dtype, 61)
    v1, v2 = _mk(1, S, 4, 128, dtype, 62), _mk(1, S, 4, 128, dtype, 63)
    o1, l1 = fwd_chunks(q, [k], [v1], 128 ** -0.5)
    o2, l2 = fwd_chunks(q, [k], [v2], 128 ** -0.5)
    o12, l12 = fwd_chunks(q, [k], [(v1.float() + v2.float()).to(dtype)], 128 ** -0.5)
    assert torch.equal(l1, l2) and torch.equal(l1, l12)  # lse does not depend on V; kernel is deterministic
    torch.testing.assert_close(o12.float(), o1.float() + o2.float(), rtol=2e-2, atol=2e-2)
    # and against torch SDPA (fp32 math on the GPU) for one head
    ref = torch.nn.functional.scaled_dot_product_attention(
        q[:, :, :1].permute(0, 2, 1, 3).float(), k[:, :, :1].permute(0, 2, 1, 3).float(),
        v1[:, :, :1].permute(0, 2, 1, 3).float()).permute(0, 2, 1, 3)
    torch.testing.assert_close(o1[:, :, :1].float(), ref, rtol=1.6e-2, atol=2e-2)
