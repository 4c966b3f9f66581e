"""Public API and ring drivers: ``burst_attn_func`` / ``burst_attn_func_striped``
with the reference's positional signature (burst_attn_interface.py:109-158), as
``torch.autograd.Function``s ``OpBurstAttn`` / ``OpBurstAttnStrip`` (:161-613).

What is the same as the reference: names, argument order and defaults, shard
layouts (contiguous / zigzag halves / striped, test/test_burst.py:44-58), the
ring schedules (forward: K/V rotate, :214-242; backward: the Q-bundle
(delta, dO, Q, lse) rotates and the partial dQ rides one hop behind it,
:291-396), output dtype, the assertion that causal needs flash == "cuda".

What is B200-native instead: every round is ONE kernel launch of the
C-ABI library (carried (O, lse) state and fp32 dQ/dK/dV accumulation fused into
the tile kernels; half-sequence and shifted cases are pointer/length views or a
causal offset, never ``.contiguous()`` copies); user tensors are never used as
receive buffers; the ring hop is grouped NCCL send/recv on a side stream posted
before the round's kernel and awaited after it.
"""
from __future__ import annotations

import math
import os
from typing import List

import torch

from .chunk_ops import get_ops
from .comm import Ring, get_rank, get_world_size

__all__ = ["burst_attn_func", "burst_attn_func_striped", "OpBurstAttn", "OpBurstAttnStrip",
           "get_partition_id", "split2_gethalf"]


class _Range:
    """NVTX range per ring round (BA_NVTX=1): `ncu --nvtx --nvtx-include "bwd_round_3/"` or a timeline tool then
    sees the post / kernel / wait of one round as a unit.  A no-op otherwise."""
    on = os.environ.get("BA_NVTX", "0") == "1"

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _Range.on:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _Range.on:
            torch.cuda.nvtx.range_pop()
        return False


def get_partition_id(double_group, r):
    """Reference :20-37.  Single ring (``double_group[0] is None``): the OFFSET ``r - 1`` of the held shard
    behind this rank.  Double ring: the RANK ID of the held shard, ``W = L*M``, rank ``= inter*L + intra``:
    node ``(inter - (r-1)//L) mod M``, slot ``(intra - (r-1)%L) mod L``."""
    if double_group[0] is None:
        return r - 1
    L, M = get_world_size(double_group[0]), get_world_size(double_group[1])
    b, a = get_rank(double_group[0]), get_rank(double_group[1])
    return ((a - (r - 1) // L) % M) * L + (b - (r - 1) % L) % L


# Hierarchical ring over NCCL: on when the caller passes double_group (BA_DOUBLE_RING=0 forces the flat ring over
# process_group).  `RING_CHECK_DOUBLE=2,4 torchrun tests/ring_check.py` passed on 8 x B200 for intra-node rings of 2
# and of 4 (profiles/ring_check_r02_n8_nccl.txt).
_DOUBLE_RING_DEFAULT = "1"


class _Topology:
    """Ring topology of one call: flat (one ring over ``process_group``) or hierarchical
    (``double_group = [intra, inter]``, each optionally a ``(group, dq_group)`` pair, reference :188-194)."""

    def __init__(self, process_group, double_group):
        self.group = process_group
        self.W, self.rank = get_world_size(process_group), get_rank(process_group)
        intra, inter = double_group[0], double_group[1]
        self.intra_dq = self.inter_dq = None
        if isinstance(intra, (tuple, list)):
            intra, self.intra_dq = intra
        if isinstance(inter, (tuple, list)):
            inter, self.inter_dq = inter
        self.intra, self.inter = intra, inter
        self.L, self.M = self.W, 1
        self.hier = False
        if intra is not None and inter is not None and os.environ.get("BA_DOUBLE_RING", _DOUBLE_RING_DEFAULT) != "0":
            L, M = get_world_size(intra), get_world_size(inter)
            if 1 < L < self.W:  # reference comm.py:215-219: a ring that is all-intra (or all-inter) is flat
                assert L * M == self.W, f"double ring: intra size {L} x inter size {M} != world {self.W}"
                self.a, self.b = get_rank(inter), get_rank(intra)
                assert self.a * L + self.b == self.rank, "double ring expects rank = inter_rank * intra_size + intra_rank"
                self.L, self.M, self.hier = L, M, True

    def source(self, r):
        """Rank whose shard (K/V forward, Q-bundle backward) is held in round r (1-based)."""
        if not self.hier:
            return (self.rank - (r - 1)) % self.W
        return get_partition_id([self.intra, self.inter], r)


def split2_gethalf(inp, first_dim, half_idx=0):
    """Half-sequence VIEW (reference :96-106); never copied here."""
    dim = 1 if first_dim else 2
    n = inp.shape[dim] // 2
    return inp.narrow(dim, 0, n) if half_idx == 0 else inp.narrow(dim, n, inp.shape[dim] - n)


def _half(t, dim, idx):
    n = t.shape[dim] // 2
    return t.narrow(dim, 0, n) if idx == 0 else t.narrow(dim, n, t.shape[dim] - n)


def _nbytes(t) -> int:
    return t.numel() * t.element_size()


def _l2_block() -> int:
    """Rows/keys per sub-launch.  One (batch, head) slice of 32768 keys is 16 MiB of K+V (or 32 MiB of
    Q, dO and fp32 dQ in the backward), so the streamed operands of a launch stay resident in B200's
    126 MB L2 while all CTAs of a head sweep them -- measured on one GPU at S=262144: forward +7 %
    (1147 vs 1069 TFLOP/s), backward unchanged, compared with one launch over the whole sequence.  The multi-GPU rounds at S_local <= 49152 are
    unaffected.  BA_L2_BLOCK overrides (tests use tiny blocks)."""
    return int(os.environ.get("BA_L2_BLOCK", "32768"))


def _bias_kw(bias, c0=None, n=None):
    """Keyword for the chunk operators: the key-bias view of keys [c0, c0+n) (nothing when there is no bias, so
    test operators without a bias parameter keep working)."""
    if bias is None:
        return {}
    return {"bias": bias if c0 is None else bias.narrow(2, c0, n)}


def _fwd_round(ops, q, k, v, o_acc, lse, out, scale, causal, off, first, last, seq_dim, bias=None):
    """One forward ring round, split over K/V blocks that fit L2 (each block is one kernel launch with
    the carried state; the state pass costs 2 x 512 B per row and head, the block > 8 MB of math)."""
    blk = _l2_block()
    Sq, Sk = q.shape[seq_dim], k.shape[seq_dim]
    if Sk <= blk + blk // 2:
        ops.fwd_chunk(q, k, v, o_acc, lse, out, scale, causal, off, first, last, seq_dim, **_bias_kw(bias))
        return
    n = (Sk + blk - 1) // blk
    for c in range(n):
        c0 = c * blk
        kc, vc = k.narrow(seq_dim, c0, min(blk, Sk - c0)), v.narrow(seq_dim, c0, min(blk, Sk - c0))
        if not causal:
            ops.fwd_chunk(q, kc, vc, o_acc, lse, out, scale, False, 0, first and c == 0, last and c == n - 1, seq_dim,
                          **_bias_kw(bias, c0, min(blk, Sk - c0)))
            continue
        # causal: rows before r_start see none of this block's keys (key c0+b visible to row a iff
        # c0 + b <= a + off); keep r_start on a tile-pair boundary
        r_start = max(0, (c0 - off) // 256 * 256)
        if r_start >= Sq:
            break
        ops.fwd_chunk(q.narrow(seq_dim, r_start, Sq - r_start), kc, vc,
                      o_acc.narrow(seq_dim, r_start, Sq - r_start), lse.narrow(2, r_start, Sq - r_start), None,
                      scale, True, r_start + off - c0, first and c == 0, False, seq_dim,
                      **_bias_kw(bias, c0, min(blk, Sk - c0)))
    if causal and last:
        ops.cast(o_acc, out, seq_dim)


def _fwd_round_needs_state(k, seq_dim) -> bool:
    blk = _l2_block()
    return k.shape[seq_dim] > blk + blk // 2


def _bwd_round(ops, g, q, k, v, delta, lse, dq_part, dk_acc, dv_acc, scale, causal, off, seq_dim, deterministic,
               bias=None):
    """One backward ring round, split over blocks of Q-bundle rows that fit L2."""
    blk = _l2_block()
    Sq, Sk = q.shape[seq_dim], k.shape[seq_dim]
    if Sq <= blk + blk // 2:
        ops.bwd_chunk(g, q, k, v, delta, lse, dq_part, dk_acc, dv_acc, scale, causal, off, seq_dim, deterministic,
                      **_bias_kw(bias))
        return
    for r0 in range(0, Sq, blk):
        n = min(blk, Sq - r0)
        kk, vv, dk, dv, o2, bkw = k, v, dk_acc, dv_acc, off, _bias_kw(bias)
        if causal:
            kmax = min(Sk, r0 + n + off)  # keys visible to the last row of this block
            if kmax <= 0:
                continue
            kk, vv, bkw = k.narrow(seq_dim, 0, kmax), v.narrow(seq_dim, 0, kmax), _bias_kw(bias, 0, kmax)
            dk, dv = dk_acc.narrow(seq_dim, 0, kmax), dv_acc.narrow(seq_dim, 0, kmax)
            o2 = off + r0
        ops.bwd_chunk(g.narrow(seq_dim, r0, n), q.narrow(seq_dim, r0, n), kk, vv, delta.narrow(2, r0, n),
                      lse.narrow(2, r0, n), dq_part.narrow(seq_dim, r0, n), dk, dv, scale, causal, o2, seq_dim,
                      deterministic, **bkw)


def _check_inputs(q, k, v, seq_dim):
    assert q.dim() == 4 and k.shape == v.shape and q.shape[0] == k.shape[0] and q.shape[3] == k.shape[3], \
        "q, k, v must be 4-D with matching batch and head_dim"
    assert q.shape[3 - seq_dim] == k.shape[3 - seq_dim], "q and k/v must have the same number of heads"
    assert q.dtype == k.dtype == v.dtype, "q, k, v must share a dtype"


def _fwd_dispatch(ops, mode, r, W, i, j, q, cur_k, cur_v, o_acc, lse, out, scale, seq_dim):
    """The kernel work of forward round r on rank i holding the K/V shard of rank j."""
    first, last = r == 1, r == W
    if mode == "none":
        _fwd_round(ops, q, cur_k, cur_v, o_acc, lse, out, scale, False, 0, first, last, seq_dim)
    elif mode == "zigzag":
        if r == 1:  # own shard: plain causal (:221-224)
            _fwd_round(ops, q, cur_k, cur_v, o_acc, lse, out, scale, True, 0, first, last, seq_dim)
        elif j < i:  # split_kv: all Q x first half of K/V (:225-231)
            _fwd_round(ops, q, _half(cur_k, seq_dim, 0), _half(cur_v, seq_dim, 0), o_acc, lse, out, scale,
                       False, 0, False, last, seq_dim)
        else:  # second half of Q x all K/V, merged into the second half of the state (:232-235)
            _fwd_round(ops, _half(q, seq_dim, 1), cur_k, cur_v, _half(o_acc, seq_dim, 1), _half(lse, 2, 1),
                       _half(out, seq_dim, 1), scale, False, 0, False, last, seq_dim)
            if last:  # rows the last round did not visit: hand their finished state over
                ops.cast(_half(o_acc, seq_dim, 0), _half(out, seq_dim, 0), seq_dim)
    elif mode == "striped":
        # source rank ahead of us -> strictly-lower-triangular (causal_shift, :454,:463-475)
        _fwd_round(ops, q, cur_k, cur_v, o_acc, lse, out, scale, True, -1 if j > i else 0, first, last, seq_dim)
    else:
        raise ValueError(mode)


# --------------------------------------------------------------------------- #
# forward ring (reference OpBurstAttn.forward :171-253, OpBurstAttnStrip.forward :411-493)
# --------------------------------------------------------------------------- #
def _ring_forward(q, k, v, scale, seq_dim, mode, topo):
    """mode: "none" (non-causal) | "zigzag" | "striped".  Returns (out, lse[B,H,S] fp32)."""
    if topo.hier:
        return _ring_forward_hier(q, k, v, scale, seq_dim, mode, topo)
    ops = get_ops()
    ring = Ring(topo.group, tag="ring")
    W, i = ring.world_size, ring.rank
    B, S, H = q.shape[0], q.shape[seq_dim], q.shape[3 - seq_dim]
    if mode == "zigzag":
        assert S % 2 == 0, "zigzag causal sharding needs an even local sequence length"
    out = torch.empty_like(q)
    lse = torch.empty((B, H, S), dtype=torch.float32, device=q.device)
    need_state = W > 1 or _fwd_round_needs_state(k, seq_dim)
    o_acc = torch.empty(q.shape, dtype=torch.float32, device=q.device) if need_state else None
    if W > 1:
        k, v = k.contiguous(), v.contiguous()
    ring.begin(q, [_nbytes(k), _nbytes(v)] * min(2, W - 1))
    recv = [[ring.empty_like(k), ring.empty_like(v)] for _ in range(min(2, W - 1))]
    cur_k, cur_v = k, v
    for r in range(1, W + 1):
        j = topo.source(r)  # source rank of the held K/V (App. B)
        with _Range(f"fwd_round_{r}"):
            if r != W:
                nxt = recv[(r - 1) % len(recv)]
                ring.post([cur_k, cur_v], nxt)
            _fwd_dispatch(ops, mode, r, W, i, j, q, cur_k, cur_v, o_acc, lse, out, scale, seq_dim)
            if r != W:
                ring.wait()
                cur_k, cur_v = nxt
    return out, lse


def _ring_forward_hier(q, k, v, scale, seq_dim, mode, topo):
    """Hierarchical ("double") ring, W = L*M (reference comm.py:187-254): rounds run
    in M cycles of L steps.  Within a cycle K/V hop round the intra-node ring; the block a cycle starts
    with is at the same time forwarded to the next node over the inter-node ring, where it starts the
    following cycle -- a prefetch with L rounds of kernel time to hide behind.  Unlike the reference no
    send-side copy is made: the cycle's starting block is never a receive target while it is in flight
    (two inter-node buffers alternate)."""
    ops = get_ops()
    # (the copy-engine transport serves the flat ring only: its receive buffers are tied to one ring's arena)
    intra, inter = Ring(topo.intra, tag="ring", transport="nccl"), Ring(topo.inter, tag="inter", transport="nccl")
    L, M, W, i = topo.L, topo.M, topo.W, topo.rank
    B, S, H = q.shape[0], q.shape[seq_dim], q.shape[3 - seq_dim]
    if mode == "zigzag":
        assert S % 2 == 0, "zigzag causal sharding needs an even local sequence length"
    out = torch.empty_like(q)
    lse = torch.empty((B, H, S), dtype=torch.float32, device=q.device)
    o_acc = torch.empty(q.shape, dtype=torch.float32, device=q.device)
    k, v = k.contiguous(), v.contiguous()
    recv = [[torch.empty_like(k), torch.empty_like(v)] for _ in range(min(2, L - 1))]
    xbuf = [[torch.empty_like(k), torch.empty_like(v)] for _ in range(min(2, M - 1))]
    cur = [k, v]
    for r in range(1, W + 1):
        c, t = divmod(r - 1, L)
        j = topo.source(r)
        if t == 0 and c != M - 1:  # next cycle's starting block, from the previous node
            inter.post(cur, xbuf[c % len(xbuf)])
        if t != L - 1:
            nxt = recv[(r - 1) % len(recv)]
            intra.post(cur, nxt)
        _fwd_dispatch(ops, mode, r, W, i, j, q, cur[0], cur[1], o_acc, lse, out, scale, seq_dim)
        if t != L - 1:
            intra.wait()
            cur = nxt
        elif c != M - 1:
            inter.wait()
            cur = xbuf[c % len(xbuf)]
    return out, lse


# --------------------------------------------------------------------------- #
# backward ring (reference OpBurstAttn.backward :256-398, OpBurstAttnStrip.backward :496-613)
# --------------------------------------------------------------------------- #
def _bwd_dispatch(ops, mode, r, i, j, bundle, dq_part, k, v, dk_acc, dv_acc, scale, seq_dim, deterministic):
    """The kernel work of backward round r: K/V at home on rank i, Q-bundle of rank j."""
    dlt, g, qq, ls = bundle
    if mode == "none":
        _bwd_round(ops, g, qq, k, v, dlt, ls, dq_part, dk_acc, dv_acc, scale, False, 0, seq_dim, deterministic)
    elif mode == "zigzag":
        if r == 1:
            _bwd_round(ops, g, qq, k, v, dlt, ls, dq_part, dk_acc, dv_acc, scale, True, 0, seq_dim, deterministic)
        elif j < i:  # split_q: second half of the bundle x all K/V (:322-345,:383-386)
            _bwd_round(ops, _half(g, seq_dim, 1), _half(qq, seq_dim, 1), k, v, _half(dlt, 2, 1), _half(ls, 2, 1),
                       _half(dq_part, seq_dim, 1), dk_acc, dv_acc, scale, False, 0, seq_dim, deterministic)
        else:  # whole bundle x first half of K/V (:347-367,:387-390)
            _bwd_round(ops, g, qq, _half(k, seq_dim, 0), _half(v, seq_dim, 0), dlt, ls, dq_part,
                       _half(dk_acc, seq_dim, 0), _half(dv_acc, seq_dim, 0), scale, False, 0, seq_dim,
                       deterministic)
    elif mode == "striped":
        # K/V home on i, bundle from j: strict iff j < i (causal_shift, :529)
        _bwd_round(ops, g, qq, k, v, dlt, ls, dq_part, dk_acc, dv_acc, scale, True, -1 if j < i else 0, seq_dim,
                   deterministic)
    else:
        raise ValueError(mode)


def _ring_backward(d_o, q, k, v, out, lse, scale, seq_dim, mode, topo, deterministic):
    ops = get_ops()
    W, i = topo.W, topo.rank
    dev = q.device
    q, k, v, d_o, out = (t.contiguous() for t in (q, k, v, d_o, out))
    B, S, H = q.shape[0], q.shape[seq_dim], q.shape[3 - seq_dim]

    # delta always travels instead of O (the reference's optimize_bwd_comm, :271-278):
    # 4 B instead of 2*D B per row and head, and the tile kernel never needs O.
    delta = torch.empty((B, H, S), dtype=torch.float32, device=dev)
    ops.delta(out, d_o, delta, seq_dim)

    f32 = dict(dtype=torch.float32, device=dev)
    dk_acc = torch.zeros(k.shape, **f32)
    dv_acc = torch.zeros(v.shape, **f32)

    def round_kernel(r, j, bundle, dq_part):
        _bwd_dispatch(ops, mode, r, i, j, bundle, dq_part, k, v, dk_acc, dv_acc, scale, seq_dim, deterministic)

    bundle = [delta, d_o, q, lse.contiguous()]
    # `part`: this round's dQ partial (the kernel reduce-adds into it)
    if W == 1:
        part = torch.zeros(q.shape, **f32)
        round_kernel(1, i, bundle, part)
        dq_final = part
    elif topo.hier:
        part = torch.zeros(q.shape, **f32)
        dq_final = _bwd_rounds_hier(ops, topo, round_kernel, bundle, part, q.shape, f32, seq_dim)
    else:
        ring = Ring(topo.group, tag="ring")
        # every buffer that is ever the destination of a hop comes from the ring (the copy-engine transport
        # keeps them in its IPC-mapped arena): two bundle sets and the three rotating fp32 dQ buffers
        ring.begin(q, [_nbytes(t) for t in bundle] * min(2, W - 1) + [4 * q.numel()] * 3)
        recv = [[ring.empty_like(t) for t in bundle] for _ in range(min(2, W - 1))]
        hold = None                      # fp32 dQ accumulated for the bundle held in the previous round
        part = ring.empty(q.shape, torch.float32, dev)
        part.zero_()
        spare = [ring.empty(q.shape, torch.float32, dev), ring.empty(q.shape, torch.float32, dev)]
        for r in range(1, W + 1):
            j = topo.source(r)
            srcs: List[torch.Tensor] = []
            dsts: List[torch.Tensor] = []
            if r != W:  # bundle hop (:295-299)
                nxt = recv[(r - 1) % len(recv)]
                srcs += bundle
                dsts += nxt
            if r != 1:  # dQ hop, one behind its bundle (:300-302)
                inbound = spare.pop()
                srcs.append(hold)
                dsts.append(inbound)
            with _Range(f"bwd_round_{r}"):
                if srcs:
                    ring.post(srcs, dsts)
                round_kernel(r, j, bundle, part)
                if srcs:
                    ring.wait()
            if r != W:
                bundle = nxt
            if r == 1:
                hold, part = part, spare.pop()
                part.zero_()
            else:
                ops.accumulate(part, inbound, seq_dim)  # dq += buf (:379-390), in fp32
                spare.append(hold)
                hold = inbound
                if r != W:
                    part.zero_()
        # final hop home (:393-396)
        dq_final = spare.pop()
        ring.post([hold], [dq_final])
        ring.wait()

    dq = torch.empty_like(q)
    dk = torch.empty_like(k)
    dv = torch.empty_like(v)
    ops.cast(dq_final, dq, seq_dim)
    ops.cast(dk_acc, dk, seq_dim)
    ops.cast(dv_acc, dv, seq_dim)
    return dq, dk, dv


def _bwd_rounds_hier(ops, topo, round_kernel, bundle, part, qshape, f32, seq_dim):
    """Backward rounds over the hierarchical ring (W = L*M); returns the fp32 dQ of this rank's own rows.

    The Q-bundle travels exactly like K/V in the forward (intra-node hops inside a cycle, the cycle's
    starting bundle prefetched to the next node).  Its dQ comes home in two levels (the reference's
    ``double_ring_send_recv_q``, comm.py:187-213, restated):
      * inside a cycle the partial rides one intra-node hop behind its bundle and picks up each rank's
        contribution (as in the flat ring); one more intra-node hop after the cycle's last step closes the
        ring, so the NODE sum for a bundle lands on the rank that started it in this node;
      * node sums chain along the inter-node ring: at the start of cycle c the node sum of the bundle
        started in cycle c-1 is added to the running sum received from the previous node and sent on --
        L rounds of kernel time to hide behind.  After the last cycle the same step is the hop home.
    """
    L, M, W = topo.L, topo.M, topo.W
    intra = Ring(topo.intra, tag="ring", transport="nccl")
    inter = Ring(topo.inter, tag="inter", transport="nccl")
    inter_q = Ring(topo.inter_dq if topo.inter_dq is not None else topo.inter, tag="inter_dq", transport="nccl")
    recv = [[torch.empty_like(t) for t in bundle] for _ in range(min(2, L - 1))]
    xbuf = [[torch.empty_like(t) for t in bundle] for _ in range(min(2, M - 1))]
    free: List[torch.Tensor] = []

    def take():
        return free.pop() if free else torch.empty(qshape, **f32)

    hold = None      # dQ accumulated in this node for the bundle held in the previous round
    running = None   # inter-node running sum in flight to the next node (kept alive until awaited)
    inter_in = None  # running sum arriving from the previous node

    def chain(node_sum):
        """node_sum (+ the sum received from the previous node) -> next node; returns the receive buffer."""
        nonlocal running, inter_in
        if inter_in is not None:
            inter_q.wait()
            ops.accumulate(inter_in, node_sum, seq_dim)
            free.extend([inter_in, running])
        running, inter_in = node_sum, take()
        inter_q.post([running], [inter_in])

    for r in range(1, W + 1):
        c, t = divmod(r - 1, L)
        j = topo.source(r)
        srcs: List[torch.Tensor] = []
        dsts: List[torch.Tensor] = []
        inbound = None
        if t != L - 1:  # bundle hop inside the node
            nxt = recv[(r - 1) % len(recv)]
            srcs += bundle
            dsts += nxt
        if r != 1:  # dQ hop: behind its bundle (t > 0), or closing the previous cycle's ring (t == 0)
            inbound = take()
            srcs.append(hold)
            dsts.append(inbound)
        if srcs:
            intra.post(srcs, dsts)
        if t == 0 and c != M - 1:  # next cycle's starting bundle, from the previous node
            inter.post(bundle, xbuf[c % len(xbuf)])
        round_kernel(r, j, bundle, part)
        if srcs:
            intra.wait()
        if t == 0:
            if r != 1:
                free.append(hold)
                chain(inbound)  # inbound = node sum of the bundle this rank started one cycle ago
            hold, part = part, take()
            part.zero_()
        else:
            ops.accumulate(part, inbound, seq_dim)
            free.append(hold)
            hold = inbound
            if r != W:
                part.zero_()
        if t != L - 1:
            bundle = nxt
        elif c != M - 1:
            inter.wait()
            bundle = xbuf[c % len(xbuf)]
    # close the last cycle's intra-node ring, then the last inter-node hop is the hop home (:393-396)
    node_sum = take()
    intra.post([hold], [node_sum])
    intra.wait()
    chain(node_sum)
    inter_q.wait()
    return inter_in


# --------------------------------------------------------------------------- #
def _prepare(ctx, q, k, v, softmax_scale, flash, causal, optimize_bwd_comm, deterministic, process_group,
             double_group):
    assert not causal or flash == "cuda", "Causal attention only supported for Flash v2"
    ctx.softmax_scale = 1 / math.sqrt(q.shape[-1]) if softmax_scale is None else softmax_scale
    ctx.flash = None if flash not in ["cuda", "triton"] else flash
    ctx.seq_dim = 1 if ctx.flash else 2
    ctx.causal = causal
    ctx.optimize_bwd_comm = optimize_bwd_comm  # delta always travels; kept for API parity
    ctx.deterministic = deterministic
    ctx.process_group = process_group
    ctx.double_group = double_group
    ctx.topo = _Topology(process_group, double_group)
    _check_inputs(q, k, v, ctx.seq_dim)


def _pad_head_dim(ops, tensors):
    """The sm_100a tile kernels exist for head_dim 64 and 128 (``ops.tile_head_dims``; the reference's
    CPU-runnable configuration C1 has 64, its benchmarks 128).  Any other head_dim <= 128 is run exactly by
    zero-padding the last axis once per call up to the next tile width: padded Q/K columns add 0 to every score,
    padded V columns produce output columns that are exactly 0 and are sliced off, and the same holds for
    dO -> dQ/dK/dV."""
    tiles = getattr(ops, "tile_head_dims", None)
    D = tensors[0].shape[-1]
    if tiles is None or D in tiles:
        return tensors, D
    bigger = [t for t in tiles if t > D]
    assert bigger, f"head_dim {D} > {max(tiles)} is not supported"
    return [torch.nn.functional.pad(t, (0, min(bigger) - D)) for t in tensors], D


def _unpad(t, D):
    return t if t.shape[-1] == D else t[..., :D].contiguous()


def _op_forward(ctx, q, k, v, mode):
    ctx.host = False
    if q.device.type == "cpu" and getattr(get_ops(), "name", "") == "sm100":  # (tests inject CPU chunk operators)
        # host-resident operands (pinned CPU tensors, one rank): copies stream under the kernels (host_stream.py)
        from . import host_stream
        if not host_stream.is_host_call(q, k, v):
            raise TypeError("burst_attn_b200 needs CUDA tensors (or pinned CPU tensors on a CUDA machine); there is "
                            "no CPU implementation")
        assert ctx.topo.W == 1, "host-resident operands are supported on a single rank only (pass device tensors)"
        assert q.shape[-1] in getattr(get_ops(), "tile_head_dims", (q.shape[-1],)), "host-resident operands need head_dim 64 or 128"
        ctx.host, ctx.mode, ctx.head_dim = True, mode, q.shape[-1]
        o_host, saved = host_stream.forward(q, k, v, ctx.softmax_scale, ctx.seq_dim, mode != "none", _l2_block())
        ctx.save_for_backward(*saved)
        return o_host
    (qp, kp, vp), ctx.head_dim = _pad_head_dim(get_ops(), [q, k, v])
    out, lse = _ring_forward(qp, kp, vp, ctx.softmax_scale, ctx.seq_dim, mode, ctx.topo)
    ctx.mode = mode
    ctx.save_for_backward(qp, kp, vp, lse, out)
    return _unpad(out, ctx.head_dim)


def _op_backward(ctx, grad_output):
    if ctx.host:
        from . import host_stream
        grads = host_stream.backward(grad_output, ctx.saved_tensors, ctx.softmax_scale, ctx.seq_dim,
                                     ctx.mode != "none", _l2_block(), ctx.deterministic)
        return tuple(grads) + (None,) * 7
    q, k, v, lse, out = ctx.saved_tensors
    (g,), _ = _pad_head_dim(get_ops(), [grad_output])
    dq, dk, dv = _ring_backward(g, q, k, v, out, lse, ctx.softmax_scale, ctx.seq_dim, ctx.mode, ctx.topo,
                                ctx.deterministic)
    return tuple(_unpad(t, ctx.head_dim) for t in (dq, dk, dv)) + (None,) * 7


class OpBurstAttn(torch.autograd.Function):
    """
    for Normal Attention (flash=None):  q, k, v: [B, N, S, H]
    for Flash ("cuda"/"triton"):        q, k, v: [B, S, N, H]
    Each rank passes its own sequence shard: contiguous when non-causal, zigzag
    halves {i, 2W-1-i} when causal.
    """

    @staticmethod
    def forward(ctx, q, k, v, softmax_scale=None, flash="cuda", causal=False, optimize_bwd_comm=False,
                deterministic=False, process_group=None, double_group=[None, None]):
        _prepare(ctx, q, k, v, softmax_scale, flash, causal, optimize_bwd_comm, deterministic, process_group,
                 double_group)
        return _op_forward(ctx, q, k, v, "zigzag" if causal else "none")

    @staticmethod
    def backward(ctx, grad_output):
        return _op_backward(ctx, grad_output)


class OpBurstAttnStrip(torch.autograd.Function):
    """Striped-causal variant: rank i owns tokens {i, i+W, i+2W, ...}."""

    @staticmethod
    def forward(ctx, q, k, v, softmax_scale=None, flash="cuda", causal=False, optimize_bwd_comm=False,
                deterministic=False, process_group=None, double_group=[None, None]):
        _prepare(ctx, q, k, v, softmax_scale, flash, causal, optimize_bwd_comm, deterministic, process_group,
                 double_group)
        return _op_forward(ctx, q, k, v, "striped" if causal else "none")

    @staticmethod
    def backward(ctx, grad_output):
        return _op_backward(ctx, grad_output)


def burst_attn_func_striped(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, softmax_scale: float = None,
                            flash: str = "cuda", causal: bool = False, optimize_bwd_comm: bool = False,
                            deterministic: bool = False, process_group=None, double_group=[None, None]):
    return OpBurstAttnStrip.apply(q, k, v, softmax_scale, flash, causal, optimize_bwd_comm, deterministic,
                                  process_group, double_group)


def burst_attn_func(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, softmax_scale: float = None,
                    flash: str = "cuda", causal: bool = False, optimize_bwd_comm: bool = False,
                    deterministic: bool = False, process_group=None, double_group=[None, None]):
    return OpBurstAttn.apply(q, k, v, softmax_scale, flash, causal, optimize_bwd_comm, deterministic,
                             process_group, double_group)
