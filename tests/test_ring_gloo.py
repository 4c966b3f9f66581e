"""World-size 2 and 4 CPU tests (gloo) of the N>1 path: the ring drivers behind
burst_attn_func / burst_attn_func_striped -- K/V rotation, the Q-bundle + dQ
ring of the backward, zigzag / striped shard views -- run with the oracle-backed
chunk operators injected (tests only) and are compared, through autograd, with
dense attention on the full sequence (the reference's protocol,
test/test_burst.py:159-219)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _install_deferred_transport():
    """Make the CPU transport behave like the asynchronous GPU one, to catch buffer hazards gloo's eager
    transfers hide: between ``post`` and ``wait`` a hop's destinations hold garbage (poisoned with NaN here, so
    any read before ``wait`` shows up in the results) and its sources must not change (snapshotted at ``post``
    and compared at ``wait``, which is when the transfer actually runs)."""
    from burst_attn import comm

    eager_commit = comm.Ring._commit_torch
    plain_empty = comm.Ring.empty

    def empty(self, shape, dtype, device):
        t = plain_empty(self, shape, dtype, device)
        self.__dict__.setdefault("_owned", set()).add(t.data_ptr())
        return t

    comm.Ring.empty = empty

    def commit(self, srcs, dsts):
        assert not getattr(self, "_deferred", None), "two hops in flight on one ring"
        if self.transport == "ce":  # copy-engine rings can only receive into buffers carved from their arena
            assert all(d.data_ptr() in getattr(self, "_owned", ()) for d in dsts), \
                "a hop destination was not allocated through Ring.empty / empty_like"
        self._deferred = (srcs, dsts, [s.clone() for s in srcs])
        for d in dsts:
            d.fill_(float("nan"))
        return ["deferred"]

    def wait(self, force_wait_inter=False):
        for r in self._reqs:
            assert r == "deferred"
            srcs, dsts, snap = self._deferred
            self._deferred = None
            for s, c in zip(srcs, snap):
                assert torch.equal(s, c), "a hop's source was modified between post and wait"
            for q in eager_commit(self, srcs, dsts):
                q.wait()
        self._reqs = []

    comm.Ring._commit_torch = commit
    comm.Ring.wait = wait


def _worker(rank, world, port, case, seq_dim, errq, intra=0, dq_groups=False):
    try:
        for p in (ROOT, os.path.join(ROOT, "burst-attention_b200"), os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from burst_attn import burst_attn_func, burst_attn_func_striped, chunk_ops
        from oracle import attention_oracle as orc
        from oracle_ops import OracleOps
        ops = OracleOps()
        chunk_ops._set_ops_for_testing(ops)
        if os.environ.get("BA_TEST_DEFERRED"):
            _install_deferred_transport()

        func, causal, layout = {
            "none": (burst_attn_func, False, "contiguous"),
            "zigzag": (burst_attn_func, True, "zigzag"),
            "striped": (burst_attn_func_striped, True, "striped"),
        }[case]
        torch.manual_seed(1234)  # same full tensors on every rank (the reference broadcasts from rank 0)
        B, S, H, D = 2, 8 * 2 * world, 3, 16
        q, k, v, do = (torch.randn(B, S, H, D, dtype=torch.float64) for _ in range(4))
        scale = D ** -0.5
        qr, kr, vr = (t.clone().requires_grad_() for t in (q, k, v))
        o_ref, _ = orc.dense_attention(qr, kr, vr, scale, causal)
        g_ref = torch.autograd.grad(o_ref, (qr, kr, vr), do)

        def sh(t):
            x = orc.shard(t, rank, world, layout)
            return x if seq_dim == 1 else x.permute(0, 2, 1, 3).contiguous()

        def unlay(t):
            return t if seq_dim == 1 else t.permute(0, 2, 1, 3)

        ql, kl, vl = (sh(t).requires_grad_() for t in (q, k, v))
        flash = "cuda" if seq_dim == 1 else None
        if causal and seq_dim == 2:
            return  # reference asserts causal needs flash == "cuda"
        double_group = [None, None]
        if intra:  # hierarchical ring: nodes of `intra` consecutive ranks (reference test/test_burst.py:120-156)
            os.environ["BA_DOUBLE_RING"] = "1"  # opt-in (flat ring over process_group is the default)
            rows = [list(range(n * intra, (n + 1) * intra)) for n in range(world // intra)]
            cols = [list(c) for c in zip(*rows)]
            mk = lambda ranks: dist.new_subgroups_by_enumeration(ranks, backend="gloo")[0]  # noqa: E731
            double_group = [mk(rows), mk(cols)]
            if dq_groups:
                double_group = [(double_group[0], mk(rows)), (double_group[1], mk(cols))]
            from burst_attn.burst_attn_interface import _Topology, get_partition_id
            topo = _Topology(None, double_group)
            assert topo.hier and (topo.L, topo.M) == (intra, world // intra)
            plain = [g[0] if isinstance(g, tuple) else g for g in double_group]
            seen = sorted(get_partition_id(plain, r) for r in range(1, world + 1))
            assert seen == list(range(world)), seen  # every shard is visited exactly once
            assert get_partition_id(plain, 1) == rank
            for r in range(1, world + 1):  # the oracle's restatement is pinned to the reference (tests/golden)
                assert get_partition_id(plain, r) == orc.get_partition_id_double(r, rank % intra, rank // intra, intra,
                                                                                world // intra)
        o = func(ql, kl, vl, None, flash, causal, True, False, None, double_group)
        g = torch.autograd.grad(o, (ql, kl, vl), sh(do))
        tol = dict(rtol=1e-5, atol=1e-5)  # fp32 carried state / accumulators in the driver
        torch.testing.assert_close(unlay(o.detach()), orc.shard(o_ref.detach(), rank, world, layout), **tol)
        for got, ref in zip(g, g_ref):
            torch.testing.assert_close(unlay(got), orc.shard(ref, rank, world, layout), **tol)
        # user inputs must not have been clobbered (the reference reuses k, v, q, dO as receive buffers)
        torch.testing.assert_close(unlay(kl.detach()), orc.shard(k, rank, world, layout))
        # one chunk launch per round, no copies: W forward rounds (+1 cast in the zigzag tail case)
        nf = sum(1 for c in ops.calls if c[0] == "fwd")
        nb = sum(1 for c in ops.calls if c[0] == "bwd")
        if not os.environ.get("BA_TEST_NO_LAUNCH_COUNT"):
            assert nf == world and nb == world, (nf, nb)
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        import traceback
        errq.put(f"rank {rank}: {type(e).__name__}: {e}\n{traceback.format_exc()}")
        raise


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("case", ["none", "zigzag", "striped"])
def test_ring_driver_matches_dense(world, case):
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, 1, errq)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    errs = []
    while not errq.empty():
        errs.append(errq.get())
    assert not errs, "\n".join(errs)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]


@pytest.mark.parametrize("world,intra,case,dq_groups", [
    (4, 2, "none", False), (4, 2, "zigzag", True), (4, 2, "striped", False),
    (6, 3, "zigzag", False), (6, 2, "none", True), (6, 2, "striped", False),
    (8, 4, "striped", True), (8, 2, "zigzag", False),
])
def test_double_ring_matches_dense(world, intra, case, dq_groups):
    """Hierarchical (double) ring, W = L*M with (L, M) in {(2,2), (3,2), (2,3), (4,2), (2,4)}: K/V and Q-bundle prefetch
    across nodes, dQ node sums chained along the inter-node ring (reference test_burst.py:239-247
    ``double_ring`` axis)."""
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, 1, errq, intra, dq_groups)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    errs = []
    while not errq.empty():
        errs.append(errq.get())
    assert not errs, "\n".join(errs)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]


@pytest.mark.parametrize("world,intra,case", [(4, 0, "none"), (4, 0, "zigzag"), (4, 0, "striped"),
                                              (4, 2, "zigzag"), (6, 2, "striped"), (6, 3, "none")])
def test_no_buffer_hazards_with_asynchronous_transport(world, intra, case, monkeypatch):
    """Flat and hierarchical rings under a transport that only moves data at ``wait`` and poisons the
    destinations at ``post`` (see _install_deferred_transport): what the side-stream transport does on GPUs."""
    monkeypatch.setenv("BA_TEST_DEFERRED", "1")
    monkeypatch.setenv("BA_RING_TRANSPORT", "ce")  # CPU tensors still travel over gloo; turns on the arena check
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, 1, errq, intra, False)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    errs = []
    while not errq.empty():
        errs.append(errq.get())
    assert not errs, "\n".join(errs)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]


def test_ring_driver_normal_layout_world2():
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, "none", 2, errq)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    errs = []
    while not errq.empty():
        errs.append(errq.get())
    assert not errs, "\n".join(errs)
    assert all(p.exitcode == 0 for p in procs)


def test_single_process_world1_cpu():
    """W=1 through the same driver (no process group needed)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from burst_attn import burst_attn_func, chunk_ops
    from oracle import attention_oracle as orc
    from oracle_ops import OracleOps
    chunk_ops._set_ops_for_testing(OracleOps())
    try:
        torch.manual_seed(0)
        q, k, v, do = (torch.randn(1, 32, 2, 16, dtype=torch.float64) for _ in range(4))
        for causal in (False, True):
            qq, kk, vv = (t.clone().requires_grad_() for t in (q, k, v))
            o = burst_attn_func(qq, kk, vv, None, "cuda", causal)
            g = torch.autograd.grad(o, (qq, kk, vv), do)
            o_ref, _, dq, dk, dv = orc.dense_attention_bwd(q, k, v, do, None, causal)
            torch.testing.assert_close(o.detach(), o_ref, rtol=1e-5, atol=1e-5)
            for a, b in zip(g, (dq, dk, dv)):
                torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)
    finally:
        chunk_ops._set_ops_for_testing(None)


def test_head_dim_padding_world1_cpu():
    """A head_dim the tile kernels are not built for (64 and 128 on the GPU; (8, 32) faked here) is zero-padded
    once per call by the driver up to the next tile width and sliced off again -- exact for O, dQ, dK, dV."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from burst_attn import burst_attn_func, burst_attn_func_striped, chunk_ops
    from oracle import attention_oracle as orc
    from oracle_ops import OracleOps
    ops = OracleOps()
    ops.tile_head_dims = (8, 32)
    chunk_ops._set_ops_for_testing(ops)
    try:
        torch.manual_seed(5)
        q, k, v, do = (torch.randn(2, 40, 3, 16, dtype=torch.float64) for _ in range(4))
        for func, causal in ((burst_attn_func, False), (burst_attn_func, True), (burst_attn_func_striped, True)):
            ops.calls.clear()
            qq, kk, vv = (t.clone().requires_grad_() for t in (q, k, v))
            o = func(qq, kk, vv, None, "cuda", causal)
            assert o.shape == q.shape and o.is_contiguous()
            g = torch.autograd.grad(o, (qq, kk, vv), do)
            assert all(c[1][-1] == 32 for c in ops.calls if c[0] in ("fwd", "bwd"))  # the tile saw padded heads
            o_ref, _, dq, dk, dv = orc.dense_attention_bwd(q, k, v, do, None, causal)
            torch.testing.assert_close(o.detach(), o_ref, rtol=1e-5, atol=1e-5)
            for a, b in zip(g, (dq, dk, dv)):
                assert a.shape == b.shape
                torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)
        # normal layout [B, H, S, D]
        p = lambda t: t.permute(0, 2, 1, 3).contiguous()  # noqa: E731
        qq, kk, vv = (p(t).requires_grad_() for t in (q, k, v))
        o = burst_attn_func(qq, kk, vv, None, None, False)
        g = torch.autograd.grad(o, (qq, kk, vv), p(do))
        o_ref, _, dq, dk, dv = orc.dense_attention_bwd(q, k, v, do, None, False)
        torch.testing.assert_close(o.detach(), p(o_ref), rtol=1e-5, atol=1e-5)
        for a, b in zip(g, (dq, dk, dv)):
            torch.testing.assert_close(a, p(b), rtol=1e-5, atol=1e-5)
    finally:
        chunk_ops._set_ops_for_testing(None)


@pytest.mark.parametrize("blk", [16, 24])
def test_l2_blocking_of_rounds_matches_dense(monkeypatch, blk):
    """The driver splits a round into L2-sized sub-launches (carried state forward, row blocks backward,
    causal offsets for views); with a tiny block size every code path runs on CPU."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from burst_attn import burst_attn_func, burst_attn_func_striped, chunk_ops
    from oracle import attention_oracle as orc
    from oracle_ops import OracleOps
    monkeypatch.setenv("BA_L2_BLOCK", str(blk))
    ops = OracleOps()
    chunk_ops._set_ops_for_testing(ops)
    try:
        torch.manual_seed(3)
        S = 100  # ragged against the block size and far above 1.5 blocks
        q, k, v, do = (torch.randn(2, S, 2, 16, dtype=torch.float64) for _ in range(4))
        for func, causal in ((burst_attn_func, False), (burst_attn_func, True), (burst_attn_func_striped, True)):
            ops.calls.clear()
            qq, kk, vv = (t.clone().requires_grad_() for t in (q, k, v))
            o = func(qq, kk, vv, None, "cuda", causal)
            g = torch.autograd.grad(o, (qq, kk, vv), do)
            o_ref, _, dq, dk, dv = orc.dense_attention_bwd(q, k, v, do, None, causal)
            torch.testing.assert_close(o.detach(), o_ref, rtol=1e-5, atol=1e-5)
            for a, b in zip(g, (dq, dk, dv)):
                torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)
            assert sum(1 for c in ops.calls if c[0] == "fwd") > 1 and sum(1 for c in ops.calls if c[0] == "bwd") > 1
    finally:
        chunk_ops._set_ops_for_testing(None)


def test_l2_blocking_random_shapes(monkeypatch):
    """Seeded sweep over (S, block size, batch, heads, causal flavour): the offset / view arithmetic of the
    L2-blocked rounds (row starts rounded to tile pairs, causal offsets of sub-views, ragged tails) against
    dense attention, forward and backward."""
    import random
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from burst_attn import burst_attn_func, burst_attn_func_striped, chunk_ops
    from oracle import attention_oracle as orc
    from oracle_ops import OracleOps
    ops = OracleOps()
    chunk_ops._set_ops_for_testing(ops)
    rng = random.Random(20240917)
    try:
        for case in range(24):
            S = rng.choice([17, 64, 100, 255, 256, 257, 300, 513, 640])
            blk = rng.choice([8, 16, 24, 100, 128, 256])
            Bn, Hn = rng.choice([1, 2]), rng.choice([1, 3])
            func, causal = rng.choice([(burst_attn_func, False), (burst_attn_func, True),
                                       (burst_attn_func_striped, True)])
            if causal and func is burst_attn_func and S % 2:
                S += 1  # zigzag halves
            monkeypatch.setenv("BA_L2_BLOCK", str(blk))
            torch.manual_seed(case)
            q, k, v, do = (torch.randn(Bn, S, Hn, 8, dtype=torch.float64) for _ in range(4))
            qq, kk, vv = (t.clone().requires_grad_() for t in (q, k, v))
            o = func(qq, kk, vv, None, "cuda", causal)
            g = torch.autograd.grad(o, (qq, kk, vv), do)
            o_ref, _, dq, dk, dv = orc.dense_attention_bwd(q, k, v, do, None, causal)
            msg = f"case {case}: S={S} blk={blk} B={Bn} H={Hn} {func.__name__} causal={causal}"
            torch.testing.assert_close(o.detach(), o_ref, rtol=1e-5, atol=1e-5, msg=lambda m: f"{msg}: {m}")
            for a, b in zip(g, (dq, dk, dv)):
                torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5, msg=lambda m: f"{msg}: {m}")
    finally:
        chunk_ops._set_ops_for_testing(None)


@pytest.mark.parametrize("case", ["none", "zigzag", "striped"])
def test_ring_driver_with_l2_blocking_world2(case, monkeypatch):
    """Ring rounds + L2 blocking together (blocks of 8 rows/keys, S_local = 16): half views of the
    zigzag rounds and the causal offsets compose."""
    monkeypatch.setenv("BA_L2_BLOCK", "8")
    monkeypatch.setenv("BA_TEST_NO_LAUNCH_COUNT", "1")
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, 1, errq)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    errs = []
    while not errq.empty():
        errs.append(errq.get())
    assert not errs, "\n".join(errs)
    assert all(p.exitcode == 0 for p in procs)


@pytest.mark.parametrize("blk", [16, 1000])
def test_single_gpu_wrappers_packed_bias_blocked_cpu(monkeypatch, blk):
    """burst_attn.flash_triton (flash_attn_func / _kvpacked_func / _qkvpacked_func): strided views of the packed
    tensors, bottom-right-aligned causal with Sq != Sk, per-key bias narrowed per L2 block -- the Python logic,
    on CPU with oracle-backed chunk operators, against dense attention."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from burst_attn import chunk_ops
    from burst_attn.flash_triton import flash_attn_func, flash_attn_kvpacked_func, flash_attn_qkvpacked_func
    from oracle import attention_oracle as orc
    from oracle_ops import OracleOps
    monkeypatch.setenv("BA_L2_BLOCK", str(blk))
    ops = OracleOps()
    ops.tile_head_dims = (16,)
    chunk_ops._set_ops_for_testing(ops)
    try:
        torch.manual_seed(11)
        B, S, H, D = 2, 70, 2, 16
        qkv = torch.randn(B, S, 3, H, D, dtype=torch.float64)
        do = torch.randn(B, S, H, D, dtype=torch.float64)
        bias = torch.randn(1, H, 1, S, dtype=torch.float64)
        bias[..., 3::5] = float("-inf")
        q, k, v = (qkv[:, :, i].contiguous() for i in range(3))
        for causal in (False, True):
            o_ref, _, dq, dk, dv = orc.dense_attention_bwd(q, k, v, do, None, causal, bias=bias)
            p = qkv.clone().requires_grad_()
            o = flash_attn_qkvpacked_func(p, bias, causal)
            (dqkv,) = torch.autograd.grad(o, (p,), do)
            torch.testing.assert_close(o.detach(), o_ref, rtol=1e-5, atol=1e-5)
            for i, r in enumerate((dq, dk, dv)):
                torch.testing.assert_close(dqkv[:, :, i], r, rtol=1e-5, atol=1e-5)
            qq, kv = q.clone().requires_grad_(), qkv[:, :, 1:].clone().requires_grad_()
            o = flash_attn_kvpacked_func(qq, kv, bias, causal)
            gq, gkv = torch.autograd.grad(o, (qq, kv), do)
            torch.testing.assert_close(gq, dq, rtol=1e-5, atol=1e-5)
            torch.testing.assert_close(gkv[:, :, 0], dk, rtol=1e-5, atol=1e-5)
            torch.testing.assert_close(gkv[:, :, 1], dv, rtol=1e-5, atol=1e-5)
        # Sq != Sk, causal aligned bottom-right (the last query sees every key), no bias
        qs, dos = q[:, 30:].contiguous(), do[:, 30:].contiguous()
        o_ref, _, dq, dk, dv = orc.dense_attention_bwd(qs, k, v, dos, None, True)
        qq, kk, vv = (t.clone().requires_grad_() for t in (qs, k, v))
        o = flash_attn_func(qq, kk, vv, None, True)
        g = torch.autograd.grad(o, (qq, kk, vv), dos)
        torch.testing.assert_close(o.detach(), o_ref, rtol=1e-5, atol=1e-5)
        for a, r in zip(g, (dq, dk, dv)):
            torch.testing.assert_close(a, r, rtol=1e-5, atol=1e-5)
        with pytest.raises(NotImplementedError):
            flash_attn_func(qq, kk, vv, torch.zeros(1, H, 40, S), False)
    finally:
        chunk_ops._set_ops_for_testing(None)
