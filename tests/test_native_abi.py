"""CPU-side checks of the C-ABI boundary: the shared library loads, exports every
symbol include/burst_attn_b200.h declares, and rejects bad arguments with an
error string instead of crashing.  No compute call is made (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "burst_attn_b200.h")


def _declared():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"\b(ba_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def nat():
    from burst_attn import native
    if not os.path.exists(native.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return native


def test_header_and_binding_agree(nat):
    assert _declared() == sorted(nat.exported_symbols())


def test_library_exports_every_declared_symbol(nat):
    L = ctypes.CDLL(nat.LIB_PATH)
    for name in _declared():
        assert hasattr(L, name), name


def test_selftest_library_is_separate_from_the_product(nat):
    """Diagnostics (ba_selftest, ba_ubench) live in their own header and library, not in the drop-in boundary."""
    hdr = open(os.path.join(ROOT, "include", "burst_attn_b200_selftest.h")).read()
    declared = sorted(set(re.findall(r"\b(ba_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(nat._SELFTEST_EXPORTS)
    T = ctypes.CDLL(nat.SELFTEST_LIB_PATH)
    for name in declared:
        assert hasattr(T, name), name
    L = ctypes.CDLL(nat.LIB_PATH)
    for name in ("ba_selftest", "ba_ubench"):
        assert not hasattr(L, name), f"{name} must not be exported by the product library"


def test_bad_arguments_return_error_string(nat):
    L = nat.lib()
    z4 = nat.ba_tensor4(None, 0, 0, 0)
    zr = nat.ba_rowstat(None, 0, 0)
    rc = L.ba_fwd_chunk(z4, z4, z4, z4, zr, z4, 1, 128, 128, 1, 96, 1.0, 0, 0, 3, 1, None)
    assert rc != 0 and b"head dim" in L.ba_last_error()
    rc = L.ba_fwd_chunk(z4, z4, z4, z4, zr, z4, 1, 128, 128, 1, 128, 1.0, 0, 0, 3, 1, None)
    assert rc != 0 and b"null" in L.ba_last_error()
    rc = L.ba_ring_post(None, None, None, None, 0, None)
    assert rc != 0
    rc = L.ba_ring_arena_create(None, 1 << 20, None, None)
    assert rc != 0 and b"ba_ring_arena_create" in L.ba_last_error()
    rc = L.ba_ring_arena_connect(None, None, None)
    assert rc != 0 and b"ba_ring_arena_connect" in L.ba_last_error()
    assert L.ba_version() >= 200


def test_missing_library_fails_loudly(nat, monkeypatch):
    monkeypatch.setattr(nat, "_lib", None)
    monkeypatch.setattr(nat, "LIB_PATH", "/nonexistent/libburst_attn_b200.so")
    with pytest.raises(nat.NativeLibraryError):
        nat.lib()


def test_receive_arena_bump_allocator_is_symmetric_and_aligned():
    """Host logic of the copy-engine transport's receive arena (burst_attn/comm.py): buffers are carved in
    call order at 1 KiB-aligned offsets, so equal call sequences give equal offsets on every rank, and the
    size announced by Ring.begin covers them."""
    import torch
    from burst_attn import comm

    ring = comm._NativeRing.__new__(comm._NativeRing)  # no library / device needed for the carving logic
    ring.ce = True
    sizes = [3 * 5 * 7 * 2, 1024, 4 * 33]
    ring.arena = torch.zeros(sum(comm._align(n) for n in sizes), dtype=torch.uint8)
    ring.arena_off = 0
    a = ring.empty((3, 5, 7), torch.bfloat16)
    b = ring.empty((256,), torch.float32)
    c = ring.empty((33,), torch.float32)
    base = ring.arena.data_ptr()
    assert [t.data_ptr() - base for t in (a, b, c)] == [0, 1024, 2048]
    assert a.shape == (3, 5, 7) and a.dtype == torch.bfloat16 and c.is_contiguous()
    with pytest.raises(AssertionError):
        ring.empty((1,), torch.float32)  # beyond what begin() announced
