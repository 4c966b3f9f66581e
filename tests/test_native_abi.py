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


def test_bad_arguments_return_error_string(nat):
    L = nat.lib()
    z4 = nat.ba_tensor4(None, 0, 0, 0)
    zr = nat.ba_rowstat(None, 0, 0)
    rc = L.ba_fwd_chunk(z4, z4, z4, z4, zr, z4, 1, 128, 128, 1, 64, 1.0, 0, 0, 3, 1, None)
    assert rc != 0 and b"head dim" in L.ba_last_error()
    rc = L.ba_fwd_chunk(z4, z4, z4, z4, zr, z4, 1, 128, 128, 1, 128, 1.0, 0, 0, 3, 1, None)
    assert rc != 0 and b"null" in L.ba_last_error()
    rc = L.ba_ring_post(None, None, None, None, 0, None)
    assert rc != 0
    assert L.ba_version() >= 100


def test_missing_library_fails_loudly(nat, monkeypatch):
    monkeypatch.setattr(nat, "_lib", None)
    monkeypatch.setattr(nat, "LIB_PATH", "/nonexistent/libburst_attn_b200.so")
    with pytest.raises(nat.NativeLibraryError):
        nat.lib()
