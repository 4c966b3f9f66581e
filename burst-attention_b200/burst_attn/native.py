"""ctypes binding of the C-ABI library ``libburst_attn_b200.so`` (include/burst_attn_b200.h).

There is deliberately NO fallback: if the shared library is missing or the
device is not sm_100, every entry point raises.  PyTorch is used only for
device memory and streams (``tensor.data_ptr()``, ``torch.cuda.current_stream()``).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# BA_LIB_PATH: developer hook to A/B a differently built library (tools/); default is the in-tree build
LIB_PATH = os.environ.get("BA_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "lib", "libburst_attn_b200.so")

BA_DTYPE_FP16, BA_DTYPE_BF16 = 0, 1
BA_MASK_NONE, BA_MASK_CAUSAL = 0, 1
BA_FWD_FIRST, BA_FWD_LAST = 1, 2
NCCL_UNIQUE_ID_BYTES = 128
IPC_HANDLE_BYTES = 64


class NativeLibraryError(RuntimeError):
    pass


class ba_tensor4(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("stride_b", ctypes.c_int64),
                ("stride_s", ctypes.c_int64), ("stride_h", ctypes.c_int64)]


class ba_rowstat(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("stride_b", ctypes.c_int64), ("stride_h", ctypes.c_int64)]


_lib = None

_EXPORTS = (
    "ba_last_error", "ba_device_check", "ba_version", "ba_fwd_chunk", "ba_fwd_chunk_bias", "ba_bwd_delta", "ba_bwd_chunk",
    "ba_bwd_chunk_bias",
    "ba_cast_from_f32", "ba_accumulate_f32", "ba_ring_unique_id", "ba_ring_create", "ba_ring_post",
    "ba_ring_wait", "ba_ring_rank", "ba_ring_world", "ba_ring_destroy", "ba_ring_arena_create",
    "ba_ring_arena_connect",
)
SELFTEST_LIB_PATH = os.path.join(os.path.dirname(LIB_PATH), "libburst_attn_b200_selftest.so")
_SELFTEST_EXPORTS = ("ba_selftest_last_error", "ba_selftest", "ba_ubench")


def exported_symbols() -> Sequence[str]:
    """Every symbol include/burst_attn_b200.h declares (checked by the CPU tests)."""
    return _EXPORTS


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C burst-attention_b200/csrc`).  There is no CPU/PyTorch fallback.")
    L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    i, f, vp = ctypes.c_int, ctypes.c_float, ctypes.c_void_p
    L.ba_last_error.restype = ctypes.c_char_p
    L.ba_last_error.argtypes = []
    L.ba_device_check.restype = i
    L.ba_version.restype = i
    L.ba_fwd_chunk.restype = i
    L.ba_fwd_chunk.argtypes = [ba_tensor4, ba_tensor4, ba_tensor4, ba_tensor4, ba_rowstat, ba_tensor4,
                               i, i, i, i, i, f, i, i, i, i, vp]
    L.ba_fwd_chunk_bias.restype = i
    L.ba_fwd_chunk_bias.argtypes = [ba_tensor4, ba_tensor4, ba_tensor4, ba_rowstat, ba_tensor4, ba_rowstat, ba_tensor4,
                                    i, i, i, i, i, f, i, i, i, i, vp]
    L.ba_bwd_chunk_bias.restype = i
    L.ba_bwd_chunk_bias.argtypes = [ba_tensor4, ba_tensor4, ba_tensor4, ba_tensor4, ba_rowstat, ba_rowstat, ba_rowstat,
                                    ba_tensor4, ba_tensor4, ba_tensor4, i, i, i, i, i, f, i, i, i, i, vp]
    L.ba_bwd_delta.restype = i
    L.ba_bwd_delta.argtypes = [ba_tensor4, ba_tensor4, ba_rowstat, i, i, i, i, i, vp]
    L.ba_bwd_chunk.restype = i
    L.ba_bwd_chunk.argtypes = [ba_tensor4, ba_tensor4, ba_tensor4, ba_tensor4, ba_rowstat, ba_rowstat,
                               ba_tensor4, ba_tensor4, ba_tensor4, i, i, i, i, i, f, i, i, i, i, vp]
    L.ba_cast_from_f32.restype = i
    L.ba_cast_from_f32.argtypes = [ba_tensor4, ba_tensor4, i, i, i, i, i, vp]
    L.ba_accumulate_f32.restype = i
    L.ba_accumulate_f32.argtypes = [ba_tensor4, ba_tensor4, i, i, i, i, vp]
    L.ba_ring_unique_id.restype = i
    L.ba_ring_unique_id.argtypes = [vp]
    L.ba_ring_create.restype = i
    L.ba_ring_create.argtypes = [vp, i, i, ctypes.POINTER(vp)]
    L.ba_ring_post.restype = i
    L.ba_ring_post.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_int64), i, vp]
    L.ba_ring_wait.restype = i
    L.ba_ring_wait.argtypes = [vp, vp]
    L.ba_ring_rank.restype = i
    L.ba_ring_rank.argtypes = [vp]
    L.ba_ring_world.restype = i
    L.ba_ring_world.argtypes = [vp]
    L.ba_ring_destroy.restype = i
    L.ba_ring_destroy.argtypes = [vp]
    L.ba_ring_arena_create.restype = i
    L.ba_ring_arena_create.argtypes = [vp, ctypes.c_int64, ctypes.POINTER(vp), vp]
    L.ba_ring_arena_connect.restype = i
    L.ba_ring_arena_connect.argtypes = [vp, vp, vp]
    _lib = L
    return L


_selftest_lib = None


def selftest_lib() -> ctypes.CDLL:
    """The diagnostics library (include/burst_attn_b200_selftest.h); loaded by tests/ and tools/ only."""
    global _selftest_lib
    if _selftest_lib is None:
        if not os.path.exists(SELFTEST_LIB_PATH):
            raise NativeLibraryError(f"{SELFTEST_LIB_PATH} not found: build with __graft_entry__.build()")
        L = ctypes.CDLL(SELFTEST_LIB_PATH)
        i, vp = ctypes.c_int, ctypes.c_void_p
        L.ba_selftest_last_error.restype = ctypes.c_char_p
        L.ba_ubench.restype = i
        L.ba_ubench.argtypes = [i, i, i, ctypes.POINTER(ctypes.c_int64), vp]
        L.ba_selftest.restype = i
        L.ba_selftest.argtypes = [i, vp, vp, vp, i, vp]
        _selftest_lib = L
    return _selftest_lib


def check_selftest(rc: int, what: str) -> None:
    if rc != 0:
        msg = selftest_lib().ba_selftest_last_error()
        raise NativeLibraryError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().ba_last_error()
        raise NativeLibraryError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.bfloat16:
        return BA_DTYPE_BF16
    if dt == torch.float16:
        return BA_DTYPE_FP16
    raise TypeError(f"burst_attn_b200 supports float16/bfloat16 inputs, got {dt}")


def t4(t: Optional[torch.Tensor], seq_dim: int) -> ba_tensor4:
    """[b,s,h,d] view descriptor of a 4-D tensor whose sequence axis is ``seq_dim``
    (1 for the flash layout [B,S,H,D], 2 for the normal layout [B,H,S,D])."""
    if t is None:
        return ba_tensor4(None, 0, 0, 0)
    assert t.dim() == 4 and t.stride(3) == 1, "last (head_dim) axis must be contiguous"
    return ba_tensor4(t.data_ptr(), t.stride(0), t.stride(seq_dim), t.stride(3 - seq_dim))


def rs(t: Optional[torch.Tensor]) -> ba_rowstat:
    """[B,H,S] fp32 row statistic (lse / delta / key bias); S contiguous.  None -> null view."""
    if t is None:
        return ba_rowstat(None, 0, 0)
    assert t.dim() == 3 and t.stride(2) == 1 and t.dtype == torch.float32
    return ba_rowstat(t.data_ptr(), t.stride(0), t.stride(1))


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream
