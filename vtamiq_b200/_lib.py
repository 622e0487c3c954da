"""ctypes binding of libvtamiq_b200.so (the C-ABI declared in include/vtamiq_b200.h).

There is deliberately no fallback: if the shared library is missing, or no sm_100 GPU is visible,
the product path raises.  Nothing here touches ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# VTQ_LIBRARY overrides the path (A/B runs of two builds of the same ABI); the default is the in-tree build
LIB_PATH = os.environ.get("VTQ_LIBRARY") or os.path.join(_HERE, "libvtamiq_b200.so")

VTQ_F16, VTQ_BF16 = 0, 1
EPI_BIAS_H, EPI_BIAS_GELU_H, EPI_BIAS_F32, EPI_BIAS_RESID_F32 = 0, 1, 2, 3
ABI_VERSION = 5

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol include/vtamiq_b200.h declares
SIGNATURES = {
    "vtq_abi_version": (_i, []),
    "vtq_create": (_i, [C.POINTER(_vp), _i]),
    "vtq_destroy": (_i, [_vp]),
    "vtq_last_error_string": (C.c_char_p, [_vp]),
    "vtq_launch_count": (C.c_ulonglong, [_vp]),
    "vtq_set_reverse": (_i, [_vp, _i]),
    "vtq_workspace_bytes": (_i64, [_vp, _i, _i]),
    "vtq_patch_gather": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp]),
    "vtq_patch_gather_u8": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp]),
    "vtq_avgpool2x2_u8": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "vtq_coord_status": (_i, [_vp, _i]),
    "vtq_tensor_map_stats": (_i, [_vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "vtq_normalize_u8": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "vtq_sample_grid": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, C.c_double, _vp, _vp]),
    "vtq_avgpool2x2": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "vtq_cast_rows": (_i, [_vp, _vp, _vp, _i64, _i, _vp]),
    "vtq_embed_assemble": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "vtq_layernorm": (_i, [_vp, _vp, _i64, _vp, _vp, _f, _i64, _i, _vp, _i, _vp]),
    "vtq_gemm": (_i, [_vp, _vp, _i64, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i64, _vp, _vp]),
    "vtq_gemm_ln": (_i, [_vp, _vp, _i64, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i64, _vp, _vp, _i, _vp, _f, _vp, _vp, _vp]),
    "vtq_gemm_ln_slots": (_i, [_i]),
    "vtq_rowstats_cast": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _i, _vp]),
    "vtq_attention_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "vtq_attention_fwd_trace": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "vtq_cls_diff": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp]),
    "vtq_diffnet_head": (_i, [_vp, _vp, C.POINTER(_vp), _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "vtq_tail_saved_floats": (_i64, [_i, _i, _i, _i, _i, _i]),
    "vtq_tail_bwd_workspace_bytes": (_i64, [_i, _i]),
    "vtq_tail_train_fwd": (_i, [_vp, _vp, _vp, C.POINTER(_vp), _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "vtq_tail_bwd": (_i, [_vp, _vp, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp), _i, _i, _i, _i, _i, _i, _i, _vp, _vp,
                          _vp, _vp, _vp, _vp]),
}

_lock = threading.Lock()
_lib = None


class VtqError(RuntimeError):
    """A C-ABI call returned a negative status."""


def load_library() -> C.CDLL:
    """dlopen the in-tree library and bind every declared symbol.  Raises if it is absent."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise VtqError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(vtamiq_b200 has no CPU or PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch
            fn.restype, fn.argtypes = res, args
        if lib.vtq_abi_version() != ABI_VERSION:
            raise VtqError("libvtamiq_b200.so ABI version mismatch; rebuild")
        _lib = lib
        return lib


class Context:
    """Owns one vtq_ctx (per device).  ``call`` raises VtqError with the library's message."""

    def __init__(self, device: int):
        self.lib = load_library()
        self.handle = _vp()
        rc = self.lib.vtq_create(C.byref(self.handle), int(device))
        if rc != 0:
            msg = self.lib.vtq_last_error_string(None)
            raise VtqError(f"vtq_create failed ({rc}): {msg.decode() if msg else '?'}")
        self.device = device

    def call(self, name: str, *args):
        rc = getattr(self.lib, name)(self.handle, *args)
        if rc != 0:
            msg = self.lib.vtq_last_error_string(self.handle)
            raise VtqError(f"{name} failed ({rc}): {msg.decode() if msg else '?'}")

    def launch_count(self) -> int:
        return int(self.lib.vtq_launch_count(self.handle))

    def coord_status(self, reset: bool = False) -> int:
        """1 if a finished gather launch saw an out-of-range patch origin since the last reset (no sync needed)."""
        return int(self.lib.vtq_coord_status(self.handle, 1 if reset else 0))

    def tensor_map_stats(self) -> tuple[int, int]:
        h, m = C.c_ulonglong(0), C.c_ulonglong(0)
        self.lib.vtq_tensor_map_stats(self.handle, C.byref(h), C.byref(m))
        return int(h.value), int(m.value)

    def workspace_bytes(self, B: int, hidden: int) -> int:
        return int(self.lib.vtq_workspace_bytes(self.handle, B, hidden))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.vtq_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


_contexts: dict[int, Context] = {}


def get_context(device: int) -> Context:
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = Context(device)
    return ctx
