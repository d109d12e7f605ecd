"""ctypes binding of libbevfront_b200.so (the C ABI declared in include/bevfront_b200.h).

There is no CPU fallback: if the library is missing and cannot be built, or a call fails, an exception is
raised.  PyTorch is used by the callers only for device memory and streams.
"""
import ctypes
import os

from . import build as _build

_lib = None


class BevfError(RuntimeError):
    pass


def _declare(lib):
    c_int, c_void_p, c_size_t = ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t
    lib.bevf_last_error.restype = ctypes.c_char_p
    for name in ("bevf_hard_voxelize_workspace_bytes", "bevf_dynamic_scatter_workspace_bytes",
                 "bevf_bev_pool_workspace_bytes", "bevf_spconv_index_bytes", "bevf_pack_sparse_rows_bytes"):
        if hasattr(lib, name):
            getattr(lib, name).restype = c_size_t
    lib.bevf_spconv_index_error_flag.restype = c_void_p
    lib.bevf_launch_count.restype = ctypes.c_longlong
    _ = c_int
    return lib


def lib():
    """Load (building first if the sources are newer) libbevfront_b200.so.  Fails loudly."""
    global _lib
    if _lib is None:
        path = os.environ.get("BEVFRONT_LIB") or _build.LIB   # BEVFRONT_LIB: an instrumented debug build of the same ABI
        if path == _build.LIB and _build.is_stale():
            if os.environ.get("BEVFRONT_NO_BUILD") and not os.path.exists(path):
                raise BevfError(f"{path} is missing and BEVFRONT_NO_BUILD is set")
            try:
                path = _build.build()
            except Exception as e:  # stale-but-present library on a box without nvcc is still usable
                if not os.path.exists(path):
                    raise BevfError(f"libbevfront_b200.so is missing and could not be built: {e}") from e
        _lib = _declare(ctypes.CDLL(path))
    return _lib


def check(rc):
    if rc != 0:
        raise BevfError(f"libbevfront_b200 error {rc}: {lib().bevf_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def cur_stream(device=None):
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def f32_array(vals):
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])


def i32_array(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])
