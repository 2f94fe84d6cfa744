"""
ctypes binding of libfbpinn_b200.so (the C ABI declared in include/fbpinn_b200.h).

There is deliberately no fallback: if the shared library is missing, cannot be loaded, or a call fails, a
`FbpError` is raised.  PyTorch is used only for device memory and streams; every pointer handed to the library
is `tensor.data_ptr()` of a contiguous CUDA tensor and the stream is torch's current stream.
"""

import ctypes as C
import os

FBP_MAX_LAYERS = 16
FBP_MAX_XD = 3
FBP_MAX_UD = 4
FBP_MAX_COMP = 10

_HERE = os.path.dirname(os.path.abspath(__file__))
# FBP_LIB selects a differently-tuned build of the same library (A/B experiments); the default is the product build
LIB_PATH = os.environ.get("FBP_LIB") or os.path.join(_HERE, "csrc", "libfbpinn_b200.so")


class FbpError(RuntimeError):
    pass


class PlanDesc(C.Structure):
    _fields_ = [("xd", C.c_int32), ("ud", C.c_int32), ("n_layers", C.c_int32),
                ("layer_sizes", C.c_int32 * (FBP_MAX_LAYERS + 1)),
                ("activation", C.c_int32), ("window", C.c_int32), ("n_comp", C.c_int32),
                ("comp_k", C.c_int32 * FBP_MAX_COMP), ("comp_l", C.c_int32 * FBP_MAX_COMP)]


FBP_HALO_MAX_WORLD = 8


class HaloPeers(C.Structure):
    "fbp_halo_peers of include/fbpinn_b200.h"
    _fields_ = [("data", C.c_void_p * FBP_HALO_MAX_WORLD), ("flags", C.c_void_p * FBP_HALO_MAX_WORLD)]


class TakesView(C.Structure):
    _fields_ = [("n", C.c_int64), ("s", C.c_int64), ("q", C.c_int64), ("s_active", C.c_int64),
                ("m_all", C.c_int32), ("m_active", C.c_int32), ("npou", C.c_int32),
                ("d_m_take", C.c_void_p), ("d_np_take", C.c_void_p), ("d_sub_ids", C.c_void_p),
                ("d_sub_off", C.c_void_p), ("d_spair_point", C.c_void_p), ("d_spair_row", C.c_void_p),
                ("d_spair_sub", C.c_void_p), ("d_pos", C.c_void_p), ("d_row_off", C.c_void_p),
                ("d_pt_row_off", C.c_void_p), ("d_items", C.c_void_p), ("d_sub_item_off", C.c_void_p),
                ("d_item_order_fwd", C.c_void_p), ("d_item_order_bwd", C.c_void_p),
                ("n_items", C.c_int32), ("n_items_active", C.c_int32),
                ("d_launch_fwd", C.c_void_p), ("d_launch_bwd", C.c_void_p)]


_P = C.c_void_p
_I32, _I64, _F = C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); every symbol include/fbpinn_b200.h declares
SIGNATURES = {
    "fbp_last_error": (C.c_char_p, []),
    "fbp_version": (C.c_int, []),
    "fbp_launch_count": (_I64, []),
    "fbp_device_info": (C.c_int, [C.POINTER(_I32), C.POINTER(_I32), C.POINTER(_I32), C.POINTER(_I64)]),
    "fbp_plan_create": (C.c_int, [C.POINTER(_P), C.POINTER(PlanDesc)]),
    "fbp_plan_destroy": (C.c_int, [_P]),
    "fbp_plan_param_count": (_I64, [_P]),
    "fbp_plan_is_fast": (_I32, [_P]),
    "fbp_plan_tile_points": (_I32, [_P]),
    "fbp_plan_set_kernel": (C.c_int, [_P, _I32]),
    "fbp_plan_has_tensor": (_I32, [_P]),
    "fbp_plan_forward_family": (_I32, [_P]),
    "fbp_plan_reverse_family": (_I32, [_P]),
    "fbp_plan_scratch_per_pair": (_I64, [_P]),
    "fbp_plan_cache_per_pair": (_I64, [_P]),
    "fbp_pack_params": (C.c_int, [_P, _I64, C.POINTER(_P), C.POINTER(_P), _P, _P]),
    "fbp_unpack_params": (C.c_int, [_P, _I64, _P, C.POINTER(_P), C.POINTER(_P), _P]),
    "fbp_pack_extra": (C.c_int, [_P, _I64, _I32, _I32, _P, _P, _I32, _P]),
    "fbp_plan_n_extra": (_I32, [_P]),
    "fbp_inside_count": (C.c_int, [_P, _I64, _I32, _P, _I32, _P, _I32, _P, _P, _P]),
    "fbp_nonzero_i32": (C.c_int, [_P, _I64, _P, C.POINTER(_I64), _P]),
    "fbp_gather_rows": (C.c_int, [_P, _P, _I64, _I32, _P, _P]),
    "fbp_takes_begin": (C.c_int, [C.POINTER(_P), _P, _I64, _I32, _P, _I32, _P, _P, _I32, _P,
                                  C.POINTER(_I64), C.POINTER(_I64)]),
    "fbp_takes_emit": (C.c_int, [_P] + [_P] * 11 + [_P]),
    "fbp_takes_destroy": (C.c_int, [_P]),
    "fbp_window_sums": (C.c_int, [_P, C.POINTER(TakesView), _P, _P, _P, _P]),
    "fbp_forward": (C.c_int, [_P, C.POINTER(TakesView), _P, _P, _P, _P, _P, _I64, _P, _P]),
    "fbp_reduce_forward": (C.c_int, [_P, C.POINTER(TakesView), _P, _P, _P, _P, _P]),
    "fbp_reduce_backward": (C.c_int, [_P, C.POINTER(TakesView), _P, _P, _P, _P, _P]),
    "fbp_row_sums": (C.c_int, [_P, C.POINTER(TakesView), _P, _P, _P]),
    "fbp_reduce_rows_forward": (C.c_int, [_P, C.POINTER(TakesView), _P, _P, _P, _P, _P, _P]),
    "fbp_backward_workspace_floats": (_I64, [_P, C.POINTER(TakesView)]),
    "fbp_backward": (C.c_int, [_P, C.POINTER(TakesView), _P, _P, _P, _P, _P, _I32, _P, _P, _I64, _P, _P]),
    "fbp_adam_step": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I64, _P, _I32, _F, _F, _F, _F, _F, _P]),
    "fbp_fma_peak": (C.c_int, [_I32, C.POINTER(_F), _P]),
    "fbp_ffma2_peak": (C.c_int, [_I32, C.POINTER(_F), _P]),
    "fbp_halo_push": (C.c_int, [_P, _I32, _P, _P, _I32, _P, _P, _I32, _I32, _P, _P, _P]),
    "fbp_halo_pull": (C.c_int, [_P, _I32, _P, _P, _P, _I32, _P, C.c_int64, C.c_uint32, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "fbp_scatter_rows": (C.c_int, [_P, _P, C.c_int64, _I32, C.c_float, _P, _P]),
    "fbp_tc_selftest": (C.c_int, [_P, _P, _P, _I32, _P]),
    "fbp_tc_selftest_g": (C.c_int, [_P, _P, _P, _I32, _P]),
}

_lib = None


def load():
    """Load the shared library (once). Raises FbpError if it is absent — there is no other code path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FbpError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       f"or `make -C fbpinns_b200/csrc` (sm_100a CUDA extension; there is no CPU fallback)")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:
        raise FbpError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise FbpError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().fbp_last_error()
        raise FbpError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL). The tensor must be contiguous, on CUDA, float32 or int32
    (the only element types the C ABI knows)."""
    if t is None:
        return None
    import torch
    if t.dtype not in (torch.float32, torch.int32):
        raise FbpError(f"libfbpinn_b200 takes float32 / int32 buffers only, got {t.dtype}")
    if not t.is_cuda:
        raise FbpError("libfbpinn_b200 needs CUDA tensors (no CPU path exists)")
    if not t.is_contiguous():
        raise FbpError("non-contiguous tensor passed to libfbpinn_b200")
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
