"""ctypes binding of csrc/libnerf_b200.so (C ABI declared in include/nerf_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing, or a call is made with
tensors that are not on a CUDA device, the call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# NERF_B200_LIB selects another build of the same library (kernel-variant sweeps); never a fallback
LIB_PATH = os.environ.get("NERF_B200_LIB") or os.path.join(_HERE, "csrc", "libnerf_b200.so")

N_PARAMS = 595844
PREC_BF16, PREC_FP32 = 0, 1
TRAIN_STATE_DOUBLES = 96
PEER_MAX = 8                  # NERF_PEER_MAX (include/nerf_b200.h)
BWD_ALL, BWD_DGRAD, BWD_WGRAD = 0, 1, 2
FWD_DENSITY_ONLY = 2

_P = c_void_p
# name -> (restype, argtypes); mirrors include/nerf_b200.h line by line
SIGNATURES = {
    "nerf_device_info": (c_int, [ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "nerf_last_error": (c_char_p, []),
    "nerf_version": (c_char_p, []),
    "nerf_launch_count": (ctypes.c_ulonglong, []),
    "nerf_packed_weight_bytes": (c_size_t, []),
    "nerf_pack_weights": (c_int, [_P, _P, _P]),
    "nerf_positional_encoding": (c_int, [_P, c_int64, c_int, _P, c_int, c_int, _P, _P]),
    "nerf_stratified_z": (c_int, [_P, _P, c_int, c_int, c_float, c_float, _P, _P]),
    "nerf_mlp_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "nerf_mlp_fwd_rays": (c_int, [_P, _P, _P, c_int, c_int, c_float, _P, _P, _P, _P, c_size_t, c_int, c_int, _P]),
    "nerf_mlp_fwd_encoded": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, c_size_t, c_int, c_int, _P]),
    "nerf_mlp_bwd": (c_int, [_P, c_int64, c_int, _P, _P, _P, _P, c_size_t, c_int, _P]),
    "nerf_mlp_bwd_stage": (c_int, [_P, c_int64, c_int, _P, _P, _P, _P, c_size_t, c_int, c_int, _P]),
    "nerf_composite_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "nerf_composite_bwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "nerf_sample_pdf": (c_int, [_P, c_int64, _P, c_int64, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P]),
    "nerf_resample_merge": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "nerf_adam_step": (c_int, [_P, _P, _P, _P, c_int64, c_double, c_double, c_double, c_double, c_int64, c_float, _P]),
    "nerf_mse_loss": (c_int, [_P, _P, c_int64, _P, _P, _P]),
    "nerf_train_prepare": (c_int, [_P, _P, _P, c_int64, _P]),
    "nerf_adam_step_dev": (c_int, [_P, _P, _P, _P, c_int64, _P, _P]),
    "nerf_adam_fused_scratch_bytes": (c_size_t, [c_int64]),
    "nerf_adam_step_fused": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, _P]),
    "nerf_adam_step_fused_peer": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, _P, c_int64, _P, _P, _P, _P]),
    "nerf_composite_train_scratch_bytes": (c_size_t, [c_int]),
    "nerf_composite_train": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "nerf_generate_rays": (c_int, [_P, c_int, c_int, c_int, c_double, _P, c_int64, c_int64, _P, _P, _P, _P, c_int, _P, _P]),
    "nerf_postprocess_rgb8": (c_int, [_P, c_int64, c_float, c_int, _P, _P]),
}

_dll = None


def bwd_fused():
    """True unless NERF_BWD_FUSED=0 selects the two-kernel backward (the library reads the same variable)."""
    return os.environ.get("NERF_BWD_FUSED", "1")[:1] != "0"


def dll():
    """Load the shared library once.  Fails loudly when it has not been built."""
    global _dll
    if _dll is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C nerf_mlp_b200/csrc`).  There is no CPU/PyTorch fallback.")
        d = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(d, name)
            fn.restype, fn.argtypes = res, args
        _dll = d
    return _dll


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = dll().nerf_last_error().decode(errors="replace")
        raise RuntimeError(f"libnerf_b200 {what} failed (rc={rc}): {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr(device=None):
    """Current torch stream of `device` as a cudaStream_t.  The library launches on the CALLING THREAD's current
    device (its per-device kernel attributes and SM count follow cudaGetDevice), so tensors on another device must be
    used under ``torch.cuda.device(...)``: refused here with a clear error instead of a failed or mis-sized launch."""
    if device is not None:
        idx = torch.device(device).index
        if idx is not None and idx != torch.cuda.current_device():
            raise RuntimeError(f"nerf_mlp_b200: tensors are on cuda:{idx} but the current device is "
                               f"cuda:{torch.cuda.current_device()}; wrap the call in `with torch.cuda.device({idx}):`")
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors, dtype=torch.float32):
    """No-fallback guard: every tensor must be a contiguous CUDA tensor of `dtype`."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("nerf_mlp_b200 runs on CUDA (sm_100a) only; got a tensor on "
                               f"'{t.device}'.  There is no CPU fallback.")
        if t.dtype != dtype:
            raise RuntimeError(f"expected dtype {dtype}, got {t.dtype}")
        if not t.is_contiguous():
            raise RuntimeError("expected a contiguous tensor")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {t.device} vs {dev}")
    return dev


def f32c(t, device=None):
    """float32 + contiguous view/copy of a CUDA tensor (errors for CPU tensors)."""
    if not t.is_cuda:
        raise RuntimeError("nerf_mlp_b200 runs on CUDA (sm_100a) only; got a tensor on "
                           f"'{t.device}'.  There is no CPU fallback.")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
