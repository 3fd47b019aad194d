"""ctypes binding of libaadg_b200.so (include/aadg_b200.h).  No fallback: a missing library raises."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libaadg_b200.so")

_lib = None

c_void_p, c_int, c_size_t, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_float
c_i64 = ctypes.c_int64

# name -> (restype, argtypes); every symbol include/aadg_b200.h declares
SIGNATURES = {
    "aadg_version": (c_int, []),
    "aadg_last_error": (ctypes.c_char_p, []),
    "aadg_u8_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "aadg_u8_apply_policy": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                      c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "aadg_u8_policy_normalize": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                          c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "aadg_sinkhorn_small_max_points": (c_int, []),
    "aadg_sinkhorn_small_batched": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "aadg_sinkhorn_rewards_workspace_bytes": (c_size_t, [c_int, c_int]),
    "aadg_sinkhorn_diversity_rewards": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                                 c_void_p, c_void_p, c_size_t, c_void_p]),
    "aadg_sinkhorn_large_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "aadg_sinkhorn_large": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p,
                                     c_void_p, c_size_t, c_void_p]),
    "aadg_conv_fprop_bf16": (c_int, [c_void_p] + [c_int] * 5 + [c_void_p] + [c_int] * 6 + [c_void_p] + [c_int] * 5 + [c_void_p]),
    "aadg_conv_dgrad_bf16": (c_int, [c_void_p] + [c_int] * 5 + [c_void_p] + [c_int] * 6 + [c_void_p] + [c_int] * 5 + [c_void_p]),
    "aadg_conv_wgrad_bf16": (c_int, [c_void_p] + [c_int] * 5 + [c_void_p] + [c_int] * 9 + [c_void_p, c_void_p]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libaadg_b200.so is not built (%s); run `python -m aadg_b200.build` — there is no "
                "CPU or PyTorch fallback for the hot path" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("libaadg_b200: error %d: %s" % (rc, lib().aadg_last_error().decode()))


def ptr(t):
    """device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


_workspaces = {}


def workspace(nbytes, device):
    """A cached, growing uint8 CUDA buffer per (device, stream) (the ABI never allocates)."""
    import torch
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf
