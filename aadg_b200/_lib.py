"""ctypes binding of libaadg_b200.so (include/aadg_b200.h).  No fallback: a missing library raises."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libaadg_b200.so")

_lib = None

c_void_p, c_int, c_size_t, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_float
c_i64 = ctypes.c_int64

ROOT = os.path.dirname(HERE)
HEADER = os.path.join(ROOT, "include", "aadg_b200.h")


def _ctype(decl):
    """C parameter / return type -> ctypes type (pointers are passed as integers / None)."""
    d = decl.replace("const", " ").strip()
    if "*" in d:
        return ctypes.c_char_p if d.replace(" ", "") == "char*" else c_void_p
    d = " ".join(d.split())
    return {"int": c_int, "float": c_float, "size_t": c_size_t, "long long": ctypes.c_longlong,
            "unsigned long long": ctypes.c_ulonglong, "double": ctypes.c_double, "void": None}[d]


def _parse_header(path=HEADER):
    """{name: (restype, [argtypes])} for every aadg_* prototype in include/aadg_b200.h."""
    import re
    text = re.sub(r"/\*.*?\*/", " ", open(path).read(), flags=re.S)
    text = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith("#"))
    sigs = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(aadg_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        ret, name, params = m.group(1), m.group(2), m.group(3)
        args = []
        for prm in params.split(","):
            prm = prm.strip()
            if prm in ("void", ""):
                continue
            prm = re.sub(r"\[.*\]", "*", prm)
            # drop the parameter name (last identifier) unless the declaration is a bare type
            mm = re.match(r"(.*?[\*\s])([A-Za-z_]\w*)$", prm)
            args.append(_ctype(mm.group(1) if mm else prm))
        sigs[name] = (_ctype(ret), args)
    return sigs


# name -> (restype, argtypes); every symbol include/aadg_b200.h declares
SIGNATURES = _parse_header()


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


CALLS = 0   # C-ABI compute calls made so far (every one launches at least one kernel of this library)


def check(rc):
    global CALLS
    CALLS += 1
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
    """raw cudaStream_t of torch's current stream on the current device (the fast accessor: this is called once
    per C-ABI launch, ~500 times per step)"""
    import torch
    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:            # older / newer torch without the private accessor
        return torch.cuda.current_stream().cuda_stream


class on_device:
    """`with on_device(t.device):` -- torch.cuda.device() only when the tensor lives on another device than the
    current one (entering the context costs ~10 us; the engine always runs on the current device)"""

    def __init__(self, device):
        import torch
        self.ctx = None
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx != torch.cuda.current_device():
            self.ctx = torch.cuda.device(idx)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)


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
