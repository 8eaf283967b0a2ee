"""ctypes binding of the C-ABI library (include/sidlsg.h -> sid_lsg_b200/_C/libsidlsg.so).

There is NO fallback: if the library is missing or a call fails, this raises.  Signatures are parsed from the
header so the header stays the single source of truth.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "sidlsg.h")
LIB_PATH = os.path.join(_HERE, "_C", "libsidlsg.so")

F32, BF16 = 0, 1
_DT = {torch.float32: F32, torch.bfloat16: BF16}

_CT = {
    "int": ctypes.c_int, "long": ctypes.c_long, "float": ctypes.c_float, "long long": ctypes.c_longlong,
    "double": ctypes.c_double,
}


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes])} for every `int|const char* sidlsg_*(...)` declaration."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"(const char\*|int)\s+(sidlsg_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        argtypes = []
        for a in [x.strip() for x in args.split(",") if x.strip()]:
            if "*" in a:
                argtypes.append(ctypes.c_void_p)
            else:
                t = re.sub(r"\b(const|unsigned)\b", "", a).strip()
                t = " ".join(t.split()[:-1])  # drop the parameter name
                argtypes.append(_CT[t])
        out[name] = (ctypes.c_char_p if ret.startswith("const char") else ctypes.c_int, argtypes)
    return out


class _Lib:
    def __init__(self):
        self._dll = None
        self._fns = {}
        self.launches = 0  # number of C-ABI calls that launched kernels (bench.py reports it)

    def load(self):
        if self._dll is not None:
            return self
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "sid_lsg_b200: %s is missing - build it with `python -m sid_lsg_b200.build` "
                "(or __graft_entry__.build()). There is no CPU / PyTorch fallback." % LIB_PATH)
        self._dll = ctypes.CDLL(LIB_PATH)
        for name, (ret, argtypes) in parse_header().items():
            fn = getattr(self._dll, name)
            fn.restype = ret
            fn.argtypes = argtypes
            self._fns[name] = fn
        return self

    def last_error(self):
        self.load()
        return self._fns["sidlsg_last_error"]().decode()

    def call(self, name, *args):
        self.load()
        st = self._fns["sidlsg_" + name](*args)
        if st != 0:
            raise RuntimeError("sidlsg_%s failed (%d): %s" % (name, st, self.last_error()))
        self.launches += 1


lib = _Lib()


def ptr(t):
    """device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def dt(t):
    try:
        return _DT[t.dtype if isinstance(t, torch.Tensor) else t]
    except KeyError:
        raise TypeError("sid_lsg_b200 kernels take float32 or bfloat16, got %s" % (t.dtype if isinstance(t, torch.Tensor) else t))


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("sid_lsg_b200 ops run on CUDA tensors only (no CPU fallback); got a %s tensor" % t.device)
