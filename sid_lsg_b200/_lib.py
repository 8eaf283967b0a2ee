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
LIB_PATH = os.environ.get("SIDLSG_LIB") or os.path.join(_HERE, "_C", "libsidlsg.so")   # SIDLSG_LIB: A/B builds side by side

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
    for m in re.finditer(r"(const char\*|int|long)\s+(sidlsg_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        argtypes = []
        for a in [x.strip() for x in args.split(",") if x.strip()]:
            if "*" in a:
                argtypes.append(ctypes.c_void_p)
            else:
                t = re.sub(r"\b(const|unsigned)\b", "", a).strip()
                t = " ".join(t.split()[:-1])  # drop the parameter name
                argtypes.append(_CT[t])
        rt = ctypes.c_char_p if ret.startswith("const char") else (ctypes.c_long if ret == "long" else ctypes.c_int)
        out[name] = (rt, argtypes)
    return out


class _Lib:
    def __init__(self):
        self._dll = None
        self._fns = {}
        self.launches = 0  # number of C-ABI calls that launched kernels (bench.py reports it)
        self.timer = None  # KernelTimer while bench.py measures per-kernel rooflines

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

    def query(self, name, *args):
        """value-returning entry points (no status code, no launch)."""
        self.load()
        return self._fns["sidlsg_" + name](*args)

    def last_error(self):
        self.load()
        return self._fns["sidlsg_last_error"]().decode()

    def try_call(self, name, *args):
        """like call(), but returns False instead of raising when the entry point reports SIDLSG_ERR_UNSUPPORTED
        (shape not eligible for that kernel: the caller then uses the general entry point)."""
        self.load()
        st = self._fns["sidlsg_" + name](*args)
        if st == -3:
            return False
        if st != 0:
            raise RuntimeError("sidlsg_%s failed (%d): %s" % (name, st, self.last_error()))
        self.launches += 1
        return True

    def call(self, name, *args):
        self.load()
        timer = self.timer
        if timer is not None and name in timer.work:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev1 = torch.cuda.Event(enable_timing=True)
            ev0.record()
            st = self._fns["sidlsg_" + name](*args)
            ev1.record()
            path = self._fns["sidlsg_last_path"]() if name in ("gemm", "conv3x3", "conv3x3_wgrad") else 1
            timer.records.append((name, path, timer.work[name](args), ev0, ev1, timer.shape_of(name, args)))
        else:
            st = self._fns["sidlsg_" + name](*args)
        if st != 0:
            raise RuntimeError("sidlsg_%s failed (%d): %s" % (name, st, self.last_error()))
        self.launches += 1


class KernelTimer:
    """Per-call CUDA-event timing of selected C-ABI entry points with their ALGORITHMIC work
    (FLOPs for the contractions, bytes for the HBM-bound kernels); used by bench.py's roofline pass."""

    def __init__(self):
        def gemm(a):      # ..., M, N, K, nb1, nb2 at positions 23..27
            return ("flop", 2.0 * a[23] * a[24] * a[25] * a[26] * a[27])

        def conv(a):      # B, Hi, Wi, Kc, Ho, Wo, N at 6..12
            return ("flop", 2.0 * a[6] * a[10] * a[11] * a[12] * 9 * a[9])

        def wgrad(a):     # B, Hi, Wi, Cin, Ho, Wo, Cout at 3..9
            return ("flop", 2.0 * a[3] * a[7] * a[8] * a[9] * 9 * a[6])

        def attn(a):      # q,k,v,o,lse,B,N,M,H,d
            return ("flop", 4.0 * a[5] * a[6] * a[7] * a[8] * a[9])

        def attn_bwd(a):  # 5 contractions (S recompute, dP, dV, dK, dQ)
            return ("flop", 10.0 * a[11] * a[12] * a[13] * a[14] * a[15])

        def gn_fwd(a):    # B, HW, C at 9..11; read x twice (stats + apply), write y
            return ("byte", 3.0 * a[9] * a[10] * a[11] * (4 if a[15] == 0 else 2))

        def gn_bwd(a):    # B, HW, C at 13..15; read dy,x twice, write dx
            return ("byte", 5.0 * a[13] * a[14] * a[15] * (4 if a[19] == 0 else 2))

        def adam(a):      # p,g,m,v,ema,shadow,ema_shadow,n: read p,g,v(,ema) write p,v(,ema,shadows)
            n = a[7]
            return ("byte", n * (20.0 + (8 if a[4] else 0) + (2 if a[5] else 0) + (2 if (a[4] and a[6]) else 0) +
                                 (8 if a[2] else 0)))

        def lsg(a):       # B, CHW at 7,8: 3 reads + 3 writes fp32
            return ("byte", 24.0 * a[7] * a[8])

        self.work = {"gemm": gemm, "conv3x3": conv, "conv3x3_wgrad": wgrad, "attention_fwd": attn,
                     "attention_bwd": attn_bwd, "groupnorm_fwd": gn_fwd, "groupnorm_bwd": gn_bwd, "adam_step": adam,
                     "lsg_loss": lsg}
        self.records = []

    @staticmethod
    def shape_of(name, a):
        """short shape tag of a timed call (per-shape table of bench.py --shapes)."""
        if name == "gemm":      # M,N,K, batches, operand majors, accumulate
            return "M%d N%d K%d nb%d %s%s acc%d" % (a[23], a[24], a[25], a[26] * a[27], "k" if a[2] == 1 else "m",
                                                    "k" if a[7] == 1 else "n", a[22])
        if name == "conv3x3":   # B,Hi,Wi,Kc,Ho,Wo,N ... stride, flip
            return "B%d H%d C%d->%d s%d%s" % (a[6], a[7], a[9], a[12], a[16], " dgrad" if a[19] else "")
        if name == "conv3x3_wgrad":
            return "B%d H%d C%d->%d s%d" % (a[3], a[4], a[6], a[9], a[13])
        if name == "attention_fwd":
            return "B%d N%d M%d H%d d%d" % tuple(a[5:10])
        if name == "attention_bwd":
            return "B%d N%d M%d H%d d%d" % tuple(a[11:16])
        if name in ("groupnorm_fwd", "groupnorm_bwd"):
            return "B%d HW%d C%d" % ((a[9], a[10], a[11]) if name == "groupnorm_fwd" else (a[13], a[14], a[15]))
        return ""

    def by_shape(self):
        """-> list of dict(name, path, shape, launches, ms, work, unit) sorted by time."""
        torch.cuda.synchronize()
        out = {}
        for name, path, (unit, work), ev0, ev1, shape in self.records:
            d = out.setdefault((name, path, shape), dict(name=name, tc=path, shape=shape, launches=0, ms=0.0, work=0.0, unit=unit))
            d["launches"] += 1
            d["ms"] += ev0.elapsed_time(ev1)
            d["work"] += work
        return sorted(out.values(), key=lambda d: -d["ms"])

    def summary(self):
        """-> {key: dict(launches, ms, work, unit)}; key = name or name+'[simt]' for CUDA-core GEMM/conv calls."""
        torch.cuda.synchronize()
        out = {}
        for name, path, (unit, work), ev0, ev1, _shape in self.records:
            key = name if path else name + "[simt]"
            d = out.setdefault(key, dict(launches=0, ms=0.0, work=0.0, unit=unit))
            d["launches"] += 1
            d["ms"] += ev0.elapsed_time(ev1)
            d["work"] += work
        return out


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
    """CUDA tensors on the CURRENT device only: the library launches on the current device's stream and keeps its
    per-process kernel attributes for one device per process (one process per GPU, as torchrun runs it)."""
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("sid_lsg_b200 ops run on CUDA tensors only (no CPU fallback); got a %s tensor" % t.device)
        if t.device.index != torch.cuda.current_device():
            raise RuntimeError("sid_lsg_b200 ops launch on the current CUDA device (%d); got a tensor on %s - call "
                               "torch.cuda.set_device first" % (torch.cuda.current_device(), t.device))
