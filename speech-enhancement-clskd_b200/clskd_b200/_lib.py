"""ctypes binding of libclskd_sm100.so (the C ABI declared in include/clskd.h).

The prototypes are parsed from the header itself so the Python side can never drift from the
C side; `EXPORTS` is the list of declared entry points (the CPU test-suite checks that the
library exports every one of them).  There is NO fallback: if the library is missing or a call
fails, a RuntimeError is raised.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG_ROOT = os.path.dirname(_HERE)
_REPO_ROOT = os.path.dirname(_PKG_ROOT)
HEADER = os.path.join(_REPO_ROOT, "include", "clskd.h")
LIB_PATH = os.path.join(_HERE, "libclskd_sm100.so")

MAX_TAPS = 16
F32, BF16 = 0, 1


class TapConv(ctypes.Structure):
    """Mirror of `ClskdTapConv` (include/clskd.h)."""
    _fields_ = [
        ("x0", ctypes.c_void_p), ("x1", ctypes.c_void_p),
        ("x0_sB", ctypes.c_int64), ("x0_sT", ctypes.c_int64), ("x0_sF", ctypes.c_int64),
        ("x1_sB", ctypes.c_int64), ("x1_sT", ctypes.c_int64), ("x1_sF", ctypes.c_int64),
        ("c0", ctypes.c_int32), ("c1", ctypes.c_int32),
        ("B", ctypes.c_int32), ("To", ctypes.c_int32), ("Fo", ctypes.c_int32),
        ("Ti", ctypes.c_int32), ("Fi", ctypes.c_int32),
        ("sf", ctypes.c_int32),
        ("ntaps", ctypes.c_int32),
        ("dt", ctypes.c_int32 * MAX_TAPS), ("df", ctypes.c_int32 * MAX_TAPS),
        ("w", ctypes.c_void_p), ("bias", ctypes.c_void_p),
        ("N", ctypes.c_int32),
        ("y", ctypes.c_void_p),
        ("y_sB", ctypes.c_int64), ("y_sT", ctypes.c_int64), ("y_sF", ctypes.c_int64),
        ("x_dtype", ctypes.c_int32), ("y_dtype", ctypes.c_int32),
        ("accumulate", ctypes.c_int32),
        ("ep_scale", ctypes.c_void_p), ("ep_shift", ctypes.c_void_p), ("ep_slope", ctypes.c_void_p),
        ("stats_sum", ctypes.c_void_p), ("stats_sumsq", ctypes.c_void_p),
    ]


_CTYPE = {
    "int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64,
    "float": ctypes.c_float, "double": ctypes.c_double, "size_t": ctypes.c_size_t,
}


def _parse_header(path):
    """-> {name: (restype, [argtypes])} for every `clskd_*` prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int)\s+(clskd_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.replace("const ", "").rsplit(" ", 1)[0].strip()
                    argtypes.append(_CTYPE[ty])
        protos[name] = (ctypes.c_char_p if "char" in ret else ctypes.c_int, argtypes)
    return protos


PROTOS = _parse_header(HEADER)
EXPORTS = sorted(PROTOS)

_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "clskd_b200: %s is missing - build it with `python speech-enhancement-clskd_b200/build.py` "
            "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOS.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


launch_count = 0  # number of C-ABI compute calls issued (bench.py reports it as gpu_launches)


def call(name, *args):
    """Call an int-returning entry point and raise on a non-zero return code."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args)
    launch_count += 1
    if rc != 0:
        msg = lib.clskd_last_error()
        raise RuntimeError("%s failed (%d): %s" % (name, rc, msg.decode() if msg else "?"))
    return rc
