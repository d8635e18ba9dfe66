"""ctypes binding of librcmvs.so (the C ABI declared in include/rcmvs.h).

The product path has NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librcmvs.so")
HEADER = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include", "rcmvs.h")

_lib = None
_lock = threading.Lock()
launches = 0          # number of C-ABI calls that enqueued a kernel (bench.py reports it as gpu_launches)

_CTYPES = {
    "const float*": ctypes.c_void_p, "float*": ctypes.c_void_p, "int32_t*": ctypes.c_void_p,
    "int": ctypes.c_int, "int64_t": ctypes.c_int64, "float": ctypes.c_float, "double": ctypes.c_double,
    "rcmvs_stream_t": ctypes.c_void_p, "uint64_t": ctypes.c_uint64, "const float* const*": ctypes.c_void_p,
    "uint8_t*": ctypes.c_void_p, "const uint8_t*": ctypes.c_void_p, "const int32_t*": ctypes.c_void_p,
    "const char*": ctypes.c_char_p, "float *": ctypes.c_void_p, "void*": ctypes.c_void_p, "double*": ctypes.c_void_p, "const double*": ctypes.c_void_p, "const void*": ctypes.c_void_p, "const int*": ctypes.c_void_p, "int*": ctypes.c_void_p,
}


def declared_functions(header=HEADER):
    """Parse include/rcmvs.h -> {name: (restype, [argtypes])}.  Keeps Python in lock-step with the header."""
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int64_t|int|const char\*)\s+(rcmvs_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                ty = a.rsplit(" ", 1)[0].replace(" *", "*")
                argtypes.append(_CTYPES[ty])
        restype = {"int": ctypes.c_int, "int64_t": ctypes.c_int64}.get(ret, ctypes.c_char_p)
        out[name] = (restype, argtypes)
    return out


def lib():
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        "librcmvs.so not found at %s — build it with `python -c 'import __graft_entry__ as g; "
                        "g.build()'` (needs nvcc). There is no CPU/PyTorch fallback." % LIB_PATH)
                L = ctypes.CDLL(LIB_PATH)
                for name, (ret, argtypes) in declared_functions().items():
                    fn = getattr(L, name)      # AttributeError here = header/library mismatch
                    fn.restype = ret
                    fn.argtypes = argtypes
                _lib = L
    return _lib


def call(name, *args):
    """Invoke an int-returning entry point; raise RuntimeError with rcmvs_last_error() on failure."""
    global launches
    L = lib()
    rc = getattr(L, name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, L.rcmvs_last_error().decode()))
    launches += 1
    return rc
