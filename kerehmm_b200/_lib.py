"""ctypes binding of libkhmm.so -- the thin C-ABI layer between the Python host classes and
the sm_100a kernels.  Prototypes are read from include/khmm.h so header and binding cannot
drift.  There is NO CPU fallback: if the library is missing, loading fails loudly."""
from __future__ import annotations

import ctypes
import os
import re

PKG = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(PKG), "include", "khmm.h")
LIB_PATH = os.path.join(PKG, "libkhmm.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3
U8, U16, I32, I64, F32, F64 = range(6)
GAMMA_REFERENCE, GAMMA_POSTERIOR = 0, 1
SMALL_N_MAX = 16

_SCALARS = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "size_t": ctypes.c_size_t,
            "int32_t": ctypes.c_int32, "uint64_t": ctypes.c_uint64, "uint32_t": ctypes.c_uint32,
            "double": ctypes.c_double, "float": ctypes.c_float}


class KhmmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libkhmm error %d: %s" % (code, msg))
        self.code = code


def _ctype(decl: str):
    decl = decl.strip()
    if "*" in decl or decl.startswith("khmm_stream_t"):
        return ctypes.c_void_p
    base = decl.replace("const", "").split()[0]
    return _SCALARS[base]


def parse_header(path: str = HEADER):
    """-> {name: (restype, [argtypes], [argnames])} for every prototype in khmm.h."""
    with open(path) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|size_t|int)\s*(khmm_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if "char" in ret else _SCALARS[ret]
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                argtypes.append(_ctype(a))
                argnames.append(re.split(r"[\s\*]+", a)[-1])
        protos[name] = (restype, argtypes, argnames)
    return protos


_lib = None
_protos = None


def lib():
    """The loaded library with argtypes set.  Raises if it has not been built."""
    global _lib, _protos
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "kerehmm_b200: %s is missing. Build it with `python -m kerehmm_b200.build` "
                "(needs nvcc). There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        _protos = parse_header()
        for name, (restype, argtypes, _) in _protos.items():
            fn = getattr(L, name)       # AttributeError if the .so does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = L
    return _lib


def call(name: str, *args):
    """Call an int-returning entry point; raise KhmmError on a non-zero status."""
    L = lib()
    rc = getattr(L, name)(*args)
    if rc != OK:
        raise KhmmError(rc, L.khmm_last_error().decode())
    return rc
