"""ctypes binding of libb200unet.so (the C-ABI declared in include/b200unet.h).

The prototypes are parsed from the header itself, so the header is the single source of truth for the
boundary.  There is no fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
HEADER = os.path.join(_ROOT, "include", "b200unet.h")
LIB_PATH = os.environ.get("B200UNET_LIB") or os.path.join(_HERE, "libb200unet.so")   # B200UNET_LIB: e.g. the debug build with wait counters
LIB_PATH_F16 = os.path.join(_HERE, "libb200unet_f16.so")   # the same sources built with fp16 activations / operands (-DB200_ACT_F16)

_CTYPES = {
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "size_t": ctypes.c_size_t,
    "b200_stream_t": ctypes.c_void_p,
}


def _ctype_of(decl: str):
    decl = decl.strip()
    if "*" in decl:
        return ctypes.c_void_p
    decl = re.sub(r"\bconst\b", "", decl).strip()
    # drop the parameter name
    for key in sorted(_CTYPES, key=len, reverse=True):
        if decl == key or decl.startswith(key + " "):
            return _CTYPES[key]
    raise ValueError(f"cannot map C type in declaration {decl!r}")


def parse_header(path: str = HEADER):
    """-> {function name: [ctypes argument types]} for every `int b200_*(...)` prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\bint\s+(b200_\w+)\s*\(([^)]*)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        if args in ("", "void"):
            protos[name] = []
        else:
            protos[name] = [_ctype_of(a) for a in args.split(",")]
    return protos


class B200Error(RuntimeError):
    pass


class _Lib:
    def __init__(self, path=None):
        path = path or LIB_PATH
        if not os.path.exists(path):
            raise B200Error(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU/PyTorch fallback for the b200 3D U-Net engine)")
        self.path = path
        self.cdll = ctypes.CDLL(path)
        self.protos = parse_header()
        for name, argtypes in self.protos.items():
            fn = getattr(self.cdll, name)  # AttributeError if the header declares something the .so lacks
            fn.argtypes = argtypes
            fn.restype = ctypes.c_int
            setattr(self, "_" + name, fn)

    def last_error(self) -> str:
        buf = ctypes.create_string_buffer(512)
        self.cdll.b200_last_error(buf, ctypes.c_size_t(512))
        return buf.value.decode(errors="replace")

    def query(self, name: str, *args) -> int:
        """functions that return a value (counts, flags) rather than a status"""
        return getattr(self, "_" + name)(*args)

    def call(self, name: str, *args) -> None:
        rc = getattr(self, "_" + name)(*args)
        if rc != 0:
            raise B200Error(f"{name} failed (rc={rc}): {self.last_error()}")


_lib = None
_lib_f16 = None


def lib(operand_dtype: str = "bf16") -> _Lib:
    """the library whose 16-bit activation / tensor-core operand type is `operand_dtype` ("bf16" default, "fp16")"""
    global _lib, _lib_f16
    if operand_dtype in ("fp16", "f16", "float16"):
        if _lib_f16 is None:
            _lib_f16 = _Lib(LIB_PATH_F16)
        return _lib_f16
    if _lib is None:
        _lib = _Lib()
    return _lib
