"""ctypes binding of tools/probes/libb200probe.so (hardware probes; test tooling)."""
import ctypes
import os

from pytorch3dunet_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "libb200probe.so")


class ProbeLib:
    def __init__(self):
        _lib.lib()  # the product library first (the probe library links against it)
        self.cdll = ctypes.CDLL(PATH)
        for name, argtypes in _lib.parse_header(os.path.join(HERE, "b200probe.h")).items():
            fn = getattr(self.cdll, name)
            fn.argtypes, fn.restype = argtypes, ctypes.c_int
            setattr(self, "_" + name, fn)

    def call(self, name, *args):
        rc = getattr(self, "_" + name)(*args)
        if rc != 0:
            raise _lib.B200Error(f"{name} failed (rc={rc}): {_lib.lib().last_error()}")


def available():
    return os.path.exists(PATH)
