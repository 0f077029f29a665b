"""Binding of an older / experimental build: symbols it does not export are skipped.  Developer tool."""
import ctypes as C

from calipso_b200 import _lib


class LooseBinding(_lib.Binding):
    def __init__(self, path):
        self.path = path
        self.lib = C.CDLL(path)
        for name, (res, argtypes) in _lib.SYMBOLS.items():
            try:
                fn = getattr(self.lib, name)
            except AttributeError:
                continue
            fn.restype = res
            fn.argtypes = argtypes
