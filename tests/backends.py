"""Backends for the parity tests: the CUDA library (needs a GPU) and the test-only host emulation of the same device
code behind the same C ABI (tests/emul/emul.cpp)."""
import os
import subprocess

import pytest

from calipso_b200 import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = {}


def binding(kind: str) -> _lib.Binding:
    if kind not in _CACHE:
        if kind == "emul":
            subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "emul")])
            _CACHE[kind] = _lib.Binding(os.path.join(_HERE, "emul", "libcb200_emul.so"))
        elif kind == "cuda":
            _CACHE[kind] = _lib.default_binding()
        else:
            raise ValueError(kind)
    return _CACHE[kind]


BACKENDS = ["emul", pytest.param("cuda", marks=pytest.mark.gpu)]
