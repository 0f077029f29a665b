"""Developer tool: supernode / schedule statistics of the symbolic analysis.  python tools/symstats.py [cfg3|cfg2|tiny] [verbose]"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from calipso_b200 import lqc

so = "/tmp/ss/libsymstats.so"
subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, os.path.join(ROOT, "tools", "symstats.cpp"),
                       os.path.join(ROOT, "calipso_b200", "csrc", "symbolic.cpp"), os.path.join(ROOT, "calipso_b200", "csrc", "amd.cpp")])
lib = C.CDLL(so)
P = getattr(lqc, sys.argv[1] if len(sys.argv) > 1 else "cfg3")()
ip = lambda a: np.ascontiguousarray(a, dtype=np.int32).ctypes.data_as(C.POINTER(C.c_int))
arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (P.soc_dims, P.W_colptr, P.W_rowval, P.G_colptr, P.G_rowval, P.C_colptr, P.C_rowval)]
lib.symstats(P.n, P.m, P.p, P.num_nonnegative, len(P.soc_dims), *[ip(a) for a in arrs], int(sys.argv[3]) if len(sys.argv) > 3 else 3000,
             int(sys.argv[2]) if len(sys.argv) > 2 else 1)
