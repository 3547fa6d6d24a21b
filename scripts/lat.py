import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torj_jl_b200 import _lib
ctx = _lib.context(0)
v = C.c_double()
for it in (1000, 10000):
    _lib.check(_lib.lib().torj_fp64_latency(ctx, it, C.byref(v)))
    print("DFMA dependent latency:", v.value, "cycles")
