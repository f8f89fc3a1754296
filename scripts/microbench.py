"""Print the measured roofline denominators (HBM copy GB/s, fp64 FMA TFLOP/s) of device 0."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pmpc_b200 import _lib
lib = _lib.load()
for kind, name in ((0, "hbm_copy_GBps"), (1, "fp64_fma_TFLOPps")):
    r = C.c_double()
    rc = lib.pmpc_microbench(0, kind, C.byref(r))
    print(name, rc, r.value)
