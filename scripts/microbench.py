"""Print the measured roofline denominators (HBM copy GB/s, fp64 FMA TFLOP/s) of device 0."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pmpc_b200 import _lib
lib = _lib.load()
for kind, name in ((0, "hbm_copy_GBps"), (1, "fp64_fma_TFLOPps"), (2, "fp64_dmma_m8n8k4_TFLOPps"),
                   (3, "fp64_fma_lds64_broadcast_TFLOPps"), (4, "fp64_fma_lds128_broadcast_TFLOPps")):
    r = C.c_double()
    rc = lib.pmpc_microbench(0, kind, C.byref(r))
    print(name, rc, r.value)
