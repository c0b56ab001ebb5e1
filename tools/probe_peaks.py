import ctypes as C, sys
sys.path.insert(0,'.')
from bayeslogit_b200 import _lib
L=_lib.lib(); _lib.check(L.bl_set_device(0))
for _ in range(3):
    b=(C.c_double*6)(); _lib.check(L.bl_probe_peaks(C.cast(b,C.c_void_p)))
    print(["%.2f"%v for v in b])
