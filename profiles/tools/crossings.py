import sys, os, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np
from pdffoam_b200 import cases
from oracle.oracle import OracleLib
lib = OracleLib()
case = cases.synthetic_box(40, 20, 20, ppc=64, closed=False, number_control=True)
h = case.make_cloud(lib); h.seed_particles()
f = lib.fn("debug_crossings"); f.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int)]
for blk in range(12):
    reps = h.evolve(10 if blk < 6 else 20)
    hist = (C.c_int64 * 64)(); mx = C.c_int(0)
    f(h.h, 64, hist, C.byref(mx))
    a = np.array(list(hist)); n = a.sum()
    mean = (a * np.arange(64)).sum() / n
    print("step", h.step_counter(), "n", n, "mean crossings %.2f" % mean, "max", mx.value, "P(>8) %.2e P(>16) %.2e P(>32) %.2e >=63: %d" % (a[9:].sum() / n, a[17:].sum() / n, a[33:].sum() / n, a[63]), "dT %.3g" % reps[-1].deltaT, "UMax %.1f" % reps[-1].UMax, "inj", reps[-1].nInjected, flush=True)
