"""Loader of the CPU oracle (test infrastructure).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs import this module; nothing under pdffoam_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from pdffoam_b200.capi import Library, StepReport, c_double_p, c_int32_p

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile oracle/src/*.cpp into oracle/libpdforacle.so (gcc, no OpenFOAM)."""
    out = os.path.join(_HERE, "libpdforacle.so")
    srcs = [os.path.join(_HERE, "src", f) for f in os.listdir(os.path.join(_HERE, "src"))]
    srcs.append(os.path.join(_HERE, "..", "include", "pdfb200.h"))
    if force or not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return out


class OracleLibrary(Library):
    def _declare(self):
        super()._declare()
        vp = C.c_void_p
        f = self.fn
        f("set_rank").argtypes = [vp, C.c_int, C.c_int]
        f("moment_block_size").argtypes = [vp]; f("moment_block_size").restype = C.c_int64
        f("evolve_begin").argtypes = [vp, C.POINTER(StepReport), c_double_p]
        f("evolve_end").argtypes = [vp, C.POINTER(StepReport), c_double_p]
        f("tet_detail").argtypes = [vp, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, c_int32_p]
        f("find_reference").argtypes = [vp, C.c_int64, c_double_p, c_int32_p, c_int32_p]
        f("locate").argtypes = [vp, C.c_int64, c_double_p, c_int32_p, c_int32_p]
        f("tet_inside").argtypes = [vp, C.c_int64, c_double_p]
        f("interpolate_scalar").argtypes = [vp, c_double_p, c_double_p, C.c_int64, c_double_p, c_int32_p, c_double_p, c_double_p]
        f("find_coeffs").argtypes = [C.c_int32, c_double_p, C.c_double, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        f("binterp").argtypes = [c_double_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_double]
        f("binterp").restype = C.c_double
        f("inlet_random").argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, c_double_p]
        f("inlet_random_sample").argtypes = [C.c_double, C.c_double, C.c_uint64, C.c_int64, c_double_p]
        f("rng_probe").argtypes = [C.c_uint64, C.c_int64, C.POINTER(C.c_uint64), C.c_uint32, C.c_uint32, C.c_uint32, c_double_p]
        f("philox").argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        f("rotation_tensor").argtypes = [c_double_p, c_double_p, c_double_p]


def OracleLib() -> OracleLibrary:
    global _LIB
    if _LIB is None:
        _LIB = OracleLibrary(build(), "pdforacle_")
    return _LIB


def dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def ip(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int32_p)
