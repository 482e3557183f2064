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
        f("inlet_sampling").argtypes = [C.c_double, C.c_double, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int64, c_double_p, c_double_p]
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


# ---------------------------------------------------------------------------
# oracle/_ref: pieces of the reference compiled unmodified against an OpenFOAM stand-in (oracle/ref/)
# ---------------------------------------------------------------------------
_REF = None


def build_ref(reference: str = "/root/reference"):
    """Compile oracle/_ref/libpdfref.so from the reference tree when it is present (this container);
    on the GPU box only the prebuilt file is used. Returns the path, or None when neither exists."""
    out = os.path.join(_HERE, "_ref", "libpdfref.so")
    if os.path.isdir(os.path.join(reference, "mcParticle")):
        r = subprocess.run(["make", "-C", os.path.join(_HERE, "ref"), "-s", f"REF={reference}"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle/_ref build failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    return out if os.path.exists(out) else None


class RefLibrary:
    """ctypes view of oracle/_ref/libpdfref.so (entry points in oracle/ref/ref_shim.cpp)."""

    def __init__(self, path: str):
        L = self.lib = C.CDLL(path)
        vp = C.c_void_p
        L.pdfref_last_error.restype = C.c_char_p
        L.pdfref_find_coeffs.argtypes = [C.c_int32, c_double_p, C.c_double, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        L.pdfref_binterp.argtypes = [c_double_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_double]
        L.pdfref_binterp.restype = C.c_double
        L.pdfref_inlet_random.argtypes = [C.c_double, C.c_double, C.c_double, c_double_p, C.c_int64, c_double_p, c_double_p, C.POINTER(C.c_int32)]
        L.pdfref_inlet_random_select.argtypes = [C.c_char_p]
        L.pdfref_inlet_sampling.argtypes = [C.c_double, C.c_double, C.c_int32, C.c_int64, c_double_p]
        L.pdfref_decomp_create.argtypes = [C.c_int32, c_double_p, C.c_int32, c_int32_p, c_int32_p, c_int32_p, C.c_int32, c_int32_p, C.c_int32,
                                           c_double_p, c_double_p]
        L.pdfref_decomp_create.restype = vp
        L.pdfref_decomp_destroy.argtypes = [vp]
        L.pdfref_decomp_num_tets.argtypes = [vp]; L.pdfref_decomp_num_tets.restype = C.c_int64
        L.pdfref_decomp_tet.argtypes = [vp, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, C.POINTER(C.c_int32), c_int32_p]
        L.pdfref_decomp_cell_tets.argtypes = [vp, c_int32_p, c_int32_p]
        L.pdfref_decomp_find.argtypes = [vp, C.c_int64, c_double_p, c_int32_p, c_int32_p]
        L.pdfref_decomp_inside.argtypes = [vp, C.c_int64, c_double_p]

    def error(self) -> str:
        return self.lib.pdfref_last_error().decode()


def RefLib():
    """The compiled reference pieces, or None when oracle/_ref is not available."""
    global _REF
    if _REF is None:
        path = build_ref()
        if path is None:
            return None
        _REF = RefLibrary(path)
    return _REF
