"""pdffoam_b200 — B200-native implementation of pdfFoam's per-time-step particle
hot path, mcParticleCloud::evolve (mcParticle/mcParticleCloud/mcParticleCloud.C:1226-1530).

The product is the CUDA library pdffoam_b200/csrc/libpdfb200.so behind the C ABI
of include/pdfb200.h. This package holds the ctypes binding (capi), structured
polyMesh generators (meshgen) and the case builders (cases) used by the tests
and bench.py. There is no CPU fallback: importing is cheap, but every compute
call needs the CUDA library and a GPU.
"""
from . import meshgen  # noqa: F401
from .capi import CloudHandle, DeviceLibrary, Library, Params, PdfError, StepReport  # noqa: F401
