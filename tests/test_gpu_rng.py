"""The keyed counter-based generator on the device against its CPU restatement.

Uniforms come from Philox4x32-10 (integer arithmetic: bit-exact by construction). The normal deviates are
drawn by a single-precision Box-Muller written with explicit fused multiply-adds only, which the oracle
restates with fmaf(): every deviate must agree to the last bit, for particle-id sized (64-bit) entities too."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _probe(lib, device, seed, ent, step, purpose, sub):
    out = np.zeros((ent.shape[0], 6))
    e = np.ascontiguousarray(ent, dtype=np.uint64)
    args = (seed, e.shape[0], e.ctypes.data_as(C.POINTER(C.c_uint64)), step, purpose, sub, out.ctypes.data_as(C.POINTER(C.c_double)))
    if device is None:
        lib.check(lib.fn("rng_probe")(*args), "rng_probe")
    else:
        lib.check(lib.fn("rng_probe")(device, *args), "rng_probe")
    return out


def test_device_generator_bit_identical_to_oracle(device_lib, oracle_lib):
    rng = np.random.default_rng(1)
    ent = np.concatenate([np.arange(200000, dtype=np.uint64), rng.integers(0, 2 ** 63, size=300000, dtype=np.uint64),
                          np.array([0, 2 ** 32 - 1, 2 ** 32, 2 ** 64 - 1], dtype=np.uint64)])
    for seed, step, purpose, sub in ((0x5EED, 0, 1, 0), (0x5EED, 7, 1, 1), (0xDEADBEEFCAFE, 123456, 7, 4 * 321 + 1), (1, 2 ** 32 - 1, 6, 1000001)):
        d = _probe(device_lib, 0, seed, ent, step, purpose, sub)
        o = _probe(oracle_lib, None, seed, ent, step, purpose, sub)
        assert np.array_equal(d[:, :2], o[:, :2]), "uniforms differ"
        bad = np.nonzero(d[:, 2:] != o[:, 2:])
        assert bad[0].size == 0, f"{bad[0].size} normal deviates differ, first: entity {ent[bad[0][0]]} {d[bad[0][0]]} vs {o[bad[0][0]]}"
    z = d[:, 2:].ravel()
    assert abs(z.mean()) < 5e-3 and abs(z.var() - 1) < 5e-3 and np.abs(z).max() < 6.7


def test_extreme_words_of_the_box_muller(oracle_lib):
    """a = 2^32-1 gives u = 1 (radius 0); a = 0 gives the largest radius sqrt(2*32*ln 2) = 6.66."""
    f = oracle_lib.fn("normal_pair")
    f.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
    out = (C.c_float * 2)()
    f(0xFFFFFFFF, 0x12345678, out)
    assert out[0] == 0.0 and out[1] == 0.0
    f(0, 0, out)   # angle 0: (r, 0)
    assert abs(out[0] - np.sqrt(2 * 32 * np.log(2.0))) < 1e-5 and out[1] == 0.0
    f(0, 1 << 30, out)   # quarter turn: (0, r) up to the cos polynomial's zero
    assert abs(out[1] - np.sqrt(2 * 32 * np.log(2.0))) < 1e-5 and abs(out[0]) < 1e-6
