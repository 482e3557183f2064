#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/.

Run in the build container (needs /root/reference for the flamelet tables):
    python tests/golden/make_golden.py

1. flamelet_sandiaD.npz / flamelet_bluffbody.npz — the reference's own flamelet
   libraries (tutorials/SandiaD/constant/flamelet/{z,chi,rho,T},
   tutorials/SydneyBluffBodyFlame/constant/flamelet/*), parsed from OpenFOAM
   scalarList / scalarListList files. These are the real table inputs of
   mcSteadyFlamelet (mcSteadyFlamelet.C:120-286).
2. lookup_golden.npz — lookups on those tables evaluated by an independent
   numpy restatement of findCoeffs/binterp (mcSteadyFlamelet.C:44-93) written
   in this script, used to pin the C++ oracle and the CUDA kernel.
4. autoignition.npz — the auto-ignition library of tutorials/AutoIgnitionBurner
   (constant/autoIgnition/{z,pv,cEq,cdot,rho,T}), the table input of
   mcKulkarniAutoIgnitionReactionModel (mcKulkarniAutoIgnition.C:111-335).
3. inlet_golden.npz — mcInletRandom PDF/CDF/Q (mcInletRandomI.H:28-39,
   mcInletRandom.C:92-108) evaluated with scipy for the parameters of
   tests/mcInletRandom (Umean=2, urms=4), plus numerical quadrature of Q.
"""
import os
import re
import sys

import numpy as np
from scipy import integrate, special

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/tutorials"


def parse_foam_list(path):
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//.*", "", txt)
    body = txt[txt.index("}") + 1:]
    toks = re.findall(r"[()]|[-+0-9.eE]+", body)
    pos = 0

    def parse():
        nonlocal pos
        n = int(toks[pos]); pos += 1
        assert toks[pos] == "("; pos += 1
        out = []
        while toks[pos] != ")":
            if pos + 1 < len(toks) and toks[pos + 1] == "(":
                out.append(parse())
            else:
                out.append(float(toks[pos])); pos += 1
        pos += 1
        assert len(out) == n, (len(out), n)
        return out

    return np.array(parse(), dtype=np.float64)


def find_coeffs(lst, v):
    """numpy restatement of findCoeffs (mcSteadyFlamelet.C:44-72)"""
    it = int(np.searchsorted(lst, v, side="left"))   # std::lower_bound
    if it == 0:
        return 0, 0.0
    if it == len(lst):
        return len(lst) - 2, 1.0
    return it - 1, (v - lst[it - 1]) / (lst[it] - lst[it - 1])


def binterp(v, ix, wx, iy, wy):
    """binterp (mcSteadyFlamelet.C:75-93) with its transposed weights"""
    r1 = min(ix + 1, v.shape[0] - 1)
    r0 = max(ix, 0)
    v00, v10, v01, v11 = v[r0, iy], v[r0, iy + 1], v[r1, iy], v[r1, iy + 1]
    return v00 * (1 - wx) * (1 - wy) + v10 * wx * (1 - wy) + v01 * (1 - wx) * wy + v11 * wx * wy


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (run in the build container)")
    tables = {}
    for name, case in (("flamelet_sandiaD", "SandiaD"), ("flamelet_bluffbody", "SydneyBluffBodyFlame")):
        d = os.path.join(REF, case, "constant", "flamelet")
        t = {k: parse_foam_list(os.path.join(d, k)) for k in ("z", "chi", "rho", "T")}
        assert t["rho"].shape == (t["chi"].size, t["z"].size) == t["T"].shape
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **t)
        tables[name] = t
        print(name, {k: v.shape for k, v in t.items()})

    # 4. autoignition.npz - the library of mcKulkarniAutoIgnitionReactionModel (tutorials/AutoIgnitionBurner/constant/
    #    autoIgnition/{z,pv,cEq,cdot,rho,T}; rows = z, columns = pv, mcKulkarniAutoIgnition.C:111-335)
    d = os.path.join(REF, "AutoIgnitionBurner", "constant", "autoIgnition")
    ai = {k: parse_foam_list(os.path.join(d, k)) for k in ("z", "pv", "cEq", "cdot", "rho", "T")}
    assert ai["cdot"].shape == (ai["z"].size, ai["pv"].size) == ai["rho"].shape == ai["T"].shape and ai["cEq"].shape == ai["z"].shape
    np.savez_compressed(os.path.join(HERE, "autoignition.npz"), **ai)
    print("autoignition", {k: v.shape for k, v in ai.items()})

    rng = np.random.default_rng(2024)
    out = {}
    for name, t in tables.items():
        z = np.concatenate([rng.uniform(-0.05, 1.05, 400), t["z"][[0, 1, 5, -1]], [t["z"][3] * (1 + 1e-15)]])
        chi = np.concatenate([10 ** rng.uniform(-3, 2.5, 400), t["chi"][[0, 1, -1, -1]], [1e-15]])
        rho = np.zeros(z.size); T = np.zeros(z.size)
        for i in range(z.size):
            iz, wz = find_coeffs(t["z"], z[i]); ic, wc = find_coeffs(t["chi"], chi[i])
            rho[i] = binterp(t["rho"], ic, wc, iz, wz)
            T[i] = binterp(t["T"], ic, wc, iz, wz)
        out[name + "_z"] = z; out[name + "_chi"] = chi; out[name + "_rho"] = rho; out[name + "_T"] = T
    np.savez_compressed(os.path.join(HERE, "lookup_golden.npz"), **out)

    # mcInletRandom known answers (tests/mcInletRandom/mcInletRandomTest.C:42-118: Umean=2, urms=4)
    res = {}
    for tag, (a, urms) in {"test": (2.0, 4.0), "fast": (10.0, 1.5), "back": (-1.0, 2.0)}.items():
        b = 1.0 / (np.sqrt(2.0) * urms)
        denom = np.exp(-a * a * b * b) + a * b * np.sqrt(np.pi) * (1 + special.erf(a * b))
        x = np.linspace(0.0, a + 6 * urms, 41)
        pdf = 2 * b * b / denom * x * np.exp(-b * b * (x - a) ** 2)
        cdf = np.array([integrate.quad(lambda s: 2 * b * b / denom * s * np.exp(-b * b * (s - a) ** 2), 0, xi, epsabs=1e-13, epsrel=1e-13)[0] for xi in x])
        gauss = lambda s: np.exp(-0.5 * ((s - a) / urms) ** 2) / (urms * np.sqrt(2 * np.pi))
        Q = integrate.quad(lambda s: s * gauss(s), 0, a + 40 * urms, epsabs=1e-14, epsrel=1e-13)[0]
        res[tag + "_params"] = np.array([a, urms]); res[tag + "_x"] = x; res[tag + "_pdf"] = pdf; res[tag + "_cdf"] = cdf
        res[tag + "_Q"] = np.array([Q])
    np.savez_compressed(os.path.join(HERE, "inlet_golden.npz"), **res)
    print("wrote goldens to", HERE)


if __name__ == "__main__":
    main()
