"""Compile the run-time specialised kernel of the bench workload locally for several CTA counts and print
registers / spills / SASS statistics.  usage: python scratch/jitcheck.py [extra nvcc flags]"""
import ctypes as C, os, subprocess, sys, re
ROOT = os.getcwd()   # run from the repository root
sys.path.insert(0, ROOT)
from pdffoam_b200 import capi, cases
case = cases.synthetic_box(4, 2, 2, ppc=64, closed=False, number_control=True)
lib = C.CDLL(capi.device_library_path())
holder = capi.Library.__new__(capi.Library)
holder.lib, holder.prefix = lib, "pdfb200_"
holder._declare()
src = capi.jit_source(holder, case.params, False, True, 0, case.mesh.geometric_d, case.mesh.solution_d)
out = os.path.join(ROOT, "scratch", "jit")   # scratch/ is git-ignored
os.makedirs(out, exist_ok=True)
cu = os.path.join(out, "bench.cu")
open(cu, "w").write(src)
csrc = os.path.join(ROOT, "pdffoam_b200", "csrc")
extra = sys.argv[1:]
for b in (3, 4, 5, 6):
    cubin = os.path.join(out, f"b{b}.cubin")
    r = subprocess.run(["nvcc", "-cubin", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
                        f"-DK1_MIN_BLOCKS={b}", "-Xptxas", "-v", "-I" + csrc, "-o", cubin, cu] + extra, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stderr[-3000:]); sys.exit(1)
    info = " ".join(l.strip() for l in r.stderr.splitlines() if "registers" in l or "spill" in l)
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    lines = [l for l in sass.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    cnt = lambda pat: sum(1 for l in lines if re.search(pat, l))
    print(f"blocks {b}: {info}\n   instr {len(lines)}  LDG {cnt(r'LDG')} (256: {cnt(r'LDG.*256')})  LDL {cnt(r'LDL')} STL {cnt(r'STL')}  DFMA/DMUL/DADD {cnt(r'D(FMA|MUL|ADD)')}  DSETP {cnt(r'DSETP')} MUFU {cnt(r'MUFU')} FFMA {cnt(r'FFMA')} IMAD {cnt(r'IMAD')} LDS {cnt(r'LDS')} STS {cnt(r'STS')}")
