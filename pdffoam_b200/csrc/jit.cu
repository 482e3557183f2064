// jit.cu — run-time specialisation of the particle kernel.
//
// pdfFoam selects its models at run time (mcModel::New, run-time selection tables); a kernel
// that carries every model is three times the code the selected ones need, and on B200 the
// instruction footprint of k_evolve is what bounds it (profiles/README.md). So at the first
// step the library compiles evolve_kernel.cuh once more (about a second) with the cloud's
// parameters and mesh flags as constants: it writes a small translation unit, runs
// `nvcc -cubin` for sm_100a (same flags as the ahead-of-time build, -fmad=false included, so
// results are bit-identical to the generic kernel), caches the cubin on disk under a hash of
// everything that went in, and loads it with cudaLibraryLoadFromFile. If any of that is not
// possible (no nvcc, sources not next to the library, PDFB200_JIT=0) the ahead-of-time generic
// kernel runs instead and says so once on stderr.
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "cloud.hpp"

namespace {

std::string hexDouble(double v) {
    if (std::isnan(v)) return "__builtin_nan(\"\")";
    if (std::isinf(v)) return v > 0 ? "__builtin_huge_val()" : "(-__builtin_huge_val())";
    char b[64];
    std::snprintf(b, sizeof b, "%a", v);
    return b;
}

// the constants the kernel may rely on: the parameters k_evolve reads and the mesh flags
std::string prologue(const pdfb200_params &p, bool axisym, bool hasOpen, int tetBits, int cellFaces, const int geoD[3], const int solD[3]) {
    std::ostringstream o;
    o << "#define K1_JIT 1\n";
    // CTAs per SM the specialised kernel is compiled for (profiles/README.md)
    o << "#ifndef K1_MIN_BLOCKS\n#define K1_MIN_BLOCKS " << K1_JIT_MIN_BLOCKS << "\n#endif\n";
    o << "#define JIT_NPHI " << (p.nPhi <= 1 ? 1 : p.nPhi == 2 ? 2 : p.nPhi == 3 ? 3 : PDFB200_MAX_PHI) << "\n";
    o << "#define JIT_AXISYM " << (axisym ? 1 : 0) << "\n";
    o << "#define JIT_HAS_OPEN " << (hasOpen ? 1 : 0) << "\n";
    o << "#define JIT_TETBITS " << tetBits << "\n";
    o << "#define JIT_CELLFACES " << cellFaces << "\n";
    o << "#define JIT_NGEOD " << ((geoD[0] > 0) + (geoD[1] > 0) + (geoD[2] > 0)) << "\n";
    for (int k = 0; k < 3; ++k) o << "#define JIT_GEOD" << k << " (" << geoD[k] << ")\n#define JIT_SOLD" << k << " (" << solD[k] << ")\n";
    o << "#include \"cloud.hpp\"\n";
    o << "struct JitP {\n";
#define JI(f) o << "    static constexpr int " #f " = " << p.f << ";\n";
#define JD(f) o << "    static constexpr double " #f " = " << hexDouble(p.f) << ";\n";
#define JA(f, n)                                                                              \
    o << "    static __host__ __device__ constexpr int " #f "At(int i) { return ";             \
    for (int k = 0; k < (n); ++k) o << "i == " << k << " ? " << p.f[k] << " : ";               \
    o << "0; }\n";
    JI(nPhi) JI(nMixed) JA(mixed, PDFB200_MAX_PHI) JI(nConserved) JA(conserved, PDFB200_MAX_PHI)
    JD(CFL) JD(kMin) JD(DNum)
    JI(interpU) JI(interpk) JI(interpDiffU) JI(interpOmega) JI(interpUPosCorr) JI(interpZVar) JI(interpEta)
    JI(velocityModel) JD(C0) JD(C1) JI(mixingModel) JD(Cmix) JI(omegaModel) JI(reactionModel) JD(coldDensity)
    JD(bsRfuel) JD(bsRox) JD(bsRstoi) JD(bsTfuel) JD(bsTox) JD(bsTstoi) JD(bsZstoi)
    JI(zIdx) JI(TIdx) JI(chiIdx) JD(Cchi) JI(nAdded) JA(addedIdx, PDFB200_MAX_ADDED)
    JI(positionCorrection) JI(localTimeStepping) JD(ltsUpperBound)
#undef JI
#undef JD
#undef JA
    o << "};\n#include \"evolve_kernel.cuh\"\n";
    return o.str();
}

uint64_t fnv1a(const std::string &s, uint64_t h = 1469598103934665603ull) {
    for (unsigned char ch : s) { h ^= ch; h *= 1099511628211ull; }
    return h;
}

bool readFile(const std::string &path, std::string &out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}

bool fileExists(const std::string &p) { struct stat st; return ::stat(p.c_str(), &st) == 0 && st.st_size > 0; }

// directory holding evolve_kernel.cuh / common.cuh: next to the shared library unless overridden
std::string sourceDir() {
    if (const char *e = std::getenv("PDFB200_SRC_DIR")) return e;
    Dl_info info;
    if (dladdr(reinterpret_cast<void *>(&fnv1a), &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        const size_t k = p.rfind('/');
        return k == std::string::npos ? "." : p.substr(0, k);
    }
    return ".";
}

std::string findNvcc() {
    if (const char *e = std::getenv("PDFB200_NVCC")) return e;
    for (const char *env : {"CUDA_HOME", "CUDA_PATH"})
        if (const char *e = std::getenv(env)) { const std::string p = std::string(e) + "/bin/nvcc"; if (fileExists(p)) return p; }
    if (fileExists("/usr/local/cuda/bin/nvcc")) return "/usr/local/cuda/bin/nvcc";
    return "nvcc";
}

bool writableDir(const std::string &d) {
    ::mkdir(d.c_str(), 0700);
    return ::access(d.c_str(), W_OK | X_OK) == 0;
}

// $PDFB200_JIT_CACHE, else $TMPDIR or /tmp, else a directory next to the library
std::string cacheDir(const std::string &srcDir) {
    const std::string leaf = "pdfb200_jit_" + std::to_string((long)getuid());
    if (const char *e = std::getenv("PDFB200_JIT_CACHE")) return writableDir(e) ? std::string(e) : std::string();
    if (const char *e = std::getenv("TMPDIR")) { const std::string d = std::string(e) + "/" + leaf; if (writableDir(d)) return d; }
    if (writableDir("/tmp/" + leaf)) return "/tmp/" + leaf;
    if (writableDir(srcDir + "/jit_cache")) return srcDir + "/jit_cache";
    return std::string();
}

}  // namespace

// Returns the specialised kernel of this cloud, or nullptr when the generic one is to be used.
void *Cloud::specialisedKernel() {
    if (jitState != 0) return jitState > 0 ? jitKernel : nullptr;
    jitState = -1;
    const char *mode = std::getenv("PDFB200_JIT");
    if (mode && std::strcmp(mode, "0") == 0) return nullptr;

    const std::string dir = sourceDir();
    std::string srcK, srcC, srcH, srcA;
    if (!readFile(dir + "/evolve_kernel.cuh", srcK) || !readFile(dir + "/common.cuh", srcC) || !readFile(dir + "/cloud.hpp", srcH) ||
        !readFile(dir + "/../../include/pdfb200.h", srcA)) {
        std::fprintf(stderr, "pdfb200: kernel sources not found in %s; running the generic particle kernel\n", dir.c_str());
        return nullptr;
    }
    const std::string pro = prologue(prm, axisym, hasOpen, tetBits, dm.cellFacesUniform, dm.geoD, dm.solD);
    std::string flags = "-cubin -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false";
    if (const char *x = std::getenv("PDFB200_JIT_FLAGS")) flags += std::string(" ") + x;   // tuning experiments (-DK1_...)
    const uint64_t h = fnv1a(srcA, fnv1a(srcH, fnv1a(srcC, fnv1a(srcK, fnv1a(flags, fnv1a(pro))))));
    char name[64];
    std::snprintf(name, sizeof name, "k_evolve_%016llx", (unsigned long long)h);
    const std::string cdir = cacheDir(dir);
    if (cdir.empty()) {
        std::fprintf(stderr, "pdfb200: no writable cache directory for the specialised kernel; running the generic kernel\n");
        return nullptr;
    }
    const std::string cubin = cdir + "/" + name + ".cubin";
    if (!fileExists(cubin)) {
        const std::string cu = cdir + "/" + name + "." + std::to_string((long)getpid()) + ".cu", tmp = cubin + "." + std::to_string((long)getpid());
        { std::ofstream f(cu); f << pro; }
        const std::string cmd = "'" + findNvcc() + "' " + flags + " -I'" + dir + "' -o '" + tmp + "' '" + cu + "' > '" + tmp + ".log' 2>&1";
        const int rc = std::system(cmd.c_str());
        if (rc != 0 || !fileExists(tmp)) {
            std::fprintf(stderr, "pdfb200: specialising the particle kernel failed (see %s.log); running the generic kernel\n", tmp.c_str());
            return nullptr;
        }
        ::rename(tmp.c_str(), cubin.c_str());   // atomic: concurrent ranks build the same file
        ::unlink(cu.c_str());
        ::unlink((tmp + ".log").c_str());
    }
    cudaLibrary_t lib = nullptr;
    cudaKernel_t fn = nullptr;
    if (cudaLibraryLoadFromFile(&lib, cubin.c_str(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess ||
        cudaLibraryGetKernel(&fn, lib, "k_evolve_jit") != cudaSuccess) {
        cudaGetLastError();
        std::fprintf(stderr, "pdfb200: could not load %s; running the generic particle kernel\n", cubin.c_str());
        return nullptr;
    }
    jitLibrary = lib;
    jitKernel = reinterpret_cast<void *>(fn);
    jitState = 1;
    return jitKernel;
}

void Cloud::dropSpecialisedKernel() {
    if (jitLibrary) cudaLibraryUnload(reinterpret_cast<cudaLibrary_t>(jitLibrary));
    jitLibrary = nullptr; jitKernel = nullptr; jitState = 0;
}

// The translation unit specialisedKernel() would compile for these parameters and mesh flags
// (host only, no device needed): lets the CPU test-suite check that the kernel source builds
// under K1_JIT. Returns the length of the text, or -1 when it does not fit.
extern "C" long pdfb200_jit_source(const pdfb200_params *p, int axisym, int hasOpen, int tetBits, int cellFacesUniform,
                                   const int32_t geoD[3], const int32_t solD[3], char *buf, long bufLen) {
    const int g[3] = {geoD[0], geoD[1], geoD[2]}, sd[3] = {solD[0], solD[1], solD[2]};
    const std::string t = prologue(*p, axisym != 0, hasOpen != 0, tetBits, cellFacesUniform, g, sd);
    if ((long)t.size() + 1 > bufLen) return -1;
    std::memcpy(buf, t.c_str(), t.size() + 1);
    return (long)t.size();
}
