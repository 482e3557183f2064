// jit.cu — run-time specialisation of the particle kernel.
//
// pdfFoam selects its models at run time (mcModel::New, run-time selection tables); a kernel
// that carries every model is three times the code the selected ones need, and on B200 the
// instruction footprint of k_evolve is what bounds it (profiles/README.md). So at the first
// step the library compiles evolve_kernel.cuh once more (about a second) with the cloud's
// parameters and mesh flags as constants.
//
// The kernel sources are embedded in the library at build time (embedded_sources.inc, written by
// the Makefile from the very files the host side was compiled with), so the specialised kernel
// always agrees with the host's K1Args / DMesh / accumulator layout and nothing has to lie next to
// the .so. They are compiled in memory with NVRTC (libnvrtc is dlopen'ed: it ships with the CUDA
// toolkit, with PyTorch and as a stand-alone redistributable), with the flags of the ahead-of-time
// build (-fmad=false included, so results are bit-identical to the generic kernel), and loaded with
// cudaLibraryLoadData. Without NVRTC the same translation unit goes through `nvcc -cubin` in a
// private temporary directory. The cubin is cached under a hash of everything that went in, in a
// directory that must belong to the caller and be closed to everyone else ($PDFB200_JIT_CACHE,
// $XDG_CACHE_HOME/pdfb200, ~/.cache/pdfb200; never a shared /tmp path); without such a directory
// it is simply rebuilt by every process. If no compiler can be found (or PDFB200_JIT=0) the
// ahead-of-time generic kernel runs instead and says so once on stderr.
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>
#include <fcntl.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <fstream>
#include <vector>
#include <sstream>

#include "cloud.hpp"
#include "embedded_sources.inc"   // SRC_common_cuh, SRC_k1_layout_h, SRC_evolve_kernel_cuh, SRC_pdfb200_h

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
    o << "#include \"common.cuh\"\n#include \"k1_layout.h\"\n";
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
    JI(zIdx) JI(TIdx) JI(chiIdx) JI(cIdx) JI(pvIdx) JD(Cchi) JI(nAdded) JA(addedIdx, PDFB200_MAX_ADDED)
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

// the headers of the translation unit: the embedded kernel sources and empty stand-ins for the
// toolkit / C library headers they name (NVRTC has the device API and the fixed-width types built in
// once these typedefs exist)
struct JitHeader { const char *name, *text; };
const char *const kStdint =
    "#pragma once\n"
    "typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;\n"
    "typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;\n";
const JitHeader kHeaders[] = {
    {"common.cuh", SRC_common_cuh},
    {"k1_layout.h", SRC_k1_layout_h},
    {"evolve_kernel.cuh", SRC_evolve_kernel_cuh},
    {"../../include/pdfb200.h", SRC_pdfb200_h},
    {"cuda_runtime.h", "#pragma once\n"},
    {"string.h", "#pragma once\n"},
    {"stdint.h", kStdint},
};
constexpr int kNumHeaders = sizeof(kHeaders) / sizeof(kHeaders[0]);

uint64_t sourcesHash(uint64_t h) {
    for (int i = 0; i < 4; ++i) h = fnv1a(kHeaders[i].text, h);
    return h;
}

// ---- NVRTC, dlopen'ed ----
struct Nvrtc {
    void *so = nullptr;
    int (*createProgram)(void **, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    int (*destroyProgram)(void **) = nullptr;
    int (*compileProgram)(void *, int, const char *const *) = nullptr;
    int (*getCUBINSize)(void *, size_t *) = nullptr;
    int (*getCUBIN)(void *, char *) = nullptr;
    int (*getProgramLogSize)(void *, size_t *) = nullptr;
    int (*getProgramLog)(void *, char *) = nullptr;
    bool ok() const { return so != nullptr; }
};

const Nvrtc &nvrtc() {
    static const Nvrtc api = [] {
        Nvrtc a;
        // The kernel uses 256-bit global loads (PTX ISA 8.8): NVRTC older than CUDA 12.9 cannot assemble it. A process
        // that has PyTorch loaded already holds that build's libnvrtc.so.12 (12.8 for cu128 wheels), which a lookup by
        // soname would hand back, so the toolkit's own copy is tried by path first and every candidate's version is checked.
        std::vector<std::string> names;
        if (const char *e = std::getenv("PDFB200_NVRTC")) names.push_back(e);
        for (const char *env : {"CUDA_HOME", "CUDA_PATH"})
            if (const char *e = std::getenv(env)) names.push_back(std::string(e) + "/lib64/libnvrtc.so");
        names.push_back("/usr/local/cuda/lib64/libnvrtc.so");
        names.insert(names.end(), {"libnvrtc.so.13", "libnvrtc.so.12", "libnvrtc.so"});
        void *so = nullptr;
        for (const std::string &n : names) {
            so = dlopen(n.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (!so) continue;
            int major = 0, minor = 0;
            auto ver = reinterpret_cast<int (*)(int *, int *)>(dlsym(so, "nvrtcVersion"));
            if (ver && ver(&major, &minor) == 0 && (major > 12 || (major == 12 && minor >= 9))) break;
            dlclose(so);
            so = nullptr;
        }
        if (!so) return a;
        bool all = true;
        auto sym = [&](const char *n) { void *p = dlsym(so, n); if (!p) all = false; return p; };
        a.createProgram = reinterpret_cast<decltype(a.createProgram)>(sym("nvrtcCreateProgram"));
        a.destroyProgram = reinterpret_cast<decltype(a.destroyProgram)>(sym("nvrtcDestroyProgram"));
        a.compileProgram = reinterpret_cast<decltype(a.compileProgram)>(sym("nvrtcCompileProgram"));
        a.getCUBINSize = reinterpret_cast<decltype(a.getCUBINSize)>(sym("nvrtcGetCUBINSize"));
        a.getCUBIN = reinterpret_cast<decltype(a.getCUBIN)>(sym("nvrtcGetCUBIN"));
        a.getProgramLogSize = reinterpret_cast<decltype(a.getProgramLogSize)>(sym("nvrtcGetProgramLogSize"));
        a.getProgramLog = reinterpret_cast<decltype(a.getProgramLog)>(sym("nvrtcGetProgramLog"));
        if (all) a.so = so; else dlclose(so);
        return a;
    }();
    return api;
}

std::vector<std::string> splitFlags(const std::string &s) {
    std::vector<std::string> out;
    std::istringstream is(s);
    std::string w;
    while (is >> w) out.push_back(w);
    return out;
}

// compile the translation unit `pro` in memory; false (and the compiler's log) when NVRTC is missing or fails
bool compileNvrtc(const std::string &pro, const std::string &extraFlags, std::vector<char> &cubin, std::string &log) {
    const Nvrtc &rt = nvrtc();
    if (!rt.ok()) { log = "libnvrtc not found"; return false; }
    const char *names[kNumHeaders], *texts[kNumHeaders];
    for (int i = 0; i < kNumHeaders; ++i) { names[i] = kHeaders[i].name; texts[i] = kHeaders[i].text; }
    void *prog = nullptr;
    if (rt.createProgram(&prog, pro.c_str(), "k_evolve_jit.cu", kNumHeaders, texts, names) != 0) { log = "nvrtcCreateProgram failed"; return false; }
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "--generate-line-info"};
    for (const std::string &f : splitFlags(extraFlags)) opts.push_back(f);
    std::vector<const char *> o;
    for (const std::string &f : opts) o.push_back(f.c_str());
    const int rc = rt.compileProgram(prog, (int)o.size(), o.data());
    size_t n = 0;
    if (rt.getProgramLogSize(prog, &n) == 0 && n > 1) { log.resize(n); rt.getProgramLog(prog, &log[0]); }
    bool ok = rc == 0;
    if (ok) {
        size_t sz = 0;
        ok = rt.getCUBINSize(prog, &sz) == 0 && sz > 0;
        if (ok) { cubin.resize(sz); ok = rt.getCUBIN(prog, cubin.data()) == 0; }
    }
    rt.destroyProgram(&prog);
    return ok;
}

bool fileExists(const std::string &p) { struct stat st; return ::stat(p.c_str(), &st) == 0 && st.st_size > 0; }

std::string findNvcc() {
    if (const char *e = std::getenv("PDFB200_NVCC")) return e;
    for (const char *env : {"CUDA_HOME", "CUDA_PATH"})
        if (const char *e = std::getenv(env)) { const std::string p = std::string(e) + "/bin/nvcc"; if (fileExists(p)) return p; }
    if (fileExists("/usr/local/cuda/bin/nvcc")) return "/usr/local/cuda/bin/nvcc";
    return "nvcc";
}

// second choice: nvcc on the embedded sources, written to a private temporary directory
bool compileNvcc(const std::string &pro, const std::string &extraFlags, std::vector<char> &cubin, std::string &log) {
    const char *t = std::getenv("TMPDIR");
    std::string tmpl = std::string(t ? t : "/tmp") + "/pdfb200_jit_XXXXXX";
    if (!mkdtemp(&tmpl[0])) { log = "mkdtemp failed"; return false; }   // 0700, owned by the caller
    const std::string d = tmpl;
    // common.cuh names "../../include/pdfb200.h": the sources go two levels below d, the header into d/include
    ::mkdir((d + "/include").c_str(), 0700); ::mkdir((d + "/x").c_str(), 0700); ::mkdir((d + "/x/y").c_str(), 0700);
    bool ok = true;
    auto put = [&](const std::string &p, const char *text) { std::ofstream f(p); f << text; ok = ok && (bool)f; };
    const std::string src = d + "/x/y";
    put(src + "/common.cuh", SRC_common_cuh); put(src + "/k1_layout.h", SRC_k1_layout_h); put(src + "/evolve_kernel.cuh", SRC_evolve_kernel_cuh);
    put(d + "/include/pdfb200.h", SRC_pdfb200_h);
    put(src + "/k.cu", pro.c_str());
    bool done = false;
    if (ok) {
        const std::string cmd = "'" + findNvcc() + "' -cubin -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false " + extraFlags +
                                " -I'" + src + "' -o '" + d + "/k.cubin' '" + src + "/k.cu' > '" + d + "/log' 2>&1";
        const int rc = std::system(cmd.c_str());
        std::string bin;
        readFile(d + "/log", log);
        if (rc == 0 && readFile(d + "/k.cubin", bin) && !bin.empty()) { cubin.assign(bin.begin(), bin.end()); done = true; }
    } else log = "could not write the kernel sources to " + d;
    for (const char *f : {"/x/y/common.cuh", "/x/y/k1_layout.h", "/x/y/evolve_kernel.cuh", "/x/y/k.cu", "/include/pdfb200.h", "/k.cubin", "/log"}) ::unlink((d + f).c_str());
    for (const char *f : {"/x/y", "/x", "/include", ""}) ::rmdir((d + f).c_str());
    return done;
}

// A cache directory is used only if it is a real directory (not a link) that belongs to the caller and that
// nobody else can write to: a cubin found there is loaded and run on the caller's data.
bool privateDir(const std::string &d, bool create) {
    if (create) {
        // create the parents with the caller's umask, the leaf closed
        for (size_t k = 1; k < d.size(); ++k) if (d[k] == '/') ::mkdir(d.substr(0, k).c_str(), 0755);
        ::mkdir(d.c_str(), 0700);
    }
    struct stat st;
    if (::lstat(d.c_str(), &st) != 0) return false;
    return S_ISDIR(st.st_mode) && st.st_uid == getuid() && (st.st_mode & (S_IWGRP | S_IWOTH)) == 0 && ::access(d.c_str(), W_OK | X_OK) == 0;
}

std::string cacheDir() {
    if (const char *e = std::getenv("PDFB200_JIT_CACHE")) return (*e && privateDir(e, true)) ? std::string(e) : std::string();
    if (const char *e = std::getenv("XDG_CACHE_HOME")) { const std::string d = std::string(e) + "/pdfb200"; if (*e && privateDir(d, true)) return d; }
    if (const char *e = std::getenv("HOME")) { const std::string d = std::string(e) + "/.cache/pdfb200"; if (*e && privateDir(d, true)) return d; }
    return std::string();
}

// a cached cubin is trusted only as a regular file of the caller that nobody else can write to
bool readPrivateFile(const std::string &p, std::vector<char> &out) {
    struct stat st;
    if (::lstat(p.c_str(), &st) != 0 || !S_ISREG(st.st_mode) || st.st_uid != getuid() || (st.st_mode & (S_IWGRP | S_IWOTH)) != 0 || st.st_size <= 0) return false;
    std::string s;
    if (!readFile(p, s) || s.empty()) return false;
    out.assign(s.begin(), s.end());
    return true;
}

std::string jitFlags() {
    const char *x = std::getenv("PDFB200_JIT_FLAGS");   // tuning experiments (-DK1_...)
    return x ? x : "";
}

// compile (NVRTC, else nvcc); which = the compiler that produced the cubin
bool compileKernel(const std::string &pro, std::vector<char> &cubin, std::string &log, const char **which) {
    const char *force = std::getenv("PDFB200_JIT_COMPILER");   // "nvrtc" / "nvcc": tests
    std::string l1, l2;
    if (!(force && std::strcmp(force, "nvcc") == 0) && compileNvrtc(pro, jitFlags(), cubin, l1)) { *which = "nvrtc"; log = l1; return true; }
    if (!(force && std::strcmp(force, "nvrtc") == 0) && compileNvcc(pro, jitFlags(), cubin, l2)) { *which = "nvcc"; log = l2; return true; }
    log = "nvrtc: " + l1 + "\nnvcc: " + l2;
    return false;
}

}  // namespace

// Returns the specialised kernel of this cloud, or nullptr when the generic one is to be used.
void *Cloud::specialisedKernel() {
    if (jitState != 0) return jitState > 0 ? jitKernel : nullptr;
    jitState = -1;
    const char *mode = std::getenv("PDFB200_JIT");
    if (mode && std::strcmp(mode, "0") == 0) return nullptr;

    const std::string pro = prologue(prm, axisym, hasOpen, tetBits, dm.cellFacesUniform, dm.geoD, dm.solD);
    const uint64_t h = sourcesHash(fnv1a(jitFlags(), fnv1a(pro)));
    char name[64];
    std::snprintf(name, sizeof name, "k_evolve_%016llx.cubin", (unsigned long long)h);
    const std::string cdir = cacheDir();
    const std::string path = cdir.empty() ? std::string() : cdir + "/" + name;
    std::vector<char> cubin;
    if (path.empty() || !readPrivateFile(path, cubin)) {
        std::string log;
        const char *which = "";
        if (!compileKernel(pro, cubin, log, &which)) {
            std::fprintf(stderr, "pdfb200: specialising the particle kernel failed; running the generic kernel\n%s\n", log.c_str());
            return nullptr;
        }
        if (!path.empty()) {
            const std::string tmp = path + "." + std::to_string((long)getpid());
            const int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600);
            if (fd >= 0) {
                const bool ok = ::write(fd, cubin.data(), cubin.size()) == (ssize_t)cubin.size();
                ::close(fd);
                if (ok) ::rename(tmp.c_str(), path.c_str());   // atomic: concurrent ranks build the same file
                else ::unlink(tmp.c_str());
            }
        }
    }
    cudaLibrary_t lib = nullptr;
    cudaKernel_t fn = nullptr;
    if (cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess ||
        cudaLibraryGetKernel(&fn, lib, "k_evolve_jit") != cudaSuccess) {
        cudaGetLastError();
        if (!path.empty()) ::unlink(path.c_str());
        std::fprintf(stderr, "pdfb200: could not load the specialised particle kernel; running the generic kernel\n");
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

// Compiles that translation unit exactly as specialisedKernel() does (embedded sources; NVRTC, else nvcc; host only,
// no device needed). compiler: 0 = first available, 1 = NVRTC, 2 = nvcc. Returns the size of the cubin, or -1 with
// the compiler's messages in log.
extern "C" long pdfb200_jit_compile(const pdfb200_params *p, int axisym, int hasOpen, int tetBits, int cellFacesUniform,
                                    const int32_t geoD[3], const int32_t solD[3], int compiler, char *log, long logLen) {
    const int g[3] = {geoD[0], geoD[1], geoD[2]}, sd[3] = {solD[0], solD[1], solD[2]};
    const std::string pro = prologue(*p, axisym != 0, hasOpen != 0, tetBits, cellFacesUniform, g, sd);
    std::vector<char> cubin;
    std::string l;
    bool ok;
    if (compiler == 1) ok = compileNvrtc(pro, jitFlags(), cubin, l);
    else if (compiler == 2) ok = compileNvcc(pro, jitFlags(), cubin, l);
    else { const char *which = ""; ok = compileKernel(pro, cubin, l, &which); }
    if (log && logLen > 0) { std::strncpy(log, l.c_str(), (size_t)logLen - 1); log[logLen - 1] = 0; }
    if (!ok) return -1;
    // the cubin must hold the entry point
    const std::string needle = "k_evolve_jit";
    if (std::search(cubin.begin(), cubin.end(), needle.begin(), needle.end()) == cubin.end()) return -1;
    return (long)cubin.size();
}

// the cache directory the library would use right now ("" when none qualifies); length or -1
extern "C" long pdfb200_jit_cache_dir(char *buf, long bufLen) {
    const std::string d = cacheDir();
    if ((long)d.size() + 1 > bufLen) return -1;
    std::memcpy(buf, d.c_str(), d.size() + 1);
    return (long)d.size();
}
