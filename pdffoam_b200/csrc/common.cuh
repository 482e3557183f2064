// common.cuh — shared device/host definitions of libpdfb200 (sm_100a).
//
// Data layout in HBM:
//   * particle state: structure of arrays, double-buffered (PState A/B),
//   * tet table: one 128-byte record per tet (three shape-function gradients,
//     four neighbours, face / point / cell labels),
//   * interpolation nodes: one 96-byte record per cell, point and face holding
//     every field the mid-point models interpolate, plus a 32-byte record for
//     the position-correction velocity,
//   * per-cell moment block [nCells][nMom] (the buffer that is all-reduced).
#pragma once

// (also compiled by NVRTC at run time, jit.cu: the toolkit and C library headers are replaced by empty
// stand-ins there, NVRTC has the device API built in)
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/pdfb200.h"

#define PDF_SMALL 1e-15
#define PDF_VSMALL 1e-300
#define PDF_GREAT 1e15
#define PDF_VGREAT 1e300

#define KEY_DEAD 0xFFFFFFFFu

// RNG purposes: must match oracle/src/pdforacle.hpp (RngPurpose)
#define RNG_STEP_NORMALS 1u
#define RNG_INLET_STEPFRAC 2u
#define RNG_NC_PARTICLE 3u
#define RNG_NC_CELL 4u
#define RNG_CLONE_POS 5u
#define RNG_SEED 6u
#define RNG_INJECT 7u
#define RNG_MIX_HASH 8u   // entity = particle id; sub = round (modifiedCurl matching order)
#define RNG_MIX_PAIR 9u   // entity = id of the pair's first particle; sub = round
#define RNG_INJECT_SAMPLE 10u   // entity = boundary face | draw << 32; sub = attempt ("sampling" inflow generator)

// The tet table is three planes of 32-byte chunks indexed by the tet label (DMesh::tetA/B/C):
//   A[t] = { nbr[4], face, nbrCell, g0 }   B[t] = { g1..g4 }   C[t] = { g5..g8 }
// (g = gradN of vertices a (face centre), b, c, row-major; g_d = -(g_a+g_b+g_c)).
// The particles of a warp sit in one or two cells (they are processed in sorted order) and the tets
// of a cell are consecutive labels, so one 256-bit load of a plane touches the 6 lines that hold the
// cell's 24 chunks instead of one line per lane: the L1 cost of a record gather is the number of
// distinct 128-byte lines per load instruction, and it was what bounded the particle kernel in
// round 1 (one 128-byte record per tet: 4 loads x up to 32 lines).
// A fourth plane, D[t] = { gd.x, gd.y, gd.z, (ptB, ptC) }, is read once per particle at the mid-point: the
// gradient of vertex d as richTetrahedron stores it (Sd/(-3V), richTetrahedronI.H:107-134 — not the
// negated sum of the other three, which differs in the last bit and, multiplied by absolute pressures
// of 1e5, showed up at 1e-9 in grad p*) and the global point labels of b and c that the interpolations
// need. The centre of the cell (vertex d) is in DMesh::cellCentre4.
struct __align__(32) TetA {
    int nbr[4];       // tet across the face opposite a, b, c, d (-1: boundary)
    int face;         // mesh face the tet stands on (tetFace_)
    int nbrCell;      // cell across that face (-1: boundary)
    double g0;
};
static_assert(sizeof(TetA) == 32, "TetA must be one 32-byte sector");

// 256-bit read-only global load (SASS: LDG.E.ENL2.256.CONSTANT, new on sm_100): one instruction per
// 32-byte chunk. The address must be 32-byte aligned.
struct __align__(32) D4 { double x, y, z, w; };
// point labels of b and c from plane D of the tet table
__host__ __device__ inline int2 tetPtsOf(const D4 &d) {
    long long bits;
    memcpy(&bits, &d.w, sizeof(bits));
    return make_int2((int)(bits & 0xffffffffll), (int)(bits >> 32));
}
__device__ __forceinline__ D4 ld256(const void *p) {
    D4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}

#define NODE_MID_STRIDE 12   // doubles per mid-point node record
// The mid-point node values are stored as NODE_MID_STRIDE/4 planes of 32-byte chunks per node
// (cells, then points, then faces): value k of node n is at nmIdx(nNodes, n, k). Points with
// neighbouring labels then share 128-byte lines in every plane.
__host__ __device__ inline size_t nmIdx(size_t nNodes, size_t node, int k) { return ((size_t)(k >> 2) * nNodes + node) * 4 + (size_t)(k & 3); }
struct NodeRef {   // nm[k] access to one node of the planes
    double *base; size_t nNodes, node;
    __host__ __device__ double &operator[](int k) const { return base[nmIdx(nNodes, node, k)]; }
};
// offsets inside a mid-point node record
#define NM_U 0
#define NM_K 3
#define NM_DU 4
#define NM_OMEGA 7
#define NM_PSTAR 8
#define NM_ETA 9
#define NM_ZVAR 10

// one entry of the list of particles lost in a step (k_lost_reduce sums them per cell in a
// fixed order: no floating-point atomics)
struct __align__(16) LostRec { uint32_t idx; int cell; double m; };

// particle state columns (SoA)
struct PState {
    double *px, *py, *pz;
    double *ux, *uy, *uz;
    double *m, *omega, *rho, *eta;
    double *phi[PDFB200_MAX_PHI];
    int *cell, *tet;
    uint64_t *id;             // persistent particle id (keys the random streams), 64 bits
};

// per-step scalars living on the device (kernels read deltaT etc. from here)
struct StepScalars {
    double deltaT;
    double pcCorr;          // limitedSimple: corr factor applied to -grad(phi)
    double omegaClip;       // min(Omega, omegaClip); HUGE when no clipping
    double etaMin, etaScale; // cell LTS rescale: eta = etaScale*(raw-etaMin)+1
    uint32_t step;
    int nSorted;            // particles reachable through perm (alive after last sort)
    int nTail;              // appended after the sort (clones, then injected)
    int nInjStart;          // tail index where this step's injected particles start
    int nSlots;             // slots in use in the current state buffer
    int nEntries;           // entries of the running step (nSorted + nTail after injection): set by k_step_begin
    int ringIdx;            // slot of the report ring the running step writes (k_step_end advances it)
    int anyLost;            // lost mass to redistribute at the next K1
    int overflow;           // capacity exceeded
    unsigned long long nextId;   // id of the next injected or cloned particle (advances in steps of nRanks)
    long long nLostTotal;
};

// mesh tables on the device
struct DMesh {
    int nCells, nFaces, nInternalFaces, nPoints, nBnd, nPatches;
    long long nTets;
    int tetBits;             // sort key = (cell << tetBits) | localTet
    int cellFacesUniform;    // faces per cell when every cell has the same number (hex meshes: 6), else 0
    int geoD[3], solD[3];
    int nGeoD;
    int axisym;
    double axis[3], centreNormal[3], openingAngle;
    const double *points;        // [nPoints*3]
    const int *faceStart, *facePoints, *owner, *neighbour;
    const int *cellFaceStart, *cellFaces;
    const int *cellTetStart;     // [nCells+1]
    const int *cellRank, *cellOfRank;   // processing order of the cells (Morton curve through the centres) and its inverse
    const int *faceTetStartOwn, *faceTetStartNei;
    const double *faceCentres, *faceAreas, *magSf;   // [nFaces*3],[nFaces*3],[nFaces]
    const double4 *cellCentre4;  // xyz + hNum
    const double *cellVolumes, *volumeOrArea;
    const double *bbMin, *bbMax; // [nCells*3]
    const double *courantCoeffs; // [nFaces*3], zero on empty patches
    const D4 *cellCo4;           // [cellFaceStart[nCells]] coefficients in cell-face order (xyz, w = 0)
    const double *linWeights;    // [nInternalFaces]
    const int *pointCellStart, *pointCells; const double *pointWeights;
    const TetA *tetA;            // tet table, three planes of 32-byte chunks (see TetA)
    const D4 *tetB, *tetC;
    const D4 *tetD;              // gradN of d + global point labels of b and c (packed in .w)
    const int *tetCell;          // cell of each tet (the particle kernel tracks it instead)
    // boundary faces
    const int *bfInfo;           // handler | reflecting<<4 | geom<<8 | patch<<16
    const double4 *bfNormal;     // unit normal xyz (w unused)
    const int *openCellStart, *openCellFaces;
    const double *patchFaceT;    // [nPatches*9] wedge rotation (identity otherwise)
};

#define BF_HANDLER(i) ((i) & 0xF)
#define BF_REFLECTING(i) (((i) >> 4) & 0xF)
#define BF_GEOM(i) (((i) >> 8) & 0xFF)
#define BF_PATCH(i) ((i) >> 16)

// ---------------------------------------------------------------------------
// small vector helpers
// ---------------------------------------------------------------------------
struct V3 { double x, y, z; };
__host__ __device__ inline V3 mk(double x, double y, double z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
__host__ __device__ inline V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ inline V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ inline V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__host__ __device__ inline V3 operator*(double s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__host__ __device__ inline V3 operator*(V3 a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
__host__ __device__ inline V3 operator/(V3 a, double s) { return mk(a.x / s, a.y / s, a.z / s); }
__host__ __device__ inline double dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ inline V3 cross3(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__host__ __device__ inline double mag3(V3 a) { return sqrt(dot3(a, a)); }
__device__ inline V3 ld3(const double *p, long long i) { return mk(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
__device__ inline void st3(double *p, long long i, V3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }
// fused dot product of the tet walk: the same fma chain as oracle dotf()
__device__ inline double dotf(double gx, double gy, double gz, V3 r) { return fma(gx, r.x, fma(gy, r.y, gz * r.z)); }

// Sums of NACC per-lane accumulators over the lanes of a warp. A butterfly per accumulator (x += shfl_xor(x, 16), 8,
// 4, 2, 1) costs 5 NACC 64-bit shuffles per cell, which made the shuffle pipe a bound of k_moments (2e6 cells x 80 at
// nPhi = 1). Recursive halving instead: at distance 16 a lane hands half of its accumulators to its partner and takes
// the partner's copies of the half it keeps, at distance 8 a quarter, ... - about NACC shuffles in total. Every
// partial sum is formed from the same two operands as in the butterfly (own + partner's; floating-point addition
// commutes), so the results are bit-identical to it. A group of C accumulators (a power of two <= 32) leaves
// accumulator k in the lanes l with (l >> log2(32 / C)) == k.
__host__ __device__ constexpr int pow2Ceil(int x) { int c = 1; while (c < x) c <<= 1; return c; }
__host__ __device__ constexpr int log2Floor(int x) { int l = 0; while ((1 << (l + 1)) <= x) ++l; return l; }
template <int C>
__device__ __forceinline__ double laneSums(double (&v)[C], int lane) {
    int n = C;   // folds to constants once the loop is unrolled
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const bool upper = (lane & o) != 0;
        if (n > 1) {
            const int half = n >> 1;
#pragma unroll
            for (int k = 0; k < C / 2; ++k) {
                if (k < half) {
                    const double send = upper ? v[k] : v[k + half];
                    const double keep = upper ? v[k + half] : v[k];
                    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            }
            n = half;
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
        }
    }
    return v[0];
}
// rotationTensor(n1,n2) (OpenFOAM transform.H), row-major
__host__ __device__ inline void rotationTensor(V3 n1, V3 n2, double T[9]) {
    const double s = dot3(n1, n2);
    const V3 n3 = cross3(n1, n2);
    const double magSqr = dot3(n3, n3);
    const double v[3] = {n3.x, n3.y, n3.z}, a[3] = {n1.x, n1.y, n1.z}, b[3] = {n2.x, n2.y, n2.z};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double val = (i == j ? s : 0.0);
            val += (1 - s) * (v[i] * v[j]) / (magSqr + PDF_VSMALL);
            val += b[i] * a[j] - a[i] * b[j];
            T[3 * i + j] = val;
        }
}
__host__ __device__ inline V3 xform(const double *T, V3 v) {
    return mk(T[0] * v.x + T[1] * v.y + T[2] * v.z, T[3] * v.x + T[4] * v.y + T[5] * v.z, T[6] * v.x + T[7] * v.y + T[8] * v.z);
}

// ---------------------------------------------------------------------------
// Philox4x32-10, keyed per (entity, step, purpose, sub)
// ---------------------------------------------------------------------------
__device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t o[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
struct RngKey { uint32_t k0, k1; };
// counter = (entity low word, entity high word, step, purpose << 24 | sub): entities are 64-bit
// (particle ids), sub < 2^24
__device__ inline void rngWords(RngKey key, uint64_t entity, uint32_t step, uint32_t purpose, uint32_t sub, uint32_t o[4]) {
    philox4x32_10((uint32_t)entity, (uint32_t)(entity >> 32), step, (purpose << 24) | sub, key.k0, key.k1, o);
}
__device__ inline void rngUniform2(RngKey key, uint64_t entity, uint32_t step, uint32_t purpose, uint32_t sub, double &u0, double &u1) {
    uint32_t o[4];
    rngWords(key, entity, step, purpose, sub, o);
    u0 = (double)((((unsigned long long)o[1] << 32) | o[0]) >> 11) * 0x1.0p-53;
    u1 = (double)((((unsigned long long)o[3] << 32) | o[2]) >> 11) * 0x1.0p-53;
}

// Two N(0,1) deviates from two 32-bit words: Box-Muller evaluated in single precision with explicit
// fused multiply-adds only (polynomials of the Cephes logf / sinf / cosf), so that the CPU oracle
// reproduces every bit with fmaf() and the double-precision pipe stays free for the physics: the
// six normals of a particle step were 15 % of the particle kernel when drawn with the
// double-precision log / sincospi. Radius from u = (a+1) 2^-32 in (0,1] (tail up to 6.66 sigma),
// angle from b: two quadrant bits and a 30-bit position inside the quadrant.
__device__ inline void normalPairF32(uint32_t a, uint32_t b, float &n0, float &n1) {
    const float x = (float)a + 1.0f;   // [1, 2^32]
    const int xi = __float_as_int(x);
    int e = (xi >> 23) - 127;
    float m = __int_as_float((xi & 0x007fffff) | 0x3f800000);   // [1, 2)
    if (m > 1.41421354f) { m *= 0.5f; e += 1; }
    const float f = m - 1.0f, z = f * f;
    float p = 7.0376836292E-2f;
    p = fmaf(p, f, -1.1514610310E-1f);
    p = fmaf(p, f, 1.1676998740E-1f);
    p = fmaf(p, f, -1.2420140846E-1f);
    p = fmaf(p, f, 1.4249322787E-1f);
    p = fmaf(p, f, -1.6668057665E-1f);
    p = fmaf(p, f, 2.0000714765E-1f);
    p = fmaf(p, f, -2.4999993993E-1f);
    p = fmaf(p, f, 3.3333331174E-1f);
    const float lnm = fmaf(f * z, p, fmaf(-0.5f, z, f));            // ln(m)
    const float L = fmaf((float)(32 - e), 0.693147182f, -lnm);      // -ln(u) >= 0
    const float r = __fsqrt_rn(2.0f * L);
    const uint32_t q = b >> 30;
    const float t = (float)(b & 0x3fffffffu) * 0x1p-30f;            // [0, 1]
    const bool sw = t > 0.5f;
    const float th = (sw ? 1.0f - t : t) * 1.57079637f;             // [0, pi/4]
    const float z2 = th * th;
    float sp = -1.9515295891E-4f;
    sp = fmaf(sp, z2, 8.3321608736E-3f);
    sp = fmaf(sp, z2, -1.6666654611E-1f);
    const float s0 = fmaf(sp * z2, th, th);
    float cp = 2.443315711809948E-5f;
    cp = fmaf(cp, z2, -1.388731625493765E-3f);
    cp = fmaf(cp, z2, 4.166664568298827E-2f);
    const float c0 = fmaf(cp * z2, z2, fmaf(-0.5f, z2, 1.0f));
    const float s1 = sw ? c0 : s0, c1 = sw ? s0 : c0;               // angle inside the quadrant
    const bool odd = (q & 1u) != 0;
    float sn = odd ? c1 : s1, cs = odd ? s1 : c1;                   // + q * pi/2
    if (q >= 2u) sn = -sn;
    if (q == 1u || q == 2u) cs = -cs;
    n0 = r * cs; n1 = r * sn;
}
__device__ inline void rngNormal2(RngKey key, uint64_t entity, uint32_t step, uint32_t purpose, uint32_t sub, double &n0, double &n1) {
    uint32_t o[4];
    rngWords(key, entity, step, purpose, sub, o);
    float a, b;
    normalPairF32(o[0], o[1], a, b);
    n0 = (double)a; n1 = (double)b;
}
__device__ inline void rngNormal4(RngKey key, uint64_t entity, uint32_t step, uint32_t purpose, uint32_t sub, double &n0, double &n1,
                                  double &n2, double &n3) {
    uint32_t o[4];
    rngWords(key, entity, step, purpose, sub, o);
    float a, b, c, d;
    normalPairF32(o[0], o[1], a, b);
    normalPairF32(o[2], o[3], c, d);
    n0 = (double)a; n1 = (double)b; n2 = (double)c; n3 = (double)d;
}

// ---------------------------------------------------------------------------
// deterministic block reductions: every CTA writes one partial row, a single
// CTA folds the rows in index order (no floating-point atomics anywhere).
// ---------------------------------------------------------------------------
template <int N, bool IsMax>
__device__ inline void blockReduce(double (&v)[N], double *smem /* [N*32] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double y = __shfl_down_sync(0xffffffffu, x, o);
            x = IsMax ? fmax(x, y) : x + y;
        }
        if (lane == 0) smem[k * 32 + warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double x = lane < nw ? smem[k * 32 + lane] : (IsMax ? -PDF_VGREAT : 0.0);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double y = __shfl_down_sync(0xffffffffu, x, o);
                x = IsMax ? fmax(x, y) : x + y;
            }
            v[k] = x;
        }
    }
    __syncthreads();
}

#ifndef __CUDACC_RTC__
#define CUDA_CHECK(expr)                                                              \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) throw CudaError(_e, #expr, __FILE__, __LINE__);        \
    } while (0)

#ifdef __cplusplus
#include <stdexcept>
#include <string>
struct CudaError : std::runtime_error {
    cudaError_t code;
    CudaError(cudaError_t e, const char *expr, const char *file, int line)
        : std::runtime_error(std::string(cudaGetErrorString(e)) + " in " + expr + " at " + file + ":" + std::to_string(line)), code(e) {}
};
#endif
#endif   // !__CUDACC_RTC__
