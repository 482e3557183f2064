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

#include <cuda_runtime.h>
#include <stdint.h>

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

struct __align__(16) TetRec {
    int nbr[4];       // tet across the face opposite a, b, c, d (-1: boundary)
    int face;         // mesh face the tet stands on (tetFace_)
    int ptB, ptC;     // global point labels of b and c
    int cell;
    double g[9];      // gradN of vertices a (face centre), b, c; g_d = -(g_a+g_b+g_c)
    double centre[3]; // centre of `cell` (vertex d): saves the dependent cell-centre load in the walk
};
static_assert(sizeof(TetRec) == 128, "TetRec must be one 128-byte line");

#define NODE_MID_STRIDE 12   // doubles per mid-point node record
// offsets inside a mid-point node record
#define NM_U 0
#define NM_K 3
#define NM_DU 4
#define NM_OMEGA 7
#define NM_PSTAR 8
#define NM_ETA 9
#define NM_ZVAR 10

// particle state columns (SoA)
struct PState {
    double *px, *py, *pz;
    double *ux, *uy, *uz;
    double *m, *omega, *rho, *eta;
    double *phi[PDFB200_MAX_PHI];
    int *cell, *tet;
    uint32_t *id;
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
    int anyLost;            // lost mass to redistribute at the next K1
    uint32_t nextId;
    int overflow;           // capacity exceeded
    long long nLostTotal;
};

// mesh tables on the device
struct DMesh {
    int nCells, nFaces, nInternalFaces, nPoints, nBnd, nPatches;
    long long nTets;
    int tetBits;             // sort key = (cell << tetBits) | localTet
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
    const double *cellCo;        // [cellFaceStart[nCells]*3] coefficients in cell-face order
    const double *linWeights;    // [nInternalFaces]
    const int *pointCellStart, *pointCells; const double *pointWeights;
    const TetRec *tets;
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

// 256-bit read-only global load (SASS: LDG.E.ENL2.256.CONSTANT, new on sm_100). One
// instruction per 32-byte chunk instead of two: half the L1 wavefronts for the
// record gathers of the particle kernel. The address must be 32-byte aligned.
struct __align__(32) D4 { double x, y, z, w; };
__device__ __forceinline__ D4 ld256(const void *p) {
    D4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
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
__device__ inline void rngUniform2(RngKey key, uint32_t entity, uint32_t step, uint32_t purpose, uint32_t sub, double &u0, double &u1) {
    uint32_t o[4];
    philox4x32_10(entity, step, purpose, sub, key.k0, key.k1, o);
    u0 = (double)((((unsigned long long)o[1] << 32) | o[0]) >> 11) * 0x1.0p-53;
    u1 = (double)((((unsigned long long)o[3] << 32) | o[2]) >> 11) * 0x1.0p-53;
}
__device__ inline void rngNormal2(RngKey key, uint32_t entity, uint32_t step, uint32_t purpose, uint32_t sub, double &n0, double &n1) {
    uint32_t o[4];
    philox4x32_10(entity, step, purpose, sub, key.k0, key.k1, o);
    const double u0 = (double)(((((unsigned long long)o[1] << 32) | o[0]) >> 11) + 1ull) * 0x1.0p-53;
    const double u1 = (double)((((unsigned long long)o[3] << 32) | o[2]) >> 11) * 0x1.0p-53;
    const double r = sqrt(-2.0 * log(u0));
    double s, c;
    sincospi(2.0 * u1, &s, &c);
    n0 = r * c; n1 = r * s;
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
