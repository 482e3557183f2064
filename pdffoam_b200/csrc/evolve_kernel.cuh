// evolve_kernel.cuh — the fused per-particle kernel of one time step.
//
// One thread carries one particle through steps E1(cull)–E11 of
// mcParticleCloud::evolve (mcParticle/mcParticleCloud/mcParticleCloud.C:1226-1352):
//   open-boundary cull (mcOpenBoundary.C:84-141), mass bookkeeping (:1238-1249),
//   position-correction gather (mcLimitedSimplePositionCorrection.C:187-192),
//   constrainParticle (:76-101), half-step tet walk (mcParticle.C:100-179),
//   Courant number (:2120-2148), Omega / IEM / reaction / SLM / local time step
//   (mcRASOmegaModel.C:125-147, mcIEMMixingModel.C:68-83, mcSteadyFlamelet.C:304-329,
//   mcBurkeSchumannReactionModel.C:77-91, mcColdReactionModel.C:62-65,
//   mcSLMFullVelocityModel.C:146-185, mc*LocalTimeStepping), the mid-point rule
//   with numerical diffusion (:1297-1331), the full-step walk and the
//   open-boundary velocity fix-up (mcOpenBoundary.C:142-155).
// The particle is read through the permutation of the previous step's sort and
// written in sorted order, so the state is read once and written once per step.
//
// Compiled with -fmad=false: every fused multiply-add below is explicit. The
// fma chains of the tet walk are the ones the CPU oracle evaluates, which makes
// cell/tet decisions bit-identical.
#pragma once

#include "common.cuh"
#include "k1_layout.h"

// Compile-time knowledge of the run. The kernel is built twice from this one source: ahead of
// time with the mesh flags and model parameters read at run time (k_evolve<NPHI>, below), and at
// run time by jit.cu for the cloud at hand with K1_JIT defined, where the generated prologue
// turns the parameters (struct JitP) and the mesh flags (JIT_*) into constants, so that every
// model the case does not select drops out of the code (instruction footprint, registers).
#ifdef K1_JIT
#define MESH_AXISYM(M) (JIT_AXISYM != 0)
#define MESH_GEOD(M, k) (JIT_GEOD##k)
#define MESH_SOLD(M, k) (JIT_SOLD##k)
#define MESH_NGEOD(M) (JIT_NGEOD)
#define MESH_TETBITS(M) (JIT_TETBITS)
#define MESH_CELLFACES(M) (JIT_CELLFACES)
// JitP (generated): the parameters the kernel reads as static constexpr members, the index lists
// as constexpr functions mixedAt(i) / conservedAt(i) / addedIdxAt(i)
typedef JitP KParams;
#define PRM_AT(P, field, i) ((P).field##At(i))
#else
typedef pdfb200_params KParams;
#define PRM_AT(P, field, i) ((P).field[i])
#define MESH_AXISYM(M) ((M).axisym != 0)
#define MESH_GEOD(M, k) ((M).geoD[k])
#define MESH_SOLD(M, k) ((M).solD[k])
#define MESH_NGEOD(M) ((M).nGeoD)
#define MESH_TETBITS(M) ((M).tetBits)
#define MESH_CELLFACES(M) ((M).cellFacesUniform)
#endif
#define MESH_GEOD3(M) MESH_GEOD(M, 0), MESH_GEOD(M, 1), MESH_GEOD(M, 2)
#define MESH_SOLD3(M) MESH_SOLD(M, 0), MESH_SOLD(M, 1), MESH_SOLD(M, 2)

struct K1Args {
    DMesh M;
    pdfb200_params P;
    PState in, out;
    const uint32_t *perm;
    const StepScalars *sc;
    const double *nodeMid, *nodePC, *cellAux;
    int auxStride;
    double pcScale;        // external position correction: a*U0 (pdfb200_poscorr_fields::scale)
    const double *Un;
    const double *lostShare;
    LostRec *lostList;     // particles lost in this step (k_lost_reduce)
    unsigned *lostCount;
    unsigned lostCap;
    const int *injPatchTail;
    const double *flChi, *flZ, *flFields;   // SteadyFlamelet: chi, z, tables [nChi][nZ]; KulkarniAutoIgnition: z, pv, tables [nZ][nPv]
    int nChi, nZ;
    const double *flCEq;   // KulkarniAutoIgnition: cEq[nZ]
    double cEqMax;
    RngKey key;
    double *partSum, *partMax;
    int *workCtr;          // dynamic work distribution (K1_DYNAMIC): next unclaimed entry (zeroed by k_step_begin)
    double *chunkSums;     // [chunk][4 * K1_NC1] mass sums of every chunk (reduced in a fixed order, folded by k_chunk_fold)
    long long *ctaClock;   // profiling hook (pdfb200_profile_ctas): per CTA {start, end, SM id}, or null
    int hasOpen;
};

// a tet record in registers (no arrays: nothing here may be indexed dynamically,
// or the compiler would move it to local memory)
struct TetL {
    int n0, n1, n2, n3;          // neighbours across the faces opposite a, b, c, d
    int face, nbrCell;           // mesh face of the tet, cell across it
    int ptB, ptC;                // global point labels of b and c      } plane D, loaded where an
    double gd0, gd1, gd2;        // gradN of d (richTetrahedron's own)  } interpolation needs it (loadPts)
    double g0, g1, g2, g3, g4, g5, g6, g7, g8;   // gradN of a, b, c
};

// three 256-bit loads, one per plane of the tet table (see TetA in common.cuh); the point labels of b and c
// are fetched where an interpolation needs them (loadPts), not with every visited record
__device__ __forceinline__ void loadTet(const DMesh &M, int t, TetL &T) {
    const D4 r0 = ld256(M.tetA + t), r1 = ld256(M.tetB + t), r2 = ld256(M.tetC + t);
    T.n0 = __double2loint(r0.x); T.n1 = __double2hiint(r0.x); T.n2 = __double2loint(r0.y); T.n3 = __double2hiint(r0.y);
    T.face = __double2loint(r0.z); T.nbrCell = __double2hiint(r0.z);
    T.g0 = r0.w; T.g1 = r1.x; T.g2 = r1.y; T.g3 = r1.z; T.g4 = r1.w; T.g5 = r2.x; T.g6 = r2.y; T.g7 = r2.z; T.g8 = r2.w;
}

__device__ __forceinline__ void loadPts(const DMesh &M, int t, TetL &T) {
    const D4 d = ld256(M.tetD + t);
    T.gd0 = d.x; T.gd1 = d.y; T.gd2 = d.z;
    T.ptB = __double2loint(d.w); T.ptC = __double2hiint(d.w);
}

// centre of a cell (vertex d of its tets) and its numerical-diffusion length: one 256-bit load
__device__ __forceinline__ V3 cellCentreH(const DMesh &M, int c, double &hNum) {
    const D4 a = ld256(M.cellCentre4 + c);
    hNum = a.w;
    return mk(a.x, a.y, a.z);
}
__device__ __forceinline__ V3 cellCentre(const DMesh &M, int c) {
    double h;
    return cellCentreH(M, c, h);
}

__device__ __forceinline__ void baryAt(const TetL &T, V3 r, double &w0, double &w1, double &w2, double &w3) {
    w0 = dotf(T.g0, T.g1, T.g2, r);
    w1 = dotf(T.g3, T.g4, T.g5, r);
    w2 = dotf(T.g6, T.g7, T.g8, r);
    w3 = 1.0 - ((w0 + w1) + w2);
}

__device__ __forceinline__ double clamp01(double x) { return fmax(0.0, fmin(1.0, x)); }

// phi[idx] without dynamic indexing
template <int NPHI>
__device__ __forceinline__ double pickPhi(const double (&phi)[NPHI], int idx) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < NPHI; ++k) if (k == idx) v = phi[k];
    return v;
}

__device__ __forceinline__ double massPerDepth(const DMesh &M, V3 pos, double m) {
    if (MESH_AXISYM(M)) {
        const V3 ax = mk(M.axis[0], M.axis[1], M.axis[2]);
        const double r = mag3(pos - dot3(pos, ax) * ax);
        return m / (fmax(r, PDF_SMALL) * M.openingAngle);
    }
    return m;
}

__device__ __forceinline__ V3 reflectVec(V3 v, V3 n) {
    const double d = dot3(v, n);
    return v - (2.0 * d) * n;
}

__device__ __forceinline__ void constrainDir(int d0, int d1, int d2, V3 &v) {
    if (d0 < 0) v.x = 0;
    if (d1 < 0) v.y = 0;
    if (d2 < 0) v.z = 0;
}

// constrainParticle (mcParticleCloud.C:76-101); the axisymmetric rotation is kept out of
// line so that 3-D cases do not carry its code in the instruction cache
__device__ __forceinline__ void constrainAxisym(const DMesh &M, double dt, V3 pos, V3 &U, V3 &Ucorr, V3 &Utrack) {
    V3 dest = pos + dt * Utrack;
    const V3 ax = mk(M.axis[0], M.axis[1], M.axis[2]);
    V3 n = cross3(ax, dest);
    n = n / mag3(n);
    double T[9];
    rotationTensor(n, mk(M.centreNormal[0], M.centreNormal[1], M.centreNormal[2]), T);
    U = xform(T, U); Ucorr = xform(T, Ucorr);
    dest = xform(T, dest);
    constrainDir(MESH_GEOD3(M), dest);
    Utrack = (dest - pos) / dt;
}
__device__ __forceinline__ void constrainParticle(const DMesh &M, double dt, V3 pos, V3 &U, V3 &Ucorr, V3 &Utrack) {
    constrainDir(MESH_SOLD3(M), Utrack);
    if (MESH_AXISYM(M)) constrainAxisym(M, dt, pos, U, Ucorr, Utrack);
}

// mcSteadyFlamelet helpers (mcSteadyFlamelet.C:44-93)
__device__ inline void findCoeffsDev(const double *lst, int n, double v, int &i, double &w) {
    int lo = 0, hi = n;   // lower_bound
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (lst[mid] < v) lo = mid + 1; else hi = mid;
    }
    if (lo == 0) { i = 0; w = 0.; }
    else if (lo == n) { i = n - 2; w = 1.; }
    else { i = lo - 1; w = (v - lst[lo - 1]) / (lst[lo] - lst[lo - 1]); }
}
__device__ inline double binterpDev(const double *v, int nRows, int nCols, int ix, double wx, int iy, double wy) {
    const int r0 = min(max(ix, 0), nRows - 1), r1 = min(max(ix + 1, 0), nRows - 1);
    const double v00 = v[r0 * nCols + iy], v10 = v[r0 * nCols + iy + 1], v01 = v[r1 * nCols + iy], v11 = v[r1 * nCols + iy + 1];
    const double wxm = 1. - wx, wym = 1. - wy;
    // transposed weights of the reference, kept verbatim
    return v00 * wxm * wym + v10 * wx * wym + v01 * wxm * wy + v11 * wx * wy;
}

__device__ __forceinline__ double stabilise(double x, double y) { return x < 0 ? x - y : x + y; }

// Reaction models on one particle: mcColdReactionModel.C:62-65,
// mcBurkeSchumannReactionModel.C:77-91, mcSteadyFlamelet.C:304-329.
struct FlameletTables { const double *chi, *z, *fields; int nChi, nZ; const double *cEq; double cEqMax; };
template <int NPHI>
__device__ __forceinline__ void reactionCorrect(const KParams &P, const FlameletTables &F, double Omega, double zVar,
                                                double pCell, double etaDeltaT, double (&phi)[NPHI], double &rho, double &Co) {
    if (P.reactionModel == PDFB200_REACTION_COLD) {
        rho = P.coldDensity;
    } else if (P.reactionModel == PDFB200_REACTION_BURKE_SCHUMANN) {
        const double z = pickPhi(phi, P.zIdx);
        const double RTox = P.bsRox * P.bsTox, RTfuel = P.bsRfuel * P.bsTfuel, RTstoi = P.bsRstoi * P.bsTstoi;
        const double RT = z < P.bsZstoi ? (RTstoi - RTox) / P.bsZstoi * z + RTox
                                        : (RTfuel - RTstoi) / (1 - P.bsZstoi) * (z - P.bsZstoi) + RTstoi;
        const double Tt = z < P.bsZstoi ? (P.bsTstoi - P.bsTox) / P.bsZstoi * z + P.bsTox
                                        : (P.bsTfuel - P.bsTstoi) / (1 - P.bsZstoi) * (z - P.bsZstoi) + P.bsTstoi;
        rho = pCell / RT;
#pragma unroll
        for (int k = 0; k < NPHI; ++k) if (k == P.TIdx) phi[k] = Tt;
    } else if (P.reactionModel == PDFB200_REACTION_NAIVE_FLAMELET) {   // mcNaiveFlameletModel.C:64-69
        const double z = pickPhi(phi, P.zIdx);
        rho = (1. - 3.2 * z * (1. - z));
        const double Tt = 1e5 / rho;
#pragma unroll
        for (int k = 0; k < NPHI; ++k) if (k == P.TIdx) phi[k] = Tt;
    } else if (P.reactionModel == PDFB200_REACTION_STEADY_FLAMELET) {
        const double z = pickPhi(phi, P.zIdx);
        const double chi = fmax(P.Cchi * Omega * zVar, PDF_SMALL);
        int iz, ichi; double wz, wchi;
        findCoeffsDev(F.z, F.nZ, z, iz, wz);
        findCoeffsDev(F.chi, F.nChi, chi, ichi, wchi);
        rho = binterpDev(F.fields, F.nChi, F.nZ, ichi, wchi, iz, wz);
#pragma unroll
        for (int k = 0; k < NPHI; ++k) if (k == P.chiIdx) phi[k] = chi;
        for (int a = 0; a < P.nAdded; ++a) {
            const double v = binterpDev(F.fields + (size_t)(1 + a) * F.nChi * F.nZ, F.nChi, F.nZ, ichi, wchi, iz, wz);
#pragma unroll
            for (int k = 0; k < NPHI; ++k) if (k == PRM_AT(P, addedIdx, a)) phi[k] = v;
        }
    } else if (P.reactionModel == PDFB200_REACTION_KULKARNI) {   // mcKulkarniAutoIgnition.C:339-380; rows = z, columns = pv
        const double z = pickPhi(phi, P.zIdx);
        double c = pickPhi(phi, P.cIdx);
        int iz, ic; double wz, wc;
        findCoeffsDev(F.chi, F.nChi, z, iz, wz);
        const double cEq = F.cEq[iz] * (1. - wz) + F.cEq[iz + 1] * wz;
        c = fmin(c, cEq);
        const double pv = cEq > 1e-5 * F.cEqMax ? c / cEq : 0.;
        findCoeffsDev(F.z, F.nZ, pv, ic, wc);
        const size_t tab = (size_t)F.nChi * F.nZ;
        const double cdot = binterpDev(F.fields, F.nChi, F.nZ, iz, wz, ic, wc);
        c += cdot * etaDeltaT;
        rho = binterpDev(F.fields + tab, F.nChi, F.nZ, iz, wz, ic, wc);
#pragma unroll
        for (int k = 0; k < NPHI; ++k) { if (k == P.cIdx) phi[k] = c; if (k == P.pvIdx) phi[k] = pv; }
        for (int a = 0; a < P.nAdded; ++a) {
            const double v = binterpDev(F.fields + (size_t)(2 + a) * tab, F.nChi, F.nZ, iz, wz, ic, wc);
#pragma unroll
            for (int k = 0; k < NPHI; ++k) if (k == PRM_AT(P, addedIdx, a)) phi[k] = v;
        }
        Co = fmax(Co, cdot);
    }
}

// interpolationCellPointFace of every mid-point field at once: the four node
// records of the tet (points b and c, face centre, cell centre) blended with
// the clamped barycentric weights (order a, b, c, d); pst receives the four p*
// node values for gradInterpolationConstantTet (order a, b, c, d).
__device__ __forceinline__ void interpMid(const DMesh &M, const KParams &P, const double *nodeMid, int face, int ptB, int ptC, int cell,
                                          double w0, double w1, double w2, double w3, double (&val)[NODE_MID_STRIDE],
                                          double (&pst)[4]) {
    const size_t nNodes = (size_t)M.nCells + M.nPoints + M.nFaces;
    const size_t ib = (size_t)M.nCells + ptB, ic = (size_t)M.nCells + ptC, ia = (size_t)M.nCells + M.nPoints + face, id = (size_t)cell;
    const double fb = clamp01(w1), fc = clamp01(w2), fa = clamp01(w0), fd = clamp01(w3);
    const bool anyCell = (P.interpU == PDFB200_INTERP_CELL) | (P.interpk == PDFB200_INTERP_CELL) |
                         (P.interpDiffU == PDFB200_INTERP_CELL) | (P.interpOmega == PDFB200_INTERP_CELL) |
                         (P.interpEta == PDFB200_INTERP_CELL) | (P.interpZVar == PDFB200_INTERP_CELL);
    double cv[NODE_MID_STRIDE];
    // one 256-bit load per node and plane
#pragma unroll
    for (int k = 0; k < NODE_MID_STRIDE / 4; ++k) {
        const double *pl = nodeMid + (size_t)k * nNodes * 4;
        const D4 vb = ld256(pl + ib * 4), vc = ld256(pl + ic * 4), va = ld256(pl + ia * 4), vd = ld256(pl + id * 4);
        val[4 * k] = fma(fd, vd.x, fma(fa, va.x, fma(fc, vc.x, fb * vb.x)));
        val[4 * k + 1] = fma(fd, vd.y, fma(fa, va.y, fma(fc, vc.y, fb * vb.y)));
        val[4 * k + 2] = fma(fd, vd.z, fma(fa, va.z, fma(fc, vc.z, fb * vb.z)));
        val[4 * k + 3] = fma(fd, vd.w, fma(fa, va.w, fma(fc, vc.w, fb * vb.w)));
        cv[4 * k] = vd.x; cv[4 * k + 1] = vd.y; cv[4 * k + 2] = vd.z; cv[4 * k + 3] = vd.w;
        if (4 * k == NM_PSTAR) { pst[0] = va.x; pst[1] = vb.x; pst[2] = vc.x; pst[3] = vd.x; }
    }
    if (anyCell) {   // interpolation scheme "cell": the cell value
        if (P.interpU == PDFB200_INTERP_CELL) { val[NM_U] = cv[NM_U]; val[NM_U + 1] = cv[NM_U + 1]; val[NM_U + 2] = cv[NM_U + 2]; }
        if (P.interpk == PDFB200_INTERP_CELL) val[NM_K] = cv[NM_K];
        if (P.interpDiffU == PDFB200_INTERP_CELL) { val[NM_DU] = cv[NM_DU]; val[NM_DU + 1] = cv[NM_DU + 1]; val[NM_DU + 2] = cv[NM_DU + 2]; }
        if (P.interpOmega == PDFB200_INTERP_CELL) val[NM_OMEGA] = cv[NM_OMEGA];
        if (P.interpEta == PDFB200_INTERP_CELL) val[NM_ETA] = cv[NM_ETA];
        if (P.interpZVar == PDFB200_INTERP_CELL) val[NM_ZVAR] = cv[NM_ZVAR];
    }
}

// shared locate rule (oracle Mesh::locate): the tet of `cell` with the largest
// minimum barycentric coordinate, first maximum wins
__device__ inline int locateTet(const DMesh &M, V3 pos, int cell) {
    const V3 r = pos - cellCentre(M, cell);
    int best = -1;
    double bestMin = -PDF_VGREAT;
    const int ts = M.cellTetStart[cell], te = M.cellTetStart[cell + 1];
    for (int t = ts; t < te; ++t) {
        TetL T;
        loadTet(M, t, T);
        const double pa = dotf(T.g0, T.g1, T.g2, r), pb = dotf(T.g3, T.g4, T.g5, r), pc = dotf(T.g6, T.g7, T.g8, r);
        const double pd = 1.0 - ((pa + pb) + pc);
        const double mn = fmin(fmin(pa, pb), fmin(pc, pd));
        if (mn > bestMin) { bestMin = mn; best = t; }
    }
    return best;
}

// primitiveMesh::pointInCell (face-plane test)
__device__ inline bool pointInCell(const DMesh &M, V3 pt, int cell) {
    for (int j = M.cellFaceStart[cell]; j < M.cellFaceStart[cell + 1]; ++j) {
        const int f = M.cellFaces[j];
        const V3 proj = pt - ld3(M.faceCentres, f);
        V3 normal = ld3(M.faceAreas, f);
        normal = normal / mag3(normal);
        if (M.owner[f] != cell) normal = -normal;
        if (dot3(normal, proj) > 0) return false;
    }
    return true;
}

// mcParticleCloud::randomPointsInCell (mcParticleCloud.C:2040-2085), one point
__device__ inline V3 randomPointInCell(const DMesh &M, RngKey key, int cell, uint64_t entity, uint32_t purpose, uint32_t stepCtr) {
    const V3 minb = ld3(M.bbMin, cell);
    const V3 dimb = ld3(M.bbMax, cell) - minb;
    V3 position = minb;
    for (uint32_t attempt = 0;; ++attempt) {
        double ux, uy, uz, dummy;
        rngUniform2(key, entity, stepCtr, purpose, 2 * attempt, ux, uy);
        rngUniform2(key, entity, stepCtr, purpose, 2 * attempt + 1, uz, dummy);
        const double rx = fmin(fmax(10.0 * PDF_SMALL, ux), 1.0 - 10.0 * PDF_SMALL);
        const double ry = fmin(fmax(10.0 * PDF_SMALL, uy), 1.0 - 10.0 * PDF_SMALL);
        const double rz = fmin(fmax(10.0 * PDF_SMALL, uz), 1.0 - 10.0 * PDF_SMALL);
        position = minb + mk(rx * dimb.x, ry * dimb.y, rz * dimb.z);
        if (MESH_NGEOD(M) <= 2) constrainDir(MESH_GEOD3(M), position);
        if (pointInCell(M, position, cell) || attempt > 10000) break;
    }
    return position;
}

// computeCourantNo (mcParticleCloud.C:2120-2148); one 256-bit load per face of the cell
__device__ __forceinline__ double courantNo(const DMesh &M, int cell, V3 Utrack) {
    double Co = 0.;
    // every cell with the same number of faces (hex meshes): the row of the cell is known without looking it up
    const int nF = MESH_CELLFACES(M);
    const int js = nF > 0 ? cell * nF : __ldg(M.cellFaceStart + cell), je = nF > 0 ? js + nF : __ldg(M.cellFaceStart + cell + 1);
    for (int j = js; j < je; ++j) {
        const D4 c4 = ld256(M.cellCo4 + j);
        Co = fmax(Co, fabs(dot3(mk(c4.x, c4.y, c4.z), Utrack)));
    }
    return Co;
}

#ifndef K1_MIN_BLOCKS
#define K1_MIN_BLOCKS 4
#endif
#ifndef K1_COMPACT_ACC
#define K1_COMPACT_ACC 1
#endif
// parking rows (the correction velocity and the Courant number; parking position, velocity, mass and local time step as
// well instead of re-reading them through the permutation was measured slower: profiles/README.md)
#define PK_UCORR 0
#define PK_CO 3
// dynamic shared memory: thread-private accumulators + K1_NPARK thread-private parking rows
// (the correction velocity and the Courant number, which must survive the tet walks without occupying registers)
// The specialised kernel carries the sum rows of the conserved scalars the case has, not of K1_MAXCONS (k1_layout.h:
// K1_SMEM_ROWS(nConserved)); with one conserved scalar five CTAs then need 105 KB instead of 140 KB of shared memory,
// which moves the SM's carve-out from 164 to 132 KB and leaves 124 instead of 92 KB of L1 for the tet and node records.
#if defined(K1_JIT) && K1_COMPACT_ACC
#define KNC1 (1 + JitP::nConserved)
#else
#define KNC1 K1_NC1
#endif
#define KACC_DM_MINUS 0
#define KACC_DM_PLUS (KACC_DM_MINUS + KNC1)
#define KACC_MASS_IN (KACC_DM_PLUS + KNC1)
#define KACC_MASS_OUT (KACC_MASS_IN + KNC1)
#define KACC_N_DELETED (KACC_MASS_OUT + KNC1)
#define KACC_N_LOST (KACC_N_DELETED + 1)
#define KACC_N_ALIVE (KACC_N_LOST + 1)
#define KACC_NSUM (KACC_N_ALIVE + 1)
#define K1_SMEM_BYTES (K1_SMEM_ROWS(K1_MAXCONS) * K1_THREADS * sizeof(double))

// Loads of this thread's own state columns that the compiler must not merge with an earlier load
// of the same address: the kernel deliberately drops cold values (m, Omega, rho, phi, id, the old
// position and velocity) during the tet walks and reads them again where they are used — the
// lines are still in L1/L2 — so that the walks run with half the live registers of round 1 and
// more warps are resident to cover their dependent loads.
__device__ __forceinline__ double reloadD(const double *p) {
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
// the same for the input state, which the kernel never writes
__device__ __forceinline__ double reloadInD(const double *p) {
    return reloadD(p);
}
__device__ __forceinline__ int reloadI(const int *p) {
    int v;
    asm volatile("ld.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// The mass sums a warp has gathered in its thread-private rows (accS: row r of this thread at accS[r * K1_THREADS]),
// reduced over the lanes in a fixed order, stored as one row of the host's layout (4 * K1_NC1 values) and reset. Out of
// line: it runs once per chunk, and the particle kernel is sensitive to its instruction footprint.
__device__ __noinline__ void flushMassSums(double *accS, int lane, double *row) {
    constexpr int NS = 4 * KNC1, NSP2 = pow2Ceil(NS), NSSHIFT = log2Floor(32 / NSP2);
    double v[NSP2];
#pragma unroll
    for (int r = 0; r < NSP2; ++r) { v[r] = r < NS ? accS[(r < NS ? r : 0) * K1_THREADS] : 0.0; if (r < NS) accS[r * K1_THREADS] = 0.0; }
    const double x = laneSums<NSP2>(v, lane);
    const int r = lane >> NSSHIFT;
    if ((lane & ((1 << NSSHIFT) - 1)) == 0 && r < NS) row[(r / KNC1) * K1_NC1 + r % KNC1] = x;
}

// NPHI = compile-time bound on the particle scalars held in registers (nPhi <= NPHI)
template <int NPHI>
__device__ __forceinline__ void evolveBody(const K1Args &A) {
    extern __shared__ double smem[];
    double *accS = smem + threadIdx.x;                         // sums: accS[k*K1_THREADS]
    double *accM = smem + KACC_NSUM * K1_THREADS + threadIdx.x;  // maxima
    double *park = smem + (KACC_NSUM + ACC_NMAX) * K1_THREADS + threadIdx.x;   // park[k*K1_THREADS]: Ucorr
#pragma unroll
    for (int k = 0; k < KACC_NSUM; ++k) accS[k * K1_THREADS] = 0.0;
#pragma unroll
    for (int k = 0; k < ACC_NMAX; ++k) accM[k * K1_THREADS] = -PDF_VGREAT;

    if (A.ctaClock && threadIdx.x == 0) {   // written at once: no register may stay occupied through the kernel
        long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        A.ctaClock[3 * blockIdx.x] = t0;
    }
    const DMesh &M = A.M;
#ifdef K1_JIT
    const JitP P{};
    constexpr bool hasOpen = JIT_HAS_OPEN != 0;
#else
    const pdfb200_params &P = A.P;
    const bool hasOpen = A.hasOpen != 0;
#endif
    const int nSorted = A.sc->nSorted, total = nSorted + A.sc->nTail, nSlotsIn = A.sc->nSlots, nInjStart = A.sc->nInjStart;
    const int anyLost = A.sc->anyLost;
    const uint32_t step = A.sc->step;
    const double deltaT = A.sc->deltaT;
    const int nPhi = P.nPhi;
    const size_t nodeP = (size_t)M.nCells, nodeF = (size_t)M.nCells + M.nPoints;

    // Work distribution. Static (K1_DYNAMIC 0): the persistent CTAs stride over the entries in waves of
    // gridDim.x * 128, every CTA the same number of trips, the four warps of a CTA side by side on 128 consecutive
    // entries. Measured with pdfb200_profile_ctas: the five CTAs of an SM do not advance at the same rate (the warp
    // schedulers favour the older warps), the CTA launched last on an SM needs 4-17 % longer for its share and the
    // kernel ends with a tail of a few busy SMs - 14 % of its duration once number control has widened the spread of
    // the per-cell populations (24.8 -> 21.3 ms there; on the freshly seeded cloud, where every cell holds exactly 64
    // particles and the warps of a CTA share their cells' records in L1, the static waves are 2.5 % faster).
    // Dynamic (K1_DYNAMIC 1, default): a warp claims the next K1_CHUNK entries from a counter (one atomic per chunk,
    // issued one chunk ahead), so that slower warps take fewer chunks and all warps end within one chunk of each other;
    // chunks are handed out in index order, which keeps the warps of the whole device on one moving window of the
    // Morton-ordered cells. (Static waves for the first 13/16 of the entries and chunks for the rest: no better. Blocks
    // of 128 entries claimed per CTA through a ring in shared memory, so that its four warps stay on neighbouring
    // entries: 19.9 / 21.8 ms, no better either - profiles/README.md.)
    // The mass sums of the report must not depend on which warp took which chunk: a warp reduces them over its lanes
    // in a fixed order at the end of every chunk and stores a row per chunk (k_chunk_fold adds the rows in a fixed
    // order); counts and maxima are exact in any order and stay in the thread-private rows.
    // The slot of a sorted entry comes through the permutation of the last sort: requested one trip ahead so
    // that the state loads of a trip do not start with a dependent load. The request is unconditional (index
    // clamped into the sorted range): a predicated load would be merged with the old value by a select right
    // behind it, which waits for the load at once (ncu: 3.7 % of the stall samples of the v3 kernel).
    const int permLast = max(nSorted - 1, 0);
    const int lane = threadIdx.x & 31;
#if K1_DYNAMIC
    // warp-uniform: cpos = first entry of this trip, cend = end of the chunk in hand (0: none yet), cfetch = lane 0's
    // claim of the chunk after it
    int cpos = 0, cend = 0, cfetch = 0;
    if (lane == 0) cfetch = atomicAdd(A.workCtr, K1_CHUNK);
    int slotNext = (int)A.perm[min(__shfl_sync(0xffffffffu, cfetch, 0) + lane, permLast)];
    for (;;) {
        if (cpos == cend) {
            if (cend > 0) flushMassSums(accS, lane, A.chunkSums + (size_t)(cend / K1_CHUNK - 1) * (4 * K1_NC1));
            cpos = __shfl_sync(0xffffffffu, cfetch, 0);
            if (cpos >= total) break;
            cend = cpos + K1_CHUNK;
            if (lane == 0) cfetch = atomicAdd(A.workCtr, K1_CHUNK);   // the chunk after this one, claimed ahead
        }
        const int i = cpos + lane;
        cpos += 32;
        const int slotPerm = slotNext;
        // (on the last trip of a chunk this waits for the claim made at the chunk's first trip)
        slotNext = (int)A.perm[min((cpos == cend ? __shfl_sync(0xffffffffu, cfetch, 0) : cpos) + lane, permLast)];
#else
    const int gstride = gridDim.x * blockDim.x;
    int slotNext = (int)A.perm[min((int)(blockIdx.x * blockDim.x + threadIdx.x), permLast)];
    for (int base = blockIdx.x * blockDim.x; base < total; base += gstride) {
        const int i = base + threadIdx.x;
        const int slotPerm = slotNext;
        slotNext = (int)A.perm[min(i + gstride, permLast)];
#endif
        if (i >= total) continue;
        const int slot = (i < nSorted) ? slotPerm : nSlotsIn + (i - nSorted);
        const int cell0 = A.in.cell[slot];
        if (cell0 < 0) { A.out.cell[i] = -1; continue; }
        V3 pos = mk(A.in.px[slot], A.in.py[slot], A.in.pz[slot]);
        const int tet0 = A.in.tet[slot];
        int tet = tet0;
        const bool injected = (i >= nSorted + nInjStart);

        // the particle's mass at the start of the step: the stored one plus its share of the mass lost in the
        // previous step (mcParticleCloud.C:1363-1369; particles injected since then get none)
        auto massIn = [&]() {
            double mm = reloadInD(A.in.m + slot);
            if (anyLost && !injected) mm += A.lostShare[cell0];
            return mm;
        };
        int status = 0;   // 0 alive, 1 deleted at an open boundary, 2 lost
        // ---- E1 + E2 need the mass and the conserved scalars; both are read again at the mid-point ----
        {
            const double m = massIn();
            // ---- E1: mcOpenBoundary::correct(false), cull behind the moving outflow plane ----
            if (hasOpen) {
                const int injPatch = injected ? A.injPatchTail[i - nSorted] : -1;
                const int ks = __ldg(M.openCellStart + cell0), ke = __ldg(M.openCellStart + cell0 + 1);
                for (int k = ks; k < ke && status == 0; ++k) {
                    const int bf = M.openCellFaces[k];
                    const int info = M.bfInfo[bf];
                    if (BF_PATCH(info) <= injPatch || !BF_REFLECTING(info)) continue;
                    const double Un = A.Un[bf];
                    if (!(Un < 0.)) {
                        const double4 n4 = M.bfNormal[bf];
                        const V3 Cf = ld3(M.faceCentres, M.nInternalFaces + bf);
                        const double xp = dot3(mk(n4.x, n4.y, n4.z), Cf - pos);
                        if (xp < A.in.eta[slot] * (Un * deltaT)) {
                            const double mpd = massPerDepth(M, pos, m);
                            accS[KACC_MASS_OUT * K1_THREADS] -= mpd;
                            for (int cs = 0; cs < P.nConserved; ++cs)
                                accS[(KACC_MASS_OUT + 1 + cs) * K1_THREADS] -= mpd * A.in.phi[PRM_AT(P, conserved, cs)][slot];
                            status = 1;
                        }
                    }
                }
            }
            if (status != 0) { accS[KACC_N_DELETED * K1_THREADS] += 1; A.out.cell[i] = -1; continue; }
            // ---- E2 ----
            const double mpd = massPerDepth(M, pos, m);
            accS[KACC_DM_MINUS * K1_THREADS] -= mpd;
            for (int cs = 0; cs < P.nConserved; ++cs)
                accS[(KACC_DM_MINUS + 1 + cs) * K1_THREADS] -= mpd * A.in.phi[PRM_AT(P, conserved, cs)][slot];
        }
        // the walk state: tet record, the cell the tet belongs to and its centre (vertex d)
        TetL T;
        loadTet(M, tet, T);
        int cell = cell0;
        V3 ctr = cellCentre(M, cell0);
        // ---- E3: position correction gather ----
        V3 Utrack;
        {
            V3 U = mk(A.in.ux[slot], A.in.uy[slot], A.in.uz[slot]);
            V3 Ucorr = mk(0, 0, 0);
            if (P.positionCorrection != PDFB200_POSCORR_NONE) {
                // the interpolated vector: limitedSimple UPosCorr (Ucorrection += it, .C:187-192), simple grad(phi), external G0
                V3 g0 = mk(0, 0, 0);
                double fb = 0, fc = 0, fa = 0, fd = 0;
                const bool blend = P.interpUPosCorr != PDFB200_INTERP_CELL;
                if (!blend) {
                    const D4 q = ld256(A.nodePC + (size_t)cell0 * 4);
                    g0 = mk(q.x, q.y, q.z);
                } else {
                    double a0, a1, a2, a3;
                    baryAt(T, pos - ctr, a0, a1, a2, a3);
                    fb = clamp01(a1); fc = clamp01(a2); fa = clamp01(a0); fd = clamp01(a3);
                    loadPts(M, tet, T);
                    // 32-byte node records: one 256-bit load per node
                    const D4 vb = ld256(A.nodePC + (nodeP + T.ptB) * 4), vc = ld256(A.nodePC + (nodeP + T.ptC) * 4);
                    const D4 va = ld256(A.nodePC + (nodeF + T.face) * 4), vd = ld256(A.nodePC + (size_t)cell0 * 4);
                    g0.x = fma(fd, vd.x, fma(fa, va.x, fma(fc, vc.x, fb * vb.x)));
                    g0.y = fma(fd, vd.y, fma(fa, va.y, fma(fc, vc.y, fb * vb.y)));
                    g0.z = fma(fd, vd.z, fma(fa, va.z, fma(fc, vc.z, fb * vb.z)));
                }
                if (P.positionCorrection == PDFB200_POSCORR_LIMITED_SIMPLE) {
                    Ucorr = g0;
                } else if (P.positionCorrection == PDFB200_POSCORR_SIMPLE) {
                    // mcSimplePositionCorrection::correct (.C:156-163): Ucorrection -= cmptMultiply(Ainv[cell], interp(gradPhi)),
                    // Ainv = (1./dy*dz, 1./dx*dz, 1./dx*dy) of the cell's bounding box (.C:127, precedence as written)
                    const V3 d = ld3(M.bbMax, cell0) - ld3(M.bbMin, cell0);
                    Ucorr = mk(0, 0, 0) - mk((1. / d.y * d.z) * g0.x, (1. / d.x * d.z) * g0.y, (1. / d.x * d.y) * g0.z);
                } else {
                    // external (pdfb200_poscorr_fields): Ucorrection -= G0 + scale l (zeta != 0 ? G1 : G2); planes 1 and 2 of
                    // the node records hold (G1, zeta) and (G2, 0), plane 0 (G0, l)
                    const size_t nNodes = (size_t)M.nCells + M.nPoints + M.nFaces;
                    const double *p1 = A.nodePC + nNodes * 4, *p2 = A.nodePC + 2 * nNodes * 4;
                    V3 g1, g2; double l, zeta;
                    if (!blend) {
                        const D4 q0 = ld256(A.nodePC + (size_t)cell0 * 4), q1 = ld256(p1 + (size_t)cell0 * 4), q2 = ld256(p2 + (size_t)cell0 * 4);
                        l = q0.w; g1 = mk(q1.x, q1.y, q1.z); zeta = q1.w; g2 = mk(q2.x, q2.y, q2.z);
                    } else {
                        const size_t ib = nodeP + T.ptB, ic = nodeP + T.ptC, ia = nodeF + T.face, id = (size_t)cell0;
                        const D4 lb = ld256(A.nodePC + ib * 4), lc = ld256(A.nodePC + ic * 4), la = ld256(A.nodePC + ia * 4), ld = ld256(A.nodePC + id * 4);
                        l = fma(fd, ld.w, fma(fa, la.w, fma(fc, lc.w, fb * lb.w)));
                        const D4 vb = ld256(p1 + ib * 4), vc = ld256(p1 + ic * 4), va = ld256(p1 + ia * 4), vd = ld256(p1 + id * 4);
                        g1.x = fma(fd, vd.x, fma(fa, va.x, fma(fc, vc.x, fb * vb.x)));
                        g1.y = fma(fd, vd.y, fma(fa, va.y, fma(fc, vc.y, fb * vb.y)));
                        g1.z = fma(fd, vd.z, fma(fa, va.z, fma(fc, vc.z, fb * vb.z)));
                        zeta = fma(fd, vd.w, fma(fa, va.w, fma(fc, vc.w, fb * vb.w)));
                        const D4 wb = ld256(p2 + ib * 4), wc = ld256(p2 + ic * 4), wa = ld256(p2 + ia * 4), wd = ld256(p2 + id * 4);
                        g2.x = fma(fd, wd.x, fma(fa, wa.x, fma(fc, wc.x, fb * wb.x)));
                        g2.y = fma(fd, wd.y, fma(fa, wa.y, fma(fc, wc.y, fb * wb.y)));
                        g2.z = fma(fd, wd.z, fma(fa, wa.z, fma(fc, wc.z, fb * wb.z)));
                    }
                    Ucorr = mk(0, 0, 0) - (g0 + (A.pcScale * l) * (zeta != 0 ? g1 : g2));
                }
            }
            // ---- E4 ----
            Utrack = U + Ucorr;
            constrainParticle(M, deltaT / 2, pos, U, Ucorr, Utrack);
            park[0] = Ucorr.x; park[K1_THREADS] = Ucorr.y; park[2 * K1_THREADS] = Ucorr.z;
            // the velocity after E4 ("Uold") equals the stored one except on axisymmetric meshes, where
            // constrainParticle has rotated it: parked in this particle's (still unwritten) output slot
            if (MESH_AXISYM(M)) { A.out.ux[i] = U.x; A.out.uy[i] = U.y; A.out.uz[i] = U.z; }
        }
        bool reflected = false;
        int reflBf = -1;   // boundary face of the last open-boundary reflection (mcOpenBoundary.C:201-204)

        // ---- E5 (phase 0: eta*deltaT/2) and E10 (phase 1: eta*deltaT): mcParticle::move on the
        //      tet walk (see oracle/src/cloud.cpp Cloud::move; identical expressions) ----
#pragma unroll 1
        for (int phase = 0; phase < 2; ++phase) {
            double tEnd;
            {
                double stepFraction = 0.0;
                if (phase == 0 && injected) { double u0, u1; rngUniform2(A.key, A.in.id[slot], step, RNG_INLET_STEPFRAC, 0, u0, u1); stepFraction = u0; }
                // phase 0 tracks with the old local time step, phase 1 with the new one (written to out.eta at the mid-point)
                const double eta = (phase == 0) ? reloadInD(A.in.eta + slot) : reloadD(A.out.eta + i);
                const double trackTime = (phase == 0) ? eta * (deltaT / 2.) : eta * deltaT;
                tEnd = (1.0 - stepFraction) * trackTime;
            }
            int entry = -1, nSteps = 0;
            while (tEnd > 0 && status == 0) {
                // mcParticle.C:137: dt = min(dtMax, tEnd) with dtMax the initial tEnd; tEnd only decreases, so this is tEnd
                double dt = tEnd;
                V3 dx;
                {
                    V3 dest = pos + dt * Utrack;
                    constrainDir(MESH_GEOD3(M), dest);
                    dx = dest - pos;
                }
                double tf = 1.0;
                // start point relative to the centre of the current cell: constant along the sub-step until the walk
                // changes cell; the centre itself is not kept through the loop (read again after it)
                V3 r = pos - ctr;
                for (;;) {
                    r = pos - ctr;
                    double p0, p1, p2, p3;
                    baryAt(T, r, p0, p1, p2, p3);
                    const double q0 = dotf(T.g0, T.g1, T.g2, dx), q1 = dotf(T.g3, T.g4, T.g5, dx), q2 = dotf(T.g6, T.g7, T.g8, dx);
                    const double q3 = -((q0 + q1) + q2);
                    // exit face: two-round tournament written with selects (no branches) - a against b, c against d, then
                    // the winners; a later face wins only with a strictly smaller crossing parameter p/(-q), compared
                    // cross-multiplied (oracle Cloud::move: the same expressions). Candidate i: not the entry face, moving
                    // towards it, and its plane lies before dest
                    const bool c0 = (entry != 0) & (q0 < 0) & (p0 + q0 < 0);
                    const bool c1 = (entry != 1) & (q1 < 0) & (p1 + q1 < 0);
                    const bool c2 = (entry != 2) & (q2 < 0) & (p2 + q2 < 0);
                    const bool c3 = (entry != 3) & (q3 < 0) & (p3 + q3 < 0);
                    const bool w1 = c1 & (!c0 | (p1 * q0 > p0 * q1));
                    const bool w3 = c3 & (!c2 | (p3 * q2 > p2 * q3));
                    const double pA = w1 ? p1 : p0, qA = w1 ? q1 : q0, pB = w3 ? p3 : p2, qB = w3 ? q3 : q2;
                    const bool cA = c0 | c1, cB = c2 | c3;
                    const bool wB = cB & (!cA | (pB * qA > pA * qB));
                    const int best = (cA | cB) ? (wB ? (w3 ? 3 : 2) : (w1 ? 1 : 0)) : -1;
                    if (best < 0) {
                        // the destination is recomputed (same expression, same bits) instead of kept in registers
                        V3 dest = pos + dt * Utrack;
                        constrainDir(MESH_GEOD3(M), dest);
                        pos = dest;
                        tf = 1.0;
                        break;
                    }
                    ++nSteps;
                    if (nSteps > 1000) { status = 2; break; }
                    if (best != 3 || T.n3 >= 0) {   // n3 < 0 marks a boundary face
                        int nb = (best == 0) ? T.n0 : T.n1;
                        nb = (best == 2) ? T.n2 : nb;
                        tet = (best == 3) ? T.n3 : nb;
                        // across the mesh face: the cell changes, and with it the centre the tets hang on
                        // (requested together with the next record, not after it)
                        if (best == 3) { cell = T.nbrCell; ctr = cellCentre(M, cell); }
                        loadTet(M, tet, T);
                        entry = best ^ (((best == 1) | (best == 2)) ? 3 : 0);   // the shared face is opposite c (b) in the tet across b (c)
                        continue;
                    }
                    // boundary face
                    double s = p3 / (-q3);
                    s = fmin(1.0, fmax(0.0, s));
                    pos = pos + s * dx;
                    tf = s;
                    const int bf = T.face - M.nInternalFaces;
                    const int info = __ldg(M.bfInfo + bf);
                    const int handler = BF_HANDLER(info);
                    if (MESH_NGEOD(M) < 3 && BF_GEOM(info) == PDFB200_PATCH_WEDGE) {
                        // mcParticle.C:220-224: a wedge patch snaps the position back onto the centre plane and the
                        // track carries on from there, in the tet the shared locate rule gives for the snapped point
                        constrainDir(MESH_GEOD3(M), pos);
                        tet = locateTet(M, pos, cell);
                        loadTet(M, tet, T);
                        entry = -1;
                        break;
                    }
                    // an empty patch aborts the reference (mcEmptyBoundary.C:61-104); a wedge patch of a 3-D mesh makes it
                    // spin past 1000 sub-steps and drop the particle (mcParticle.C:147-158): both end as a lost particle
                    if (BF_GEOM(info) != PDFB200_PATCH_GENERIC || handler == PDFB200_BC_EMPTY) { status = 2; break; }
                    const double4 n4 = M.bfNormal[bf];
                    const V3 nw = mk(n4.x, n4.y, n4.z);
                    if (handler != PDFB200_BC_SLIP) {
                        const double Un = A.Un[bf];
                        if (Un < 0 || !BF_REFLECTING(info)) {
                            // the particle leaves through an open face: its mass and scalars as they are at this point
                            const double mpd = massPerDepth(M, pos, massIn());
                            const int baseAcc = (Un < 0) ? KACC_MASS_IN : KACC_MASS_OUT;
                            accS[baseAcc * K1_THREADS] -= mpd;
                            for (int cs = 0; cs < P.nConserved; ++cs) {
                                const int sIdx = PRM_AT(P, conserved, cs);
                                const double ph = (phase == 0) ? reloadInD(A.in.phi[sIdx] + slot) : reloadD(A.out.phi[sIdx] + i);
                                accS[(baseAcc + 1 + cs) * K1_THREADS] -= mpd * ph;
                            }
                            status = 1;
                            break;
                        }
                        reflBf = bf;
                    }
                    // reflection (rare): the particle velocity and the correction velocity are not held in registers
                    // during the walk. Phase 0: U lives in out.p* (parked there by the first reflection), Ucorr in
                    // shared memory. Phase 1: U is this step's new velocity in out.u*.
                    {
                        V3 Uc = mk(park[0], park[K1_THREADS], park[2 * K1_THREADS]);
                        Uc = reflectVec(Uc, nw);
                        park[0] = Uc.x; park[K1_THREADS] = Uc.y; park[2 * K1_THREADS] = Uc.z;
                    }
                    if (phase == 0) {
                        V3 Ur;
                        if (reflected) Ur = mk(reloadD(A.out.px + i), reloadD(A.out.py + i), reloadD(A.out.pz + i));
                        else if (MESH_AXISYM(M)) Ur = mk(reloadD(A.out.ux + i), reloadD(A.out.uy + i), reloadD(A.out.uz + i));
                        else Ur = mk(reloadInD(A.in.ux + slot), reloadInD(A.in.uy + slot), reloadInD(A.in.uz + slot));
                        Ur = reflectVec(Ur, nw);
                        A.out.px[i] = Ur.x; A.out.py[i] = Ur.y; A.out.pz[i] = Ur.z;
                    } else {
                        V3 Ur = mk(reloadD(A.out.ux + i), reloadD(A.out.uy + i), reloadD(A.out.uz + i));
                        Ur = reflectVec(Ur, nw);
                        A.out.ux[i] = Ur.x; A.out.uy[i] = Ur.y; A.out.uz[i] = Ur.z;
                    }
                    Utrack = reflectVec(Utrack, nw);
                    reflected = true;
                    entry = 3;
                    break;
                }
                dt *= tf;
                tEnd -= dt;
            }
            if (status != 0) break;
            if (phase == 1) break;
            // barycentric coordinates of the mid-point in its tet (oracle Cloud::baryWeights: the same expression)
            double w0, w1, w2, w3;
            baryAt(T, pos - ctr, w0, w1, w2, w3);
            loadPts(M, tet, T);

            // ================= mid-point: E6, E7, E9 =================
            double Co = courantNo(M, cell, Utrack);
            double val[NODE_MID_STRIDE];
            V3 gradP;
            {
                double pst[4];
                interpMid(M, P, A.nodeMid, T.face, T.ptB, T.ptC, cell, w0, w1, w2, w3, val, pst);
                // gradInterpolationConstantTet: g_d*cell + g_b*ptB + g_c*ptC + g_a*face (the record is dead afterwards)
                const V3 ga = mk(T.g0, T.g1, T.g2), gb = mk(T.g3, T.g4, T.g5), gc = mk(T.g6, T.g7, T.g8);
                const V3 gdv = mk(T.gd0, T.gd1, T.gd2);
                gradP = (gdv * pst[3] + gb * pst[1]) + gc * pst[2];
                gradP = gradP + ga * pst[0];
            }
            const double *aux = A.cellAux + (size_t)cell * A.auxStride;
            // the particle's scalars, read again (see reloadD)
            const double etaOld = reloadInD(A.in.eta + slot);
            double Omega = reloadInD(A.in.omega + slot), rho = reloadInD(A.in.rho + slot);
            double phi[NPHI];
#pragma unroll
            for (int k = 0; k < NPHI; ++k) phi[k] = (k < nPhi) ? reloadInD(A.in.phi[k] + slot) : 0.0;
            const uint64_t id = (uint64_t)__double_as_longlong(reloadInD(reinterpret_cast<const double *>(A.in.id) + slot));
            // mcRASOmegaModel::correct
            if (P.omegaModel != PDFB200_OMEGA_NONE) Omega = fmax(val[NM_OMEGA], PDF_SMALL);
            // mcIEMMixingModel::correct
            if (P.mixingModel == PDFB200_MIXING_IEM) {
                const double Cmix2 = 0.5 * P.Cmix;
                const double fac = 1.0 - exp(-Cmix2 * Omega * etaOld * deltaT);
                for (int mI = 0; mI < P.nMixed; ++mI) {
                    const int s = PRM_AT(P, mixed, mI);
                    const double mean = __ldg(aux + 2 + s);
#pragma unroll
                    for (int k = 0; k < NPHI; ++k)
                        if (k == s) phi[k] -= fac * (phi[k] - mean);
                }
            }
            // reaction model
            {
                FlameletTables FT;
                FT.chi = A.flChi; FT.z = A.flZ; FT.fields = A.flFields; FT.nChi = A.nChi; FT.nZ = A.nZ; FT.cEq = A.flCEq; FT.cEqMax = A.cEqMax;
                reactionCorrect(P, FT, Omega, val[NM_ZVAR], aux[1], etaOld * deltaT, phi, rho, Co);
            }
            // the scalars have their final values: written now, not carried through the second walk
            A.out.omega[i] = Omega; A.out.rho[i] = rho;
#pragma unroll
            for (int k = 0; k < NPHI; ++k) if (k < nPhi) A.out.phi[k][i] = phi[k];
            A.out.id[i] = id;
            A.out.m[i] = massIn();
            // the six normals of the step: two Philox calls, single-precision Box-Muller (common.cuh)
            // (= rngNormal4(sub 0) -> xi0 xi1 xi2 gd0, rngNormal2(sub 1) -> gd1 gd2; written as a rolled loop so that
            // the kernel holds one copy of the Philox rounds and of the Box-Muller code: its instruction footprint
            // is what the round-1 kernel was most sensitive to)
            double xi0 = 0, xi1 = 0, xi2 = 0, gd0 = 0, gd1 = 0, gd2 = 0;
            {
                uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
#pragma unroll 1
                for (int pair = 0; pair < 3; ++pair) {
                    if (pair != 1) {
                        uint32_t o[4];
                        rngWords(A.key, id, step, RNG_STEP_NORMALS, (uint32_t)(pair >> 1), o);
                        o0 = o[0]; o1 = o[1]; o2 = o[2]; o3 = o[3];
                    }
                    float na, nb;
                    normalPairF32(pair == 1 ? o2 : o0, pair == 1 ? o3 : o1, na, nb);
                    if (pair == 0) { xi0 = (double)na; xi1 = (double)nb; }
                    else if (pair == 1) { xi2 = (double)na; gd0 = (double)nb; }
                    else { gd1 = (double)na; gd2 = (double)nb; }
                }
            }
            // the velocity before the half step ("Uold") and the velocity the half step ended with
            V3 Uold;
            if (MESH_AXISYM(M)) Uold = mk(reloadD(A.out.ux + i), reloadD(A.out.uy + i), reloadD(A.out.uz + i));
            else Uold = mk(reloadInD(A.in.ux + slot), reloadInD(A.in.uy + slot), reloadInD(A.in.uz + slot));
            V3 U = Uold;
            if (reflected) U = mk(reloadD(A.out.px + i), reloadD(A.out.py + i), reloadD(A.out.pz + i));
            // mcSLMFullVelocityModel::correct
            if (P.velocityModel != PDFB200_VELOCITY_NONE) {
                const double dTp = etaOld * deltaT;
                const V3 UFap = mk(val[NM_U], val[NM_U + 1], val[NM_U + 2]);
                const double kFap = fmax(val[NM_K], P.kMin);
                const V3 diffUap = mk(val[NM_DU], val[NM_DU + 1], val[NM_DU + 2]);
                const double Acoef = -(0.5 * P.C1 + 0.75 * P.C0) * Omega;
                const V3 B = -(gradP / rho + Acoef * UFap);   // mcSLMFullVelocityModel.C:174
                const double sq = sqrt(P.C0 * kFap * Omega * dTp);
                const V3 deltaU = (B + Acoef * U) * dTp + sq * mk(xi0, xi1, xi2);
                const V3 Ubefore = U;
                U = U + (((deltaU + 0.5 * Acoef * deltaU * dTp) + diffUap * deltaT) + (Ubefore - UFap) * aux[0] * deltaT);
                Co = fmax(Co, Omega);
            }
            park[PK_CO * K1_THREADS] = Co;
            // local time stepping
            double eta;
            if (P.localTimeStepping == PDFB200_LTS_CELL) eta = val[NM_ETA];
            else if (P.localTimeStepping == PDFB200_LTS_PARTICLE) eta = fmin(P.CFL / stabilise(deltaT * Co, PDF_VSMALL), P.ltsUpperBound);
            else eta = 1.0;
            A.out.eta[i] = eta;

            // ---- E9: mid-point rule from the old position ----
            pos = mk(reloadInD(A.in.px + slot), reloadInD(A.in.py + slot), reloadInD(A.in.pz + slot));
            tet = tet0;
            loadTet(M, tet, T);
            cell = cell0;
            double hNum;   // numerical-diffusion length of the start cell (mcParticleCloud.C:1937-1971)
            ctr = cellCentreH(M, cell0, hNum);
            if (reflected) Utrack = Uold; else Utrack = 0.5 * (Uold + U);
            const double C = P.DNum * sqrt(hNum * deltaT * mag3(Utrack)) / deltaT;
            Utrack = Utrack + mk(C * gd0, C * gd1, C * gd2);
            {
                V3 Ucorr = mk(park[0], park[K1_THREADS], park[2 * K1_THREADS]);
                Utrack = Utrack + Ucorr;
                constrainParticle(M, deltaT, pos, U, Ucorr, Utrack);
                park[0] = Ucorr.x; park[K1_THREADS] = Ucorr.y; park[2 * K1_THREADS] = Ucorr.z;
            }
            A.out.ux[i] = U.x; A.out.uy[i] = U.y; A.out.uz[i] = U.z;
            reflected = false;   // from here on: "reflected during the full step" (not used further)
        }
        if (status != 0) {
            if (status == 2) {
                // lost (mcParticleCloud.C:1257-1260, 1346-1351, notifyLostParticle :2088-2092): its mass goes on the
                // list that k_lost_reduce sums per cell in a fixed order (no floating-point atomics). Lost during
                // the half step: the mass has not been copied to `out` yet, so it is read from `in` either way.
                const double m = massIn();
                const unsigned k = atomicAdd(A.lostCount, 1u);
                if (k < A.lostCap) { LostRec r; r.idx = (uint32_t)i; r.cell = cell; r.m = m; A.lostList[k] = r; }
                accS[KACC_N_LOST * K1_THREADS] += 1;
            } else accS[KACC_N_DELETED * K1_THREADS] += 1;
            A.out.cell[i] = -1;
            continue;
        }
        // ---- E11: mcOpenBoundary::correct(true) ----
        V3 U = mk(reloadD(A.out.ux + i), reloadD(A.out.uy + i), reloadD(A.out.uz + i));
        if (reflBf >= 0) {
            const double4 n4 = M.bfNormal[reflBf];
            const V3 reflBV = mk(n4.x, n4.y, n4.z) * A.Un[reflBf];
            U = U + 2 * reflBV;
            A.out.ux[i] = U.x; A.out.uy[i] = U.y; A.out.uz[i] = U.z;
        }

        // ---- write back in sorted order (the scalars were written at the mid-point) ----
        A.out.px[i] = pos.x; A.out.py[i] = pos.y; A.out.pz[i] = pos.z;
        A.out.cell[i] = cell; A.out.tet[i] = tet;
        // (the sort key — rank of the cell in the processing order — is formed by the sort kernels from out.cell and
        // out.tet, so that this kernel does not end on a dependent look-up)

        // ---- E15 particle sums / maxima (before particle-number control) ----
        {
            const double m = reloadD(A.out.m + i), eta = reloadD(A.out.eta + i);
            const double mpd = massPerDepth(M, pos, m);
            accS[KACC_DM_PLUS * K1_THREADS] += mpd;
            for (int cs = 0; cs < P.nConserved; ++cs)
                accS[(KACC_DM_PLUS + 1 + cs) * K1_THREADS] += mpd * reloadD(A.out.phi[PRM_AT(P, conserved, cs)] + i);
            accS[KACC_N_ALIVE * K1_THREADS] += 1;
            const V3 Ucorr = mk(park[0], park[K1_THREADS], park[2 * K1_THREADS]);
            accM[ACC_UMAX * K1_THREADS] = fmax(accM[ACC_UMAX * K1_THREADS], mag3(U));
            accM[ACC_UCORRMAX * K1_THREADS] = fmax(accM[ACC_UCORRMAX * K1_THREADS], mag3(Ucorr));
            accM[ACC_ETAMAX * K1_THREADS] = fmax(accM[ACC_ETAMAX * K1_THREADS], eta);
            accM[ACC_NEG_ETAMIN * K1_THREADS] = fmax(accM[ACC_NEG_ETAMIN * K1_THREADS], -eta);
            accM[ACC_COMAX * K1_THREADS] = fmax(accM[ACC_COMAX * K1_THREADS], park[PK_CO * K1_THREADS]);
        }
    }

    // ---- deterministic CTA reduction of the thread-private accumulators ----
    __syncthreads();
    if (A.ctaClock && threadIdx.x == 0) {
        long long t1; unsigned sm;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        A.ctaClock[3 * blockIdx.x + 1] = t1; A.ctaClock[3 * blockIdx.x + 2] = (long long)sm;
    }
    {
        const int warp = threadIdx.x >> 5;
        // warp w reduces rows w, w+4, ... of the host's layout (k1_layout.h); the rows of conserved scalars the
        // specialised kernel does not carry are written as zeros
        for (int k = warp; k < ACC_NSUM + ACC_NMAX; k += K1_THREADS / 32) {
            const bool isMax = k >= ACC_NSUM;
            int row;   // row of this kernel's accumulators, -1: none
            if (isMax) row = KACC_NSUM + (k - ACC_NSUM);
            else if (k >= 4 * K1_NC1) row = 4 * KNC1 + (k - 4 * K1_NC1);
            else row = (k % K1_NC1 < KNC1) ? (k / K1_NC1) * KNC1 + k % K1_NC1 : -1;
            double x = isMax ? -PDF_VGREAT : 0.0;
            if (row >= 0) {
                for (int t = lane; t < K1_THREADS; t += 32) {
                    const double y = smem[row * K1_THREADS + t];
                    x = isMax ? fmax(x, y) : x + y;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double y = __shfl_down_sync(0xffffffffu, x, o);
                    x = isMax ? fmax(x, y) : x + y;
                }
            }
            if (lane == 0) {
                if (isMax) A.partMax[blockIdx.x * ACC_NMAX + (k - ACC_NSUM)] = x;
                else A.partSum[blockIdx.x * ACC_NSUM + k] = x;
            }
        }
    }
}

#ifndef K1_JIT
template <int NPHI>
__global__ void __launch_bounds__(K1_THREADS, K1_MIN_BLOCKS) k_evolve(const K1Args A) { evolveBody<NPHI>(A); }
#else
// the specialised kernel of this cloud (loaded by jit.cu through cudaLibraryLoadFromFile)
extern "C" __global__ void __launch_bounds__(K1_THREADS, K1_MIN_BLOCKS) k_evolve_jit(const K1Args A) { evolveBody<JIT_NPHI>(A); }
#endif
