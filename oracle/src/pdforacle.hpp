// pdforacle.hpp — CPU ORACLE (test infrastructure, not product code).
//
// A single-threaded C++ restatement of pdfFoam's mcParticleCloud::evolve and
// everything it calls, written against plain arrays instead of OpenFOAM.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
// legs may load this library; the product (pdffoam_b200/csrc) never does.
//
// PARITY STATUS: pinned in part by the reference's own code compiled here, unpinned for the rest.
//   * PINNED (oracle/_ref/libpdfref.so, built by oracle/ref/Makefile from the unmodified sources under
//     /root/reference against a stand-in for the OpenFOAM headers; compared bit for bit by
//     tests/test_oracle_vs_ref.py): findCoeffs/binterp (mcSteadyFlamelet.C:44-93), mcInletRandom::
//     updateCoeffs/PDF/CDF + mcInversionInletRandom::newton/value, richTetrahedron::update/inside,
//     tetFacePointCellDecomposition (order, face, point pair of every tet, cellTetrahedra, find).
//   * UNPINNED ("parity unpinned"): evolve() as a whole. pdfFoam needs OpenFOAM 1.7.x/2.1.x and its wmake
//     build, neither of which exists here, and the reference ships no golden vectors.
//     What holds that part instead: closed-form known answers derived from the cited reference lines
//     (tests/test_oracle_known_answers.py, test_oracle_evolve.py), the reference's own flamelet tables and
//     the analytic inlet PDF/CDF used by tests/mcInletRandom, independent numpy restatements
//     (tests/golden/make_golden.py).
// OpenFOAM primitives (trackToFace, interpolationCellPointFace,
// volPointInterpolation, linearInterpolate, fvc::grad, Random) are *defined*
// here per SURVEY.md Appendix B.
//
// Each function cites the reference file:line it follows (paths relative to
// the pdfFoam tree).
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/pdfb200.h"

namespace pdforacle {

// OpenFOAM double-precision constants
constexpr double SMALL = 1e-15;
constexpr double VSMALL = 1e-300;
constexpr double GREAT = 1e15;
constexpr double VGREAT = 1e300;

struct Vec3 {
    double x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
    double &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 operator+(const Vec3 &a, const Vec3 &b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(const Vec3 &a, const Vec3 &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator-(const Vec3 &a) { return {-a.x, -a.y, -a.z}; }
inline Vec3 operator*(double s, const Vec3 &a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 operator*(const Vec3 &a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator/(const Vec3 &a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline Vec3 &operator+=(Vec3 &a, const Vec3 &b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
inline Vec3 &operator-=(Vec3 &a, const Vec3 &b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
inline Vec3 &operator*=(Vec3 &a, double s) { a.x *= s; a.y *= s; a.z *= s; return a; }
inline Vec3 &operator/=(Vec3 &a, double s) { a.x /= s; a.y /= s; a.z /= s; return a; }
// OpenFOAM's operator& (inner product): x*x + y*y + z*z, left to right
inline double dot(const Vec3 &a, const Vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// OpenFOAM's operator^ (cross product)
inline Vec3 cross(const Vec3 &a, const Vec3 &b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double mag(const Vec3 &a) { return std::sqrt(dot(a, a)); }
// fused dot product used by the tet walk; the CUDA kernel evaluates the very
// same fma chain so that tet/cell decisions are bit-identical.
inline double dotf(const Vec3 &a, const Vec3 &b) {
    return std::fma(a.x, b.x, std::fma(a.y, b.y, a.z * b.z));
}

struct Tensor {  // row-major 3x3
    double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    static Tensor identity() { Tensor t; t.m[0] = t.m[4] = t.m[8] = 1; return t; }
    Tensor T() const {
        Tensor r;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[3 * i + j] = m[3 * j + i];
        return r;
    }
};
// transform(T, v) = T & v
inline Vec3 transform(const Tensor &t, const Vec3 &v) {
    return {t.m[0] * v.x + t.m[1] * v.y + t.m[2] * v.z,
            t.m[3] * v.x + t.m[4] * v.y + t.m[5] * v.z,
            t.m[6] * v.x + t.m[7] * v.y + t.m[8] * v.z};
}
// OpenFOAM rotationTensor(n1, n2) (transform.H), SURVEY Appendix B.4
Tensor rotationTensor(const Vec3 &n1, const Vec3 &n2);

struct SymmTensor {  // xx xy xz yy yz zz
    double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
};
// transform(T, symmTensor) = T & S & T^T
SymmTensor transform(const Tensor &t, const SymmTensor &s);

struct Patch {
    int start = 0, size = 0;
    int geom = PDFB200_PATCH_GENERIC;
    int handler = PDFB200_BC_SLIP;
    int reflecting = 1;
};

// richTetrahedron (richTetrahedron/richTetrahedronI.H:107-134) plus the
// bookkeeping of tetFacePointCellDecomposition (…Decomposition.C:31-67) and
// the adjacency of SURVEY Appendix D.
struct Tet {
    Vec3 a, b, c, d;          // face centre, two face points, cell centre
    Vec3 Sa, Sb, Sc, Sd;      // face area vectors (outward)
    double vol = 0;           // volume (tetrahedron::mag())
    Vec3 gradN[4];            // shape-function gradients, S_i/(-3 mag)
    int face = -1;            // tetFace_
    int ptFirst = -1, ptSecond = -1;  // tetPoints_ (local face-point indices)
    int ptB = -1, ptC = -1;   // global point labels of b and c
    int cell = -1;
    int nbr[4] = {-1, -1, -1, -1};  // tet across the face opposite a,b,c,d (-1: boundary)
    // literal restatement of richTetrahedron::inside (richTetrahedron.C:40-91)
    bool inside(const Vec3 &pt) const;
};

struct Mesh {
    // raw polyMesh
    std::vector<Vec3> points;
    std::vector<int> faceStart, facePoints, owner, neighbour;
    std::vector<Patch> patches;
    int nCells = 0, nFaces = 0, nInternalFaces = 0;
    int geometricD[3] = {1, 1, 1}, solutionD[3] = {1, 1, 1};
    int nBoundaryFaces() const { return nFaces - nInternalFaces; }
    int nGeometricD() const { return (geometricD[0] > 0) + (geometricD[1] > 0) + (geometricD[2] > 0); }

    // derived geometry (primitiveMesh::makeFaceCentresAndAreas / makeCellCentresAndVols)
    std::vector<Vec3> faceCentres, faceAreas, cellCentres;
    std::vector<double> magSf, cellVolumes;
    std::vector<int> cellFaceStart, cellFaces;     // primitiveMesh::calcCells order
    std::vector<int> facePatch;                    // boundary face -> patch
    std::vector<int> pointCellStart, pointCells;   // point -> cells (ascending)
    std::vector<double> pointWeights;              // volPointInterpolation weights
    std::vector<double> linWeights;                // surfaceInterpolation::weights (internal faces)
    std::vector<Vec3> courantCoeffs;               // mcParticleCloud.C:526-535,649 (zero on empty patches)
    std::vector<double> hNum;                      // mcParticleCloud.C:1937-1971
    std::vector<Vec3> bbMin, bbMax;                // cell bounding boxes

    // tetFacePointCellDecomposition
    std::vector<Tet> tets;
    std::vector<int> cellTetStart;                 // [nCells+1]
    std::vector<int> faceTetStartOwn, faceTetStartNei;  // first tet of face f in its owner / neighbour cell
    int maxTetsPerCell = 0;

    // axisymmetry (mcParticleCloud.C:661-717)
    bool axisym = false;
    Vec3 axis, centreNormal, wedgePatchNormal;
    double openingAngle = 0;
    std::vector<double> volumeOrArea;              // cell volume, or wedge face area

    void build(const pdfb200_mesh &m);
    bool isInternalFace(int f) const { return f < nInternalFaces; }
    int faceSize(int f) const { return faceStart[f + 1] - faceStart[f]; }
    int facePoint(int f, int i) const { return facePoints[faceStart[f] + i]; }
    // tetFacePointCellDecomposition::find (…Decomposition.C:125-151), reference semantics
    int findReference(const Vec3 &pt, int cellHint) const;
    // locate rule shared with the device: tet of `cell` maximising min_i phi_i
    int locate(const Vec3 &pt, int cell) const;
    // polyMesh::pointInCell (face-plane test)
    bool pointInCell(const Vec3 &pt, int cell) const;
    void constrainDirection(const int dirs[3], Vec3 &v) const {
        for (int i = 0; i < 3; ++i) if (dirs[i] < 0) v[i] = 0;
    }

private:
    void makeFaceGeometry();
    void makeCells();
    void makeCellGeometry();
    void makeTets();
    void makeInterpolationWeights();
    void makeCourantCoeffs();
    void makeHNum();
    void detectAxisymmetry();
};

// ---------------------------------------------------------------------------
// Random numbers
// ---------------------------------------------------------------------------

// counter-based Philox4x32-10 (Salmon et al. 2011); identical on the device
void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void normalPairF32(uint32_t a, uint32_t b, float &n0, float &n1);

enum RngPurpose : uint32_t {
    RNG_STEP_NORMALS = 1,   // entity = particle id; sub 0 -> 4 normals, sub 1 -> 2 normals
    RNG_INLET_STEPFRAC = 2, // entity = particle id
    RNG_NC_PARTICLE = 3,    // entity = particle id (eliminate: selection draws)
    RNG_NC_CELL = 4,        // entity = cell (eliminate: which population dies)
    RNG_CLONE_POS = 5,      // entity = particle id; sub = 2*attempt(+1)
    RNG_SEED = 6,           // entity = initial particle index
    RNG_INJECT = 7,         // entity = boundary face; sub = 4*k + {0..3}
    RNG_MIX_HASH = 8,       // entity = particle id; sub = round (modifiedCurl: matching order)
    RNG_MIX_PAIR = 9,       // entity = id of the pair's first particle; sub = round (mix?, a)
    RNG_INJECT_SAMPLE = 10  // entity = boundary face | draw << 32; sub = attempt ("sampling" inflow generator)
};

struct KeyedRng {
    uint32_t key[2] = {0, 0};
    void setSeed(uint64_t seed) {
        key[0] = (uint32_t)(seed & 0xffffffffu);
        key[1] = (uint32_t)(seed >> 32);
    }
    void words(uint64_t entity, uint32_t step, uint32_t purpose, uint32_t sub, uint32_t o[4]) const;
    // two uniforms in [0,1) from one Philox call
    void uniform2(uint64_t entity, uint32_t step, uint32_t purpose, uint32_t sub,
                  double &u0, double &u1) const;
    // two / four N(0,1) deviates from one Philox call (single-precision Box-Muller, normalPairF32)
    void normal2(uint64_t entity, uint32_t step, uint32_t purpose, uint32_t sub,
                 double &n0, double &n1) const;
    void normal4(uint64_t entity, uint32_t step, uint32_t purpose, uint32_t sub, double n[4]) const;
};

// Sequential generator with OpenFOAM Random's interface (B.3): scalar01 is a
// 48-bit LCG (drand48 recurrence), GaussNormal is Marsaglia polar with a
// cached second deviate.
struct SeqRandom {
    uint64_t state = 0;
    bool haveCached = false;
    double cached = 0;
    void seed(uint64_t s) { state = ((s & 0xffffffffu) << 16) | 0x330Eu; haveCached = false; }
    double scalar01();
    double GaussNormal();
};

// ---------------------------------------------------------------------------
// Fields
// ---------------------------------------------------------------------------

// volScalarField-like: cell values + boundary-face values
struct ScalarField {
    std::vector<double> c, b;
    std::vector<uint8_t> fixed;   // per boundary face: fixedValue (1) or zeroGradient-like (0)
    void resize(int nc, int nb) { c.assign(nc, 0.0); b.assign(nb, 0.0); if (fixed.empty()) fixed.assign(nb, 0); }
};
struct VectorField {
    std::vector<Vec3> c, b;
    std::vector<uint8_t> fixed;
    void resize(int nc, int nb) { c.assign(nc, Vec3()); b.assign(nb, Vec3()); if (fixed.empty()) fixed.assign(nb, 0); }
};
struct SymmTensorField {
    std::vector<SymmTensor> c, b;
    std::vector<uint8_t> fixed;
};

// node values of one interpolated field (cellPointFace): cell / point / face
template <class T> struct NodeField {
    std::vector<T> cell, point, face;
    int scheme = PDFB200_INTERP_CELL_POINT_FACE;
};

// ---------------------------------------------------------------------------
// Particle (mcParticle/mcParticle/mcParticle.H:88-156 + Particle<T> base)
// ---------------------------------------------------------------------------
struct Particle {
    Vec3 position;
    int cell = -1;
    int tet = -1;             // carried by the tet walk (new design, Appendix D)
    double stepFraction = 0;
    double m = 0;
    Vec3 UParticle, UParticleOld, Ucorrection, Utracking;
    double Omega = 0, rho = 0, eta = 1;
    Vec3 positionOld;
    int cellOld = -1, tetOld = -1;
    double Co = 0;
    Vec3 reflectionBoundaryVelocity;
    int nSteps = 0;
    bool reflected = false, isOnInletBoundary = false, reflectedAtOpenBoundary = false;
    bool keep = true;
    int injPatch = -1;        // patch that injected it this step (cull ordering, see cloud.cpp)
    uint64_t id = 0;
    std::array<double, PDFB200_MAX_PHI> Phi{};
};

struct InletRandom;  // mcInletRandom + inversion

// ---------------------------------------------------------------------------
// The cloud
// ---------------------------------------------------------------------------
struct Cloud {
    Mesh mesh;
    pdfb200_params prm{};
    int nCov() const { return prm.nPhi * (prm.nPhi + 1) / 2; }
    int nBnd() const { return mesh.nBoundaryFaces(); }

    // FV-owned
    VectorField Ufv; ScalarField pfv, kfv, epsfv; SymmTensorField Rfv;
    bool haveFv = false, haveCloudFields = false, haveMoments = false;
    // cloud fields
    ScalarField rhocPdf, rhocPdfInst, pndcPdf, pndcPdfInst, kcPdf;
    VectorField UcPdf; SymmTensorField TaucPdf;
    std::vector<ScalarField> PhicPdf, PhiPhicPdf;
    std::vector<double> PaNIC, lostMass;
    // moments
    std::vector<double> mMom, VMom; std::vector<Vec3> UMom; std::vector<SymmTensor> UUMom;
    std::vector<std::vector<Vec3>> UPhiMom, UPhiFlux;   // [nPhi][nCells]: time-averaged sum of eta m U phi, and <u'' phi''>
    std::vector<std::vector<double>> PhiMom, PhiPhiMom;

    // flamelet tables
    std::vector<double> flZ, flChi; std::vector<std::vector<double>> flFields; int nChi = 0, nZ = 0;
    // KulkarniAutoIgnition library: z, pv, cEq[z], tables [nZ][nPv] (0 cdot, 1 rho, 2.. added)
    std::vector<double> aiZ, aiPv, aiCEq; std::vector<std::vector<double>> aiFields; double aiCEqMax = 0;

    // model internals (updateInternals products)
    NodeField<double> OmegaN, kN, pStarN, etaN, zVarN;
    NodeField<Vec3> UN, diffUN, UPosCorrN;
    // externally solved position correction (PDFB200_POSCORR_EXTERNAL): uploaded fields and their node values
    VectorField pcG0, pcG1, pcG2; ScalarField pcL, pcZeta; double pcScale = 0; bool havePcExt = false;
    NodeField<Vec3> pcG1N, pcG2N; NodeField<double> pcLN, pcZetaN;
    std::vector<double> diffk;

    // inlet caches (mcInletOutletBoundary::getU/getR cache on first use)
    std::vector<Vec3> inletU; std::vector<SymmTensor> inletR; bool inletCached = false;
    std::vector<Tensor> fwdTrans, revTrans;   // per boundary face
    std::vector<std::vector<double>> probVec; // per boundary face
    std::vector<double> Un;                   // per boundary face, open patches

    // open-boundary lookup: cell -> (boundary faces on open patches)
    std::vector<int> openCellStart, openCellFaces;

    std::vector<Particle> particles;
    double deltaT = 100 * SMALL;
    uint32_t step = 0;
    uint64_t nextId = 0;
    int rank = 0, nRanks = 1;
    KeyedRng krng; SeqRandom srng;
    int64_t nLostTotal = 0;

    // per-patch flux bookkeeping of the current step
    std::vector<std::array<double, 1 + PDFB200_MAX_PHI>> patchMassIn, patchMassOut;

    // last-step counters
    int64_t nInjected = 0, nDeleted = 0, nCloned = 0, nEliminated = 0, nLost = 0;

    std::string init(const pdfb200_mesh &m, const pdfb200_params &p);
    std::string setParams(const pdfb200_params &p);

    // ---- fields ----
    void applyBC(ScalarField &f) const;   // correctBoundaryConditions()
    void applyBC(VectorField &f) const;
    void applyBC(SymmTensorField &f) const;
    void initMoments();                    // mcParticleCloud.C:1633-1701
    void makeNodes(const ScalarField &f, NodeField<double> &n, int scheme) const;
    void makeNodes(const VectorField &f, NodeField<Vec3> &n, int scheme) const;
    void gaussGrad(const ScalarField &f, std::vector<Vec3> &g) const;  // fvc::grad, Gauss linear
    void updateModelInternals();           // all updateInternals() of one step

    // ---- interpolation at a particle ----
    void baryWeights(const Particle &p, double w[4]) const;
    double interp(const NodeField<double> &n, const Particle &p) const;
    Vec3 interp(const NodeField<Vec3> &n, const Particle &p) const;
    Vec3 gradInterp(const NodeField<double> &n, const Particle &p) const;

    // ---- the step ----
    double massPerDepth(const Particle &p) const;       // mcParticleCloudI.H:196-208
    void constrainParticle(double dt, Particle &p) const; // mcParticleCloud.C:76-101
    void computeCourantNo(Particle &p) const;           // mcParticleCloud.C:2120-2148
    bool move(Particle &p, double trackTimeGlobal);     // mcParticle.C:100-179 on the tet walk
    bool hitPatch(Particle &p, int face);               // mcParticle.C:211-235 -> handlers; true: snapped + relocated (wedge)
    void boundaryCorrectBefore();                       // mcOpenBoundary/mcInletOutletBoundary::correct(false)
    void injectPatch(int patchI);
    void cullPatch(int patchI);
    void boundaryCorrectAfter();                        // correct(true)
    void updateCloudPDF(double existWt);                // mcParticleCloud.C:824-1009
    void accumulateInstantMoments(std::vector<double> &block) const;
    void finaliseMoments(const std::vector<double> &block, double existWt);
    void particleNumberControl();                       // mcParticleCloud.C:1014-1221
    void evolve(pdfb200_step_report *rep);              // mcParticleCloud.C:1226-1530
    // split step for particle-sharded multi-rank runs
    void evolveBegin(pdfb200_step_report *rep, std::vector<double> &momentBlock);
    void evolveEnd(pdfb200_step_report *rep, const std::vector<double> &momentBlock);
    int nMoments() const { return 12 + prm.nPhi + nCov() + 3 * prm.nPhi; }   // N, M, V, MU, MPhi, MPhiPhi, MUU, MUPhi
    int momentBlockSize() const { return mesh.nCells * nMoments() + 32; }

    // models on one particle (mcModel::correct(mcParticle&))
    void omegaCorrect(Particle &p) const;               // mcRASOmegaModel.C:125-147
    void mixingCorrect(Particle &p) const;              // mcIEMMixingModel.C:68-83
    void curlMixing();                                  // "modifiedCurl" (new; see include/pdfb200.h)
    void reactionCorrect(Particle &p) const;            // cold / BurkeSchumann / SteadyFlamelet
    void velocityCorrect(Particle &p, const double xi[3]) const;  // mcSLMFullVelocityModel.C:146-185
    void ltsCorrect(Particle &p) const;                 // mcCellLocalTimeStepping.C:164-174 etc.
    void positionCorrectionCorrect(Particle &p) const;  // mcLimitedSimplePositionCorrection.C:187-192

    void seedParticles();                               // mcParticleCloud.C:1534-1629
    std::vector<Vec3> randomPointsInCell(int n, int cell, uint64_t entityBase, uint32_t purpose, uint32_t stepCtr);
    void adjustAxiSymmetricMass(std::vector<Particle *> &ps) const;  // mcParticleCloud.C:2095-2117

    // step-local diagnostics
    std::array<double, 1 + PDFB200_MAX_PHI> deltaMassInst{};
    void stepNormals(const Particle &p, double out[6], int phase);
};

// flamelet helpers (mcSteadyFlamelet.C:44-93)
void findCoeffs(const std::vector<double> &lst, double v, int &i, double &w);
double binterp(const std::vector<double> &v, int nCols, int ix, double wx, int iy, double wy);

// inflow velocity PDF (mcInletRandom.C:92-108, mcInletRandomI.H:28-39,
// mcInversionInletRandom.C:63-119)
struct InletRandom {
    double UMean = 0, uRms = 0, b = 0, b2 = 0, expmU2b2 = 0, UbSqrtPi = 0, erfUb = 0, denom = 0,
           b22denom = 0, Q = 0, m = 0;
    void updateCoeffs(double UMean, double uRms);
    double PDF(double x) const;
    double CDF(double x) const;
    double newton(double x0, double y, int &nIter) const;
    double value(double U01) const;
    // "sampling" (mcSamplingInletRandom.C:37-102) on the keyed generator; see InletRnd::sample in pdffoam_b200/csrc/particles.cu
    double sample(const KeyedRng &k, uint32_t bf, uint32_t step, uint32_t kDraw) const;
    // the same on a sequential stream (rngMode 1)
    double sample(SeqRandom &s) const;
};

}  // namespace pdforacle
