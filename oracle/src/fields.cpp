// fields.cpp — CPU ORACLE (test infrastructure): random numbers, boundary
// conditions, cell->node interpolation tables and the updateInternals() part
// of every sub-model.
#include <algorithm>
#include <cstring>
#include <stdexcept>

#include "pdforacle.hpp"

namespace pdforacle {

// ---------------------------------------------------------------------------
// Philox4x32-10
// ---------------------------------------------------------------------------
void philox4x32_10(const uint32_t ctrIn[4], const uint32_t keyIn[2], uint32_t out[4]) {
    uint32_t c[4] = {ctrIn[0], ctrIn[1], ctrIn[2], ctrIn[3]};
    uint32_t k[2] = {keyIn[0], keyIn[1]};
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
        const uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
    for (int i = 0; i < 4; ++i) out[i] = c[i];
}

static inline uint64_t pack64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// counter = (entity low word, entity high word, step, purpose << 24 | sub): 64-bit entities, sub < 2^24
void KeyedRng::words(uint64_t entity, uint32_t step, uint32_t purpose, uint32_t sub, uint32_t o[4]) const {
    const uint32_t ctr[4] = {(uint32_t)entity, (uint32_t)(entity >> 32), step, (purpose << 24) | sub};
    philox4x32_10(ctr, key, o);
}

void KeyedRng::uniform2(uint64_t entity, uint32_t step, uint32_t purpose, uint32_t sub,
                        double &u0, double &u1) const {
    uint32_t o[4];
    words(entity, step, purpose, sub, o);
    u0 = (double)(pack64(o[0], o[1]) >> 11) * 0x1.0p-53;
    u1 = (double)(pack64(o[2], o[3]) >> 11) * 0x1.0p-53;
}

// Two N(0,1) deviates from two 32-bit words: single-precision Box-Muller. The device evaluates the
// same sequence of single-precision operations (explicit fused multiply-adds only; both sides are
// compiled without contraction), so the deviates agree bit for bit. Radius from u = (a+1) 2^-32,
// angle from b: two quadrant bits and a 30-bit position inside the quadrant; log, sin and cos by the
// Cephes single-precision polynomials.
static inline float bitsToFloat(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t floatToBits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
void normalPairF32(uint32_t a, uint32_t b, float &n0, float &n1) {
    const float x = (float)a + 1.0f;
    const int32_t xi = (int32_t)floatToBits(x);
    int e = (xi >> 23) - 127;
    float m = bitsToFloat(((uint32_t)xi & 0x007fffffu) | 0x3f800000u);
    if (m > 1.41421354f) { m *= 0.5f; e += 1; }
    const float f = m - 1.0f, z = f * f;
    float p = 7.0376836292E-2f;
    p = std::fmaf(p, f, -1.1514610310E-1f);
    p = std::fmaf(p, f, 1.1676998740E-1f);
    p = std::fmaf(p, f, -1.2420140846E-1f);
    p = std::fmaf(p, f, 1.4249322787E-1f);
    p = std::fmaf(p, f, -1.6668057665E-1f);
    p = std::fmaf(p, f, 2.0000714765E-1f);
    p = std::fmaf(p, f, -2.4999993993E-1f);
    p = std::fmaf(p, f, 3.3333331174E-1f);
    const float lnm = std::fmaf(f * z, p, std::fmaf(-0.5f, z, f));
    const float L = std::fmaf((float)(32 - e), 0.693147182f, -lnm);
    const float r = std::sqrt(2.0f * L);
    const uint32_t q = b >> 30;
    const float t = (float)(b & 0x3fffffffu) * 0x1p-30f;
    const bool sw = t > 0.5f;
    const float th = (sw ? 1.0f - t : t) * 1.57079637f;
    const float z2 = th * th;
    float sp = -1.9515295891E-4f;
    sp = std::fmaf(sp, z2, 8.3321608736E-3f);
    sp = std::fmaf(sp, z2, -1.6666654611E-1f);
    const float s0 = std::fmaf(sp * z2, th, th);
    float cp = 2.443315711809948E-5f;
    cp = std::fmaf(cp, z2, -1.388731625493765E-3f);
    cp = std::fmaf(cp, z2, 4.166664568298827E-2f);
    const float c0 = std::fmaf(cp * z2, z2, std::fmaf(-0.5f, z2, 1.0f));
    const float s1 = sw ? c0 : s0, c1 = sw ? s0 : c0;
    const bool odd = (q & 1u) != 0;
    float sn = odd ? c1 : s1, cs = odd ? s1 : c1;
    if (q >= 2u) sn = -sn;
    if (q == 1u || q == 2u) cs = -cs;
    n0 = r * cs; n1 = r * sn;
}

void KeyedRng::normal2(uint64_t entity, uint32_t step, uint32_t purpose, uint32_t sub,
                       double &n0, double &n1) const {
    uint32_t o[4];
    words(entity, step, purpose, sub, o);
    float a, b;
    normalPairF32(o[0], o[1], a, b);
    n0 = (double)a; n1 = (double)b;
}

void KeyedRng::normal4(uint64_t entity, uint32_t step, uint32_t purpose, uint32_t sub, double n[4]) const {
    uint32_t o[4];
    words(entity, step, purpose, sub, o);
    float a, b, c, d;
    normalPairF32(o[0], o[1], a, b);
    normalPairF32(o[2], o[3], c, d);
    n[0] = (double)a; n[1] = (double)b; n[2] = (double)c; n[3] = (double)d;
}

double SeqRandom::scalar01() {
    state = (state * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (double)state * 0x1.0p-48;
}

// OpenFOAM Random::GaussNormal: Marsaglia polar with a cached second deviate
double SeqRandom::GaussNormal() {
    if (haveCached) { haveCached = false; return cached; }
    double v1, v2, rsq;
    do {
        v1 = 2.0 * scalar01() - 1.0;
        v2 = 2.0 * scalar01() - 1.0;
        rsq = v1 * v1 + v2 * v2;
    } while (rsq >= 1.0 || rsq == 0.0);
    const double fac = std::sqrt(-2.0 * std::log(rsq) / rsq);
    cached = v1 * fac; haveCached = true;
    return v2 * fac;
}

// ---------------------------------------------------------------------------
// correctBoundaryConditions(): fixedValue faces keep their value, every other
// face takes the cell value (zeroGradient); wedge faces rotate vectors/tensors
// onto the patch (wedgeFvPatchField::evaluate), empty faces carry the cell
// value so that "cell value on an empty patch"
// (gradInterpolationConstantTet.C:94-97) needs no special case downstream.
// ---------------------------------------------------------------------------
static Tensor wedgeFaceT(const Mesh &mesh, const Patch &pp) {
    Vec3 n = mesh.faceAreas[pp.start];
    n /= mag(n);
    Vec3 cn;
    for (int k = 0; k < 3; ++k) {
        const double s = n[k] > 0 ? 1.0 : (n[k] < 0 ? -1.0 : 0.0);
        cn[k] = s * (std::max(std::fabs(n[k]), 0.5) - 0.5);
    }
    cn /= mag(cn);
    return rotationTensor(cn, n);
}

void Cloud::applyBC(ScalarField &f) const {
    const int nI = mesh.nInternalFaces;
    for (int bf = 0; bf < nBnd(); ++bf) {
        const Patch &pp = mesh.patches[mesh.facePatch[bf]];
        if (pp.geom == PDFB200_PATCH_GENERIC && f.fixed[bf]) continue;
        f.b[bf] = f.c[mesh.owner[nI + bf]];
    }
}

void Cloud::applyBC(VectorField &f) const {
    const int nI = mesh.nInternalFaces;
    for (const Patch &pp : mesh.patches) {
        Tensor T;
        if (pp.geom == PDFB200_PATCH_WEDGE && pp.size > 0) T = wedgeFaceT(mesh, pp);
        for (int j = 0; j < pp.size; ++j) {
            const int bf = pp.start + j - nI;
            if (pp.geom == PDFB200_PATCH_GENERIC && f.fixed[bf]) continue;
            const Vec3 &cv = f.c[mesh.owner[nI + bf]];
            f.b[bf] = (pp.geom == PDFB200_PATCH_WEDGE) ? transform(T, cv) : cv;
        }
    }
}

void Cloud::applyBC(SymmTensorField &f) const {
    const int nI = mesh.nInternalFaces;
    for (const Patch &pp : mesh.patches) {
        Tensor T;
        if (pp.geom == PDFB200_PATCH_WEDGE && pp.size > 0) T = wedgeFaceT(mesh, pp);
        for (int j = 0; j < pp.size; ++j) {
            const int bf = pp.start + j - nI;
            if (pp.geom == PDFB200_PATCH_GENERIC && f.fixed[bf]) continue;
            const SymmTensor &cv = f.c[mesh.owner[nI + bf]];
            f.b[bf] = (pp.geom == PDFB200_PATCH_WEDGE) ? transform(T, cv) : cv;
        }
    }
}

// ---------------------------------------------------------------------------
// interpolationCellPointFace node values:
// volPointInterpolation for points, linearInterpolate for faces (patch value on
// boundary faces, cell value on empty patches).
// ---------------------------------------------------------------------------
void Cloud::makeNodes(const ScalarField &f, NodeField<double> &n, int scheme) const {
    n.scheme = scheme;
    n.cell = f.c;
    const int nP = (int)mesh.points.size();
    n.point.assign(nP, 0.0); n.face.assign(mesh.nFaces, 0.0);
    for (int p = 0; p < nP; ++p) {
        double s = 0;
        for (int k = mesh.pointCellStart[p]; k < mesh.pointCellStart[p + 1]; ++k)
            s += mesh.pointWeights[k] * f.c[mesh.pointCells[k]];
        n.point[p] = s;
    }
    for (int fc = 0; fc < mesh.nFaces; ++fc) {
        if (fc < mesh.nInternalFaces) {
            const double w = mesh.linWeights[fc];
            n.face[fc] = w * f.c[mesh.owner[fc]] + (1 - w) * f.c[mesh.neighbour[fc]];
        } else {
            const int bf = fc - mesh.nInternalFaces;
            const bool empty = mesh.patches[mesh.facePatch[bf]].geom == PDFB200_PATCH_EMPTY;
            n.face[fc] = empty ? f.c[mesh.owner[fc]] : f.b[bf];
        }
    }
}

void Cloud::makeNodes(const VectorField &f, NodeField<Vec3> &n, int scheme) const {
    n.scheme = scheme;
    n.cell = f.c;
    const int nP = (int)mesh.points.size();
    n.point.assign(nP, Vec3()); n.face.assign(mesh.nFaces, Vec3());
    for (int p = 0; p < nP; ++p) {
        Vec3 s;
        for (int k = mesh.pointCellStart[p]; k < mesh.pointCellStart[p + 1]; ++k)
            s += mesh.pointWeights[k] * f.c[mesh.pointCells[k]];
        n.point[p] = s;
    }
    for (int fc = 0; fc < mesh.nFaces; ++fc) {
        if (fc < mesh.nInternalFaces) {
            const double w = mesh.linWeights[fc];
            n.face[fc] = w * f.c[mesh.owner[fc]] + (1 - w) * f.c[mesh.neighbour[fc]];
        } else {
            const int bf = fc - mesh.nInternalFaces;
            const bool empty = mesh.patches[mesh.facePatch[bf]].geom == PDFB200_PATCH_EMPTY;
            n.face[fc] = empty ? f.c[mesh.owner[fc]] : f.b[bf];
        }
    }
}

// fvc::grad with "Gauss linear" (SURVEY B.5): (1/V) sum_f Sf psi_f; empty
// patches have no faces in the fvMesh and contribute nothing.
void Cloud::gaussGrad(const ScalarField &f, std::vector<Vec3> &g) const {
    g.assign(mesh.nCells, Vec3());
    for (int fc = 0; fc < mesh.nFaces; ++fc) {
        if (fc < mesh.nInternalFaces) {
            const double w = mesh.linWeights[fc];
            const double pf = w * f.c[mesh.owner[fc]] + (1 - w) * f.c[mesh.neighbour[fc]];
            const Vec3 Sfpf = mesh.faceAreas[fc] * pf;
            g[mesh.owner[fc]] += Sfpf;
            g[mesh.neighbour[fc]] -= Sfpf;
        } else {
            const int bf = fc - mesh.nInternalFaces;
            if (mesh.patches[mesh.facePatch[bf]].geom == PDFB200_PATCH_EMPTY) continue;
            g[mesh.owner[fc]] += mesh.faceAreas[fc] * f.b[bf];
        }
    }
    for (int c = 0; c < mesh.nCells; ++c) g[c] /= mesh.cellVolumes[c];
}

// mcParticleCloud::initMoments (mcParticleCloud.C:1633-1701)
void Cloud::initMoments() {
    const int nc = mesh.nCells, nb = nBnd();
    mMom.resize(nc); VMom.resize(nc); UMom.resize(nc); UUMom.resize(nc);
    PhiMom.assign(prm.nPhi, std::vector<double>(nc)); PhiPhiMom.assign(nCov(), std::vector<double>(nc));
    UPhiMom.assign(prm.nPhi, std::vector<Vec3>(nc)); UPhiFlux.assign(prm.nPhi, std::vector<Vec3>(nc));
    rhocPdfInst = rhocPdf; pndcPdf = rhocPdf; pndcPdfInst = rhocPdf;
    UcPdf.c = Ufv.c; UcPdf.b = Ufv.b; UcPdf.fixed = Ufv.fixed;
    applyBC(UcPdf);
    kcPdf.c = kfv.c; kcPdf.b = kfv.b; kcPdf.fixed = kfv.fixed;
    applyBC(kcPdf);
    TaucPdf.c = Rfv.c; TaucPdf.b = Rfv.b; TaucPdf.fixed = kfv.fixed;   // turbModel_.R(): patch types of k
    for (int c = 0; c < nc; ++c) {
        mMom[c] = rhocPdf.c[c] * mesh.volumeOrArea[c];
        VMom[c] = mMom[c] / rhocPdf.c[c];
        UMom[c] = mMom[c] * Ufv.c[c];
        const Vec3 &u = UcPdf.c[c];
        const SymmTensor &R = Rfv.c[c];
        SymmTensor uu;
        uu.xx = mMom[c] * (R.xx + u.x * u.x); uu.xy = mMom[c] * (R.xy + u.x * u.y);
        uu.xz = mMom[c] * (R.xz + u.x * u.z); uu.yy = mMom[c] * (R.yy + u.y * u.y);
        uu.yz = mMom[c] * (R.yz + u.y * u.z); uu.zz = mMom[c] * (R.zz + u.z * u.z);
        UUMom[c] = uu;
        int pp = 0;
        for (int i = 0; i < prm.nPhi; ++i) {
            PhiMom[i][c] = mMom[c] * PhicPdf[i].c[c];
            UPhiMom[i][c] = (mMom[c] * PhicPdf[i].c[c]) * u;   // zero flux to start with
            UPhiFlux[i][c] = Vec3();
            for (int j = i; j < prm.nPhi; ++j, ++pp)
                PhiPhiMom[pp][c] = mMom[c] * (PhiPhicPdf[pp].c[c] + PhicPdf[i].c[c] * PhicPdf[j].c[c]);
        }
    }
    PaNIC.assign(nc, 0.0); lostMass.assign(nc, 0.0);
    (void)nb;
    haveMoments = true;
}

// ---------------------------------------------------------------------------
// All updateInternals() of one step. They only read FV fields, the cloud mean
// fields of the previous updateCloudPDF() and deltaT, none of which change
// before the particle passes of this step, so evaluating them together at the
// start of the step is equivalent to the reference's interleaved calls
// (mcModel::correct, mcModel.C:74-81).
// ---------------------------------------------------------------------------
void Cloud::updateModelInternals() {
    const int nc = mesh.nCells, nb = nBnd(), nI = mesh.nInternalFaces;

    // --- mcLimitedSimplePositionCorrection::updateInternals (.C:88-184) ---
    {
        VectorField UPC; UPC.resize(nc, nb);
        if (prm.positionCorrection == PDFB200_POSCORR_LIMITED_SIMPLE) {
            ScalarField phi; phi.resize(nc, nb);
            for (int c = 0; c < nc; ++c) phi.c[c] = (pndcPdf.c[c] - rhocPdf.c[c]) / rhocPdf.c[c];
            applyBC(phi);                                 // zeroGradient
            double phiLim = 0.;
            for (int c = 0; c < nc; ++c) phiLim = std::max(phiLim, std::fabs(phi.c[c]));
            phiLim = std::tanh(10. * phiLim);
            std::vector<Vec3> g;
            gaussGrad(phi, g);
            for (int c = 0; c < nc; ++c) UPC.c[c] = -g[c];
            double Co = 0;
            for (int c = 0; c < nc; ++c)
                for (int j = mesh.cellFaceStart[c]; j < mesh.cellFaceStart[c + 1]; ++j) {
                    const Vec3 coeff = deltaT * mesh.courantCoeffs[mesh.cellFaces[j]];
                    Co = std::max(Co, std::fabs(dot(coeff, UPC.c[c])));
                }
            double corr = 0.;
            if (Co > SMALL) corr = prm.pcCFL / Co * prm.pcC * phiLim;
            for (int c = 0; c < nc; ++c) UPC.c[c] *= corr;
            // boundary: patch types of rho, fixedValue faces forced to zero (.C:76-84)
            UPC.fixed = rhocPdf.fixed;
            applyBC(UPC);
        }
        if (prm.positionCorrection == PDFB200_POSCORR_SIMPLE) {
            // mcSimplePositionCorrection::updateInternals (.C:135-163): phi = pnd/rho - 1 (zeroGradient),
            // gradPhi = C |Ufv| grad(phi). Boundary values of gradPhi: zeroGradient (its declared patch type, .C:92-109;
            // OpenFOAM's assignment copies fvc::grad's patch values instead, which differ by the normal component -
            // an OpenFOAM primitive that is defined here).
            ScalarField phi; phi.resize(nc, nb);
            for (int c = 0; c < nc; ++c) phi.c[c] = pndcPdf.c[c] / rhocPdf.c[c] - 1.;
            applyBC(phi);
            std::vector<Vec3> g;
            gaussGrad(phi, g);
            for (int c = 0; c < nc; ++c) UPC.c[c] = (prm.pcC * mag(Ufv.c[c])) * g[c];
            UPC.fixed.assign(nb, 0);
            applyBC(UPC);
        }
        if (prm.positionCorrection == PDFB200_POSCORR_EXTERNAL) {
            if (!havePcExt) throw std::logic_error("evolve: positionCorrection external selected but no fields uploaded (upload_poscorr_fields)");
            makeNodes(pcG0, UPosCorrN, prm.interpUPosCorr);
            makeNodes(pcG1, pcG1N, prm.interpUPosCorr); makeNodes(pcG2, pcG2N, prm.interpUPosCorr);
            makeNodes(pcL, pcLN, prm.interpUPosCorr); makeNodes(pcZeta, pcZetaN, prm.interpUPosCorr);
        } else {
            makeNodes(UPC, UPosCorrN, prm.interpUPosCorr);
        }
    }

    // --- mcRASOmegaModel::updateInternals (.C:90-122) ---
    {
        ScalarField Om; Om.resize(nc, nb);
        for (int c = 0; c < nc; ++c) Om.c[c] = epsfv.c[c] / kfv.c[c];
        for (int bf = 0; bf < nb; ++bf) Om.b[bf] = epsfv.b[bf] / kfv.b[bf];
        double maxOm = -VGREAT;
        for (int c = 0; c < nc; ++c) maxOm = std::max(maxOm, Om.c[c]);
        for (int bf = 0; bf < nb; ++bf)
            if (mesh.patches[mesh.facePatch[bf]].geom != PDFB200_PATCH_EMPTY) maxOm = std::max(maxOm, Om.b[bf]);
        if (maxOm > prm.Omega0) {
            for (auto &v : Om.c) v = std::min(v, prm.Omega0);
            for (auto &v : Om.b) v = std::min(v, prm.Omega0);
        }
        makeNodes(Om, OmegaN, prm.interpOmega);
    }

    // --- mcSteadyFlamelet::updateInternals (.C:291-301) ---
    if (prm.reactionModel == PDFB200_REACTION_STEADY_FLAMELET) {
        int pp = 0, zz = -1;
        for (int i = 0; i < prm.nPhi; ++i)
            for (int j = i; j < prm.nPhi; ++j, ++pp)
                if (i == prm.zIdx && j == prm.zIdx) zz = pp;
        makeNodes(PhiPhicPdf[zz], zVarN, prm.interpZVar);
    }

    // --- mcSLMFullVelocityModel::updateInternals (.C:104-143) ---
    {
        ScalarField ps; ps.resize(nc, nb);
        for (int c = 0; c < nc; ++c) ps.c[c] = pfv.c[c] - 2. / 3. * rhocPdf.c[c] * kfv.c[c];
        for (int bf = 0; bf < nb; ++bf) ps.b[bf] = pfv.b[bf] - 2. / 3. * rhocPdf.b[bf] * kfv.b[bf];
        makeNodes(ps, pStarN, PDFB200_INTERP_CELL_POINT_FACE);
        const double TU = std::max(prm.relaxTimeU, prm.minRelaxTime);
        const double Tk = std::max(prm.relaxTimek, prm.minRelaxTime);
        VectorField dU; dU.resize(nc, nb);
        for (int c = 0; c < nc; ++c) dU.c[c] = (Ufv.c[c] - UcPdf.c[c]) / TU;
        applyBC(dU);                                      // zeroGradient
        makeNodes(dU, diffUN, prm.interpDiffU);
        diffk.resize(nc);
        for (int c = 0; c < nc; ++c) diffk[c] = (std::sqrt(kfv.c[c] / kcPdf.c[c]) - 1.0) / Tk;
        makeNodes(Ufv, UN, prm.interpU);
        makeNodes(kfv, kN, prm.interpk);
    }

    // --- mcCellLocalTimeStepping::updateInternals (.C:81-161) ---
    if (prm.localTimeStepping == PDFB200_LTS_CELL) {
        const double etaMaxP = std::max(prm.ltsUpperBound, 1.0);
        ScalarField eta; eta.resize(nc, nb);
        const double k0 = prm.k0;
        for (int c = 0; c < nc; ++c) {
            const SymmTensor &R = Rfv.c[c];
            const Vec3 urms2(2 * std::sqrt(std::max(k0, R.xx)), 2 * std::sqrt(std::max(k0, R.yy)),
                             2 * std::sqrt(std::max(k0, R.zz)));
            double dtc = GREAT;
            for (int j = mesh.cellFaceStart[c]; j < mesh.cellFaceStart[c + 1]; ++j) {
                const int f = mesh.cellFaces[j];
                if (f >= nI && mesh.patches[mesh.facePatch[f - nI]].geom == PDFB200_PATCH_EMPTY) continue;
                const Vec3 &coeff = mesh.courantCoeffs[f];
                dtc = std::min(dtc, 1. / (std::fabs(dot(coeff, Ufv.c[c])) + std::fabs(dot(coeff, urms2)) + SMALL));
            }
            eta.c[c] = dtc / deltaT;
        }
        double etaMin = VGREAT;
        for (double v : eta.c) etaMin = std::min(etaMin, v);
        for (double &v : eta.c) v -= etaMin;
        double mx = -VGREAT;
        for (double v : eta.c) mx = std::max(mx, v);
        for (double &v : eta.c) v = (etaMaxP - 1) / mx * v + 1;
        applyBC(eta);                                     // zeroGradient
        makeNodes(eta, etaN, prm.interpEta);
    }
}

// ---------------------------------------------------------------------------
// mcSteadyFlamelet helpers (mcSteadyFlamelet.C:44-93) — verbatim semantics,
// including the transposed weights of binterp.
// ---------------------------------------------------------------------------
void findCoeffs(const std::vector<double> &lst, double v, int &i, double &w) {
    auto iter = std::lower_bound(lst.begin(), lst.end(), v);
    if (iter == lst.begin()) { i = 0; w = 0.; }
    else if (iter == lst.end()) { i = (int)lst.size() - 2; w = 1.; }
    else {
        auto prev = iter - 1;
        i = (int)(prev - lst.begin());
        w = (v - *prev) / (*iter - *prev);
    }
    if (w < -VSMALL || w > (1. + VSMALL)) throw std::runtime_error("findCoeffs: weight not within [0, 1]");
}

double binterp(const std::vector<double> &v, int nCols, int ix, double wx, int iy, double wy) {
    // Nchi == 1 and chi > chi[0] gives ix = -1 in the reference (out-of-bounds
    // row read multiplied by weight 0); the guard keeps the read in bounds.
    const int nRows = (int)v.size() / nCols;
    auto at = [&](int r, int c) {
        r = std::min(std::max(r, 0), nRows - 1);
        return v[(size_t)r * nCols + c];
    };
    const double v00 = at(ix, iy), v10 = at(ix, iy + 1), v01 = at(ix + 1, iy), v11 = at(ix + 1, iy + 1);
    const double wxm = 1. - wx, wym = 1. - wy;
    return v00 * wxm * wym + v10 * wx * wym + v01 * wxm * wy + v11 * wx * wy;
}

// ---------------------------------------------------------------------------
// mcInletRandom (mcInletRandom.C:92-108, mcInletRandomI.H:28-39) and the
// inversion generator (mcInversionInletRandom.C:63-119)
// ---------------------------------------------------------------------------
void InletRandom::updateCoeffs(double UMean_, double uRms_) {
    UMean = UMean_; uRms = uRms_;
    b = M_SQRT1_2 / uRms;
    b2 = b * b;
    expmU2b2 = std::exp(-UMean * UMean * b2);
    UbSqrtPi = UMean * b * std::sqrt(M_PI);
    erfUb = std::erf(UMean * b);
    denom = expmU2b2 + UbSqrtPi * (1. + erfUb);
    b22denom = 2. * b2 / denom;
    Q = 0.5 * (UMean + expmU2b2 / (b * std::sqrt(M_PI)) + UMean * erfUb);
    m = 0.5 * (UMean * b + std::sqrt(2.0 + UMean * UMean * b2)) / b;
}
double InletRandom::PDF(double x) const { const double xp = x - UMean; return b22denom * std::exp(-b2 * xp * xp) * x; }
double InletRandom::CDF(double x) const {
    x -= UMean;
    return (expmU2b2 - std::exp(-b2 * x * x) + UbSqrtPi * (erfUb + std::erf(b * x))) / denom;
}
double InletRandom::newton(double x0, double y, int &nIter) const {
    const int nMaxIter = 50; const double tol = 1e-8;
    nIter = 0;
    do {
        const double fval = CDF(x0) - y, fder = PDF(x0);
        if (std::fabs(fder) < VSMALL) return x0;
        const double x = x0 - fval / fder;
        if (std::fabs(x - x0) < tol) return x;
        x0 = x;
        ++nIter;
    } while (nIter < nMaxIter);
    return x0;
}
double InletRandom::value(double U01) const {
    const double diff = U01 - CDF(m);
    const double d = (diff >= 0 ? 1.0 : -1.0) * SMALL;   // Foam::sign: +1 for >= 0
    int nIter;
    return newton(m + d, U01, nIter);
}

double InletRandom::sample(const KeyedRng &k, uint32_t bf, uint32_t step, uint32_t kDraw) const {
    const double xMax = std::max(UMean + 6.7 * uRms, SMALL);
    const uint64_t entity = (uint64_t)bf | ((uint64_t)kDraw << 32);
    for (uint32_t attempt = 0; attempt < (1u << 20); ++attempt) {
        uint32_t o[4];
        k.words(entity, step, RNG_INJECT_SAMPLE, attempt, o);
        float n0, n1;
        normalPairF32(o[0], o[1], n0, n1);
        const double u = (double)((((unsigned long long)o[3] << 32) | o[2]) >> 11) * 0x1.0p-53;
        const double x = UMean + uRms * (double)n0;
        if (u * xMax < x) return x;
    }
    return m;
}

double InletRandom::sample(SeqRandom &s) const {
    // mcSamplingInletRandom.C:37-102 with the bound of the scale known in advance instead of taken from a batch
    const double xMax = std::max(UMean + 8.0 * uRms, SMALL);
    for (;;) {
        const double x = UMean + uRms * s.GaussNormal();
        if (s.scalar01() * xMax < x) return x;
    }
}

}  // namespace pdforacle
