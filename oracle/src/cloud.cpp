// cloud.cpp — CPU ORACLE (test infrastructure): mcParticleCloud::evolve and the
// per-particle part of every sub-model, restated pass by pass in the
// reference's order (SURVEY.md §3.2 E1–E16).
#include <algorithm>
#include <cstring>
#include <numeric>
#include <stdexcept>

#include "pdforacle.hpp"

namespace pdforacle {

static inline double stabilise(double x, double y) { return x < 0 ? x - y : x + y; }

std::string Cloud::init(const pdfb200_mesh &m, const pdfb200_params &p) {
    try {
        mesh.build(m);
    } catch (const std::exception &e) {
        return e.what();
    }
    std::string err = setParams(p);
    if (!err.empty()) return err;
    deltaT = p.deltaT0 > 0 ? p.deltaT0 : 100 * SMALL;
    krng.setSeed(p.seed);
    srng.seed(p.seed);

    const int nI = mesh.nInternalFaces, nb = nBnd();
    // mcOpenBoundary ctor (mcOpenBoundary.C:56-69): cell -> patch faces
    std::vector<std::vector<int>> cf(mesh.nCells);
    for (int bf = 0; bf < nb; ++bf) {
        const Patch &pp = mesh.patches[mesh.facePatch[bf]];
        if (pp.handler == PDFB200_BC_OPEN || pp.handler == PDFB200_BC_INLET_OUTLET) cf[mesh.owner[nI + bf]].push_back(bf);
    }
    openCellStart.assign(mesh.nCells + 1, 0);
    for (int c = 0; c < mesh.nCells; ++c) openCellStart[c + 1] = openCellStart[c] + (int)cf[c].size();
    openCellFaces.clear();
    for (int c = 0; c < mesh.nCells; ++c) openCellFaces.insert(openCellFaces.end(), cf[c].begin(), cf[c].end());
    Un.assign(nb, 0.0);

    // mcInletOutletBoundary ctor (mcInletOutletBoundary.C:57-103)
    fwdTrans.assign(nb, Tensor::identity()); revTrans.assign(nb, Tensor::identity());
    probVec.assign(nb, {});
    const Vec3 ex(1.0, 0.0, 0.0);
    for (int bf = 0; bf < nb; ++bf) {
        const Patch &pp = mesh.patches[mesh.facePatch[bf]];
        if (pp.handler != PDFB200_BC_INLET_OUTLET) continue;
        const int f = nI + bf, n = mesh.faceSize(f);
        const Vec3 &C = mesh.faceCentres[f];
        probVec[bf].resize(n);
        double area = 0;
        for (int i1 = 0, i2 = 1; i1 < n; ++i1, ++i2) {
            if (i2 == n) i2 = 0;
            const Vec3 v1 = mesh.points[mesh.facePoint(f, i1)] - C, v2 = mesh.points[mesh.facePoint(f, i2)] - C;
            area += mag(cross(v1, v2));
            probVec[bf][i1] = area;
        }
        for (int i = 0; i < n; ++i) probVec[bf][i] /= area;
        fwdTrans[bf] = rotationTensor(-(mesh.faceAreas[f] / mesh.magSf[f]), ex);
        revTrans[bf] = fwdTrans[bf].T();
    }
    patchMassIn.assign(mesh.patches.size(), {}); patchMassOut.assign(mesh.patches.size(), {});
    return "";
}

std::string Cloud::setParams(const pdfb200_params &p) {
    if (p.nPhi < 0 || p.nPhi > PDFB200_MAX_PHI) return "nPhi out of range";
    if (p.nMixed > p.nPhi || p.nConserved > p.nPhi) return "too many mixed/conserved scalars";
    if (!(p.cloneAt >= 0 && p.cloneAt < 1.)) return "cloneAt must be in the range [0, 1)";   // mcSolution.C:146-152
    if (!(p.eliminateAt > 1)) return "eliminateAt must be > 1";                              // mcSolution.C:155-161
    prm = p;
    prm.C0 = std::max(0.0, prm.C0);                       // mcSLMFullVelocityModel.C:97-98
    prm.C1 = std::max(0.0, std::min(1.0, prm.C1));
    return "";
}

// ---------------------------------------------------------------------------
// interpolation at a particle
// ---------------------------------------------------------------------------
void Cloud::baryWeights(const Particle &p, double w[4]) const {
    const Tet &T = mesh.tets[p.tet];
    const Vec3 r = p.position - mesh.cellCentres[T.cell];
    w[0] = dotf(T.gradN[0], r); w[1] = dotf(T.gradN[1], r); w[2] = dotf(T.gradN[2], r);
    w[3] = 1.0 - ((w[0] + w[1]) + w[2]);
}

// interpolationCellPointFace::interpolate (B.2): two points, face, cell, each
// weight clamped to [0,1]; scheme "cell" returns the cell value.
double Cloud::interp(const NodeField<double> &n, const Particle &p) const {
    if (n.scheme == PDFB200_INTERP_CELL) return n.cell[p.cell];
    const Tet &T = mesh.tets[p.tet];
    double w[4];
    baryWeights(p, w);
    const double ts[4] = {n.point[T.ptB], n.point[T.ptC], n.face[T.face], n.cell[T.cell]};
    const double ph[4] = {w[1], w[2], w[0], w[3]};
    double t = 0;
    for (int i = 0; i < 4; ++i) {
        double f = std::min(1.0, ph[i]);
        f = std::max(0.0, f);
        t += f * ts[i];
    }
    return t;
}

Vec3 Cloud::interp(const NodeField<Vec3> &n, const Particle &p) const {
    if (n.scheme == PDFB200_INTERP_CELL) return n.cell[p.cell];
    const Tet &T = mesh.tets[p.tet];
    double w[4];
    baryWeights(p, w);
    const Vec3 ts[4] = {n.point[T.ptB], n.point[T.ptC], n.face[T.face], n.cell[T.cell]};
    const double ph[4] = {w[1], w[2], w[0], w[3]};
    Vec3 t;
    for (int i = 0; i < 4; ++i) {
        double f = std::min(1.0, ph[i]);
        f = std::max(0.0, f);
        t += f * ts[i];
    }
    return t;
}

// gradInterpolationConstantTet<scalar>::interpolate (.C:67-139); the face value
// on a boundary face is the patch value, or the cell value on an empty patch
// (both already folded into n.face by makeNodes).
Vec3 Cloud::gradInterp(const NodeField<double> &n, const Particle &p) const {
    const Tet &T = mesh.tets[p.tet];
    Vec3 grad = (T.gradN[3] * n.cell[T.cell] + T.gradN[1] * n.point[T.ptB]) + T.gradN[2] * n.point[T.ptC];
    grad += T.gradN[0] * n.face[T.face];
    return grad;
}

// ---------------------------------------------------------------------------
// helpers of evolve
// ---------------------------------------------------------------------------
double Cloud::massPerDepth(const Particle &p) const {
    if (mesh.axisym) {
        const double r = mag(p.position - dot(p.position, mesh.axis) * mesh.axis);
        return p.m / (std::max(r, SMALL) * mesh.openingAngle);
    }
    return p.m;
}

void Cloud::constrainParticle(double dt, Particle &p) const {
    Vec3 &u = p.Utracking;
    mesh.constrainDirection(mesh.solutionD, u);
    if (mesh.axisym) {
        Vec3 destPos = p.position + dt * u;
        Vec3 n = cross(mesh.axis, destPos);
        n /= mag(n);
        const Tensor T = rotationTensor(n, mesh.centreNormal);
        // mcParticle::transformProperties (mcParticle.C:238-245)
        p.UParticle = transform(T, p.UParticle);
        p.Ucorrection = transform(T, p.Ucorrection);
        p.Utracking = transform(T, p.Utracking);
        destPos = transform(T, destPos);
        mesh.constrainDirection(mesh.geometricD, destPos);
        u = (destPos - p.position) / dt;
    }
}

void Cloud::computeCourantNo(Particle &p) const {
    p.Co = 0.;
    const Vec3 &U = p.Utracking;
    for (int j = mesh.cellFaceStart[p.cell]; j < mesh.cellFaceStart[p.cell + 1]; ++j) {
        const int f = mesh.cellFaces[j];
        if (f >= mesh.nInternalFaces && mesh.patches[mesh.facePatch[f - mesh.nInternalFaces]].geom == PDFB200_PATCH_EMPTY) continue;
        p.Co = std::max(p.Co, std::fabs(dot(mesh.courantCoeffs[f], U)));
    }
}

static inline Vec3 reflectVec(const Vec3 &v, const Vec3 &n) {
    // transform(I - 2 n n, v)
    const double d = dot(v, n);
    return v - (2.0 * d) * n;
}

// mcParticle::hitPatch (mcParticle.C:211-235) -> mcParticleCloud::hitPatch
// (mcParticleCloudI.H:211-230) -> mcSlipBoundary::hitPatch (.C:61-79) /
// mcOpenBoundary::hitPatch (.C:159-205).
bool Cloud::hitPatch(Particle &p, int face) {
    const int bf = face - mesh.nInternalFaces;
    const int patchI = mesh.facePatch[bf];
    const Patch &pp = mesh.patches[patchI];
    if (pp.geom == PDFB200_PATCH_WEDGE && mesh.nGeometricD() < 3) {
        // mcParticle.C:220-224: a wedge patch snaps the position back onto the centre plane and the track
        // carries on from there; the tet of the snapped point is found with the shared locate rule.
        mesh.constrainDirection(mesh.geometricD, p.position);
        p.tet = mesh.locate(p.position, p.cell);
        return true;
    }
    if (pp.geom != PDFB200_PATCH_GENERIC || pp.handler == PDFB200_BC_EMPTY) {
        // mcEmptyBoundary aborts the run (.C:61-104); a wedge patch of a 3-D mesh (nothing to snap) makes the
        // reference spin until nSteps > 1000 and drop the particle (mcParticle.C:147-158). Both end as a lost particle.
        lostMass[p.cell] += p.m; ++nLost; p.keep = false;
        return false;
    }
    const Vec3 nw = mesh.faceAreas[face] / mesh.magSf[face];
    if (pp.handler == PDFB200_BC_SLIP) {
        p.UParticle = reflectVec(p.UParticle, nw);
        p.Ucorrection = reflectVec(p.Ucorrection, nw);
        p.Utracking = reflectVec(p.Utracking, nw);
        p.reflected = true;
        return false;
    }
    // open / inletOutlet
    if (Un[bf] < 0) {
        const double mpd = massPerDepth(p);
        patchMassIn[patchI][0] -= mpd;
        for (int cs = 0; cs < prm.nConserved; ++cs) patchMassIn[patchI][cs + 1] -= mpd * p.Phi[prm.conserved[cs]];
        p.keep = false; ++nDeleted;
        return false;
    }
    if (!pp.reflecting) {
        const double mpd = massPerDepth(p);
        patchMassOut[patchI][0] -= mpd;
        for (int cs = 0; cs < prm.nConserved; ++cs) patchMassOut[patchI][cs + 1] -= mpd * p.Phi[prm.conserved[cs]];
        p.keep = false; ++nDeleted;
        return false;
    }
    p.UParticle = reflectVec(p.UParticle, nw);
    p.Ucorrection = reflectVec(p.Ucorrection, nw);
    p.Utracking = reflectVec(p.Utracking, nw);
    p.reflected = true;
    p.reflectedAtOpenBoundary = true;
    p.reflectionBoundaryVelocity = nw * Un[bf];
    return false;
}

// mcParticle::move (mcParticle.C:100-179). OpenFOAM's trackToFace is replaced
// by the tet walk of SURVEY Appendix D: the path x0 -> dest is followed through
// the face-centre tets; a sub-step ends at a boundary face (patch handler) or
// at the destination. The exit face of a tet is the one with the smallest
// crossing parameter p_i/(-q_i) among faces with q_i < 0 whose plane lies
// before the destination (p_i + q_i < 0), excluding the face just entered;
// ties keep the lower face index (two-round tournament, see below). The CUDA
// kernel evaluates the same expressions in the same order.
bool Cloud::move(Particle &p, double tdTrackTime) {
    p.keep = true;
    if (p.isOnInletBoundary) {
        if (prm.rngMode == 1) p.stepFraction = srng.scalar01();
        else { double u0, u1; krng.uniform2(p.id, step, RNG_INLET_STEPFRAC, 0, u0, u1); p.stepFraction = u0; }
    } else {
        p.stepFraction = 0;   // Cloud<T>::move resets the step fraction (B.1)
    }
    const double trackTime = p.eta * tdTrackTime;
    double tEnd = (1.0 - p.stepFraction) * trackTime;
    const double dtMax = tEnd;
    p.isOnInletBoundary = false;
    int entry = -1;
    static const int entryMap[4] = {0, 2, 1, 3};

    while (p.keep && tEnd > 0) {
        double dt = std::min(dtMax, tEnd);
        Vec3 destPos = p.position + dt * p.Utracking;
        mesh.constrainDirection(mesh.geometricD, destPos);
        const Vec3 x0 = p.position;
        const Vec3 dx = destPos - x0;
        double tf = 1.0;
        for (;;) {
            const Tet &T = mesh.tets[p.tet];
            const Vec3 r = x0 - mesh.cellCentres[T.cell];
            double pp[4], qq[4];
            pp[0] = dotf(T.gradN[0], r); pp[1] = dotf(T.gradN[1], r); pp[2] = dotf(T.gradN[2], r);
            pp[3] = 1.0 - ((pp[0] + pp[1]) + pp[2]);
            qq[0] = dotf(T.gradN[0], dx); qq[1] = dotf(T.gradN[1], dx); qq[2] = dotf(T.gradN[2], dx);
            qq[3] = -((qq[0] + qq[1]) + qq[2]);
            // exit face: a two-round tournament (a against b, c against d, then the winners); a later face wins only
            // with a strictly smaller crossing parameter p/(-q), compared cross-multiplied. Same result as a scan in
            // face order wherever the comparison is transitive; the kernel evaluates exactly this.
            bool cand[4];
            for (int i = 0; i < 4; ++i) cand[i] = (i != entry) && (qq[i] < 0) && (pp[i] + qq[i] < 0);
            const int w01 = (cand[1] && (!cand[0] || pp[1] * qq[0] > pp[0] * qq[1])) ? 1 : 0;
            const int w23 = (cand[3] && (!cand[2] || pp[3] * qq[2] > pp[2] * qq[3])) ? 3 : 2;
            const bool c01 = cand[0] || cand[1], c23 = cand[2] || cand[3];
            int best = -1;
            if (c01 || c23) best = (c23 && (!c01 || pp[w23] * qq[w01] > pp[w01] * qq[w23])) ? w23 : w01;
            if (best < 0) {          // destination lies in this tet
                p.position = destPos;
                tf = 1.0;
                break;
            }
            ++p.nSteps;
            if (p.nSteps > 1000) {   // mcParticle.C:147-158 -> notifyLostParticle (:2088-2092)
                lostMass[p.cell] += p.m; ++nLost; p.keep = false;
                break;
            }
            if (best != 3 || T.face < mesh.nInternalFaces) {
                p.tet = T.nbr[best];
                p.cell = mesh.tets[p.tet].cell;
                entry = entryMap[best];
                continue;
            }
            // boundary face: stop there and let the patch handler act
            double s = pp[3] / (-qq[3]);
            s = std::min(1.0, std::max(0.0, s));
            p.position = x0 + s * dx;
            tf = s;
            entry = hitPatch(p, T.face) ? -1 : 3;   // relocated on the centre plane: no entry face
            break;
        }
        if (!p.keep) break;
        dt *= tf;
        tEnd -= dt;
        p.stepFraction = 1.0 - tEnd / trackTime;
    }
    return p.keep;
}

// ---------------------------------------------------------------------------
// per-particle model passes
// ---------------------------------------------------------------------------
void Cloud::positionCorrectionCorrect(Particle &p) const {
    if (prm.positionCorrection == PDFB200_POSCORR_NONE) return;
    if (prm.positionCorrection == PDFB200_POSCORR_SIMPLE) {
        // mcSimplePositionCorrection::correct (.C:156-163): Ucorrection -= cmptMultiply(Ainv[cell], interp(gradPhi)) with
        // Ainv = (1./dy*dz, 1./dx*dz, 1./dx*dy) of the cell's bounding box (.C:127) - operator precedence as written
        const Vec3 d = mesh.bbMax[p.cell] - mesh.bbMin[p.cell];
        const Vec3 Ainv(1. / d.y * d.z, 1. / d.x * d.z, 1. / d.x * d.y);
        const Vec3 g = interp(UPosCorrN, p);
        p.Ucorrection -= Vec3(Ainv.x * g.x, Ainv.y * g.y, Ainv.z * g.z);
        return;
    }
    if (prm.positionCorrection == PDFB200_POSCORR_EXTERNAL) {
        // mcMuradogluPositionCorrection::correct (.C:158-169) / mcEllipticRelaxation... (.C:238-248) / mcIntegrated... (.C:132-138)
        const double z = interp(pcZetaN, p), l = interp(pcLN, p);
        p.Ucorrection -= interp(UPosCorrN, p) + (pcScale * l) * (z != 0 ? interp(pcG1N, p) : interp(pcG2N, p));
        return;
    }
    p.Ucorrection += interp(UPosCorrN, p);
}

void Cloud::omegaCorrect(Particle &p) const {
    if (prm.omegaModel == PDFB200_OMEGA_NONE) return;
    p.Omega = std::max(interp(OmegaN, p), SMALL);
}

void Cloud::mixingCorrect(Particle &p) const {
    if (prm.mixingModel != PDFB200_MIXING_IEM) return;
    const double Cmix2 = 0.5 * prm.Cmix;
    for (int mI = 0; mI < prm.nMixed; ++mI) {
        double &phi = p.Phi[prm.mixed[mI]];
        const double phiap = PhicPdf[prm.mixed[mI]].c[p.cell];
        phi -= (1.0 - std::exp(-Cmix2 * p.Omega * p.eta * deltaT)) * (phi - phiap);
    }
}

// "modifiedCurl" pair mixing (no reference; definition in include/pdfb200.h and DESIGN.md).
// Runs on the particles of every cell after tracking, before the moments are taken.
// Per cell and round: order the particles by (hash(id, step, round), id), pair consecutive
// ones, mix pair (p,q) with probability Ppq/R and a ~ U(0,1) towards the weighted pair mean.
void Cloud::curlMixing() {
    if (prm.mixingModel != PDFB200_MIXING_MODCURL || prm.nMixed == 0) return;
    const int RMAX = 8;
    std::vector<std::vector<size_t>> addr(mesh.nCells);
    for (size_t i = 0; i < particles.size(); ++i) addr[particles[i].cell].push_back(i);
    struct Ent { uint32_t h; uint64_t id; size_t i; };
    std::vector<Ent> ord;
    for (int c = 0; c < mesh.nCells; ++c) {
        const std::vector<size_t> &L = addr[c];
        const int n = (int)L.size();
        if (n < 2) continue;
        double wSum = 0, wMax = 0, oeMax = 0;
        for (size_t i : L) {
            const Particle &p = particles[i];
            const double w = p.eta * p.m;
            wSum += w; wMax = std::max(wMax, w); oeMax = std::max(oeMax, p.Omega * p.eta);
        }
        const double wBar = wSum / n;
        const double pc = 3.0 * prm.Cmix * deltaT * oeMax * (wMax / wBar);
        const int R = std::min(RMAX, std::max(1, (int)std::ceil(pc)));
        for (int r = 0; r < R; ++r) {
            ord.clear();
            for (size_t i : L) {
                uint32_t h;
                if (prm.rngMode == 1) h = (uint32_t)(srng.scalar01() * 4294967296.0);
                else {
                    uint32_t o[4];
                    krng.words(particles[i].id, step, RNG_MIX_HASH, (uint32_t)r, o);
                    h = o[0];
                }
                ord.push_back({h, particles[i].id, i});
            }
            std::sort(ord.begin(), ord.end(), [](const Ent &a, const Ent &b) { return a.h < b.h || (a.h == b.h && a.id < b.id); });
            for (int j = 0; j + 1 < n; j += 2) {
                Particle &p = particles[ord[j].i], &q = particles[ord[j + 1].i];
                double u, a;
                if (prm.rngMode == 1) { u = srng.scalar01(); a = srng.scalar01(); }
                else krng.uniform2(p.id, step, RNG_MIX_PAIR, (uint32_t)r, u, a);
                const double wp = p.eta * p.m, wq = q.eta * q.m;
                const double Om = 0.5 * (p.Omega + q.Omega), et = 0.5 * (p.eta + q.eta);
                const double Ppq = 3.0 * prm.Cmix * Om * et * deltaT * ((wp + wq) / (2.0 * wBar)) / R;
                if (!(u < Ppq)) continue;
                for (int mI = 0; mI < prm.nMixed; ++mI) {
                    const int s = prm.mixed[mI];
                    const double mean = (wp * p.Phi[s] + wq * q.Phi[s]) / (wp + wq);
                    p.Phi[s] += a * (mean - p.Phi[s]);
                    q.Phi[s] += a * (mean - q.Phi[s]);
                }
            }
        }
    }
}

void Cloud::reactionCorrect(Particle &p) const {
    if (prm.reactionModel == PDFB200_REACTION_KULKARNI && aiFields.size() < (size_t)(2 + prm.nAdded))
        throw std::invalid_argument("KulkarniAutoIgnition selected but no auto-ignition tables were set");
    switch (prm.reactionModel) {
    case PDFB200_REACTION_COLD:                  // mcColdReactionModel.C:62-65
        p.rho = prm.coldDensity;
        break;
    case PDFB200_REACTION_BURKE_SCHUMANN: {      // mcBurkeSchumannReactionModel.C:77-91
        const double z = p.Phi[prm.zIdx];
        const double RTox = prm.bsRox * prm.bsTox, RTfuel = prm.bsRfuel * prm.bsTfuel, RTstoi = prm.bsRstoi * prm.bsTstoi;
        const double RT = z < prm.bsZstoi ? (RTstoi - RTox) / prm.bsZstoi * z + RTox
                                          : (RTfuel - RTstoi) / (1 - prm.bsZstoi) * (z - prm.bsZstoi) + RTstoi;
        const double T = z < prm.bsZstoi ? (prm.bsTstoi - prm.bsTox) / prm.bsZstoi * z + prm.bsTox
                                         : (prm.bsTfuel - prm.bsTstoi) / (1 - prm.bsZstoi) * (z - prm.bsZstoi) + prm.bsTstoi;
        p.rho = pfv.c[p.cell] / RT;
        p.Phi[prm.TIdx] = T;
        break;
    }
    case PDFB200_REACTION_NAIVE_FLAMELET: {      // mcNaiveFlameletModel.C:64-69
        const double z = p.Phi[prm.zIdx];
        p.rho = (1. - 3.2 * z * (1. - z));
        p.Phi[prm.TIdx] = 1e5 / p.rho;
        break;
    }
    case PDFB200_REACTION_STEADY_FLAMELET: {     // mcSteadyFlamelet.C:304-329
        const double z = p.Phi[prm.zIdx];
        const double zVar = interp(zVarN, p);
        const double chi = std::max(prm.Cchi * p.Omega * zVar, SMALL);
        int iz, ichi; double wz, wchi;
        findCoeffs(flZ, z, iz, wz);
        findCoeffs(flChi, chi, ichi, wchi);
        p.Phi[prm.chiIdx] = chi;
        p.rho = binterp(flFields[0], nZ, ichi, wchi, iz, wz);
        for (int i = 0; i < prm.nAdded; ++i) p.Phi[prm.addedIdx[i]] = binterp(flFields[1 + i], nZ, ichi, wchi, iz, wz);
        break;
    }
    case PDFB200_REACTION_KULKARNI: {            // mcKulkarniAutoIgnition.C:339-380
        const double z = p.Phi[prm.zIdx];
        double &c = p.Phi[prm.cIdx];
        int iz, ic; double wz, wc;
        findCoeffs(aiZ, z, iz, wz);
        const double cEq = aiCEq[iz] * (1. - wz) + aiCEq[iz + 1] * wz;
        c = std::min(c, cEq);
        const double pv = cEq > 1e-5 * aiCEqMax ? c / cEq : 0.;
        p.Phi[prm.pvIdx] = pv;
        findCoeffs(aiPv, pv, ic, wc);
        const int nPv = (int)aiPv.size();
        const double cdot = binterp(aiFields[0], nPv, iz, wz, ic, wc);
        c += cdot * p.eta * deltaT;
        findCoeffs(aiPv, pv, ic, wc);
        p.rho = binterp(aiFields[1], nPv, iz, wz, ic, wc);
        for (int i = 0; i < prm.nAdded; ++i) p.Phi[prm.addedIdx[i]] = binterp(aiFields[2 + i], nPv, iz, wz, ic, wc);
        p.Co = std::max(p.Co, cdot);
        break;
    }
    default: break;
    }
}

// mcSLMFullVelocityModel::correct (.C:146-185)
void Cloud::velocityCorrect(Particle &p, const double xiIn[3]) const {
    if (prm.velocityModel == PDFB200_VELOCITY_NONE) return;
    const double deltaTg = deltaT;
    const double dT = p.eta * deltaTg;
    const Vec3 UFap = interp(UN, p);
    const Vec3 gradPFap = gradInterp(pStarN, p);
    const double kFap = std::max(interp(kN, p), prm.kMin);
    const Vec3 diffUap = interp(diffUN, p);
    const Vec3 xi(xiIn[0], xiIn[1], xiIn[2]);
    const double A = -(0.5 * prm.C1 + 0.75 * prm.C0) * p.Omega;
    const Vec3 B = -(gradPFap / p.rho + A * UFap);
    const Vec3 deltaU = (B + A * p.UParticle) * dT + std::sqrt(prm.C0 * kFap * p.Omega * dT) * xi;
    const Vec3 Uold = p.UParticle;
    p.UParticle += ((deltaU + 0.5 * A * deltaU * dT) + diffUap * deltaTg) + (Uold - UFap) * diffk[p.cell] * deltaTg;
    p.Co = std::max(p.Co, p.Omega);
}

void Cloud::ltsCorrect(Particle &p) const {
    switch (prm.localTimeStepping) {
    case PDFB200_LTS_CELL: p.eta = interp(etaN, p); break;                    // mcCellLocalTimeStepping.C:164-174
    case PDFB200_LTS_PARTICLE:                                                 // mcParticleLocalTimeStepping.C:71-80
        p.eta = std::min(prm.CFL / stabilise(deltaT * p.Co, VSMALL), prm.ltsUpperBound);
        break;
    default: p.eta = 1.0; break;                                               // mcLocalTimeStepping.C:101-104
    }
}

// The six N(0,1) deviates of a particle step: out[0..2] for the Langevin term
// (mcSLMFullVelocityModel.C:165-171), out[3..5] for the numerical diffusion of
// E9 (mcParticleCloud.C:1319-1325). phase 0 is called in the velocity pass,
// phase 1 in the E9 loop, which is also the order in which the reference
// consumes its sequential stream.
void Cloud::stepNormals(const Particle &p, double out[6], int phase) {
    if (prm.rngMode == 1) {
        for (int i = 3 * phase; i < 3 * phase + 3; ++i) out[i] = srng.GaussNormal();
    } else if (phase == 0) {
        krng.normal4(p.id, step, RNG_STEP_NORMALS, 0, out);
        krng.normal2(p.id, step, RNG_STEP_NORMALS, 1, out[4], out[5]);
    }
}

// ---------------------------------------------------------------------------
// boundaries before the move: mcOpenBoundary::correct(false) (.C:84-141) and
// mcInletOutletBoundary::correct(false) (.C:232-354)
// ---------------------------------------------------------------------------
void Cloud::cullPatch(int patchI) {
    const Patch &pp = mesh.patches[patchI];
    const int nI = mesh.nInternalFaces;
    // Un_ = n & U_patch
    for (int j = 0; j < pp.size; ++j) {
        const int bf = pp.start + j - nI;
        const Vec3 n = mesh.faceAreas[nI + bf] / mesh.magSf[nI + bf];
        Un[bf] = dot(n, Ufv.b[bf]);
    }
    for (Particle &p : particles) {
        if (!p.keep) continue;
        for (int k = openCellStart[p.cell]; k < openCellStart[p.cell + 1]; ++k) {
            const int bf = openCellFaces[k];
            if (mesh.facePatch[bf] != patchI) continue;
            if (!(Un[bf] < 0.)) {
                const Vec3 n = mesh.faceAreas[nI + bf] / mesh.magSf[nI + bf];
                const double xp = dot(n, mesh.faceCentres[nI + bf] - p.position);
                if (xp < p.eta * (Un[bf] * deltaT)) {
                    const double mpd = massPerDepth(p);
                    patchMassOut[patchI][0] -= mpd;
                    for (int cs = 0; cs < prm.nConserved; ++cs) patchMassOut[patchI][cs + 1] -= mpd * p.Phi[prm.conserved[cs]];
                    p.keep = false; ++nDeleted;
                    break;   // the reference would delete the same particle twice here
                }
            }
        }
    }
}

void Cloud::injectPatch(int patchI) {
    const Patch &pp = mesh.patches[patchI];
    const int nI = mesh.nInternalFaces;
    const int Npc = prm.particlesPerCell;
    for (int j = 0; j < pp.size; ++j) {
        const int bf = pp.start + j - nI, f = nI + bf;
        const int cellI = mesh.owner[f];
        if ((bf % nRanks) != rank) continue;   // particle-sharded runs: inflow faces are dealt round-robin to the ranks
        if (Un[bf] > 0) continue;   // outlet face
        InletRandom inrnd;
        inrnd.updateCoeffs(inletU[bf].x, std::sqrt(inletR[bf].xx));
        const double mp = rhocPdf.b[bf] * mesh.cellVolumes[cellI] / Npc;
        const double mIn = rhocPdf.b[bf] * inrnd.Q * mesh.magSf[f] * deltaT;
        std::vector<size_t> gen;
        double mGen = 0;
        uint32_t kDraw = 0;
        while (mGen < mIn) {
            double uVal, uTri, n1, n2, a1, a2, uAcc;
            if (prm.rngMode == 1) {
                uVal = srng.scalar01(); n1 = srng.GaussNormal(); n2 = srng.GaussNormal();
                uTri = srng.scalar01(); a1 = srng.scalar01(); a2 = srng.scalar01();
                uAcc = -1;   // drawn lazily below
            } else {
                double dummy;
                krng.uniform2((uint32_t)bf, step, RNG_INJECT, 4 * kDraw + 0, uVal, uTri);
                krng.normal2((uint32_t)bf, step, RNG_INJECT, 4 * kDraw + 1, n1, n2);
                krng.uniform2((uint32_t)bf, step, RNG_INJECT, 4 * kDraw + 2, a1, a2);
                krng.uniform2((uint32_t)bf, step, RNG_INJECT, 4 * kDraw + 3, uAcc, dummy);
            }
            ++kDraw;
            // randomVelocity (.C:215-229)
            const Vec3 &U = inletU[bf]; const SymmTensor &R = inletR[bf];
            const double uIn = prm.inletRandom != PDFB200_INLET_RANDOM_SAMPLING ? inrnd.value(uVal)
                             : (prm.rngMode == 1 ? inrnd.sample(srng) : inrnd.sample(krng, (uint32_t)bf, step, kDraw - 1));
            const Vec3 UpLocal(uIn, U.y + std::sqrt(R.yy) * n1, U.z + std::sqrt(R.zz) * n2);
            const Vec3 Up = transform(revTrans[bf], UpLocal);
            // randomPoint (.C:179-212)
            const std::vector<double> &w = probVec[bf];
            int idx1 = (int)(std::lower_bound(w.begin(), w.end(), uTri) - w.begin());
            int idx2 = idx1 + 1;
            if (idx2 == (int)w.size()) idx2 = 0;
            if (a1 + a2 > 1.0) { a1 = 1.0 - a1; a2 = 1.0 - a2; }
            Vec3 pos = ((a1 * mesh.points[mesh.facePoint(f, idx1)] + a2 * mesh.points[mesh.facePoint(f, idx2)])
                        + (1.0 - a1 - a2) * mesh.faceCentres[f]) - SMALL * mesh.faceAreas[f];
            if (mesh.nGeometricD() <= 2) mesh.constrainDirection(mesh.geometricD, pos);
            // new mcParticle(...) (mcParticle.C:61-96)
            Particle p;
            p.position = pos; p.cell = cellI; p.tet = mesh.locate(pos, cellI);
            p.m = mp; p.UParticle = Up; p.Utracking = Up;
            mesh.constrainDirection(mesh.geometricD, p.Utracking);
            computeCourantNo(p);
            for (int i = 0; i < prm.nPhi; ++i) p.Phi[i] = PhicPdf[i].b[bf];
            p.isOnInletBoundary = true;
            p.injPatch = patchI;
            ltsCorrect(p);
            p.m /= p.eta;
            if (mGen + p.m > mIn) {
                if (prm.rngMode == 1) uAcc = srng.scalar01();
                if (uAcc > (mIn - mGen) / p.m) break;
            }
            p.id = nextId; nextId += (uint64_t)nRanks;
            particles.push_back(p);
            gen.push_back(particles.size() - 1);
            mGen += p.m;
            ++nInjected;
        }
        if (!gen.empty()) {
            std::vector<Particle *> ptrs;
            for (size_t g : gen) ptrs.push_back(&particles[g]);
            adjustAxiSymmetricMass(ptrs);
            for (Particle *q : ptrs) {
                const double mpd = massPerDepth(*q);
                patchMassIn[patchI][0] += mpd;
                for (int cs = 0; cs < prm.nConserved; ++cs) patchMassIn[patchI][cs + 1] += mpd * q->Phi[prm.conserved[cs]];
            }
        }
    }
}

void Cloud::boundaryCorrectBefore() {
    const int nI = mesh.nInternalFaces;
    // mcInletOutletBoundary::getU()/getR() cache their first evaluation (.C:107-151)
    if (!inletCached) {
        inletU.assign(nBnd(), Vec3()); inletR.assign(nBnd(), SymmTensor());
        for (int bf = 0; bf < nBnd(); ++bf) {
            if (mesh.patches[mesh.facePatch[bf]].handler != PDFB200_BC_INLET_OUTLET) continue;
            inletU[bf] = transform(fwdTrans[bf], Ufv.b[bf]);
            SymmTensor R = transform(fwdTrans[bf], TaucPdf.b[bf]);
            R.xx = std::max(R.xx, prm.k0); R.yy = std::max(R.yy, prm.k0); R.zz = std::max(R.zz, prm.k0);
            R.xy = std::max(R.xy, -GREAT); R.xz = std::max(R.xz, -GREAT); R.yz = std::max(R.yz, -GREAT);
            inletR[bf] = R;
        }
        inletCached = true;
    }
    (void)nI;
    for (size_t patchI = 0; patchI < mesh.patches.size(); ++patchI) {
        patchMassIn[patchI].fill(0.); patchMassOut[patchI].fill(0.);
        const Patch &pp = mesh.patches[patchI];
        if (pp.handler != PDFB200_BC_OPEN && pp.handler != PDFB200_BC_INLET_OUTLET) continue;
        if (pp.reflecting) cullPatch((int)patchI);   // mcOpenBoundary::correct returns early if !reflecting_
        if (pp.handler == PDFB200_BC_INLET_OUTLET) injectPatch((int)patchI);
    }
    // drop culled particles (Cloud::deleteParticle)
    particles.erase(std::remove_if(particles.begin(), particles.end(), [](const Particle &p) { return !p.keep; }), particles.end());
}

void Cloud::boundaryCorrectAfter() {
    // mcOpenBoundary::correct(true) (.C:142-155)
    for (Particle &p : particles) {
        if (p.reflectedAtOpenBoundary) {
            p.UParticle += 2 * p.reflectionBoundaryVelocity;
            p.reflectedAtOpenBoundary = false;
        }
    }
}

void Cloud::adjustAxiSymmetricMass(std::vector<Particle *> &ps) const {
    if (!mesh.axisym) return;
    std::vector<double> r(ps.size());
    double sumR = 0.;
    for (size_t i = 0; i < ps.size(); ++i) {
        const Particle &p = *ps[i];
        r[i] = mag(p.position - dot(p.position, mesh.axis) * mesh.axis);
        sumR += r[i];
    }
    const double alpha = ps.size() / sumR;
    for (size_t i = 0; i < ps.size(); ++i) ps[i]->m *= alpha * r[i];
}

// ---------------------------------------------------------------------------
// updateCloudPDF (mcParticleCloud.C:824-1009), split into the particle loop
// (:933-956) and the time-averaging / derived fields (:958-1008) so that a
// particle-sharded run can sum the instantaneous block across ranks in between.
// Block layout per cell: N, M, V, MU(3), MPhi(nPhi), MPhiPhi(nCov), MUU(6), MUPhi(3 nPhi)
// (the last group is the velocity-scalar flux moment, new: not in the reference).
// ---------------------------------------------------------------------------
void Cloud::accumulateInstantMoments(std::vector<double> &block) const {
    const int nM = nMoments();
    block.assign(momentBlockSize(), 0.0);
    for (const Particle &p : particles) {
        double *b = &block[(size_t)p.cell * nM];
        const Vec3 &Up = p.UParticle;
        b[0] += 1;
        const double mpd = p.eta * massPerDepth(p);
        b[1] += mpd;
        b[2] += mpd / p.rho;
        b[3] += mpd * Up.x; b[4] += mpd * Up.y; b[5] += mpd * Up.z;
        int o = 6, oc = 6 + prm.nPhi;
        for (int i = 0; i < prm.nPhi; ++i) {
            b[o + i] += mpd * p.Phi[i];
            for (int j = i; j < prm.nPhi; ++j, ++oc) b[oc] += mpd * p.Phi[i] * p.Phi[j];
        }
        double *uu = b + 6 + prm.nPhi + nCov();
        uu[0] += mpd * (Up.x * Up.x); uu[1] += mpd * (Up.x * Up.y); uu[2] += mpd * (Up.x * Up.z);
        uu[3] += mpd * (Up.y * Up.y); uu[4] += mpd * (Up.y * Up.z); uu[5] += mpd * (Up.z * Up.z);
        double *up = uu + 6;
        for (int i = 0; i < prm.nPhi; ++i) {
            const double mp = mpd * p.Phi[i];
            up[3 * i] += mp * Up.x; up[3 * i + 1] += mp * Up.y; up[3 * i + 2] += mp * Up.z;
        }
    }
}

void Cloud::finaliseMoments(const std::vector<double> &block, double existWt) {
    const int nc = mesh.nCells, nM = nMoments();
    const double newWt = 1.0 - existWt;
    const double SMALL_MASS = SMALL, SMALL_VOLUME = SMALL;
    for (int c = 0; c < nc; ++c) {
        const double *b = &block[(size_t)c * nM];
        PaNIC[c] = b[0];
        const double mI = b[1], VI = b[2];
        mMom[c] = existWt * mMom[c] + newWt * mI;
        const double mB = std::max(mMom[c], SMALL_MASS);
        pndcPdfInst.c[c] = mI / mesh.volumeOrArea[c];
        pndcPdf.c[c] = mMom[c] / mesh.volumeOrArea[c];
        VMom[c] = existWt * VMom[c] + newWt * VI;
        const double VB = std::max(VMom[c], SMALL_VOLUME);
        rhocPdfInst.c[c] = mI / std::max(VI, SMALL_VOLUME);
        rhocPdf.c[c] = mMom[c] / VB;
        UMom[c] = existWt * UMom[c] + newWt * Vec3(b[3], b[4], b[5]);
        UcPdf.c[c] = UMom[c] / mB;
        int oc = 6 + prm.nPhi, pp = 0;
        for (int i = 0; i < prm.nPhi; ++i) {
            PhiMom[i][c] = existWt * PhiMom[i][c] + newWt * b[6 + i];
            PhicPdf[i].c[c] = PhiMom[i][c] / mB;
        }
        for (int i = 0; i < prm.nPhi; ++i)
            for (int j = i; j < prm.nPhi; ++j, ++pp, ++oc) {
                PhiPhiMom[pp][c] = existWt * PhiPhiMom[pp][c] + newWt * b[oc];
                PhiPhicPdf[pp].c[c] = PhiPhiMom[pp][c] / mB - PhicPdf[i].c[c] * PhicPdf[j].c[c];
            }
        const double *uu = b + 6 + prm.nPhi + nCov();
        SymmTensor &M = UUMom[c];
        M.xx = existWt * M.xx + newWt * uu[0]; M.xy = existWt * M.xy + newWt * uu[1]; M.xz = existWt * M.xz + newWt * uu[2];
        M.yy = existWt * M.yy + newWt * uu[3]; M.yz = existWt * M.yz + newWt * uu[4]; M.zz = existWt * M.zz + newWt * uu[5];
        const Vec3 &u = UcPdf.c[c];
        SymmTensor &T = TaucPdf.c[c];
        T.xx = M.xx / mB - u.x * u.x; T.xy = M.xy / mB - u.x * u.y; T.xz = M.xz / mB - u.x * u.z;
        T.yy = M.yy / mB - u.y * u.y; T.yz = M.yz / mB - u.y * u.z; T.zz = M.zz / mB - u.z * u.z;
        kcPdf.c[c] = 0.5 * (T.xx + T.yy + T.zz);
        const double *up = uu + 6;
        for (int i = 0; i < prm.nPhi; ++i) {
            Vec3 &F = UPhiMom[i][c];
            F = existWt * F + newWt * Vec3(up[3 * i], up[3 * i + 1], up[3 * i + 2]);
            UPhiFlux[i][c] = F / mB - PhicPdf[i].c[c] * u;
        }
    }
    applyBC(pndcPdfInst); applyBC(pndcPdf); applyBC(rhocPdfInst); applyBC(rhocPdf);
    applyBC(UcPdf);
    for (auto &f : PhicPdf) applyBC(f);
    for (auto &f : PhiPhicPdf) applyBC(f);
    applyBC(TaucPdf);
    applyBC(kcPdf);
    // bound(kcPdf_, kMin): the tutorials' fields never go below kMin by more
    // than round-off; OpenFOAM's bound() replaces offending values by
    // max(k, local average of positive neighbours, kMin). Restated as a floor.
    for (double &v : kcPdf.c) v = std::max(v, prm.kMin);
    for (double &v : kcPdf.b) v = std::max(v, prm.kMin);
}

void Cloud::updateCloudPDF(double existWt) {
    std::vector<double> block;
    accumulateInstantMoments(block);
    finaliseMoments(block, existWt);
}

// mcParticleCloud::randomPointsInCell (mcParticleCloud.C:2040-2085)
std::vector<Vec3> Cloud::randomPointsInCell(int n, int cell, uint64_t entityBase, uint32_t purpose, uint32_t stepCtr) {
    std::vector<Vec3> pts(n);
    const Vec3 minb = mesh.bbMin[cell];
    const Vec3 dimb = mesh.bbMax[cell] - minb;
    for (int i = 0; i < n; ++i) {
        for (uint32_t attempt = 0;; ++attempt) {
            double ux, uy, uz, dummy;
            if (prm.rngMode == 1) { ux = srng.scalar01(); uy = srng.scalar01(); uz = srng.scalar01(); }
            else {
                krng.uniform2(entityBase + i, stepCtr, purpose, 2 * attempt, ux, uy);
                krng.uniform2(entityBase + i, stepCtr, purpose, 2 * attempt + 1, uz, dummy);
            }
            const double rx = std::min(std::max(10.0 * SMALL, ux), 1.0 - 10.0 * SMALL);
            const double ry = std::min(std::max(10.0 * SMALL, uy), 1.0 - 10.0 * SMALL);
            const double rz = std::min(std::max(10.0 * SMALL, uz), 1.0 - 10.0 * SMALL);
            Vec3 position = minb + Vec3(rx * dimb.x, ry * dimb.y, rz * dimb.z);
            if (mesh.nGeometricD() <= 2) mesh.constrainDirection(mesh.geometricD, position);
            if (mesh.pointInCell(position, cell) || attempt > 10000) { pts[i] = position; break; }
        }
    }
    return pts;
}

// mcParticleCloud::particleNumberControl / cloneParticles / eliminateParticles
// (mcParticleCloud.C:1014-1221). Keyed-RNG mode draws per particle id so that
// the outcome does not depend on list order; ties of eta*m sort by id.
void Cloud::particleNumberControl() {
    if (!prm.particleNumberControl) return;
    const int Npc = prm.particlesPerCell;
    const int nc = mesh.nCells;
    std::vector<std::vector<size_t>> addr(nc);
    std::vector<int> flag(nc, 0);   // 0 normal, 1 too few, 2 too many, -1 empty
    // local population per cell (multi-rank: decisions use the global PaNIC, actions the local list)
    std::vector<int> nLocal(nc, 0);
    for (const Particle &p : particles) nLocal[p.cell]++;
    for (int c = 0; c < nc; ++c) {
        const int np = (int)std::round(PaNIC[c]);
        if (np < 1) flag[c] = -1;
        else if (np < std::round(Npc * prm.cloneAt)) flag[c] = 1;
        else if (np > std::round(Npc * prm.eliminateAt)) flag[c] = 2;
    }
    for (size_t i = 0; i < particles.size(); ++i)
        if (flag[particles[i].cell] > 0) addr[particles[i].cell].push_back(i);

    std::vector<Particle> clones;
    for (int c = 0; c < nc; ++c) {
        if (flag[c] == 1) {
            // cloneParticles (:1107-1139)
            const int npGlobal = (int)std::round(PaNIC[c]);
            int n = Npc - npGlobal;
            n = std::min(npGlobal, n);
            // particle-sharded runs (no counterpart in the reference): every rank steers its own share of the cell towards
            // 1/nRanks of the population the cell has after cloning. (Shares proportional to the local counts have no
            // restoring force: the ranks' holdings drift apart like competing species - 6 % after 400 steps on 4 ranks.)
            if (nRanks > 1) n = (int)std::min<double>(nLocal[c], std::max(0., std::floor((double)(npGlobal + n) / nRanks - nLocal[c] + 0.5)));
            auto &lst = addr[c];
            std::sort(lst.begin(), lst.end(), [&](size_t a, size_t b) {
                // weights rounded to single precision, ties by id: see cloneWeight in pdffoam_b200/csrc/particles.cu
                const float ma = (float)(particles[a].eta * particles[a].m), mb = (float)(particles[b].eta * particles[b].m);
                if (ma != mb) return ma > mb;
                return particles[a].id < particles[b].id;
            });
            for (int i = 0; i < n; ++i) {
                Particle &p = particles[lst[i]];
                std::vector<Vec3> pos = randomPointsInCell(1, c, p.id, RNG_CLONE_POS, step);
                p.m /= 2.0;
                Particle q = p;
                q.position = pos[0];
                q.tet = mesh.locate(q.position, c);
                q.id = nextId; nextId += (uint64_t)nRanks;
                clones.push_back(q);
            }
            PaNIC[c] += n;
            nCloned += n;
        } else if (flag[c] == 2) {
            // eliminateParticles (:1143-1221)
            const int ncur = (int)std::round(PaNIC[c]);
            const int nx = ncur - Npc;
            std::vector<size_t> popA, popB;
            double mA = 0., mB = 0.;
            double P = (2. * nx) / ncur;
            // particle-sharded runs: the rank removes what it holds above its share Npc / nRanks of the cell
            if (nRanks > 1) P = std::min(1., 2. * std::max(0., nLocal[c] - (double)Npc / nRanks) / std::max(nLocal[c], 1));
            for (size_t idx : addr[c]) {
                Particle &p = particles[idx];
                double u1, u2;
                if (prm.rngMode == 1) { u1 = srng.scalar01(); u2 = -1; }
                else krng.uniform2(p.id, step, RNG_NC_PARTICLE, 0, u1, u2);
                if (u1 < P) {
                    const double meta = p.eta * p.m;
                    if (prm.rngMode == 1) u2 = srng.scalar01();
                    if (u2 < 0.5) { popA.push_back(idx); mA += meta; }
                    else { popB.push_back(idx); mB += meta; }
                }
            }
            if (mA > 0 || mB > 0) {
                P = mB / (mA + mB);
                double s, u, dummy;
                if (prm.rngMode == 1) u = srng.scalar01();
                else krng.uniform2((uint32_t)c, step, RNG_NC_CELL, (uint32_t)rank, u, dummy);
                std::vector<size_t> *popDel, *popKeep;
                if (u < P) { s = 1. / P; popDel = &popA; popKeep = &popB; }
                else { s = (mA + mB) / mA; popDel = &popB; popKeep = &popA; }
                for (size_t idx : *popKeep) particles[idx].m *= s;
                int nKilled = 0;
                for (size_t idx : *popDel) { particles[idx].keep = false; ++nKilled; }
                PaNIC[c] -= nKilled;
                nEliminated += nKilled;
            }
        }
    }
    particles.erase(std::remove_if(particles.begin(), particles.end(), [](const Particle &p) { return !p.keep; }), particles.end());
    particles.insert(particles.end(), clones.begin(), clones.end());
}

// mcParticleCloud::initReleaseParticles + particleGenInCell
// (mcParticleCloud.C:1534-1629)
void Cloud::seedParticles() {
    particles.clear();
    const int Npc = prm.particlesPerCell;
    // multi-rank: rank r seeds particles j = r, r+nRanks, ... of every cell
    for (int c = 0; c < mesh.nCells; ++c) {
        const double m = mesh.cellVolumes[c] * rhocPdf.c[c] / Npc;
        const Vec3 Updf = Ufv.c[c];
        const double urms = std::sqrt(2. / 3. * kfv.c[c]);
        std::vector<Particle *> gen;
        const size_t first = particles.size();
        for (int j = rank; j < Npc; j += nRanks) {
            const uint64_t idx = (uint64_t)((int64_t)c * Npc + j);
            // every draw is keyed by the *global* particle index, so the union of
            // all ranks' particles equals the single-rank population
            std::vector<Vec3> pos = randomPointsInCell(1, c, idx, RNG_SEED, 0);
            double g0, g1, g2, g3;
            if (prm.rngMode == 1) { g0 = srng.GaussNormal(); g1 = srng.GaussNormal(); g2 = srng.GaussNormal(); }
            else { krng.normal2(idx, 0, RNG_SEED, 1000000, g0, g1); krng.normal2(idx, 0, RNG_SEED, 1000001, g2, g3); }
            Particle p;
            p.position = pos[0]; p.cell = c; p.tet = mesh.locate(p.position, c);
            p.m = m;
            p.UParticle = Vec3(g0 * urms, g1 * urms, g2 * urms) + Updf;
            p.Utracking = p.UParticle;
            mesh.constrainDirection(mesh.geometricD, p.Utracking);
            computeCourantNo(p);
            for (int i = 0; i < prm.nPhi; ++i) p.Phi[i] = PhicPdf[i].c[c];
            p.id = idx;
            particles.push_back(p);
        }
        for (size_t i = first; i < particles.size(); ++i) gen.push_back(&particles[i]);
        adjustAxiSymmetricMass(gen);
    }
    nextId = (uint64_t)((int64_t)mesh.nCells * Npc + rank);   // ids of later particles: stride nRanks
    // correct mass according to local time-stepping, then Omega and reaction
    updateModelInternals();
    for (Particle &p : particles) { ltsCorrect(p); p.m /= p.eta; }
    for (Particle &p : particles) omegaCorrect(p);
    for (Particle &p : particles) reactionCorrect(p);
}

// ---------------------------------------------------------------------------
// mcParticleCloud::evolve (mcParticleCloud.C:1226-1530)
// ---------------------------------------------------------------------------
void Cloud::evolveBegin(pdfb200_step_report *rep, std::vector<double> &block) {
    nInjected = nDeleted = nCloned = nEliminated = nLost = 0;
    updateModelInternals();

    // E1 (:1230-1236)
    boundaryCorrectBefore();
    // E2 (:1238-1249)
    deltaMassInst.fill(0.);
    for (Particle &p : particles) {
        const double mpd = massPerDepth(p);
        deltaMassInst[0] -= mpd;
        for (int cs = 0; cs < prm.nConserved; ++cs) deltaMassInst[cs + 1] -= mpd * p.Phi[prm.conserved[cs]];
        p.Ucorrection = Vec3();
    }
    // E3 (:1250)
    for (Particle &p : particles) positionCorrectionCorrect(p);
    std::fill(lostMass.begin(), lostMass.end(), 0.0);
    // E4 (:1257-1270)
    for (Particle &p : particles) {
        p.nSteps = 0;
        p.Utracking = p.UParticle + p.Ucorrection;
        constrainParticle(deltaT / 2, p);
        p.reflected = false;
        p.UParticleOld = p.UParticle;
        p.positionOld = p.position;
        p.cellOld = p.cell; p.tetOld = p.tet;
    }
    // E5 (:1272-1277)
    for (Particle &p : particles) move(p, deltaT / 2.);
    particles.erase(std::remove_if(particles.begin(), particles.end(), [](const Particle &p) { return !p.keep; }), particles.end());
    // E6 (:1280-1284)
    for (Particle &p : particles) { p.nSteps = 0; computeCourantNo(p); }
    // E7 (:1285-1289)
    for (Particle &p : particles) omegaCorrect(p);
    for (Particle &p : particles) mixingCorrect(p);
    for (Particle &p : particles) reactionCorrect(p);
    // the six normals of a step are drawn together per particle (3 for the
    // Langevin term, 3 for the numerical diffusion of E9)
    std::vector<std::array<double, 6>> normals(particles.size());
    for (size_t i = 0; i < particles.size(); ++i) {
        stepNormals(particles[i], normals[i].data(), 0);
        velocityCorrect(particles[i], normals[i].data());
    }
    for (Particle &p : particles) ltsCorrect(p);
    // E9 (:1297-1331)
    for (size_t i = 0; i < particles.size(); ++i) {
        Particle &p = particles[i];
        p.position = p.positionOld;
        p.cell = p.cellOld; p.tet = p.tetOld;
        if (p.reflected) p.Utracking = p.UParticleOld;
        else p.Utracking = 0.5 * (p.UParticleOld + p.UParticle);
        const double C = prm.DNum * std::sqrt(mesh.hNum[p.cell] * deltaT * mag(p.Utracking)) / deltaT;
        stepNormals(p, normals[i].data(), 1);
        p.Utracking += Vec3(C * normals[i][3], C * normals[i][4], C * normals[i][5]);
        p.Utracking += p.Ucorrection;
        constrainParticle(deltaT, p);
    }
    // E10 (:1336-1341)
    for (Particle &p : particles) move(p, deltaT);
    particles.erase(std::remove_if(particles.begin(), particles.end(), [](const Particle &p) { return !p.keep; }), particles.end());
    // E11 (:1344-1352)
    boundaryCorrectAfter();
    std::array<double, 1 + PDFB200_MAX_PHI> massInInst{}, massOutInst{};
    for (size_t pI = 0; pI < mesh.patches.size(); ++pI)
        for (int k = 0; k <= prm.nConserved; ++k) { massInInst[k] += patchMassIn[pI][k]; massOutInst[k] += patchMassOut[pI][k]; }
    // pair mixing of the modifiedCurl model (new; IEM already acted at the mid-point)
    curlMixing();
    // E12 first half (:929-956)
    accumulateInstantMoments(block);

    // maxima of E15 (:1434-1450) are taken here, before particle-number
    // control (documented deviation: the reference takes them after it)
    double UMax = 0, UcorrMax = 0, etaMax = 0, etaMin = VGREAT, CoMax = 0;
    for (const Particle &p : particles) {
        UMax = std::max(UMax, mag(p.UParticle));
        UcorrMax = std::max(UcorrMax, mag(p.Ucorrection));
        etaMax = std::max(etaMax, p.eta); etaMin = std::min(etaMin, p.eta);
        CoMax = std::max(CoMax, p.Co);
    }
    if (rep) {
        std::memset(rep, 0, sizeof(*rep));
        rep->deltaT = deltaT;
        rep->UMax = UMax; rep->UcorrMax = UcorrMax; rep->etaMax = etaMax; rep->etaMin = etaMin; rep->CoMax = CoMax;
        for (int k = 0; k <= prm.nConserved; ++k) { rep->massInInst[k] = massInInst[k]; rep->massOutInst[k] = massOutInst[k]; }
    }
}

void Cloud::evolveEnd(pdfb200_step_report *rep, const std::vector<double> &block) {
    // E12 second half (:958-1008)
    const double existWt = (prm.averagingCoeff - 1.) / prm.averagingCoeff;
    finaliseMoments(block, existWt);
    // E13 (:1359)
    particleNumberControl();
    // E14 (:1363-1369)
    // lostMass_/max(PaNIC,1): with one rank PaNIC is the number of particles now in the cell; particle-sharded
    // ranks each re-spread what they lost over the particles they hold (same number when nRanks == 1)
    {
        std::vector<double> nLocal(mesh.nCells, 0.0);
        for (const Particle &p : particles) nLocal[p.cell] += 1.0;
        for (int c = 0; c < mesh.nCells; ++c) lostMass[c] = lostMass[c] / std::max(nRanks > 1 ? nLocal[c] : PaNIC[c], 1.);
    }
    for (Particle &p : particles) p.m += lostMass[p.cell];
    std::fill(lostMass.begin(), lostMass.end(), 0.0);
    nLostTotal += nLost;
    // E15 (:1376-1512)
    double rhoRes = 0, eMin = VGREAT, eMax = -VGREAT, totalMass = 0, totalParticleMass = 0;
    for (int c = 0; c < mesh.nCells; ++c) {
        const double e = (pndcPdf.c[c] - rhocPdf.c[c]) / rhocPdf.c[c];
        rhoRes += std::fabs(e); eMin = std::min(eMin, e); eMax = std::max(eMax, e);
        totalMass += rhocPdfInst.c[c] * mesh.cellVolumes[c];
        totalParticleMass += pndcPdfInst.c[c] * mesh.cellVolumes[c];
    }
    rhoRes /= mesh.nCells;
    for (const Particle &p : particles) {
        const double mpd = massPerDepth(p);
        deltaMassInst[0] += mpd;
        for (int cs = 0; cs < prm.nConserved; ++cs) deltaMassInst[cs + 1] += mpd * p.Phi[prm.conserved[cs]];
    }
    const double CoMax = rep ? rep->CoMax : 0;
    const double dtUsed = deltaT;
    if (rep) {
        rep->rhoRes = rhoRes; rep->rhoErrMin = eMin; rep->rhoErrMax = eMax;
        rep->totalMass = totalMass; rep->totalParticleMass = totalParticleMass;
        for (int k = 0; k <= prm.nConserved; ++k) rep->deltaMassInst[k] = deltaMassInst[k];
        rep->nParticles = (int64_t)particles.size();
        rep->nInjected = nInjected; rep->nDeleted = nDeleted; rep->nCloned = nCloned;
        rep->nEliminated = nEliminated; rep->nLost = nLost;
    }
    // next time step (:1512)
    if (CoMax > 0) deltaT = prm.CFL / CoMax;
    if (rep) { rep->deltaT = dtUsed; rep->deltaTNext = deltaT; }
    ++step;
}

void Cloud::evolve(pdfb200_step_report *rep) {
    pdfb200_step_report local;
    if (!rep) rep = &local;
    std::vector<double> block;
    evolveBegin(rep, block);
    evolveEnd(rep, block);
}

}  // namespace pdforacle
