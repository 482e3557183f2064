// mesh.cpp — CPU ORACLE (test infrastructure): polyMesh geometry, the
// tetFacePointCellDecomposition and the interpolation tables.
#include <algorithm>
#include <stdexcept>
#include <tuple>

#include "pdforacle.hpp"

namespace pdforacle {

// OpenFOAM transform.H rotationTensor(n1, n2); SURVEY Appendix B.4
Tensor rotationTensor(const Vec3 &n1, const Vec3 &n2) {
    const double s = dot(n1, n2);
    const Vec3 n3 = cross(n1, n2);
    const double magSqrN3 = dot(n3, n3);
    Tensor t;
    const double v[3] = {n3.x, n3.y, n3.z}, a[3] = {n1.x, n1.y, n1.z}, b[3] = {n2.x, n2.y, n2.z};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double val = (i == j ? s : 0.0);
            val += (1 - s) * (v[i] * v[j]) / (magSqrN3 + VSMALL);
            val += b[i] * a[j] - a[i] * b[j];
            t.m[3 * i + j] = val;
        }
    return t;
}

SymmTensor transform(const Tensor &t, const SymmTensor &s) {
    const double S[9] = {s.xx, s.xy, s.xz, s.xy, s.yy, s.yz, s.xz, s.yz, s.zz};
    double TS[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double v = 0;
            for (int k = 0; k < 3; ++k) v += t.m[3 * i + k] * S[3 * k + j];
            TS[3 * i + j] = v;
        }
    double R[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double v = 0;
            for (int k = 0; k < 3; ++k) v += TS[3 * i + k] * t.m[3 * j + k];
            R[3 * i + j] = v;
        }
    SymmTensor r;
    r.xx = R[0]; r.xy = R[1]; r.xz = R[2]; r.yy = R[4]; r.yz = R[5]; r.zz = R[8];
    return r;
}

// richTetrahedron::inside (richTetrahedron/richTetrahedron.C:40-91), verbatim
// plane order and tolerances: (Sa,b) (Sb,c) (Sc,b) (Sd,b).
bool Tet::inside(const Vec3 &pt) const {
    const Vec3 *S[4] = {&Sa, &Sb, &Sc, &Sd};
    const Vec3 *base[4] = {&b, &c, &b, &b};
    for (int i = 0; i < 4; ++i) {
        Vec3 n = *S[i];
        n /= (mag(n) + VSMALL);
        if (dot(pt - *base[i], n) > SMALL) return false;
    }
    return true;
}

void Mesh::build(const pdfb200_mesh &m) {
    nCells = m.nCells; nFaces = m.nFaces; nInternalFaces = m.nInternalFaces;
    points.resize(m.nPoints);
    for (int i = 0; i < m.nPoints; ++i) points[i] = Vec3(m.points[3 * i], m.points[3 * i + 1], m.points[3 * i + 2]);
    faceStart.assign(m.faceStart, m.faceStart + nFaces + 1);
    facePoints.assign(m.facePoints, m.facePoints + faceStart[nFaces]);
    owner.assign(m.owner, m.owner + nFaces);
    neighbour.assign(m.neighbour, m.neighbour + nInternalFaces);
    patches.resize(m.nPatches);
    for (int i = 0; i < m.nPatches; ++i) {
        patches[i].start = m.patchStart[i]; patches[i].size = m.patchSize[i];
        patches[i].geom = m.patchGeom[i]; patches[i].handler = m.patchHandler[i];
        patches[i].reflecting = m.patchReflecting ? m.patchReflecting[i] : 1;
    }
    for (int i = 0; i < 3; ++i) { geometricD[i] = m.geometricD[i]; solutionD[i] = m.solutionD[i]; }
    facePatch.assign(nBoundaryFaces(), -1);
    for (size_t pI = 0; pI < patches.size(); ++pI)
        for (int j = 0; j < patches[pI].size; ++j) facePatch[patches[pI].start + j - nInternalFaces] = (int)pI;
    for (int v : facePatch) if (v < 0) throw std::runtime_error("boundary face not covered by a patch");

    makeFaceGeometry();
    makeCells();
    makeCellGeometry();
    makeTets();
    makeInterpolationWeights();
    makeCourantCoeffs();
    makeHNum();
    detectAxisymmetry();
}

// primitiveMesh::makeFaceCentresAndAreas (OpenFOAM, external)
void Mesh::makeFaceGeometry() {
    faceCentres.resize(nFaces); faceAreas.resize(nFaces); magSf.resize(nFaces);
    for (int f = 0; f < nFaces; ++f) {
        const int n = faceSize(f);
        if (n == 3) {
            const Vec3 &p0 = points[facePoint(f, 0)], &p1 = points[facePoint(f, 1)], &p2 = points[facePoint(f, 2)];
            faceCentres[f] = (1.0 / 3.0) * ((p0 + p1) + p2);
            faceAreas[f] = 0.5 * cross(p1 - p0, p2 - p0);
        } else {
            Vec3 sumN, sumAc; double sumA = 0;
            Vec3 fC = points[facePoint(f, 0)];
            for (int i = 1; i < n; ++i) fC += points[facePoint(f, i)];
            fC /= (double)n;
            for (int i = 0; i < n; ++i) {
                const Vec3 &pt = points[facePoint(f, i)];
                const Vec3 &nx = points[facePoint(f, (i + 1) % n)];
                Vec3 c = (pt + nx) + fC;
                Vec3 nn = cross(nx - pt, fC - pt);
                double a = mag(nn);
                sumN += nn; sumA += a; sumAc += a * c;
            }
            faceCentres[f] = ((1.0 / 3.0) * sumAc) / (sumA + VSMALL);
            faceAreas[f] = 0.5 * sumN;
        }
        magSf[f] = mag(faceAreas[f]);
    }
}

// primitiveMesh::calcCells: owned faces first (ascending), then neighbour faces
void Mesh::makeCells() {
    std::vector<int> ncf(nCells, 0);
    for (int f = 0; f < nFaces; ++f) ncf[owner[f]]++;
    for (int f = 0; f < nInternalFaces; ++f) ncf[neighbour[f]]++;
    cellFaceStart.assign(nCells + 1, 0);
    for (int c = 0; c < nCells; ++c) cellFaceStart[c + 1] = cellFaceStart[c] + ncf[c];
    cellFaces.assign(cellFaceStart[nCells], -1);
    std::fill(ncf.begin(), ncf.end(), 0);
    for (int f = 0; f < nFaces; ++f) { int c = owner[f]; cellFaces[cellFaceStart[c] + ncf[c]++] = f; }
    for (int f = 0; f < nInternalFaces; ++f) { int c = neighbour[f]; cellFaces[cellFaceStart[c] + ncf[c]++] = f; }
}

// primitiveMesh::makeCellCentresAndVols (OpenFOAM, external)
void Mesh::makeCellGeometry() {
    cellCentres.assign(nCells, Vec3()); cellVolumes.assign(nCells, 0.0);
    std::vector<Vec3> cEst(nCells);
    for (int c = 0; c < nCells; ++c) {
        Vec3 s; int nf = 0;
        for (int j = cellFaceStart[c]; j < cellFaceStart[c + 1]; ++j) { s += faceCentres[cellFaces[j]]; ++nf; }
        cEst[c] = s / (double)nf;
    }
    for (int c = 0; c < nCells; ++c) {
        Vec3 ctr; double vol = 0;
        for (int j = cellFaceStart[c]; j < cellFaceStart[c + 1]; ++j) {
            const int f = cellFaces[j];
            double pyr3Vol;
            if (owner[f] == c) pyr3Vol = std::max(dot(faceAreas[f], faceCentres[f] - cEst[c]), VSMALL);
            else pyr3Vol = std::max(dot(faceAreas[f], cEst[c] - faceCentres[f]), VSMALL);
            Vec3 pc = (3.0 / 4.0) * faceCentres[f] + (1.0 / 4.0) * cEst[c];
            ctr += pyr3Vol * pc;
            vol += pyr3Vol;
        }
        cellCentres[c] = ctr / vol;
        cellVolumes[c] = vol * (1.0 / 3.0);
    }
    // bounding boxes over the cell's vertices
    bbMin.assign(nCells, Vec3(GREAT, GREAT, GREAT)); bbMax.assign(nCells, Vec3(-GREAT, -GREAT, -GREAT));
    for (int c = 0; c < nCells; ++c)
        for (int j = cellFaceStart[c]; j < cellFaceStart[c + 1]; ++j) {
            const int f = cellFaces[j];
            for (int i = 0; i < faceSize(f); ++i) {
                const Vec3 &p = points[facePoint(f, i)];
                for (int k = 0; k < 3; ++k) { bbMin[c][k] = std::min(bbMin[c][k], p[k]); bbMax[c][k] = std::max(bbMax[c][k], p[k]); }
            }
        }
}

// tetFacePointCellDecomposition ctor + decompose()
// (tetFacePointCellDecomposition.C:31-120, …I.H:30-63) and richTetrahedron::update
// (richTetrahedronI.H:107-134); adjacency per SURVEY Appendix D.
void Mesh::makeTets() {
    cellTetStart.assign(nCells + 1, 0);
    faceTetStartOwn.assign(nFaces, -1); faceTetStartNei.assign(nFaces, -1);
    int nT = 0;
    maxTetsPerCell = 0;
    for (int c = 0; c < nCells; ++c) {
        cellTetStart[c] = nT;
        for (int j = cellFaceStart[c]; j < cellFaceStart[c + 1]; ++j) {
            const int f = cellFaces[j];
            if (owner[f] == c) faceTetStartOwn[f] = nT; else faceTetStartNei[f] = nT;
            nT += faceSize(f);
        }
        maxTetsPerCell = std::max(maxTetsPerCell, nT - cellTetStart[c]);
    }
    cellTetStart[nCells] = nT;
    tets.resize(nT);
    for (int c = 0; c < nCells; ++c) {
        int t = cellTetStart[c];
        for (int j = cellFaceStart[c]; j < cellFaceStart[c + 1]; ++j) {
            const int f = cellFaces[j];
            const int n = faceSize(f);
            const bool own = (owner[f] == c);
            const int t0 = t;
            for (int i = 0; i < n; ++i, ++t) {
                Tet &T = tets[t];
                const int rc = (i == 0 ? n - 1 : i - 1);   // face::rcIndex
                int ptBI = facePoint(f, i), ptCI = facePoint(f, rc);
                T.ptFirst = i; T.ptSecond = rc;
                if (!own) { std::swap(ptBI, ptCI); std::swap(T.ptFirst, T.ptSecond); }
                T.a = faceCentres[f]; T.b = points[ptBI]; T.c = points[ptCI]; T.d = cellCentres[c];
                T.face = f; T.ptB = ptBI; T.ptC = ptCI; T.cell = c;
                // tetrahedron<Point,PointRef>::{Sa,Sb,Sc,Sd,mag}
                T.Sa = 0.5 * cross(T.c - T.b, T.d - T.b);
                T.Sb = 0.5 * cross(T.d - T.a, T.c - T.a);
                T.Sc = 0.5 * cross(T.b - T.a, T.d - T.a);
                T.Sd = 0.5 * cross(T.c - T.a, T.b - T.a);
                T.vol = (1.0 / 6.0) * dot(cross(T.b - T.a, T.c - T.a), T.d - T.a);
                if (T.vol > 0) {
                    const double den = -3. * T.vol;
                    T.gradN[0] = T.Sa / den; T.gradN[1] = T.Sb / den;
                    T.gradN[2] = T.Sc / den; T.gradN[3] = T.Sd / den;
                } else {
                    for (auto &g : T.gradN) g = Vec3();
                }
                // around the face: the face opposite b contains c, opposite c contains b
                const int prev = t0 + (i == 0 ? n - 1 : i - 1), next = t0 + (i == n - 1 ? 0 : i + 1);
                T.nbr[1] = own ? prev : next;
                T.nbr[2] = own ? next : prev;
                // through the mesh face
                if (f < nInternalFaces) T.nbr[3] = -2;  // fixed below once both starts are known
                else T.nbr[3] = -1;
            }
        }
        // within the cell: the other face sharing mesh edge (b,c)
        std::vector<std::tuple<int, int, int>> edges;
        for (int tt = cellTetStart[c]; tt < cellTetStart[c + 1]; ++tt)
            edges.emplace_back(std::min(tets[tt].ptB, tets[tt].ptC), std::max(tets[tt].ptB, tets[tt].ptC), tt);
        std::sort(edges.begin(), edges.end());
        for (size_t e = 0; e + 1 < edges.size(); ++e) {
            if (std::get<0>(edges[e]) == std::get<0>(edges[e + 1]) && std::get<1>(edges[e]) == std::get<1>(edges[e + 1])) {
                const int t1 = std::get<2>(edges[e]), t2 = std::get<2>(edges[e + 1]);
                tets[t1].nbr[0] = t2; tets[t2].nbr[0] = t1;
                ++e;
            }
        }
    }
    for (int t = 0; t < nT; ++t) {
        Tet &T = tets[t];
        if (T.nbr[3] == -2) {
            const bool own = (owner[T.face] == T.cell);
            const int local = t - (own ? faceTetStartOwn[T.face] : faceTetStartNei[T.face]);
            T.nbr[3] = (own ? faceTetStartNei[T.face] : faceTetStartOwn[T.face]) + local;
        }
        if (T.nbr[0] < 0) throw std::runtime_error("open cell: tet edge without a second face");
    }
}

// tetFacePointCellDecomposition::find (…Decomposition.C:125-151)
int Mesh::findReference(const Vec3 &pt, int cellHint) const {
    if (cellHint > -1)
        for (int t = cellTetStart[cellHint]; t < cellTetStart[cellHint + 1]; ++t)
            if (tets[t].inside(pt)) return t;
    for (size_t t = 0; t < tets.size(); ++t)
        if (tets[t].inside(pt)) return (int)t;
    return -1;
}

// Shared locate rule (device: k_locate): the tet of `cell` with the largest
// minimum barycentric coordinate; first maximum wins.
int Mesh::locate(const Vec3 &pt, int cell) const {
    int best = -1; double bestMin = -VGREAT;
    const Vec3 r = pt - cellCentres[cell];
    for (int t = cellTetStart[cell]; t < cellTetStart[cell + 1]; ++t) {
        const Tet &T = tets[t];
        const double pa = dotf(T.gradN[0], r), pb = dotf(T.gradN[1], r), pc = dotf(T.gradN[2], r);
        const double pd = 1.0 - ((pa + pb) + pc);
        const double mn = std::min(std::min(pa, pb), std::min(pc, pd));
        if (mn > bestMin) { bestMin = mn; best = t; }
    }
    return best;
}

// primitiveMesh::pointInCell: inside iff on the inner side of every face plane
bool Mesh::pointInCell(const Vec3 &pt, int cell) const {
    for (int j = cellFaceStart[cell]; j < cellFaceStart[cell + 1]; ++j) {
        const int f = cellFaces[j];
        Vec3 proj = pt - faceCentres[f];
        Vec3 normal = faceAreas[f];
        normal /= mag(normal);
        if (owner[f] != cell) normal = -normal;
        if (dot(normal, proj) > 0) return false;
    }
    return true;
}

// volPointInterpolation::makeWeights (inverse distance to cell centres) and
// surfaceInterpolation::makeWeights (linear)
void Mesh::makeInterpolationWeights() {
    const int nP = (int)points.size();
    std::vector<std::vector<int>> pc(nP);
    for (int c = 0; c < nCells; ++c)
        for (int j = cellFaceStart[c]; j < cellFaceStart[c + 1]; ++j) {
            const int f = cellFaces[j];
            for (int i = 0; i < faceSize(f); ++i) {
                auto &lst = pc[facePoint(f, i)];
                if (lst.empty() || lst.back() != c) lst.push_back(c);   // cells ascend, so back() suffices
            }
        }
    pointCellStart.assign(nP + 1, 0);
    for (int p = 0; p < nP; ++p) pointCellStart[p + 1] = pointCellStart[p] + (int)pc[p].size();
    pointCells.resize(pointCellStart[nP]); pointWeights.resize(pointCellStart[nP]);
    for (int p = 0; p < nP; ++p) {
        double sumW = 0;
        for (size_t k = 0; k < pc[p].size(); ++k) {
            const int c = pc[p][k];
            pointCells[pointCellStart[p] + k] = c;
            const double w = 1.0 / mag(points[p] - cellCentres[c]);
            pointWeights[pointCellStart[p] + k] = w;
            sumW += w;
        }
        for (size_t k = 0; k < pc[p].size(); ++k) pointWeights[pointCellStart[p] + k] /= sumW;
    }
    linWeights.resize(nInternalFaces);
    for (int f = 0; f < nInternalFaces; ++f) {
        const double SfdOwn = std::fabs(dot(faceAreas[f], faceCentres[f] - cellCentres[owner[f]]));
        const double SfdNei = std::fabs(dot(faceAreas[f], cellCentres[neighbour[f]] - faceCentres[f]));
        linWeights[f] = SfdNei / (SfdOwn + SfdNei);
    }
}

// CourantCoeffs_ = deltaCoeffs*Sf/magSf, boundary values halved
// (mcParticleCloud.C:526-535,649); empty patches carry no values and are
// skipped by every consumer, stored as zero here.
void Mesh::makeCourantCoeffs() {
    courantCoeffs.assign(nFaces, Vec3());
    for (int f = 0; f < nFaces; ++f) {
        const Vec3 nHat = faceAreas[f] / magSf[f];
        if (f < nInternalFaces) {
            const double dc = 1.0 / mag(cellCentres[neighbour[f]] - cellCentres[owner[f]]);
            courantCoeffs[f] = dc * nHat;
        } else {
            if (patches[facePatch[f - nInternalFaces]].geom == PDFB200_PATCH_EMPTY) continue;
            const double dc = 1.0 / dot(nHat, faceCentres[f] - cellCentres[owner[f]]);
            courantCoeffs[f] = (dc * nHat) / 2.;
        }
    }
}

// mcParticleCloud::initHNum (mcParticleCloud.C:1937-1971), including the
// dx.y()/dx.z() slip for reduced-dimension cases.
void Mesh::makeHNum() {
    hNum.resize(nCells);
    for (int c = 0; c < nCells; ++c) {
        Vec3 dx(bbMax[c].x - bbMin[c].x, bbMax[c].y - bbMin[c].y, bbMax[c].z - bbMin[c].z);
        if (nGeometricD() < 3) {
            dx.x = geometricD[0] < 0 ? 1. : dx.x;
            dx.y = geometricD[1] < 0 ? 1. : dx.y;
            dx.y = geometricD[2] < 0 ? 1. : dx.z;
            hNum[c] = std::sqrt(dx.x * dx.y * dx.z);
        } else {
            hNum[c] = std::cbrt(dx.x * dx.y * dx.z);
        }
    }
}

// mcParticleCloud.C:661-717 with wedgePolyPatch::initTransforms (OpenFOAM)
void Mesh::detectAxisymmetry() {
    axisym = false;
    volumeOrArea = cellVolumes;
    if (nGeometricD() > 2) return;
    int nWedge = 0;
    for (const Patch &pp : patches) {
        if (pp.geom != PDFB200_PATCH_WEDGE || pp.size == 0) continue;
        if (!axisym) {
            axisym = true;
            Vec3 n = faceAreas[pp.start];
            n /= mag(n);
            wedgePatchNormal = n;
            Vec3 cn;
            for (int k = 0; k < 3; ++k) {
                const double s = n[k] > 0 ? 1.0 : (n[k] < 0 ? -1.0 : 0.0);
                cn[k] = s * (std::max(std::fabs(n[k]), 0.5) - 0.5);
            }
            cn /= mag(cn);
            centreNormal = cn;
            axis = cross(centreNormal, wedgePatchNormal);
            axis /= mag(axis);
            openingAngle = 2. * std::acos(dot(wedgePatchNormal, centreNormal));
            volumeOrArea.assign(nCells, 0.0);
            for (int j = 0; j < pp.size; ++j) volumeOrArea[owner[pp.start + j]] = magSf[pp.start + j];
        }
        ++nWedge;
    }
    if (axisym && nWedge != 2) throw std::runtime_error("Only one pair of wedge patches allowed");
}

}  // namespace pdforacle
