// mesh_build.cu — geometry, tetFacePointCellDecomposition and interpolation
// tables, computed on the device once per cloud (compiled with -fmad=false so
// the arithmetic is the plain IEEE sequence the CPU oracle also performs).
//
// Reference: mcParticleCloud ctor (mcParticle/mcParticleCloud/mcParticleCloud.C:301-735),
// tetFacePointCellDecomposition.C:31-120, …I.H:30-63, richTetrahedronI.H:107-134,
// initHNum (:1937-1971), CourantCoeffs (:526-535,649), axisymmetry (:661-717).
#include <algorithm>
#include <cmath>

#include "cloud.hpp"

// ---- primitiveMesh::makeFaceCentresAndAreas --------------------------------
__global__ void k_face_geom(int nFaces, const int *faceStart, const int *facePoints, const double *points,
                            double *faceCentres, double *faceAreas, double *magSf) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nFaces) return;
    const int s = faceStart[f], n = faceStart[f + 1] - s;
    V3 ctr, area;
    if (n == 3) {
        const V3 p0 = ld3(points, facePoints[s]), p1 = ld3(points, facePoints[s + 1]), p2 = ld3(points, facePoints[s + 2]);
        ctr = (1.0 / 3.0) * ((p0 + p1) + p2);
        area = 0.5 * cross3(p1 - p0, p2 - p0);
    } else {
        V3 sumN = mk(0, 0, 0), sumAc = mk(0, 0, 0);
        double sumA = 0;
        V3 fC = ld3(points, facePoints[s]);
        for (int i = 1; i < n; ++i) fC = fC + ld3(points, facePoints[s + i]);
        fC = fC / (double)n;
        for (int i = 0; i < n; ++i) {
            const V3 pt = ld3(points, facePoints[s + i]);
            const V3 nx = ld3(points, facePoints[s + (i + 1) % n]);
            const V3 c = (pt + nx) + fC;
            const V3 nn = cross3(nx - pt, fC - pt);
            const double a = mag3(nn);
            sumN = sumN + nn; sumA += a; sumAc = sumAc + a * c;
        }
        ctr = ((1.0 / 3.0) * sumAc) / (sumA + PDF_VSMALL);
        area = 0.5 * sumN;
    }
    st3(faceCentres, f, ctr);
    st3(faceAreas, f, area);
    magSf[f] = mag3(area);
}

// ---- primitiveMesh::makeCellCentresAndVols + bounding box + hNum -----------
__global__ void k_cell_geom(int nCells, const int *cellFaceStart, const int *cellFaces, const int *owner,
                            const int *faceStart, const int *facePoints, const double *points,
                            const double *faceCentres, const double *faceAreas, double4 *cellCentre4,
                            double *cellVolumes, double *bbMin, double *bbMax, int g0, int g1, int g2) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    const int s = cellFaceStart[c], e = cellFaceStart[c + 1];
    V3 sum = mk(0, 0, 0);
    for (int j = s; j < e; ++j) sum = sum + ld3(faceCentres, cellFaces[j]);
    const V3 cEst = sum / (double)(e - s);
    V3 ctr = mk(0, 0, 0);
    double vol = 0;
    V3 lo = mk(PDF_GREAT, PDF_GREAT, PDF_GREAT), hi = mk(-PDF_GREAT, -PDF_GREAT, -PDF_GREAT);
    for (int j = s; j < e; ++j) {
        const int f = cellFaces[j];
        const V3 fc = ld3(faceCentres, f), fa = ld3(faceAreas, f);
        double pyr3Vol;
        if (owner[f] == c) pyr3Vol = fmax(dot3(fa, fc - cEst), PDF_VSMALL);
        else pyr3Vol = fmax(dot3(fa, cEst - fc), PDF_VSMALL);
        const V3 pc = (3.0 / 4.0) * fc + (1.0 / 4.0) * cEst;
        ctr = ctr + pyr3Vol * pc;
        vol += pyr3Vol;
        for (int i = faceStart[f]; i < faceStart[f + 1]; ++i) {
            const V3 p = ld3(points, facePoints[i]);
            lo.x = fmin(lo.x, p.x); lo.y = fmin(lo.y, p.y); lo.z = fmin(lo.z, p.z);
            hi.x = fmax(hi.x, p.x); hi.y = fmax(hi.y, p.y); hi.z = fmax(hi.z, p.z);
        }
    }
    ctr = ctr / vol;
    cellVolumes[c] = vol * (1.0 / 3.0);
    st3(bbMin, c, lo); st3(bbMax, c, hi);
    // initHNum, including the dx.y()/dx.z() slip of the reference
    V3 dx = hi - lo;
    double h;
    const int nGeo = (g0 > 0) + (g1 > 0) + (g2 > 0);
    if (nGeo < 3) {
        dx.x = g0 < 0 ? 1. : dx.x;
        dx.y = g1 < 0 ? 1. : dx.y;
        dx.y = g2 < 0 ? 1. : dx.z;
        h = sqrt(dx.x * dx.y * dx.z);
    } else {
        h = cbrt(dx.x * dx.y * dx.z);
    }
    cellCentre4[c] = make_double4(ctr.x, ctr.y, ctr.z, h);
}

// ---- tetFacePointCellDecomposition + richTetrahedron + adjacency -----------
__global__ void k_build_tets(int nCells, int nInternalFaces, const int *cellFaceStart, const int *cellFaces,
                             const int *cellTetStart, const int *faceTetStartOwn, const int *faceTetStartNei,
                             const int *owner, const int *faceStart, const int *facePoints, const double *points,
                             const int *neighbour, const double *faceCentres, const double4 *cellCentre4, TetA *tetA, D4 *tetB, D4 *tetC,
                             D4 *tetD, int *tetCell) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    const double4 cc = cellCentre4[c];
    const V3 d = mk(cc.x, cc.y, cc.z);
    int t = cellTetStart[c];
    const int tBegin = t, tEnd = cellTetStart[c + 1];
    for (int j = cellFaceStart[c]; j < cellFaceStart[c + 1]; ++j) {
        const int f = cellFaces[j];
        const int s = faceStart[f], n = faceStart[f + 1] - s;
        const bool own = owner[f] == c;
        const int t0 = t;
        const V3 a = ld3(faceCentres, f);
        for (int i = 0; i < n; ++i, ++t) {
            const int rc = (i == 0 ? n - 1 : i - 1);
            int ptBI = facePoints[s + i], ptCI = facePoints[s + rc];
            if (!own) { const int tmp = ptBI; ptBI = ptCI; ptCI = tmp; }
            const V3 b = ld3(points, ptBI), cpt = ld3(points, ptCI);
            const V3 Sa = 0.5 * cross3(cpt - b, d - b);
            const V3 Sb = 0.5 * cross3(d - a, cpt - a);
            const V3 Sc = 0.5 * cross3(b - a, d - a);
            const V3 Sd = 0.5 * cross3(cpt - a, b - a);
            const double vol = (1.0 / 6.0) * dot3(cross3(b - a, cpt - a), d - a);
            double g[9];
            D4 d4; d4.x = d4.y = d4.z = 0;
            if (vol > 0) {
                const double den = -3. * vol;
                const V3 ga = Sa / den, gb = Sb / den, gc = Sc / den, gd = Sd / den;
                d4.x = gd.x; d4.y = gd.y; d4.z = gd.z;
                g[0] = ga.x; g[1] = ga.y; g[2] = ga.z;
                g[3] = gb.x; g[4] = gb.y; g[5] = gb.z;
                g[6] = gc.x; g[7] = gc.y; g[8] = gc.z;
            } else {
                for (int k = 0; k < 9; ++k) g[k] = 0;
            }
            const int prev = t0 + (i == 0 ? n - 1 : i - 1), next = t0 + (i == n - 1 ? 0 : i + 1);
            TetA A;
            A.nbr[0] = -1;
            A.nbr[1] = own ? prev : next;
            A.nbr[2] = own ? next : prev;
            if (f < nInternalFaces) { A.nbr[3] = (own ? faceTetStartNei[f] : faceTetStartOwn[f]) + i; A.nbrCell = own ? neighbour[f] : owner[f]; }
            else { A.nbr[3] = -1; A.nbrCell = -1; }
            A.face = f; A.g0 = g[0];
            tetA[t] = A;
            D4 b4, c4;
            b4.x = g[1]; b4.y = g[2]; b4.z = g[3]; b4.w = g[4];
            c4.x = g[5]; c4.y = g[6]; c4.z = g[7]; c4.w = g[8];
            tetB[t] = b4; tetC[t] = c4;
            d4.w = __hiloint2double(ptCI, ptBI);
            tetD[t] = d4;
            tetCell[t] = c;
        }
    }
    // the other face of this cell sharing mesh edge (b,c)
    for (int t1 = tBegin; t1 < tEnd; ++t1) {
        const int2 p1 = tetPtsOf(tetD[t1]);
        const int lo1 = min(p1.x, p1.y), hi1 = max(p1.x, p1.y);
        for (int t2 = tBegin; t2 < tEnd; ++t2) {
            if (t2 == t1) continue;
            const int2 p2 = tetPtsOf(tetD[t2]);
            if (min(p2.x, p2.y) == lo1 && max(p2.x, p2.y) == hi1) { tetA[t1].nbr[0] = t2; break; }
        }
    }
}

// ---- volPointInterpolation weights -----------------------------------------
__global__ void k_point_weights(int nPoints, const int *pointCellStart, const int *pointCells, const double *points,
                                const double4 *cellCentre4, double *pointWeights) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nPoints) return;
    const V3 x = ld3(points, p);
    double sumW = 0;
    for (int k = pointCellStart[p]; k < pointCellStart[p + 1]; ++k) {
        const double4 cc = cellCentre4[pointCells[k]];
        const double w = 1.0 / mag3(x - mk(cc.x, cc.y, cc.z));
        pointWeights[k] = w;
        sumW += w;
    }
    for (int k = pointCellStart[p]; k < pointCellStart[p + 1]; ++k) pointWeights[k] /= sumW;
}

// ---- linear weights, CourantCoeffs, boundary normals -----------------------
__global__ void k_face_tables(int nFaces, int nInternalFaces, const int *owner, const int *neighbour,
                              const double *faceCentres, const double *faceAreas, const double *magSf,
                              const double4 *cellCentre4, const int *bfInfo, double *linWeights, double *courantCoeffs,
                              double4 *bfNormal) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nFaces) return;
    const V3 Sf = ld3(faceAreas, f), Cf = ld3(faceCentres, f);
    const V3 nHat = Sf / magSf[f];
    const double4 co = cellCentre4[owner[f]];
    const V3 CP = mk(co.x, co.y, co.z);
    if (f < nInternalFaces) {
        const double4 cn = cellCentre4[neighbour[f]];
        const V3 CN = mk(cn.x, cn.y, cn.z);
        const double SfdOwn = fabs(dot3(Sf, Cf - CP)), SfdNei = fabs(dot3(Sf, CN - Cf));
        linWeights[f] = SfdNei / (SfdOwn + SfdNei);
        const double dc = 1.0 / mag3(CN - CP);
        st3(courantCoeffs, f, dc * nHat);
    } else {
        const int bf = f - nInternalFaces;
        bfNormal[bf] = make_double4(nHat.x, nHat.y, nHat.z, 0.0);
        if (BF_GEOM(bfInfo[bf]) == PDFB200_PATCH_EMPTY) {
            st3(courantCoeffs, f, mk(0, 0, 0));
        } else {
            const double dc = 1.0 / dot3(nHat, Cf - CP);
            st3(courantCoeffs, f, (dc * nHat) / 2.);
        }
    }
}

__global__ void k_cell_co(int nSlots, const int *cellFaces, const double *courantCoeffs, D4 *cellCo4) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nSlots) return;
    const V3 co = ld3(courantCoeffs, cellFaces[j]);
    D4 v;
    v.x = co.x; v.y = co.y; v.z = co.z; v.w = 0.0;
    cellCo4[j] = v;
}

static int ilog2ceil(long long v) {
    int b = 0;
    while ((1LL << b) < v) ++b;
    return b;
}

// Sort key: cell only (one radix pass fewer; the stable sort keeps the previous order inside a
// cell) unless this rank holds so many particles per cell that whole warps share a tet, where
// ordering by (cell, local tet) makes the lanes of a warp read the same tet records.
void Cloud::chooseSortKey() {
    const int localPpc = prm.particlesPerCell / std::max(nRanks, 1);
    tetBits = (localPpc >= 256) ? ilog2ceil(maxTetsPerCell) : 0;
    keyBits = ilog2ceil(nCells) + tetBits;
    if (keyBits > 31) { tetBits = 0; keyBits = ilog2ceil(nCells); }
    dm.tetBits = tetBits;
}

void Cloud::buildMesh(const pdfb200_mesh &m) {
    nCells = m.nCells; nFaces = m.nFaces; nInternalFaces = m.nInternalFaces; nPoints = m.nPoints;
    nBnd = nFaces - nInternalFaces;
    if (nCells <= 0 || nFaces <= 0 || nPoints <= 0) throw std::runtime_error("mesh: empty");
    patches.resize(m.nPatches);
    for (int i = 0; i < m.nPatches; ++i)
        patches[i] = {m.patchStart[i], m.patchSize[i], m.patchGeom[i], m.patchHandler[i], m.patchReflecting ? m.patchReflecting[i] : 1};
    hOwner.assign(m.owner, m.owner + nFaces);
    hFaceStart.assign(m.faceStart, m.faceStart + nFaces + 1);
    hFacePoints.assign(m.facePoints, m.facePoints + hFaceStart[nFaces]);
    const int *own = m.owner, *nei = m.neighbour;

    // boundary face -> patch info
    std::vector<int> hBfInfo(nBnd, -1);
    hasOpen = hasInlet = false;
    for (int pI = 0; pI < m.nPatches; ++pI) {
        const HostPatch &pp = patches[pI];
        if (pp.start < nInternalFaces || pp.start + pp.size > nFaces) throw std::runtime_error("mesh: patch range outside the boundary faces");
        for (int j = 0; j < pp.size; ++j)
            hBfInfo[pp.start + j - nInternalFaces] = (pp.handler & 0xF) | ((pp.reflecting & 0xF) << 4) | ((pp.geom & 0xFF) << 8) | (pI << 16);
        if (pp.handler == PDFB200_BC_OPEN || pp.handler == PDFB200_BC_INLET_OUTLET) hasOpen = true;
        if (pp.handler == PDFB200_BC_INLET_OUTLET) hasInlet = true;
    }
    for (int v : hBfInfo) if (v < 0) throw std::runtime_error("mesh: boundary face not covered by a patch");

    // primitiveMesh::calcCells: owned faces first, then neighbour faces
    std::vector<int> hCFS(nCells + 1, 0), hCF;
    {
        std::vector<int> ncf(nCells, 0);
        for (int f = 0; f < nFaces; ++f) ncf[own[f]]++;
        for (int f = 0; f < nInternalFaces; ++f) ncf[nei[f]]++;
        for (int c = 0; c < nCells; ++c) hCFS[c + 1] = hCFS[c] + ncf[c];
        hCF.assign(hCFS[nCells], -1);
        std::fill(ncf.begin(), ncf.end(), 0);
        for (int f = 0; f < nFaces; ++f) { const int c = own[f]; hCF[hCFS[c] + ncf[c]++] = f; }
        for (int f = 0; f < nInternalFaces; ++f) { const int c = nei[f]; hCF[hCFS[c] + ncf[c]++] = f; }
    }
    // tet numbering
    std::vector<int> hCTS(nCells + 1, 0), hFTO(nFaces, -1), hFTN(nFaces, -1);
    long long nT = 0;
    maxTetsPerCell = 0;
    for (int c = 0; c < nCells; ++c) {
        hCTS[c] = (int)nT;
        for (int j = hCFS[c]; j < hCFS[c + 1]; ++j) {
            const int f = hCF[j];
            if (own[f] == c) hFTO[f] = (int)nT; else hFTN[f] = (int)nT;
            nT += hFaceStart[f + 1] - hFaceStart[f];
        }
        maxTetsPerCell = std::max(maxTetsPerCell, (int)(nT - hCTS[c]));
        if (nT > 2000000000LL) throw std::runtime_error("mesh: more than 2e9 tets");
    }
    hCTS[nCells] = (int)nT;
    nTets = nT;
    chooseSortKey();
    if (keyBits > 31) throw std::runtime_error("mesh: sort key does not fit 31 bits");

    // point -> cells (ascending cells)
    std::vector<int> hPCS(nPoints + 1, 0), hPC;
    {
        std::vector<int> last(nPoints, -1), cnt(nPoints, 0);
        for (int c = 0; c < nCells; ++c)
            for (int j = hCFS[c]; j < hCFS[c + 1]; ++j) {
                const int f = hCF[j];
                for (int i = hFaceStart[f]; i < hFaceStart[f + 1]; ++i) {
                    const int p = hFacePoints[i];
                    if (last[p] != c) { last[p] = c; cnt[p]++; }
                }
            }
        for (int p = 0; p < nPoints; ++p) hPCS[p + 1] = hPCS[p] + cnt[p];
        hPC.assign(hPCS[nPoints], -1);
        std::fill(last.begin(), last.end(), -1); std::fill(cnt.begin(), cnt.end(), 0);
        for (int c = 0; c < nCells; ++c)
            for (int j = hCFS[c]; j < hCFS[c + 1]; ++j) {
                const int f = hCF[j];
                for (int i = hFaceStart[f]; i < hFaceStart[f + 1]; ++i) {
                    const int p = hFacePoints[i];
                    if (last[p] != c) { last[p] = c; hPC[hPCS[p] + cnt[p]++] = c; }
                }
            }
    }
    // open-boundary lookup (mcOpenBoundary ctor, mcOpenBoundary.C:56-69)
    std::vector<int> hOCS(nCells + 1, 0), hOCF;
    {
        std::vector<int> cnt(nCells, 0);
        for (int bf = 0; bf < nBnd; ++bf) {
            const int h = BF_HANDLER(hBfInfo[bf]);
            if (h == PDFB200_BC_OPEN || h == PDFB200_BC_INLET_OUTLET) cnt[own[nInternalFaces + bf]]++;
        }
        for (int c = 0; c < nCells; ++c) hOCS[c + 1] = hOCS[c] + cnt[c];
        hOCF.assign(std::max(hOCS[nCells], 1), -1);
        std::fill(cnt.begin(), cnt.end(), 0);
        for (int bf = 0; bf < nBnd; ++bf) {
            const int h = BF_HANDLER(hBfInfo[bf]);
            if (h == PDFB200_BC_OPEN || h == PDFB200_BC_INLET_OUTLET) { const int c = own[nInternalFaces + bf]; hOCF[hOCS[c] + cnt[c]++] = bf; }
        }
    }

    // ---- upload + device geometry ----
    cudaStream_t s = stream;
    points.upload(m.points, (size_t)nPoints * 3, s);
    faceStart.upload(m.faceStart, nFaces + 1, s); facePoints.upload(m.facePoints, hFaceStart[nFaces], s);
    owner.upload(own, nFaces, s); neighbour.upload(nei, std::max(nInternalFaces, 1), s);
    cellFaceStart.upload(hCFS.data(), hCFS.size(), s); cellFaces.upload(hCF.data(), hCF.size(), s);
    cellTetStart.upload(hCTS.data(), hCTS.size(), s);
    faceTetStartOwn.upload(hFTO.data(), nFaces, s); faceTetStartNei.upload(hFTN.data(), nFaces, s);
    pointCellStart.upload(hPCS.data(), hPCS.size(), s); pointCells.upload(hPC.data(), hPC.size(), s);
    bfInfo.upload(hBfInfo.data(), nBnd, s);
    openCellStart.upload(hOCS.data(), hOCS.size(), s); openCellFaces.upload(hOCF.data(), hOCF.size(), s);
    faceCentres.alloc((size_t)nFaces * 3); faceAreas.alloc((size_t)nFaces * 3); magSf.alloc(nFaces);
    cellCentre4.alloc(nCells); cellVolumes.alloc(nCells); volumeOrArea.alloc(nCells);
    bbMin.alloc((size_t)nCells * 3); bbMax.alloc((size_t)nCells * 3);
    courantCoeffs.alloc((size_t)nFaces * 3); cellCo4.alloc(hCF.size());
    linWeights.alloc(std::max(nInternalFaces, 1)); pointWeights.alloc(hPC.size());
    bfNormal.alloc(nBnd);
    tetA.alloc((size_t)nTets); tetB.alloc((size_t)nTets); tetC.alloc((size_t)nTets); tetD.alloc((size_t)nTets); tetCell.alloc((size_t)nTets);

    const int B = 128;
    LAUNCH(this, k_face_geom, gridFor(nFaces, B), B, 0, nFaces, faceStart.p, facePoints.p, points.p, faceCentres.p, faceAreas.p, magSf.p);
    LAUNCH(this, k_cell_geom, gridFor(nCells, B), B, 0, nCells, cellFaceStart.p, cellFaces.p, owner.p, faceStart.p, facePoints.p,
           points.p, faceCentres.p, faceAreas.p, cellCentre4.p, cellVolumes.p, bbMin.p, bbMax.p, m.geometricD[0], m.geometricD[1], m.geometricD[2]);
    LAUNCH(this, k_build_tets, gridFor(nCells, 64), 64, 0, nCells, nInternalFaces, cellFaceStart.p, cellFaces.p, cellTetStart.p,
           faceTetStartOwn.p, faceTetStartNei.p, owner.p, faceStart.p, facePoints.p, points.p, neighbour.p, faceCentres.p, cellCentre4.p,
           tetA.p, tetB.p, tetC.p, tetD.p, tetCell.p);
    LAUNCH(this, k_point_weights, gridFor(nPoints, B), B, 0, nPoints, pointCellStart.p, pointCells.p, points.p, cellCentre4.p, pointWeights.p);
    LAUNCH(this, k_face_tables, gridFor(nFaces, B), B, 0, nFaces, nInternalFaces, owner.p, neighbour.p, faceCentres.p, faceAreas.p,
           magSf.p, cellCentre4.p, bfInfo.p, linWeights.p, courantCoeffs.p, bfNormal.p);
    LAUNCH(this, k_cell_co, gridFor((long long)hCF.size(), B), B, 0, (int)hCF.size(), cellFaces.p, courantCoeffs.p, cellCo4.p);
    CUDA_CHECK(cudaMemcpyAsync(volumeOrArea.p, cellVolumes.p, nCells * sizeof(double), cudaMemcpyDeviceToDevice, s));

    // ---- processing order of the cells ----
    // Particles are sorted by the rank of their cell along a Morton curve through the cell
    // centres rather than by the caller's cell label: cells that are neighbours in space are
    // then processed close together in time, so the tet/node records and the state sectors a
    // crossing particle touches are still in L2. PDFB200_CELL_ORDER=native keeps label order.
    {
        std::vector<double4> hC(nCells);
        CUDA_CHECK(cudaMemcpyAsync(hC.data(), cellCentre4.p, (size_t)nCells * sizeof(double4), cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        std::vector<int> hRank(nCells), hOfRank(nCells);
        const char *ord = std::getenv("PDFB200_CELL_ORDER");
        if (ord && std::string(ord) == "native") {
            for (int c = 0; c < nCells; ++c) hRank[c] = hOfRank[c] = c;
        } else {
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
            for (int c = 0; c < nCells; ++c) {
                const double x[3] = {hC[c].x, hC[c].y, hC[c].z};
                for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], x[k]); hi[k] = std::max(hi[k], x[k]); }
            }
            const double ext = std::max(std::max(hi[0] - lo[0], hi[1] - lo[1]), std::max(hi[2] - lo[2], 1e-300));
            auto spread = [](unsigned long long v) {   // 21 bits -> every third bit
                v &= 0x1fffffULL;
                v = (v | v << 32) & 0x1f00000000ffffULL;
                v = (v | v << 16) & 0x1f0000ff0000ffULL;
                v = (v | v << 8) & 0x100f00f00f00f00fULL;
                v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
                v = (v | v << 2) & 0x1249249249249249ULL;
                return v;
            };
            std::vector<std::pair<unsigned long long, int>> code(nCells);
            for (int c = 0; c < nCells; ++c) {
                const double x[3] = {hC[c].x, hC[c].y, hC[c].z};
                unsigned long long q[3];
                for (int k = 0; k < 3; ++k) q[k] = (unsigned long long)std::min(2097151.0, std::max(0.0, (x[k] - lo[k]) / ext * 2097151.0));
                code[c] = {spread(q[0]) | (spread(q[1]) << 1) | (spread(q[2]) << 2), c};
            }
            std::sort(code.begin(), code.end());
            for (int r = 0; r < nCells; ++r) { hOfRank[r] = code[r].second; hRank[code[r].second] = r; }
        }
        cellRank.upload(hRank.data(), nCells, s); cellOfRank.upload(hOfRank.data(), nCells, s);
        CUDA_CHECK(cudaStreamSynchronize(s));
    }

    // ---- axisymmetry (mcParticleCloud.C:661-717, wedgePolyPatch::initTransforms) ----
    const int nGeoD = (m.geometricD[0] > 0) + (m.geometricD[1] > 0) + (m.geometricD[2] > 0);
    axisym = false;
    std::vector<double> hFaceT((size_t)std::max(m.nPatches, 1) * 9, 0.0);
    for (int pI = 0; pI < m.nPatches; ++pI) { hFaceT[9 * pI] = hFaceT[9 * pI + 4] = hFaceT[9 * pI + 8] = 1.0; }
    int nWedge = 0;
    for (int pI = 0; pI < m.nPatches; ++pI) {
        const HostPatch &pp = patches[pI];
        if (pp.geom != PDFB200_PATCH_WEDGE || pp.size == 0) continue;
        double a[3];
        CUDA_CHECK(cudaMemcpyAsync(a, faceAreas.p + (size_t)pp.start * 3, 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        V3 n = mk(a[0], a[1], a[2]);
        n = n / mag3(n);
        const double nn[3] = {n.x, n.y, n.z};
        double cnv[3];
        for (int k = 0; k < 3; ++k) {
            const double sg = nn[k] > 0 ? 1.0 : (nn[k] < 0 ? -1.0 : 0.0);
            cnv[k] = sg * (std::max(std::fabs(nn[k]), 0.5) - 0.5);
        }
        V3 cn = mk(cnv[0], cnv[1], cnv[2]);
        cn = cn / mag3(cn);
        rotationTensor(cn, n, &hFaceT[9 * pI]);
        if (nGeoD <= 2) {
            if (!axisym) {
                axisym = true;
                V3 ax = cross3(cn, n);
                ax = ax / mag3(ax);
                axis[0] = ax.x; axis[1] = ax.y; axis[2] = ax.z;
                centreNormal[0] = cn.x; centreNormal[1] = cn.y; centreNormal[2] = cn.z;
                openingAngle = 2. * std::acos(dot3(n, cn));
                // area_[faceCells] = mag(faceAreas) of the first wedge patch
                std::vector<double> hMag(pp.size), hVoA(nCells, 0.0);
                CUDA_CHECK(cudaMemcpyAsync(hMag.data(), magSf.p + pp.start, pp.size * sizeof(double), cudaMemcpyDeviceToHost, s));
                CUDA_CHECK(cudaStreamSynchronize(s));
                for (int j = 0; j < pp.size; ++j) hVoA[own[pp.start + j]] = hMag[j];
                volumeOrArea.upload(hVoA.data(), nCells, s);
                CUDA_CHECK(cudaStreamSynchronize(s));
            }
            ++nWedge;
        }
    }
    if (axisym && nWedge != 2) throw std::runtime_error("Only one pair of wedge patches allowed");
    patchFaceT.upload(hFaceT.data(), hFaceT.size(), s);

    // ---- fill the device mesh view ----
    dm.nCells = nCells; dm.nFaces = nFaces; dm.nInternalFaces = nInternalFaces; dm.nPoints = nPoints; dm.nBnd = nBnd;
    dm.nPatches = m.nPatches; dm.nTets = nTets; dm.tetBits = tetBits;
    {
        int nF = hCFS[1] - hCFS[0];
        for (int c = 1; c < nCells && nF > 0; ++c) if (hCFS[c + 1] - hCFS[c] != nF) nF = 0;
        dm.cellFacesUniform = nF;
    }
    for (int k = 0; k < 3; ++k) { dm.geoD[k] = m.geometricD[k]; dm.solD[k] = m.solutionD[k]; dm.axis[k] = axis[k]; dm.centreNormal[k] = centreNormal[k]; }
    dm.nGeoD = nGeoD; dm.axisym = axisym; dm.openingAngle = openingAngle;
    dm.points = points.p; dm.faceStart = faceStart.p; dm.facePoints = facePoints.p; dm.owner = owner.p; dm.neighbour = neighbour.p;
    dm.cellFaceStart = cellFaceStart.p; dm.cellFaces = cellFaces.p; dm.cellTetStart = cellTetStart.p;
    dm.cellRank = cellRank.p; dm.cellOfRank = cellOfRank.p;
    dm.faceTetStartOwn = faceTetStartOwn.p; dm.faceTetStartNei = faceTetStartNei.p;
    dm.faceCentres = faceCentres.p; dm.faceAreas = faceAreas.p; dm.magSf = magSf.p; dm.cellCentre4 = cellCentre4.p;
    dm.cellVolumes = cellVolumes.p; dm.volumeOrArea = volumeOrArea.p; dm.bbMin = bbMin.p; dm.bbMax = bbMax.p;
    dm.courantCoeffs = courantCoeffs.p; dm.cellCo4 = cellCo4.p; dm.linWeights = linWeights.p;
    dm.pointCellStart = pointCellStart.p; dm.pointCells = pointCells.p; dm.pointWeights = pointWeights.p;
    dm.tetA = tetA.p; dm.tetB = tetB.p; dm.tetC = tetC.p; dm.tetD = tetD.p; dm.tetCell = tetCell.p; dm.bfInfo = bfInfo.p; dm.bfNormal = bfNormal.p;
    dm.openCellStart = openCellStart.p; dm.openCellFaces = openCellFaces.p; dm.patchFaceT = patchFaceT.p;
    CUDA_CHECK(cudaStreamSynchronize(s));
}
