// ref_shim.cpp — C entry points around four pieces of pdfFoam compiled UNMODIFIED from /root/reference
// against foam_stub/ (see Makefile). TEST INFRASTRUCTURE ONLY: tests/test_oracle_vs_ref.py holds the CPU oracle
// (oracle/src) to these functions; nothing in pdffoam_b200/ or bench.py touches this library.
//
//   pdfref_find_coeffs / pdfref_binterp  -> mcSteadyFlamelet.C:36-95 (anonymous-namespace templates; the Makefile
//                                           cuts that block out of the file into _ref/gen/steadyFlameletHelpers.inc)
//   pdfref_inlet_*                       -> mcInletRandom.C:92-108, mcInletRandomI.H:28-39,
//                                           mcInversionInletRandom.C:63-119 through mcInletRandom::New("inversion"),
//                                           mcSamplingInletRandom.C through New("sampling")
//   pdfref_decomp_*                      -> tetFacePointCellDecomposition{.H,I.H,.C}<richTetPointRef>,
//                                           richTetrahedron{.H,I.H,.C}
#include "foamStub.H"

const Foam::Vector Foam::Vector::zero(0, 0, 0);

// ---- the reference's sources, where they lie ----
#define NoRepository   // OpenFOAM's switch for "include the template definitions" (richTetrahedron.H, tetFacePointCellDecomposition.H)
#include "richTetPointRef.H"
#include "tetFacePointCellDecomposition.H"
#include "mcInletRandom.C"
#include "mcInversionInletRandom.C"
#include "mcSamplingInletRandom.C"
#include "steadyFlameletHelpers.inc"

#include <cstdint>
#include <cstring>

using namespace Foam;

namespace
{
thread_local std::string g_err;
template<class F> int guarded(F f)
{
    try { f(); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
}
}

struct pdfref_decomp
{
    polyMesh mesh;
    tetFacePointCellDecomposition<richTetPointRef>* d;
    pdfref_decomp() : d(0) {}
    ~pdfref_decomp() { delete d; }
};

extern "C"
{

const char* pdfref_last_error() { return g_err.c_str(); }

// ---- mcSteadyFlamelet helpers ----
int pdfref_find_coeffs(int32_t n, const double* lst, double v, int32_t* i, double* w)
{
    return guarded([&] {
        scalarList l(n);
        for (int k = 0; k < n; ++k) l[k] = lst[k];
        int ii; scalar ww;
        findCoeffs(l, v, ii, ww);
        *i = ii; *w = ww;
    });
}
double pdfref_binterp(const double* v, int32_t nRows, int32_t nCols, int32_t ix, double wx, int32_t iy, double wy)
{
    List<scalarList> t(nRows);
    for (int r = 0; r < nRows; ++r) { t[r].setSize(nCols); for (int c = 0; c < nCols; ++c) t[r][c] = v[(size_t)r*nCols + c]; }
    return binterp(t, ix, wx, iy, wy);
}

// ---- mcInletRandom ("inversion"): out = {Q, PDF(x), CDF(x)}; values[k] = value() with u[k] as the uniform ----
int pdfref_inlet_random(double UMean, double uRms, double x, double* out3, int64_t n, const double* u, double* values, int32_t* warnings)
{
    return guarded([&] {
        Random rnd(0);
        dictionary d;
        d.add("type", "inversion");
        autoPtr<mcInletRandom> r = mcInletRandom::New(rnd, UMean, uRms, d);
        out3[0] = r().Q(); out3[1] = r().PDF(x); out3[2] = r().CDF(x);
        const int w0 = messageStream::warnings();
        rnd.queue(u, (size_t)n);
        for (int64_t k = 0; k < n; ++k) values[k] = r().value();
        if (warnings) *warnings = messageStream::warnings() - w0;
    });
}
// n values of the "sampling" generator (mcSamplingInletRandom.C:37-102) on a free-running Random
int pdfref_inlet_sampling(double UMean, double uRms, int32_t seed, int64_t n, double* values)
{
    return guarded([&] {
        Random rnd(seed);
        rnd.freeRunning();
        dictionary d;
        d.add("type", "sampling");
        autoPtr<mcInletRandom> r = mcInletRandom::New(rnd, UMean, uRms, d);
        for (int64_t k = 0; k < n; ++k) values[k] = r().value();
    });
}
// unknown generator types fail like the reference (mcInletRandom.C:76-88)
int pdfref_inlet_random_select(const char* type)
{
    return guarded([&] {
        Random rnd(0);
        dictionary d;
        d.add("type", type);
        mcInletRandom::New(rnd, 1.0, 1.0, d);
    });
}

// ---- tetFacePointCellDecomposition<richTetPointRef> on a polyMesh given by its arrays. Face and cell centres are
//      inputs (OpenFOAM computes them; here they come from the oracle so that the same points feed both); the cells'
//      face lists follow primitiveMesh::calcCells (owner faces in face order, then neighbour faces in face order). ----
pdfref_decomp* pdfref_decomp_create(int32_t nPoints, const double* points, int32_t nFaces, const int32_t* faceStart, const int32_t* facePoints,
                                    const int32_t* owner, int32_t nInternalFaces, const int32_t* neighbour, int32_t nCells,
                                    const double* faceCentres, const double* cellCentres)
{
    pdfref_decomp* h = new pdfref_decomp;
    const int rc = guarded([&] {
        polyMesh& m = h->mesh;
        m.points_.setSize(nPoints);
        for (int i = 0; i < nPoints; ++i) m.points_[i] = point(points[3*i], points[3*i + 1], points[3*i + 2]);
        m.faces_.setSize(nFaces); m.faceCentres_.setSize(nFaces); m.owner_.setSize(nFaces); m.neighbour_.setSize(nInternalFaces);
        for (int f = 0; f < nFaces; ++f)
        {
            m.faces_[f].setSize(faceStart[f + 1] - faceStart[f]);
            for (int k = faceStart[f]; k < faceStart[f + 1]; ++k) m.faces_[f][k - faceStart[f]] = facePoints[k];
            m.faceCentres_[f] = point(faceCentres[3*f], faceCentres[3*f + 1], faceCentres[3*f + 2]);
            m.owner_[f] = owner[f];
        }
        for (int f = 0; f < nInternalFaces; ++f) m.neighbour_[f] = neighbour[f];
        m.cellCentres_.setSize(nCells); m.cells_.setSize(nCells);
        for (int c = 0; c < nCells; ++c) m.cellCentres_[c] = point(cellCentres[3*c], cellCentres[3*c + 1], cellCentres[3*c + 2]);
        for (int f = 0; f < nFaces; ++f) m.cells_[owner[f]].append(f);
        for (int f = 0; f < nInternalFaces; ++f) m.cells_[neighbour[f]].append(f);
        h->d = new tetFacePointCellDecomposition<richTetPointRef>(m);
    });
    if (rc) { delete h; return 0; }
    return h;
}
void pdfref_decomp_destroy(pdfref_decomp* h) { delete h; }
int64_t pdfref_decomp_num_tets(pdfref_decomp* h) { return h->d->tetrahedra().size(); }
// per tet: S[12] (Sa..Sd), mag, gradNi[12], verts[12] (a,b,c,d), face label, the two face-point indices
int pdfref_decomp_tet(pdfref_decomp* h, int64_t t, double* S, double* magV, double* gradN4, double* verts, int32_t* faceLabel, int32_t* ptPair)
{
    return guarded([&] {
        const richTetPointRef& T = h->d->tetrahedra()[label(t)];
        const vector s[4] = {T.Sa(), T.Sb(), T.Sc(), T.Sd()};
        const point v[4] = {T.a(), T.b(), T.c(), T.d()};
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 3; ++k) { S[3*i + k] = s[i][k]; gradN4[3*i + k] = T.gradNi()[i][k]; verts[3*i + k] = v[i][k]; }
        *magV = T.mag();
        *faceLabel = h->d->tetrahedronFace()[label(t)];
        ptPair[0] = h->d->tetrahedronPoints()[label(t)].first();
        ptPair[1] = h->d->tetrahedronPoints()[label(t)].second();
    });
}
// cellTetrahedra(): offsets [nCells+1] and the labels
int pdfref_decomp_cell_tets(pdfref_decomp* h, int32_t* start, int32_t* tets)
{
    return guarded([&] {
        const labelListList& ct = h->d->cellTetrahedra();
        int k = 0;
        forAll(ct, c) { start[c] = k; forAll(ct[c], i) tets[k++] = ct[c][i]; }
        start[ct.size()] = k;
    });
}
int pdfref_decomp_find(pdfref_decomp* h, int64_t n, const double* pts, const int32_t* cellHint, int32_t* out)
{
    return guarded([&] { for (int64_t i = 0; i < n; ++i) out[i] = h->d->find(point(pts[3*i], pts[3*i + 1], pts[3*i + 2]), cellHint[i]); });
}
int pdfref_decomp_inside(pdfref_decomp* h, int64_t t, const double* pt)
{
    return h->d->tetrahedra()[label(t)].inside(point(pt[0], pt[1], pt[2])) ? 1 : 0;
}

}  // extern "C"
