// fields.cu — cell-side kernels: the updateInternals() of every sub-model
// (evaluated once at the start of a step), cell->node interpolation tables,
// moment time-averaging and the derived mean fields.
//
// Reference: mcLimitedSimplePositionCorrection::updateInternals (.C:88-184),
// mcRASOmegaModel::updateInternals (.C:90-122), mcSLMFullVelocityModel::updateInternals
// (.C:104-143), mcCellLocalTimeStepping::updateInternals (.C:81-161),
// mcSteadyFlamelet::updateInternals (.C:291-301), mcParticleCloud::updateCloudPDF
// (mcParticleCloud.C:958-1008), initMoments (:1633-1701); OpenFOAM's
// volPointInterpolation / linearInterpolate / fvc::grad per SURVEY Appendix B.
#include <cstring>
#include <vector>

#include "cloud.hpp"

struct FieldPtrs {
    // FV
    const double *Ufv_c, *Ufv_b, *p_c, *p_b, *kfv_c, *kfv_b, *eps_c, *eps_b, *R_c;
    // cloud
    double *rho_c, *rho_b, *rhoInst_c, *pnd_c, *pndInst_c, *Updf_c, *Tau_c, *Tau_b, *kpdf_c;
    double *Phi_c, *Phi_b, *Cov_c, *Cov_b;
    const uint8_t *rhoFixed, *kbFixed, *PhiFixed, *CovFixed;
    double *PaNIC, *lostMass, *lostShare;
    double *mMom, *VMom, *UMom, *UUMom, *PhiMom, *PhiPhiMom, *UPhiMom, *UPhi_c;   // UPhi*: [nPhi][nCells][3]
    // work
    double *phiPC, *UPC_c, *OmegaRaw_c, *etaRaw_c;
    double *nodeMid, *nodePC, *cellAux;
    const double *pcExt_c, *pcExt_b;   // external position-correction rows (12 per cell / boundary face)
    int pcPlanes;                       // planes of nodePC: plane q of node n at nodePC[(q * nNodes + n) * 4]
};

static FieldPtrs fieldPtrs(Cloud &c) {
    FieldPtrs f;
    f.Ufv_c = c.Ufv_c.p; f.Ufv_b = c.Ufv_b.p; f.p_c = c.p_c.p; f.p_b = c.p_b.p; f.kfv_c = c.kfv_c.p; f.kfv_b = c.kfv_b.p;
    f.eps_c = c.eps_c.p; f.eps_b = c.eps_b.p; f.R_c = c.R_c.p;
    f.rho_c = c.rho_c.p; f.rho_b = c.rho_b.p; f.rhoInst_c = c.rhoInst_c.p; f.pnd_c = c.pnd_c.p; f.pndInst_c = c.pndInst_c.p;
    f.Updf_c = c.Updf_c.p; f.Tau_c = c.Tau_c.p; f.Tau_b = c.Tau_b.p; f.kpdf_c = c.kpdf_c.p;
    f.Phi_c = c.Phi_c.p; f.Phi_b = c.Phi_b.p; f.Cov_c = c.Cov_c.p; f.Cov_b = c.Cov_b.p;
    f.rhoFixed = c.rhoFixed.p; f.kbFixed = c.kbFixed.p; f.PhiFixed = c.PhiFixed.p; f.CovFixed = c.CovFixed.p;
    f.PaNIC = c.PaNIC.p; f.lostMass = c.lostMass.p; f.lostShare = c.lostShare.p;
    f.mMom = c.mMom.p; f.VMom = c.VMom.p; f.UMom = c.UMom.p; f.UUMom = c.UUMom.p; f.PhiMom = c.PhiMom.p; f.PhiPhiMom = c.PhiPhiMom.p; f.UPhiMom = c.UPhiMom.p; f.UPhi_c = c.UPhi_c.p;
    f.phiPC = c.phiPC.p; f.UPC_c = c.UPC_c.p; f.OmegaRaw_c = c.OmegaRaw_c.p; f.etaRaw_c = c.etaRaw_c.p;
    f.nodeMid = c.nodeMid.p; f.nodePC = c.nodePC.p; f.cellAux = c.cellAux.p;
    f.pcExt_c = c.pcExt_c.p; f.pcExt_b = c.pcExt_b.p; f.pcPlanes = c.pcPlanes;
    return f;
}

// ---------------------------------------------------------------------------
// step 1 of the field preparation: raw cell fields + maxima
// partial row: [max|phi|, max Omega, max etaRaw, max(-etaRaw)]
// ---------------------------------------------------------------------------
__global__ void k_prep_cells(DMesh M, FieldPtrs F, pdfb200_params P, const StepScalars *sc, double *partials) {
    __shared__ double sm[4 * 32];
    double v[4] = {0.0, -PDF_VGREAT, -PDF_VGREAT, -PDF_VGREAT};
    const double deltaT = sc->deltaT;
    const int n = M.nCells + M.nBnd;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (i < M.nCells) {
            const int c = i;
            if (P.positionCorrection == PDFB200_POSCORR_LIMITED_SIMPLE) {
                const double ph = (F.pnd_c[c] - F.rho_c[c]) / F.rho_c[c];
                F.phiPC[c] = ph;
                v[0] = fmax(v[0], fabs(ph));
            } else if (P.positionCorrection == PDFB200_POSCORR_SIMPLE) {
                F.phiPC[c] = F.pnd_c[c] / F.rho_c[c] - 1.;   // mcSimplePositionCorrection.C:143-145
            }
            const double Om = F.eps_c[c] / F.kfv_c[c];
            F.OmegaRaw_c[c] = Om;
            v[1] = fmax(v[1], Om);
            if (P.localTimeStepping == PDFB200_LTS_CELL) {
                const double k0 = P.k0;
                const V3 urms2 = mk(2 * sqrt(fmax(k0, F.R_c[6 * c])), 2 * sqrt(fmax(k0, F.R_c[6 * c + 3])), 2 * sqrt(fmax(k0, F.R_c[6 * c + 5])));
                const V3 U = ld3(F.Ufv_c, c);
                double dtc = PDF_GREAT;
                for (int j = M.cellFaceStart[c]; j < M.cellFaceStart[c + 1]; ++j) {
                    const int f = M.cellFaces[j];
                    if (f >= M.nInternalFaces && BF_GEOM(M.bfInfo[f - M.nInternalFaces]) == PDFB200_PATCH_EMPTY) continue;
                    const D4 co4 = M.cellCo4[j];
                    const V3 co = mk(co4.x, co4.y, co4.z);
                    dtc = fmin(dtc, 1. / (fabs(dot3(co, U)) + fabs(dot3(co, urms2)) + PDF_SMALL));
                }
                const double e = dtc / deltaT;
                F.etaRaw_c[c] = e;
                v[2] = fmax(v[2], e); v[3] = fmax(v[3], -e);
            }
        } else {
            const int bf = i - M.nCells;
            if (BF_GEOM(M.bfInfo[bf]) != PDFB200_PATCH_EMPTY) v[1] = fmax(v[1], F.eps_b[bf] / F.kfv_b[bf]);
        }
    }
    blockReduce<4, true>(v, sm);
    if (threadIdx.x == 0) for (int k = 0; k < 4; ++k) partials[blockIdx.x * 4 + k] = v[k];
}

__global__ void k_prep_reduce1(int nRows, const double *partials, pdfb200_params P, StepScalars *sc, double *scratch) {
    // one warp; maxima are order-independent, so the lanes may share the rows
    double v[4] = {0.0, -PDF_VGREAT, -PDF_VGREAT, -PDF_VGREAT};
    for (int r = threadIdx.x; r < nRows; r += 32) for (int k = 0; k < 4; ++k) v[k] = fmax(v[k], partials[r * 4 + k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        for (int k = 0; k < 4; ++k) v[k] = fmax(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
    if (threadIdx.x != 0) return;
    scratch[0] = tanh(10. * v[0]);                                  // phiLim
    sc->omegaClip = (v[1] > P.Omega0) ? P.Omega0 : 1.7976931348623157e308;
    if (P.localTimeStepping == PDFB200_LTS_CELL) {
        const double etaMin = -v[3];
        const double mx = v[2] - etaMin;
        const double etaMaxP = fmax(P.ltsUpperBound, 1.0);
        sc->etaMin = etaMin;
        sc->etaScale = (etaMaxP - 1) / mx;
    } else {
        sc->etaMin = 0; sc->etaScale = 0;
    }
}

// UPosCorr = -fvc::grad(phi) (Gauss linear) and its Courant number
__global__ void k_prep_grad(DMesh M, FieldPtrs F, const StepScalars *sc, double *partials) {
    __shared__ double sm[32];
    double v[1] = {0.0};
    const double deltaT = sc->deltaT;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < M.nCells; c += gridDim.x * blockDim.x) {
        V3 g = mk(0, 0, 0);
        const double phc = F.phiPC[c];
        for (int j = M.cellFaceStart[c]; j < M.cellFaceStart[c + 1]; ++j) {
            const int f = M.cellFaces[j];
            const V3 Sf = ld3(M.faceAreas, f);
            if (f < M.nInternalFaces) {
                const double w = M.linWeights[f];
                const int o = M.owner[f], nb = M.neighbour[f];
                const double pf = w * F.phiPC[o] + (1 - w) * F.phiPC[nb];
                if (o == c) g = g + Sf * pf; else g = g - Sf * pf;
            } else {
                if (BF_GEOM(M.bfInfo[f - M.nInternalFaces]) == PDFB200_PATCH_EMPTY) continue;
                g = g + Sf * phc;   // zeroGradient boundary value
            }
        }
        g = g / M.cellVolumes[c];
        const V3 upc = -g;
        st3(F.UPC_c, c, upc);
        for (int j = M.cellFaceStart[c]; j < M.cellFaceStart[c + 1]; ++j) {
            const D4 co4 = M.cellCo4[j];
            const V3 co = deltaT * mk(co4.x, co4.y, co4.z);
            v[0] = fmax(v[0], fabs(dot3(co, upc)));
        }
    }
    blockReduce<1, true>(v, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = v[0];
}

__global__ void k_prep_reduce2(int nRows, const double *partials, pdfb200_params P, StepScalars *sc, const double *scratch) {
    double Co = 0;
    for (int r = threadIdx.x; r < nRows; r += 32) Co = fmax(Co, partials[r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Co = fmax(Co, __shfl_xor_sync(0xffffffffu, Co, o));
    if (threadIdx.x != 0) return;
    double corr = 0.;
    if (Co > PDF_SMALL) corr = P.pcCFL / Co * P.pcC * scratch[0];
    sc->pcCorr = corr;
}

// ---------------------------------------------------------------------------
// node records: cells, then points (volPointInterpolation), then faces
// (linearInterpolate; patch values on boundary faces)
// ---------------------------------------------------------------------------
__global__ void k_nodes_cell(DMesh M, FieldPtrs F, pdfb200_params P, const StepScalars *sc, int covZZ, int auxStride) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M.nCells) return;
    const double TU = fmax(P.relaxTimeU, P.minRelaxTime), Tk = fmax(P.relaxTimek, P.minRelaxTime);
    const NodeRef nm{F.nodeMid, (size_t)M.nCells + M.nPoints + M.nFaces, (size_t)c};
    const V3 U = ld3(F.Ufv_c, c), Up = ld3(F.Updf_c, c);
    const double k = F.kfv_c[c];
    nm[NM_U] = U.x; nm[NM_U + 1] = U.y; nm[NM_U + 2] = U.z;
    nm[NM_K] = k;
    const V3 dU = (U - Up) / TU;
    nm[NM_DU] = dU.x; nm[NM_DU + 1] = dU.y; nm[NM_DU + 2] = dU.z;
    nm[NM_OMEGA] = fmin(F.OmegaRaw_c[c], sc->omegaClip);
    nm[NM_PSTAR] = F.p_c[c] - 2. / 3. * F.rho_c[c] * k;
    nm[NM_ETA] = (P.localTimeStepping == PDFB200_LTS_CELL) ? sc->etaScale * (F.etaRaw_c[c] - sc->etaMin) + 1 : 1.0;
    nm[NM_ZVAR] = covZZ >= 0 ? F.Cov_c[(size_t)covZZ * M.nCells + c] : 0.0;
    nm[11] = 0;
    double *pc = F.nodePC + (size_t)c * 4;
    if (P.positionCorrection == PDFB200_POSCORR_LIMITED_SIMPLE) {
        const V3 u = ld3(F.UPC_c, c) * sc->pcCorr;
        pc[0] = u.x; pc[1] = u.y; pc[2] = u.z;
    } else if (P.positionCorrection == PDFB200_POSCORR_SIMPLE) {
        // gradPhi = C |Ufv| grad(phi) (mcSimplePositionCorrection.C:149); UPC_c holds -grad(phi)
        const V3 g = (P.pcC * mag3(U)) * (-ld3(F.UPC_c, c));
        pc[0] = g.x; pc[1] = g.y; pc[2] = g.z;
    } else { pc[0] = pc[1] = pc[2] = 0; }
    pc[3] = 0;
    if (P.positionCorrection == PDFB200_POSCORR_EXTERNAL) {
        const size_t nNodes = (size_t)M.nCells + M.nPoints + M.nFaces;
        const double *row = F.pcExt_c + (size_t)c * 12;
        for (int q = 0; q < 3; ++q)
            for (int k = 0; k < 4; ++k) F.nodePC[(q * nNodes + c) * 4 + k] = row[4 * q + k];
    }
    double *aux = F.cellAux + (size_t)c * auxStride;
    aux[0] = (sqrt(k / F.kpdf_c[c]) - 1.0) / Tk;   // diffk
    aux[1] = F.p_c[c];
    for (int i = 0; i < P.nPhi; ++i) aux[2 + i] = F.Phi_c[(size_t)i * M.nCells + c];
}

__global__ void k_nodes_point(DMesh M, FieldPtrs F) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M.nPoints) return;
    double acc[NODE_MID_STRIDE], pc[12];
#pragma unroll
    for (int k = 0; k < NODE_MID_STRIDE; ++k) acc[k] = 0;
#pragma unroll
    for (int k = 0; k < 12; ++k) pc[k] = 0;
    const size_t nNodes = (size_t)M.nCells + M.nPoints + M.nFaces;
    const int nPl = F.pcPlanes;
    for (int j = M.pointCellStart[p]; j < M.pointCellStart[p + 1]; ++j) {
        const double w = M.pointWeights[j];
        const int c = M.pointCells[j];
        const NodeRef nm{F.nodeMid, nNodes, (size_t)c};
#pragma unroll
        for (int k = 0; k < NODE_MID_STRIDE; ++k) acc[k] += w * nm[k];
        for (int q = 0; q < nPl; ++q) {
            const double *r = F.nodePC + (q * nNodes + c) * 4;
            const int nk = nPl > 1 ? 4 : 3;
            for (int k = 0; k < nk; ++k) pc[4 * q + k] += w * r[k];
        }
    }
    const NodeRef o{F.nodeMid, nNodes, (size_t)(M.nCells + p)};
#pragma unroll
    for (int k = 0; k < NODE_MID_STRIDE; ++k) o[k] = acc[k];
    for (int q = 0; q < nPl; ++q) {
        double *oq = F.nodePC + (q * nNodes + M.nCells + p) * 4;
        oq[0] = pc[4 * q]; oq[1] = pc[4 * q + 1]; oq[2] = pc[4 * q + 2]; oq[3] = pc[4 * q + 3];
    }
}

__global__ void k_nodes_face(DMesh M, FieldPtrs F, pdfb200_params P, const StepScalars *sc, int covZZ) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= M.nFaces) return;
    const size_t nNodes = (size_t)M.nCells + M.nPoints + M.nFaces, nodeF = (size_t)M.nCells + M.nPoints + f;
    const NodeRef o{F.nodeMid, nNodes, nodeF};
    double *oq = F.nodePC + nodeF * 4;
    const int ow = M.owner[f];
    const NodeRef no{F.nodeMid, nNodes, (size_t)ow};
    const double *qo = F.nodePC + (size_t)ow * 4;
    const int nPl = F.pcPlanes;
    if (f < M.nInternalFaces) {
        const double w = M.linWeights[f];
        const int nb = M.neighbour[f];
        const NodeRef nn{F.nodeMid, nNodes, (size_t)nb};
#pragma unroll
        for (int k = 0; k < NODE_MID_STRIDE; ++k) o[k] = w * no[k] + (1 - w) * nn[k];
        for (int q = 0; q < nPl; ++q) {
            const double *a = F.nodePC + (q * nNodes + ow) * 4, *b = F.nodePC + (q * nNodes + nb) * 4;
            double *r = F.nodePC + (q * nNodes + nodeF) * 4;
            for (int k = 0; k < 4; ++k) r[k] = w * a[k] + (1 - w) * b[k];
        }
        return;
    }
    const int bf = f - M.nInternalFaces;
    const int info = M.bfInfo[bf];
    const int geom = BF_GEOM(info);
    if (geom == PDFB200_PATCH_EMPTY) {     // "cell value on an empty patch"
#pragma unroll
        for (int k = 0; k < NODE_MID_STRIDE; ++k) o[k] = no[k];
        for (int q = 0; q < nPl; ++q) {
            const double *a = F.nodePC + (q * nNodes + ow) * 4;
            double *r = F.nodePC + (q * nNodes + nodeF) * 4;
            for (int k = 0; k < 4; ++k) r[k] = a[k];
        }
        return;
    }
    const double *T = M.patchFaceT + 9 * BF_PATCH(info);
    const V3 Ub = ld3(F.Ufv_b, bf);
    o[NM_U] = Ub.x; o[NM_U + 1] = Ub.y; o[NM_U + 2] = Ub.z;
    o[NM_K] = F.kfv_b[bf];
    V3 dU = mk(no[NM_DU], no[NM_DU + 1], no[NM_DU + 2]);       // zeroGradient
    if (geom == PDFB200_PATCH_WEDGE) dU = xform(T, dU);
    o[NM_DU] = dU.x; o[NM_DU + 1] = dU.y; o[NM_DU + 2] = dU.z;
    o[NM_OMEGA] = fmin(F.eps_b[bf] / F.kfv_b[bf], sc->omegaClip);
    o[NM_PSTAR] = F.p_b[bf] - 2. / 3. * F.rho_b[bf] * F.kfv_b[bf];
    o[NM_ETA] = no[NM_ETA];                                     // zeroGradient
    o[NM_ZVAR] = covZZ >= 0 ? F.Cov_b[(size_t)covZZ * M.nBnd + bf] : 0.0;
    o[11] = 0;
    if (P.positionCorrection == PDFB200_POSCORR_EXTERNAL) {
        // boundary values as uploaded (the OpenFOAM side has run correctBoundaryConditions())
        const double *row = F.pcExt_b + (size_t)bf * 12;
        for (int q = 0; q < 3; ++q) {
            double *r = F.nodePC + (q * nNodes + nodeF) * 4;
            for (int k = 0; k < 4; ++k) r[k] = row[4 * q + k];
        }
        return;
    }
    // limitedSimple: UPosCorr has the patch types of rho, fixedValue faces are zero (.C:76-84); simple: grad(phi) is
    // zeroGradient (mcSimplePositionCorrection.C:92-109)
    V3 u = mk(qo[0], qo[1], qo[2]);
    if (geom == PDFB200_PATCH_WEDGE) u = xform(T, u);
    else if (P.positionCorrection == PDFB200_POSCORR_LIMITED_SIMPLE && F.rhoFixed[bf]) u = mk(0, 0, 0);
    oq[0] = u.x; oq[1] = u.y; oq[2] = u.z; oq[3] = 0;
}

// ---------------------------------------------------------------------------
// correctBoundaryConditions() of the cloud fields that are read on boundary
// faces (rho, scalars, covariances, Tau)
// ---------------------------------------------------------------------------
__device__ inline void xformSymm(const double *T, const double *s, double *r) {
    const double S[9] = {s[0], s[1], s[2], s[1], s[3], s[4], s[2], s[4], s[5]};
    double TS[9], R[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double v = 0; for (int k = 0; k < 3; ++k) v += T[3 * i + k] * S[3 * k + j]; TS[3 * i + j] = v; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double v = 0; for (int k = 0; k < 3; ++k) v += TS[3 * i + k] * T[3 * j + k]; R[3 * i + j] = v; }
    r[0] = R[0]; r[1] = R[1]; r[2] = R[2]; r[3] = R[4]; r[4] = R[5]; r[5] = R[8];
}

__global__ void k_bnd_apply(DMesh M, FieldPtrs F, int nPhi, int nCov) {
    const int bf = blockIdx.x * blockDim.x + threadIdx.x;
    if (bf >= M.nBnd) return;
    const int info = M.bfInfo[bf], geom = BF_GEOM(info);
    const int o = M.owner[M.nInternalFaces + bf];
    const bool generic = geom == PDFB200_PATCH_GENERIC;
    if (!(generic && F.rhoFixed[bf])) F.rho_b[bf] = F.rho_c[o];
    for (int i = 0; i < nPhi; ++i)
        if (!(generic && F.PhiFixed[(size_t)i * M.nBnd + bf])) F.Phi_b[(size_t)i * M.nBnd + bf] = F.Phi_c[(size_t)i * M.nCells + o];
    for (int i = 0; i < nCov; ++i)
        if (!(generic && F.CovFixed[(size_t)i * M.nBnd + bf])) F.Cov_b[(size_t)i * M.nBnd + bf] = F.Cov_c[(size_t)i * M.nCells + o];
    if (!(generic && F.kbFixed[bf])) {
        if (geom == PDFB200_PATCH_WEDGE) xformSymm(M.patchFaceT + 9 * BF_PATCH(info), F.Tau_c + 6 * (size_t)o, F.Tau_b + 6 * (size_t)bf);
        else for (int k = 0; k < 6; ++k) F.Tau_b[6 * (size_t)bf + k] = F.Tau_c[6 * (size_t)o + k];
    }
}

// ---------------------------------------------------------------------------
// initMoments (mcParticleCloud.C:1633-1701)
// ---------------------------------------------------------------------------
__global__ void k_init_moments(DMesh M, FieldPtrs F, int nPhi) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M.nCells) return;
    const double rho = F.rho_c[c];
    const double mM = rho * M.volumeOrArea[c];
    F.mMom[c] = mM;
    F.VMom[c] = mM / rho;
    F.rhoInst_c[c] = rho; F.pnd_c[c] = rho; F.pndInst_c[c] = rho;
    const V3 u = ld3(F.Ufv_c, c);
    st3(F.Updf_c, c, u);
    st3(F.UMom, c, mM * u);
    F.kpdf_c[c] = F.kfv_c[c];
    const double *R = F.R_c + 6 * (size_t)c;
    double *uu = F.UUMom + 6 * (size_t)c, *tau = F.Tau_c + 6 * (size_t)c;
    uu[0] = mM * (R[0] + u.x * u.x); uu[1] = mM * (R[1] + u.x * u.y); uu[2] = mM * (R[2] + u.x * u.z);
    uu[3] = mM * (R[3] + u.y * u.y); uu[4] = mM * (R[4] + u.y * u.z); uu[5] = mM * (R[5] + u.z * u.z);
    for (int k = 0; k < 6; ++k) tau[k] = R[k];
    int pp = 0;
    for (int i = 0; i < nPhi; ++i) {
        const double pi = F.Phi_c[(size_t)i * M.nCells + c];
        F.PhiMom[(size_t)i * M.nCells + c] = mM * pi;
        st3(F.UPhiMom, (size_t)i * M.nCells + c, (mM * pi) * u);   // zero flux to start with
        st3(F.UPhi_c, (size_t)i * M.nCells + c, mk(0, 0, 0));
        for (int j = i; j < nPhi; ++j, ++pp)
            F.PhiPhiMom[(size_t)pp * M.nCells + c] = mM * (F.Cov_c[(size_t)pp * M.nCells + c] + pi * F.Phi_c[(size_t)j * M.nCells + c]);
    }
    F.PaNIC[c] = 0; F.lostMass[c] = 0; F.lostShare[c] = 0;
}

// ---------------------------------------------------------------------------
// updateCloudPDF second half (mcParticleCloud.C:958-1008): time-average the
// (all-reduced) instantaneous block and derive the mean fields. Also the cell
// sums of E15 (:1376-1381,1432-1433). Partial row: sums [ |rhoErr|, totalMass,
// totalParticleMass ], maxima [ -rhoErrMin, rhoErrMax ].
// ---------------------------------------------------------------------------
// deriveOnly: restart from uploaded moments (pdfb200_upload_moments): existing weight 1, the block is not read.
__global__ void k_finalise(DMesh M, FieldPtrs F, pdfb200_params P, const double *block, int nMom, int nCov,
                           double *partSum, double *partMax, int deriveOnly) {
    __shared__ double sm[3 * 32];
    double s[3] = {0, 0, 0}, mx[2] = {-PDF_VGREAT, -PDF_VGREAT};
    const double existWt = deriveOnly ? 1.0 : (P.averagingCoeff - 1.) / P.averagingCoeff, newWt = 1.0 - existWt;
    const int nPhi = P.nPhi;
    double zeros[12 + 2 * PDFB200_MAX_PHI + PDFB200_MAX_COV + 3 * PDFB200_MAX_PHI];
    if (deriveOnly) for (int k = 0; k < nMom; ++k) zeros[k] = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < M.nCells; c += gridDim.x * blockDim.x) {
        const double *b = deriveOnly ? zeros : block + (size_t)c * nMom;
        F.PaNIC[c] = b[0];
        const double mI = b[1], VI = b[2];
        const double voa = M.volumeOrArea[c];
        const double mM = existWt * F.mMom[c] + newWt * mI;
        F.mMom[c] = mM;
        const double mB = fmax(mM, PDF_SMALL);
        const double pndI = mI / voa, pnd = mM / voa;
        F.pndInst_c[c] = pndI; F.pnd_c[c] = pnd;
        const double VM = existWt * F.VMom[c] + newWt * VI;
        F.VMom[c] = VM;
        const double rhoI = mI / fmax(VI, PDF_SMALL), rho = mM / fmax(VM, PDF_SMALL);
        F.rhoInst_c[c] = rhoI; F.rho_c[c] = rho;
        V3 UM = existWt * ld3(F.UMom, c) + newWt * mk(b[3], b[4], b[5]);
        st3(F.UMom, c, UM);
        const V3 u = UM / mB;
        st3(F.Updf_c, c, u);
        double phi[PDFB200_MAX_PHI];
        for (int i = 0; i < nPhi; ++i) {
            const size_t o = (size_t)i * M.nCells + c;
            const double pm = existWt * F.PhiMom[o] + newWt * b[6 + i];
            F.PhiMom[o] = pm;
            phi[i] = pm / mB;
            F.Phi_c[o] = phi[i];
        }
        int pp = 0;
        for (int i = 0; i < nPhi; ++i)
            for (int j = i; j < nPhi; ++j, ++pp) {
                const size_t o = (size_t)pp * M.nCells + c;
                const double pm = existWt * F.PhiPhiMom[o] + newWt * b[6 + nPhi + pp];
                F.PhiPhiMom[o] = pm;
                F.Cov_c[o] = pm / mB - phi[i] * phi[j];
            }
        const double *uuI = b + 6 + nPhi + nCov;
        double *uu = F.UUMom + 6 * (size_t)c, *tau = F.Tau_c + 6 * (size_t)c;
        for (int k = 0; k < 6; ++k) uu[k] = existWt * uu[k] + newWt * uuI[k];
        tau[0] = uu[0] / mB - u.x * u.x; tau[1] = uu[1] / mB - u.x * u.y; tau[2] = uu[2] / mB - u.x * u.z;
        tau[3] = uu[3] / mB - u.y * u.y; tau[4] = uu[4] / mB - u.y * u.z; tau[5] = uu[5] / mB - u.z * u.z;
        // bound(kcPdf_, kMin) restated as a floor
        F.kpdf_c[c] = fmax(0.5 * (tau[0] + tau[3] + tau[5]), P.kMin);
        const double *upI = uuI + 6;
        for (int i = 0; i < nPhi; ++i) {
            const size_t o = (size_t)i * M.nCells + c;
            const V3 fm = existWt * ld3(F.UPhiMom, o) + newWt * mk(upI[3 * i], upI[3 * i + 1], upI[3 * i + 2]);
            st3(F.UPhiMom, o, fm);
            st3(F.UPhi_c, o, fm / mB - phi[i] * u);
        }
        const double e = (pnd - rho) / rho;
        s[0] += fabs(e);
        const double V = M.cellVolumes[c];
        s[1] += rhoI * V; s[2] += pndI * V;
        mx[0] = fmax(mx[0], -e); mx[1] = fmax(mx[1], e);
    }
    blockReduce<3, false>(s, sm);
    blockReduce<2, true>(mx, sm);
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; ++k) partSum[blockIdx.x * 3 + k] = s[k];
        for (int k = 0; k < 2; ++k) partMax[blockIdx.x * 2 + k] = mx[k];
    }
}

// ===========================================================================
// host side
// ===========================================================================
static void uploadOrZero(DBuf<double> &d, const double *h, size_t n, cudaStream_t s) {
    if (d.n < n) d.alloc(n);
    if (h) d.upload(h, n, s); else d.zero(s);
}
static void uploadFlags(DBuf<uint8_t> &d, const uint8_t *h, size_t n, cudaStream_t s, size_t offset = 0, size_t total = 0) {
    if (total == 0) total = n;
    if (d.n < total) { d.alloc(total); d.zero(s); }
    if (h) CUDA_CHECK(cudaMemcpyAsync(d.p + offset, h, n, cudaMemcpyHostToDevice, s));
    else CUDA_CHECK(cudaMemsetAsync(d.p + offset, 0, n, s));
}

void Cloud::uploadFv(const pdfb200_fv_fields &f) {
    dropGraphs();
    if (!f.U || !f.Ub || !f.p || !f.pb || !f.k || !f.kb || !f.epsilon || !f.epsilonb || !f.R || !f.Rb)
        throw std::runtime_error("upload_fv_fields: null field pointer");
    cudaStream_t s = stream;
    const size_t nc = nCells, nb = nBnd;
    uploadOrZero(Ufv_c, f.U, nc * 3, s); uploadOrZero(Ufv_b, f.Ub, nb * 3, s);
    uploadOrZero(p_c, f.p, nc, s); uploadOrZero(p_b, f.pb, nb, s);
    uploadOrZero(kfv_c, f.k, nc, s); uploadOrZero(kfv_b, f.kb, nb, s);
    uploadOrZero(eps_c, f.epsilon, nc, s); uploadOrZero(eps_b, f.epsilonb, nb, s);
    uploadOrZero(R_c, f.R, nc * 6, s); uploadOrZero(R_b, f.Rb, nb * 6, s);
    uploadFlags(UbFixed, f.UbFixed, nb, s); uploadFlags(kbFixed, f.kbFixed, nb, s);
    CUDA_CHECK(cudaStreamSynchronize(s));
    haveFv = true;
}

void Cloud::stageFv(const pdfb200_fv_fields &f) {
    if (!f.U || !f.Ub || !f.p || !f.pb || !f.k || !f.kb || !f.epsilon || !f.epsilonb || !f.R || !f.Rb)
        throw std::runtime_error("stage_fv_fields: null field pointer");
    if (!copyStream) {
        CUDA_CHECK(cudaStreamCreateWithFlags(&copyStream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&evStaged, cudaEventDisableTiming));
    }
    cudaStream_t s = copyStream;
    const size_t nc = nCells, nb = nBnd;
    uploadOrZero(stg_U_c, f.U, nc * 3, s); uploadOrZero(stg_U_b, f.Ub, nb * 3, s);
    uploadOrZero(stg_p_c, f.p, nc, s); uploadOrZero(stg_p_b, f.pb, nb, s);
    uploadOrZero(stg_k_c, f.k, nc, s); uploadOrZero(stg_k_b, f.kb, nb, s);
    uploadOrZero(stg_eps_c, f.epsilon, nc, s); uploadOrZero(stg_eps_b, f.epsilonb, nb, s);
    uploadOrZero(stg_R_c, f.R, nc * 6, s); uploadOrZero(stg_R_b, f.Rb, nb * 6, s);
    uploadFlags(stg_UbFixed, f.UbFixed, nb, s); uploadFlags(stg_kbFixed, f.kbFixed, nb, s);
    CUDA_CHECK(cudaEventRecord(evStaged, s));
    stagedPending = true;
}

template <class T>
static void swapBuf(DBuf<T> &a, DBuf<T> &b) { std::swap(a.p, b.p); std::swap(a.n, b.n); }

void Cloud::commitFv() {
    if (!stagedPending) throw std::logic_error("commit_fv_fields: nothing staged (stage_fv_fields first)");
    // every later launch on the library's stream sees the staged copy complete; the host does not wait
    CUDA_CHECK(cudaStreamWaitEvent(stream, evStaged, 0));
    swapBuf(Ufv_c, stg_U_c); swapBuf(Ufv_b, stg_U_b); swapBuf(p_c, stg_p_c); swapBuf(p_b, stg_p_b);
    swapBuf(kfv_c, stg_k_c); swapBuf(kfv_b, stg_k_b); swapBuf(eps_c, stg_eps_c); swapBuf(eps_b, stg_eps_b);
    swapBuf(R_c, stg_R_c); swapBuf(R_b, stg_R_b); swapBuf(UbFixed, stg_UbFixed); swapBuf(kbFixed, stg_kbFixed);
    stagedPending = false;
    haveFv = true;
    dropGraphs();   // the captured steps hold the old pointers
}

void Cloud::uploadCloudFields(const pdfb200_cloud_fields &f) {
    orderAfterDownloads();
    dropGraphs();
    if (!f.rho || !f.rhob) throw std::runtime_error("upload_cloud_fields: rho missing");
    cudaStream_t s = stream;
    const size_t nc = nCells, nb = nBnd;
    uploadOrZero(rho_c, f.rho, nc, s); uploadOrZero(rho_b, f.rhob, nb, s);
    uploadFlags(rhoFixed, f.rhobFixed, nb, s);
    const int nPhi = prm.nPhi;
    if (Phi_c.n < nc * std::max(nPhi, 1)) { Phi_c.alloc(nc * std::max(nPhi, 1)); Phi_b.alloc(nb * std::max(nPhi, 1)); }
    if (Cov_c.n < nc * std::max(nCov, 1)) { Cov_c.alloc(nc * std::max(nCov, 1)); Cov_b.alloc(nb * std::max(nCov, 1)); }
    Cov_c.zero(s); Cov_b.zero(s);
    for (int i = 0; i < nPhi; ++i) {
        if (!f.Phi[i] || !f.Phib[i]) throw std::runtime_error("upload_cloud_fields: scalar field missing");
        CUDA_CHECK(cudaMemcpyAsync(Phi_c.p + i * nc, f.Phi[i], nc * sizeof(double), cudaMemcpyHostToDevice, s));
        CUDA_CHECK(cudaMemcpyAsync(Phi_b.p + i * nb, f.Phib[i], nb * sizeof(double), cudaMemcpyHostToDevice, s));
        uploadFlags(PhiFixed, f.PhibFixed[i], nb, s, i * nb, nb * std::max(nPhi, 1));
    }
    for (int i = 0; i < nCov; ++i) {
        if (f.Cov[i]) {
            CUDA_CHECK(cudaMemcpyAsync(Cov_c.p + i * nc, f.Cov[i], nc * sizeof(double), cudaMemcpyHostToDevice, s));
            if (f.Covb[i]) CUDA_CHECK(cudaMemcpyAsync(Cov_b.p + i * nb, f.Covb[i], nb * sizeof(double), cudaMemcpyHostToDevice, s));
        }
        uploadFlags(CovFixed, f.Cov[i] ? f.CovbFixed[i] : nullptr, nb, s, i * nb, nb * std::max(nCov, 1));
    }
    CUDA_CHECK(cudaStreamSynchronize(s));
    haveCloudFields = true;
}

void Cloud::initMoments() {
    orderAfterDownloads();
    dropGraphs();
    if (!haveFv || !haveCloudFields) throw std::logic_error("init_moments: upload the FV and cloud fields first");
    const size_t nc = nCells, nb = nBnd;
    const int nPhi = std::max(prm.nPhi, 1), nC = std::max(nCov, 1);
    rhoInst_c.alloc(nc); pnd_c.alloc(nc); pndInst_c.alloc(nc); Updf_c.alloc(nc * 3); Tau_c.alloc(nc * 6); Tau_b.alloc(nb * 6);
    kpdf_c.alloc(nc); PaNIC.alloc(nc); lostMass.alloc(nc); lostShare.alloc(nc);
    mMom.alloc(nc); VMom.alloc(nc); UMom.alloc(nc * 3); UUMom.alloc(nc * 6); PhiMom.alloc(nc * nPhi); PhiPhiMom.alloc(nc * nC);
    UPhiMom.alloc(nc * nPhi * 3); UPhi_c.alloc(nc * nPhi * 3);
    momentBlock.alloc(nc * nMom + BLOCK_TAIL); momentBlock.zero(stream);
    phiPC.alloc(nc); UPC_c.alloc(nc * 3); OmegaRaw_c.alloc(nc); etaRaw_c.alloc(nc);
    const size_t nNodes = nc + nPoints + nFaces;
    nodeMid.alloc(nNodes * NODE_MID_STRIDE); nodePC.alloc(nNodes * 4 * pcPlanes); cellAux.alloc(nc * cellAuxStride);
    phiPC.zero(stream); UPC_c.zero(stream); etaRaw_c.zero(stream);
    // TauCloudPDF boundary = R boundary (turbModel_.R()), patch types of k
    CUDA_CHECK(cudaMemcpyAsync(Tau_b.p, R_b.p, nb * 6 * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    FieldPtrs F = fieldPtrs(*this);
    LAUNCH(this, k_init_moments, gridFor(nCells, 128), 128, 0, dm, F, prm.nPhi);
    CUDA_CHECK(cudaStreamSynchronize(stream));
    haveMoments = true;
    inletCached = false;
}

static int covIndexZZ(const pdfb200_params &p) {
    if (p.reactionModel != PDFB200_REACTION_STEADY_FLAMELET) return -1;
    int pp = 0;
    for (int i = 0; i < p.nPhi; ++i)
        for (int j = i; j < p.nPhi; ++j, ++pp)
            if (i == p.zIdx && j == p.zIdx) return pp;
    return -1;
}

void Cloud::prepareFields() {
    FieldPtrs F = fieldPtrs(*this);
    const int B = 256;
    const int gC = std::min(gridFor((long long)nCells + nBnd, B), numSMs * 8);
    if ((int)k1PartMax.n < gC * 4) k1PartMax.alloc((size_t)std::max(gC * 4, numSMs * 64));
    double *scratch = finalSums.p + 64;   // device scratch scalars
    LAUNCH(this, k_prep_cells, gC, B, 0, dm, F, prm, scal.p, k1PartMax.p);
    LAUNCH(this, k_prep_reduce1, 1, 32, 0, gC, k1PartMax.p, prm, scal.p, scratch);
    if (prm.positionCorrection == PDFB200_POSCORR_LIMITED_SIMPLE || prm.positionCorrection == PDFB200_POSCORR_SIMPLE) {
        const int gG = std::min(gridFor(nCells, B), numSMs * 8);
        LAUNCH(this, k_prep_grad, gG, B, 0, dm, F, scal.p, k1PartMax.p);
        if (prm.positionCorrection == PDFB200_POSCORR_LIMITED_SIMPLE) LAUNCH(this, k_prep_reduce2, 1, 32, 0, gG, k1PartMax.p, prm, scal.p, scratch);
    }
    if (prm.positionCorrection == PDFB200_POSCORR_EXTERNAL && !havePcExt)
        throw std::logic_error("evolve: positionCorrection external selected but no fields uploaded (upload_poscorr_fields)");
    const int zz = covIndexZZ(prm);
    LAUNCH(this, k_nodes_cell, gridFor(nCells, 128), 128, 0, dm, F, prm, scal.p, zz, cellAuxStride);
    LAUNCH(this, k_nodes_point, gridFor(nPoints, 128), 128, 0, dm, F);
    LAUNCH(this, k_nodes_face, gridFor(nFaces, 128), 128, 0, dm, F, prm, scal.p, zz);
}

// called by Cloud::finaliseStep (particles.cu) once the moment block is complete
void fieldsFinalise(Cloud &c, double *partSum, double *partMax, int grid) {
    FieldPtrs F = fieldPtrs(c);
    LAUNCH(&c, k_finalise, grid, 256, 0, c.dm, F, c.prm, c.momentBlock.p, c.nMom, c.nCov, partSum, partMax, 0);
    LAUNCH(&c, k_bnd_apply, gridFor(c.nBnd, 128), 128, 0, c.dm, F, c.prm.nPhi, c.nCov);
}

static void copyCellFields(Cloud &c, const pdfb200_cell_fields &o, cudaStream_t s) {
    const size_t nc = c.nCells;
    auto dl = [&](double *dst, const double *src, size_t n) { if (dst) CUDA_CHECK(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, s)); };
    dl(o.rho, c.rho_c.p, nc); dl(o.rhoInst, c.rhoInst_c.p, nc); dl(o.pnd, c.pnd_c.p, nc); dl(o.pndInst, c.pndInst_c.p, nc);
    dl(o.U, c.Updf_c.p, nc * 3); dl(o.Tau, c.Tau_c.p, nc * 6); dl(o.k, c.kpdf_c.p, nc); dl(o.PaNIC, c.PaNIC.p, nc);
    dl(o.mMom, c.mMom.p, nc); dl(o.VMom, c.VMom.p, nc); dl(o.UMom, c.UMom.p, nc * 3); dl(o.UUMom, c.UUMom.p, nc * 6);
    for (int i = 0; i < c.prm.nPhi; ++i) { dl(o.Phi[i], c.Phi_c.p + i * nc, nc); dl(o.PhiMom[i], c.PhiMom.p + i * nc, nc); }
    for (int i = 0; i < c.nCov; ++i) { dl(o.Cov[i], c.Cov_c.p + i * nc, nc); dl(o.PhiPhiMom[i], c.PhiPhiMom.p + i * nc, nc); }
    for (int i = 0; i < c.prm.nPhi; ++i) { dl(o.UPhi[i], c.UPhi_c.p + i * nc * 3, nc * 3); dl(o.UPhiMom[i], c.UPhiMom.p + i * nc * 3, nc * 3); }
}

void Cloud::downloadCellFields(pdfb200_cell_fields &o) {
    if (!haveMoments) throw std::logic_error("download_cell_fields: moments not initialised");
    copyCellFields(*this, o, stream);
    CUDA_CHECK(cudaStreamSynchronize(stream));
}

void Cloud::downloadCellFieldsAsync(const pdfb200_cell_fields &o) {
    if (!haveMoments) throw std::logic_error("download_cell_fields_async: moments not initialised");
    if (!dlStream) {
        CUDA_CHECK(cudaStreamCreateWithFlags(&dlStream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&evStepDone, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&evDownloaded, cudaEventDisableTiming));
    }
    // behind everything issued on the library's stream so far, beside whatever is issued next
    CUDA_CHECK(cudaEventRecord(evStepDone, stream));
    CUDA_CHECK(cudaStreamWaitEvent(dlStream, evStepDone, 0));
    copyCellFields(*this, o, dlStream);
    CUDA_CHECK(cudaEventRecord(evDownloaded, dlStream));
    downloadPending = true;
}

void Cloud::syncDownloads() {
    if (dlStream) CUDA_CHECK(cudaStreamSynchronize(dlStream));
    downloadPending = false;
}

void Cloud::orderAfterDownloads() {
    if (!downloadPending) return;
    CUDA_CHECK(cudaStreamWaitEvent(stream, evDownloaded, 0));
    downloadPending = false;
}

// fields of an externally solved position correction (pdfb200_poscorr_fields): packed into rows of 12
void Cloud::uploadPoscorr(const pdfb200_poscorr_fields &f) {
    dropGraphs();
    const size_t nc = nCells, nb = nBnd;
    std::vector<double> hc(nc * 12, 0.0), hb(std::max<size_t>(nb, 1) * 12, 0.0);
    auto put = [&](const double *cv, const double *bv, int off, int width) {
        if (!cv) return;
        if (!bv) throw std::invalid_argument("upload_poscorr_fields: cell values without boundary values");
        for (size_t i = 0; i < nc; ++i) for (int k = 0; k < width; ++k) hc[i * 12 + off + k] = cv[i * width + k];
        for (size_t i = 0; i < nb; ++i) for (int k = 0; k < width; ++k) hb[i * 12 + off + k] = bv[i * width + k];
    };
    put(f.G0, f.G0b, 0, 3); put(f.l, f.lb, 3, 1);
    put(f.G1, f.G1b, 4, 3); put(f.zeta, f.zetab, 7, 1);
    put(f.G2, f.G2b, 8, 3);
    pcExt_c.upload(hc.data(), hc.size(), stream); pcExt_b.upload(hb.data(), hb.size(), stream);
    CUDA_CHECK(cudaStreamSynchronize(stream));
    pcScale = f.scale;
    havePcExt = true;
    if (pcPlanes != 3) {
        pcPlanes = 3;
        const size_t nNodes = nc + nPoints + nFaces;
        if (nodePC.n) { nodePC.alloc(nNodes * 4 * 3); nodePC.zero(stream); CUDA_CHECK(cudaStreamSynchronize(stream)); }
    }
}

// the moment fields of a previous run instead of initMoments (mcParticleCloud.C:339-517, checkMoments :739-820)
__global__ void k_flux_from_means(int n, int nCells, const double *PhiMom, const double *mMom, const double *UMom, double *UPhiMom) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int c = j % nCells;
    st3(UPhiMom, j, (PhiMom[j] / fmax(mMom[c], PDF_SMALL)) * ld3(UMom, c));
}
void Cloud::uploadMoments(const pdfb200_cell_fields &m) {
    orderAfterDownloads();
    if (!haveFv || !haveCloudFields) throw std::logic_error("upload_moments: upload the FV and cloud fields first");
    bool complete = m.mMom && m.VMom && m.UMom && m.UUMom;
    for (int i = 0; i < prm.nPhi; ++i) complete = complete && m.PhiMom[i];
    for (int i = 0; i < nCov; ++i) complete = complete && m.PhiPhiMom[i];
    if (!complete) throw std::invalid_argument("upload_moments: Moments are missing (use init_moments)");
    initMoments();   // allocations, boundary values of the derived fields
    const size_t nc = nCells;
    auto ul = [&](double *dst, const double *src, size_t n) { CUDA_CHECK(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, stream)); };
    ul(mMom.p, m.mMom, nc); ul(VMom.p, m.VMom, nc); ul(UMom.p, m.UMom, nc * 3); ul(UUMom.p, m.UUMom, nc * 6);
    for (int i = 0; i < prm.nPhi; ++i) ul(PhiMom.p + i * nc, m.PhiMom[i], nc);
    for (int i = 0; i < nCov; ++i) ul(PhiPhiMom.p + i * nc, m.PhiPhiMom[i], nc);
    bool haveFlux = prm.nPhi > 0;
    for (int i = 0; i < prm.nPhi; ++i) haveFlux = haveFlux && m.UPhiMom[i];
    if (haveFlux) for (int i = 0; i < prm.nPhi; ++i) ul(UPhiMom.p + i * nc * 3, m.UPhiMom[i], nc * 3);
    else if (prm.nPhi > 0) LAUNCH(this, k_flux_from_means, gridFor((long long)nc * prm.nPhi, 256), 256, 0, (int)(nc * prm.nPhi), nCells, PhiMom.p, mMom.p, UMom.p, UPhiMom.p);
    FieldPtrs F = fieldPtrs(*this);
    const int gF = std::min(gridFor(nCells, 256), numSMs * 8);
    if (cellPartSum.n < (size_t)gF * 8) cellPartSum.alloc((size_t)std::max(gF, numSMs * 16) * 8);
    LAUNCH(this, k_finalise, gF, 256, 0, dm, F, prm, momentBlock.p, nMom, nCov, cellPartSum.p, cellPartSum.p + (size_t)gF * 3, 1);
    LAUNCH(this, k_bnd_apply, gridFor(nBnd, 128), 128, 0, dm, F, prm.nPhi, nCov);
    // the instantaneous fields are diagnostics of the last step: they restart from the averaged ones
    CUDA_CHECK(cudaMemcpyAsync(rhoInst_c.p, rho_c.p, nc * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    CUDA_CHECK(cudaMemcpyAsync(pndInst_c.p, pnd_c.p, nc * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    if (m.PaNIC) ul(PaNIC.p, m.PaNIC, nc);
    CUDA_CHECK(cudaStreamSynchronize(stream));
}
