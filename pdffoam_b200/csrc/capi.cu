// capi.cu — the extern "C" entry points of include/pdfb200.h, plus handle
// construction. No exception crosses the ABI: every call returns a status and
// stores the message for pdfb200_last_error().
#include <cstring>
#include <new>
#include <string>

#include "cloud.hpp"

struct pdfb200_cloud { Cloud c; };

static thread_local std::string g_err;

template <class F>
static int guarded(F &&f) {
    try {
        f();
        return PDFB200_OK;
    } catch (const CudaError &e) { g_err = e.what(); return PDFB200_ERR_CUDA; }
    catch (const std::length_error &e) { g_err = e.what(); return PDFB200_ERR_CAPACITY; }
    catch (const std::logic_error &e) { g_err = e.what(); return dynamic_cast<const std::invalid_argument *>(&e) ? PDFB200_ERR_INVALID : PDFB200_ERR_STATE; }
    catch (const std::bad_alloc &) { g_err = "out of host memory"; return PDFB200_ERR_CAPACITY; }
    catch (const std::exception &e) { g_err = e.what(); return PDFB200_ERR_INVALID; }
}

void commDestroy(Cloud &c);   // comm.cu

Cloud::~Cloud() {
    if (device >= 0) cudaSetDevice(device);
    commDestroy(*this);
    dropSpecialisedKernel();
    for (void *p : stateAllocs) cudaFree(p);
    if (hScal) cudaFreeHost(hScal);
    if (hFinal) cudaFreeHost(hFinal);
    if (hRing) cudaFreeHost(hRing);
    for (auto &e : evMarks) if (e) cudaEventDestroy(e);
    dropGraphs();
    if (evStart) cudaEventDestroy(evStart);
    if (evStop) cudaEventDestroy(evStop);
    if (evArA) cudaEventDestroy(evArA);
    if (evArB) cudaEventDestroy(evArB);
    for (auto &e : evMark) if (e) cudaEventDestroy(e);
    if (dlStream) { cudaStreamSynchronize(dlStream); cudaStreamDestroy(dlStream); }
    if (evStepDone) cudaEventDestroy(evStepDone);
    if (evDownloaded) cudaEventDestroy(evDownloaded);
    if (evStaged) cudaEventDestroy(evStaged);
    if (copyStream) cudaStreamDestroy(copyStream);
    if (stream) cudaStreamDestroy(stream);
}

static void validateParams(const pdfb200_params &p) {
    if (p.nPhi < 0 || p.nPhi > PDFB200_MAX_PHI) throw std::invalid_argument("nPhi out of range");
    if (p.nMixed < 0 || p.nMixed > p.nPhi || p.nConserved < 0 || p.nConserved > p.nPhi) throw std::invalid_argument("too many mixed/conserved scalars");
    for (int i = 0; i < p.nMixed; ++i) if (p.mixed[i] < 0 || p.mixed[i] >= p.nPhi) throw std::invalid_argument("No such scalar field (mixedScalars)");
    for (int i = 0; i < p.nConserved; ++i) if (p.conserved[i] < 0 || p.conserved[i] >= p.nPhi) throw std::invalid_argument("No such scalar field (conservedScalars)");
    if (!(p.cloneAt >= 0 && p.cloneAt < 1.)) throw std::invalid_argument("cloneAt must be in the range [0, 1)");      // mcSolution.C:146-152
    if (!(p.eliminateAt > 1)) throw std::invalid_argument("eliminateAt must be > 1");                                   // mcSolution.C:155-161
    if (p.particlesPerCell <= 0) throw std::invalid_argument("particlesPerCell must be positive");
    if (p.velocityModel < 0 || p.velocityModel > PDFB200_VELOCITY_SLMFULL) throw std::invalid_argument("Unknown mcVelocityModel type");
    if (p.mixingModel < 0 || p.mixingModel > PDFB200_MIXING_MODCURL) throw std::invalid_argument("Unknown mcMixingModel type");
    if (p.omegaModel < 0 || p.omegaModel > PDFB200_OMEGA_RAS) throw std::invalid_argument("Unknown mcOmegaModel type");
    if (p.reactionModel < 0 || p.reactionModel > PDFB200_REACTION_KULKARNI) throw std::invalid_argument("Unknown mcReactionModel type");
    if (p.positionCorrection < 0 || p.positionCorrection > PDFB200_POSCORR_EXTERNAL) throw std::invalid_argument("Unknown mcPositionCorrection type");
    if (p.inletRandom < 0 || p.inletRandom > PDFB200_INLET_RANDOM_SAMPLING) throw std::invalid_argument("Unknown mcInletRandom type");
    if (p.localTimeStepping < 0 || p.localTimeStepping > PDFB200_LTS_PARTICLE) throw std::invalid_argument("Unknown mcLocalTimeStepping type");
    if (p.reactionModel == PDFB200_REACTION_KULKARNI) {
        if (p.cIdx < 0 || p.cIdx >= p.nPhi || p.pvIdx < 0 || p.pvIdx >= p.nPhi) throw std::invalid_argument("KulkarniAutoIgnition: cIdx / pvIdx out of range");
        if (p.nAdded < 0 || p.nAdded > PDFB200_MAX_ADDED) throw std::invalid_argument("KulkarniAutoIgnition: too many added scalars");
        for (int i = 0; i < p.nAdded; ++i) if (p.addedIdx[i] < 0 || p.addedIdx[i] >= p.nPhi) throw std::invalid_argument("KulkarniAutoIgnition: added scalar out of range");
    }
    if (p.reactionModel == PDFB200_REACTION_BURKE_SCHUMANN || p.reactionModel == PDFB200_REACTION_STEADY_FLAMELET ||
        p.reactionModel == PDFB200_REACTION_NAIVE_FLAMELET || p.reactionModel == PDFB200_REACTION_KULKARNI)
        if (p.zIdx < 0 || p.zIdx >= p.nPhi) throw std::invalid_argument("reaction model: zIdx out of range");
    if ((p.reactionModel == PDFB200_REACTION_BURKE_SCHUMANN || p.reactionModel == PDFB200_REACTION_NAIVE_FLAMELET) &&
        (p.TIdx < 0 || p.TIdx >= p.nPhi))
        throw std::invalid_argument("reaction model: TIdx out of range");
    if (p.reactionModel == PDFB200_REACTION_STEADY_FLAMELET) {
        if (p.chiIdx < 0 || p.chiIdx >= p.nPhi) throw std::invalid_argument("SteadyFlamelet: chiIdx out of range");
        if (p.nAdded < 0 || p.nAdded > PDFB200_MAX_ADDED) throw std::invalid_argument("SteadyFlamelet: too many added scalars");
        for (int i = 0; i < p.nAdded; ++i) if (p.addedIdx[i] < 0 || p.addedIdx[i] >= p.nPhi) throw std::invalid_argument("SteadyFlamelet: added scalar out of range");
    }
}

void Cloud::setParams(const pdfb200_params &p) {
    validateParams(p);
    if (haveCloudFields && p.nPhi != prm.nPhi) throw std::invalid_argument("set_params: nPhi cannot change after the fields were uploaded");
    prm = p;
    dropSpecialisedKernel();
    dropGraphs();
    prm.C0 = fmax(0.0, prm.C0);                    // mcSLMFullVelocityModel.C:97-98
    prm.C1 = fmax(0.0, fmin(1.0, prm.C1));
    nCov = prm.nPhi * (prm.nPhi + 1) / 2;
    nMom = 12 + prm.nPhi + nCov + 3 * prm.nPhi;   // N, M, V, MU(3), MPhi, MPhiPhi, MUU(6), MUPhi(3 nPhi)
    cellAuxStride = 2 + prm.nPhi;
}

void Cloud::create(const pdfb200_mesh &m, const pdfb200_params &p, int dev, int64_t cap) {
    int nDev = 0;
    cudaError_t e = cudaGetDeviceCount(&nDev);
    if (e != cudaSuccess || nDev == 0) throw CudaError(e == cudaSuccess ? cudaErrorNoDevice : e, "cudaGetDeviceCount (libpdfb200 has no CPU fallback)", __FILE__, __LINE__);
    if (dev < 0 || dev >= nDev) throw std::invalid_argument("create: no such CUDA device");
    if (cap <= 0) throw std::invalid_argument("create: capacity must be positive");
    if (cap > 2000000000LL) throw std::invalid_argument("create: capacity above 2e9 particles per GPU");
    device = dev;
    capacity = cap;
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    numSMs = prop.multiProcessorCount;
    CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreate(&evStart)); CUDA_CHECK(cudaEventCreate(&evStop));
    CUDA_CHECK(cudaEventCreate(&evArA)); CUDA_CHECK(cudaEventCreate(&evArB));
    for (auto &e : evMark) CUDA_CHECK(cudaEventCreate(&e));
    setParams(p);
    CUDA_CHECK(cudaMallocHost((void **)&hScal, sizeof(StepScalars)));
    CUDA_CHECK(cudaMallocHost((void **)&hFinal, 64 * sizeof(double)));
    std::memset(hScal, 0, sizeof(StepScalars));
    hScal->deltaT = p.deltaT0 > 0 ? p.deltaT0 : 100 * PDF_SMALL;
    hScal->omegaClip = 1.7976931348623157e308;
    scal.alloc(1);
    CUDA_CHECK(cudaMemcpyAsync(scal.p, hScal, sizeof(StepScalars), cudaMemcpyHostToDevice, stream));
    deltaTHost = hScal->deltaT;
    finalSums.alloc(256); finalSums.zero(stream);
    k1PartMax.alloc((size_t)numSMs * 64);
    buildMesh(m);
}

// mcSteadyFlamelet's constructor fails cleanly when its library is missing or too small (mcSteadyFlamelet.C:120-286);
// the kernels index the tables without bounds checks, so the same is demanded before anything runs
void Cloud::requireTables() const {
    if (prm.reactionModel == PDFB200_REACTION_KULKARNI) {
        if (tableKind != 2) throw std::invalid_argument("KulkarniAutoIgnition selected but no auto-ignition tables were set (pdfb200_set_autoignition_tables)");
        if (nFlFields < 2 + prm.nAdded) throw std::invalid_argument("KulkarniAutoIgnition: the library holds fewer tables than cdot + rho + the added scalars");
        return;
    }
    if (prm.reactionModel != PDFB200_REACTION_STEADY_FLAMELET) return;
    if (tableKind != 1 || nChi < 1 || nZ < 2) throw std::invalid_argument("SteadyFlamelet selected but no flamelet tables were set (pdfb200_set_flamelet_tables)");
    if (nFlFields < 1 + prm.nAdded) throw std::invalid_argument("SteadyFlamelet: the flamelet library holds fewer tables than rho + the added scalars");
}

// sanity checks of mcKulkarniAutoIgnitionReactionModel's constructor (mcKulkarniAutoIgnition.C:262-335)
void Cloud::setAutoIgnitionTables(int nZ_, int nPv_, const double *z, const double *pv, const double *cEq, int nFields, const double *const *fields) {
    if (nZ_ < 2 || nPv_ < 2) throw std::invalid_argument("auto-ignition z and pv must have at least 2 elements");
    for (int i = 1; i < nZ_; ++i) if (!(z[i] - z[i - 1] > 0)) throw std::invalid_argument("auto-ignition z not strictly monotonically increasing");
    for (int i = 1; i < nPv_; ++i) if (!(pv[i] - pv[i - 1] > 0)) throw std::invalid_argument("auto-ignition pv not strictly monotonically increasing");
    if (nFields < 2) throw std::invalid_argument("auto-ignition tables: cdot and rho are required");
    nChi = nZ_; nZ = nPv_; nFlFields = nFields; tableKind = 2;
    dropGraphs();
    cEqMax = cEq[0];
    for (int i = 1; i < nZ_; ++i) cEqMax = std::max(cEqMax, cEq[i]);
    flChi.upload(z, nZ_, stream); flZ.upload(pv, nPv_, stream); flCEq.upload(cEq, nZ_, stream);
    flFields.alloc((size_t)nFields * nZ_ * nPv_);
    for (int f = 0; f < nFields; ++f)
        CUDA_CHECK(cudaMemcpyAsync(flFields.p + (size_t)f * nZ_ * nPv_, fields[f], (size_t)nZ_ * nPv_ * sizeof(double), cudaMemcpyHostToDevice, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
}

void Cloud::setTables(int nChi_, int nZ_, const double *chi, const double *z, int nFields, const double *const *fields) {
    // sanity checks of mcSteadyFlamelet's constructor (mcSteadyFlamelet.C:222-286)
    if (nZ_ < 2) throw std::invalid_argument("flamelet z must have at least 2 elements");
    if (nChi_ < 1) throw std::invalid_argument("flamelet chi must have at least 1 element");
    for (int i = 1; i < nZ_; ++i) if (!(z[i] - z[i - 1] > 0)) throw std::invalid_argument("flamelet z not strictly monotonically increasing");
    for (int i = 1; i < nChi_; ++i) if (!(chi[i] - chi[i - 1] > 0)) throw std::invalid_argument("flamelet chi not strictly monotonically increasing");
    if (nFields < 1) throw std::invalid_argument("flamelet tables: at least rho is required");
    nChi = nChi_; nZ = nZ_; nFlFields = nFields; tableKind = 1;
    dropGraphs();
    flChi.upload(chi, nChi, stream); flZ.upload(z, nZ, stream);
    flFields.alloc((size_t)nFields * nChi * nZ);
    for (int f = 0; f < nFields; ++f)
        CUDA_CHECK(cudaMemcpyAsync(flFields.p + (size_t)f * nChi * nZ, fields[f], (size_t)nChi * nZ * sizeof(double), cudaMemcpyHostToDevice, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
}

#define H(c) (&(c)->c)
#define USE_DEVICE(c) cudaSetDevice((c)->c.device)

extern "C" {

const char *pdfb200_last_error(void) { return g_err.c_str(); }

int pdfb200_create(const pdfb200_mesh *mesh, const pdfb200_params *params, int device, int64_t capacity, pdfb200_cloud **out) {
    if (!mesh || !params || !out) { g_err = "create: null argument"; return PDFB200_ERR_INVALID; }
    pdfb200_cloud *h = new (std::nothrow) pdfb200_cloud();
    if (!h) { g_err = "out of host memory"; return PDFB200_ERR_CAPACITY; }
    const int rc = guarded([&] { h->c.create(*mesh, *params, device, capacity); });
    if (rc != PDFB200_OK) { delete h; return rc; }
    *out = h;
    return PDFB200_OK;
}

void pdfb200_destroy(pdfb200_cloud *c) { if (c) { USE_DEVICE(c); delete c; } }

int pdfb200_set_params(pdfb200_cloud *c, const pdfb200_params *p) { USE_DEVICE(c); return guarded([&] { H(c)->setParams(*p); }); }

int pdfb200_set_flamelet_tables(pdfb200_cloud *c, int32_t nChi, int32_t nZ, const double *chi, const double *z, int32_t nFields,
                                const double *const *fields) {
    USE_DEVICE(c);
    return guarded([&] { H(c)->setTables(nChi, nZ, chi, z, nFields, fields); });
}

int pdfb200_set_autoignition_tables(pdfb200_cloud *c, int32_t nZ, int32_t nPv, const double *z, const double *pv, const double *cEq,
                                    int32_t nFields, const double *const *fields) {
    USE_DEVICE(c);
    return guarded([&] { H(c)->setAutoIgnitionTables(nZ, nPv, z, pv, cEq, nFields, fields); });
}

int pdfb200_upload_fv_fields(pdfb200_cloud *c, const pdfb200_fv_fields *f) { USE_DEVICE(c); return guarded([&] { H(c)->uploadFv(*f); }); }
int pdfb200_stage_fv_fields(pdfb200_cloud *c, const pdfb200_fv_fields *f) { USE_DEVICE(c); return guarded([&] { H(c)->stageFv(*f); }); }
int pdfb200_commit_fv_fields(pdfb200_cloud *c) { USE_DEVICE(c); return guarded([&] { H(c)->commitFv(); }); }
int pdfb200_upload_cloud_fields(pdfb200_cloud *c, const pdfb200_cloud_fields *f) { USE_DEVICE(c); return guarded([&] { H(c)->uploadCloudFields(*f); }); }
int pdfb200_upload_poscorr_fields(pdfb200_cloud *c, const pdfb200_poscorr_fields *f) {
    if (!f) { g_err = "upload_poscorr_fields: null argument"; return PDFB200_ERR_INVALID; }
    USE_DEVICE(c);
    return guarded([&] { H(c)->uploadPoscorr(*f); });
}
int pdfb200_init_moments(pdfb200_cloud *c) { USE_DEVICE(c); return guarded([&] { H(c)->initMoments(); }); }
int pdfb200_seed_particles(pdfb200_cloud *c) { USE_DEVICE(c); return guarded([&] { H(c)->seedParticles(); }); }
int pdfb200_upload_particles(pdfb200_cloud *c, int64_t n, const pdfb200_particles *p) { USE_DEVICE(c); return guarded([&] { H(c)->uploadParticles(n, *p); }); }
int64_t pdfb200_num_particles(pdfb200_cloud *c) { return c->c.nParticlesHost; }
int pdfb200_download_particles(pdfb200_cloud *c, int64_t maxn, pdfb200_particles *out, int64_t *n) {
    USE_DEVICE(c);
    return guarded([&] { H(c)->downloadParticles(maxn, *out, n); });
}
int pdfb200_evolve(pdfb200_cloud *c, int32_t nSteps, pdfb200_step_report *reports) { USE_DEVICE(c); return guarded([&] { H(c)->evolve(nSteps, reports); }); }
int pdfb200_download_cell_fields(pdfb200_cloud *c, pdfb200_cell_fields *o) { USE_DEVICE(c); return guarded([&] { H(c)->downloadCellFields(*o); }); }
int pdfb200_download_cell_fields_async(pdfb200_cloud *c, const pdfb200_cell_fields *o) {
    if (!o) { g_err = "download_cell_fields_async: null argument"; return PDFB200_ERR_INVALID; }
    USE_DEVICE(c);
    return guarded([&] { H(c)->downloadCellFieldsAsync(*o); });
}
int pdfb200_sync_downloads(pdfb200_cloud *c) { USE_DEVICE(c); return guarded([&] { H(c)->syncDownloads(); }); }
int pdfb200_upload_moments(pdfb200_cloud *c, const pdfb200_cell_fields *m) {
    if (!m) { g_err = "upload_moments: null argument"; return PDFB200_ERR_INVALID; }
    USE_DEVICE(c);
    return guarded([&] { H(c)->uploadMoments(*m); });
}
int pdfb200_set_graph_mode(pdfb200_cloud *c, int mode) { c->c.graphMode = mode; return PDFB200_OK; }
int64_t pdfb200_id_counter(pdfb200_cloud *c) {
    USE_DEVICE(c);
    unsigned long long v = 0;
    if (cudaMemcpy(&v, &c->c.scal.p->nextId, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)v;
}
int pdfb200_set_id_counter(pdfb200_cloud *c, int64_t id) {
    USE_DEVICE(c);
    return guarded([&] {
        Cloud &C = c->c;
        if (id < 0) throw std::invalid_argument("set_id_counter: out of range");
        if (id % C.nRanks != C.rank) throw std::invalid_argument("set_id_counter: the id must be congruent to the rank modulo the number of ranks");
        const unsigned long long v = (unsigned long long)id;
        C.hScal->nextId = v;
        CUDA_CHECK(cudaMemcpy(&C.scal.p->nextId, &v, sizeof(v), cudaMemcpyHostToDevice));
    });
}
int64_t pdfb200_step_counter(pdfb200_cloud *c) { return (int64_t)c->c.stepHost; }
int pdfb200_set_step_counter(pdfb200_cloud *c, int64_t step) {
    USE_DEVICE(c);
    return guarded([&] {
        Cloud &C = c->c;
        if (step < 0 || step > 0xffffffffLL) throw std::invalid_argument("set_step_counter: out of range");
        const uint32_t s32 = (uint32_t)step;
        C.stepHost = s32;
        C.hScal->step = s32;   // seed/upload push the whole mirror back to the device
        CUDA_CHECK(cudaMemcpy(&C.scal.p->step, &s32, sizeof(uint32_t), cudaMemcpyHostToDevice));
    });
}

double pdfb200_delta_t(pdfb200_cloud *c) { return c->c.deltaTHost; }
int pdfb200_set_delta_t(pdfb200_cloud *c, double dt) {
    USE_DEVICE(c);
    return guarded([&] {
        Cloud &C = c->c;
        if (!(dt > 0)) throw std::invalid_argument("set_delta_t: the time step must be positive");
        C.deltaTHost = dt;
        C.hScal->deltaT = dt;   // seed/upload push the whole mirror back to the device
        CUDA_CHECK(cudaMemcpy(&C.scal.p->deltaT, &dt, sizeof(double), cudaMemcpyHostToDevice));
    });
}

int pdfb200_comm_init(pdfb200_cloud *c, int32_t nRanks, int32_t rank, const char id[128]) {
    USE_DEVICE(c);
    return guarded([&] { H(c)->commInit(nRanks, rank, id); });
}

int pdfb200_mesh_info_get(pdfb200_cloud *c, pdfb200_mesh_info *info) {
    const Cloud &C = c->c;
    info->nTets = C.nTets; info->maxTetsPerCell = C.maxTetsPerCell; info->isAxiSymmetric = C.axisym;
    for (int k = 0; k < 3; ++k) { info->axis[k] = C.axis[k]; info->centrePlaneNormal[k] = C.centreNormal[k]; }
    info->openingAngle = C.openingAngle;
    return PDFB200_OK;
}

int pdfb200_download_tets(pdfb200_cloud *c, double *gradN, int32_t *nbr, int32_t *faceLabel, int32_t *ptB, int32_t *ptC, int32_t *cell) {
    USE_DEVICE(c);
    return guarded([&] {
        Cloud &C = c->c;
        const size_t n = (size_t)C.nTets;
        std::vector<TetA> hA(n);
        std::vector<D4> hB(n), hC(n);
        std::vector<D4> hP(n);
        CUDA_CHECK(cudaMemcpy(hA.data(), C.tetA.p, n * sizeof(TetA), cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(hB.data(), C.tetB.p, n * sizeof(D4), cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(hC.data(), C.tetC.p, n * sizeof(D4), cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(hP.data(), C.tetD.p, n * sizeof(D4), cudaMemcpyDeviceToHost));
        if (cell) CUDA_CHECK(cudaMemcpy(cell, C.tetCell.p, n * sizeof(int), cudaMemcpyDeviceToHost));
        for (size_t t = 0; t < n; ++t) {
            if (gradN) {
                double *g = gradN + 9 * t;
                g[0] = hA[t].g0; g[1] = hB[t].x; g[2] = hB[t].y; g[3] = hB[t].z; g[4] = hB[t].w;
                g[5] = hC[t].x; g[6] = hC[t].y; g[7] = hC[t].z; g[8] = hC[t].w;
            }
            if (nbr) for (int k = 0; k < 4; ++k) nbr[4 * t + k] = hA[t].nbr[k];
            if (faceLabel) faceLabel[t] = hA[t].face;
            if (ptB) ptB[t] = tetPtsOf(hP[t]).x;
            if (ptC) ptC[t] = tetPtsOf(hP[t]).y;
        }
    });
}

int pdfb200_download_geometry(pdfb200_cloud *c, double *fc, double *fa, double *cc, double *cv, double *co, double *hn) {
    USE_DEVICE(c);
    return guarded([&] {
        Cloud &C = c->c;
        const size_t nf = C.nFaces, nc = C.nCells;
        if (fc) CUDA_CHECK(cudaMemcpy(fc, C.faceCentres.p, nf * 3 * sizeof(double), cudaMemcpyDeviceToHost));
        if (fa) CUDA_CHECK(cudaMemcpy(fa, C.faceAreas.p, nf * 3 * sizeof(double), cudaMemcpyDeviceToHost));
        if (co) CUDA_CHECK(cudaMemcpy(co, C.courantCoeffs.p, nf * 3 * sizeof(double), cudaMemcpyDeviceToHost));
        if (cv) CUDA_CHECK(cudaMemcpy(cv, C.cellVolumes.p, nc * sizeof(double), cudaMemcpyDeviceToHost));
        if (cc || hn) {
            std::vector<double4> h(nc);
            CUDA_CHECK(cudaMemcpy(h.data(), C.cellCentre4.p, nc * sizeof(double4), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < nc; ++i) {
                if (cc) { cc[3 * i] = h[i].x; cc[3 * i + 1] = h[i].y; cc[3 * i + 2] = h[i].z; }
                if (hn) hn[i] = h[i].w;
            }
        }
    });
}

int64_t pdfb200_kernel_launches(pdfb200_cloud *c) { return c->c.launches; }

int pdfb200_profile_ctas(pdfb200_cloud *c, int64_t maxCtas, int64_t *out, int64_t *nCtas) {
    USE_DEVICE(c);
    return guarded([&] {
        Cloud &C = c->c;
        if (!C.ctaClock.n) {                 // first call switches the hook on: the next evolve() records
            C.ctaClock.alloc((size_t)3 * C.numSMs * 16); C.ctaClock.zero(C.stream);
            CUDA_CHECK(cudaStreamSynchronize(C.stream));
            C.dropGraphs();
            *nCtas = 0;
            return;
        }
        const int64_t n = std::min<int64_t>(maxCtas, C.ctaClockGrid);
        CUDA_CHECK(cudaMemcpy(out, C.ctaClock.p, (size_t)n * 3 * sizeof(long long), cudaMemcpyDeviceToHost));
        *nCtas = n;
    });
}
int pdfb200_profile(pdfb200_cloud *c, double out[8], int reset) {
    Cloud &C = c->c;
    for (int r = 0; r < 6; ++r) out[r] = C.profMs[r];
    out[6] = (double)C.profSteps; out[7] = C.profParticleSteps;
    if (reset) { for (double &v : C.profMs) v = 0; C.profSteps = 0; C.profParticleSteps = 0; }
    return PDFB200_OK;
}
double pdfb200_last_evolve_ms(pdfb200_cloud *c) { return c->c.lastEvolveMs; }

}  // extern "C"
