// cloud.hpp — host-side state of one device cloud (libpdfb200).
#pragma once

#include <string>
#include <vector>

#include "common.cuh"

template <class T>
struct DBuf {
    T *p = nullptr;
    size_t n = 0;
    void alloc(size_t count) {
        release();
        n = count;
        if (count) CUDA_CHECK(cudaMalloc((void **)&p, count * sizeof(T)));
    }
    void zero(cudaStream_t s) { if (n) CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    void upload(const T *h, size_t count, cudaStream_t s) {
        if (count > n) alloc(count);
        if (count) CUDA_CHECK(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void download(T *h, size_t count, cudaStream_t s) const {
        if (count) CUDA_CHECK(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DBuf() { release(); }
    DBuf() = default;
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
};

struct HostPatch { int start, size, geom, handler, reflecting; };

// number of doubles in the trailing global section of the moment block
#define BLOCK_TAIL 32
// steps issued without a host synchronisation in between (their reports wait in a ring on the device)
#define REPORT_RING 64

#include "k1_layout.h"

struct NcclApi;   // dlopen'ed NCCL entry points (comm.cu)

struct Cloud {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t evStart = nullptr, evStop = nullptr, evArA = nullptr, evArB = nullptr;
    int64_t capacity = 0;
    pdfb200_params prm{};
    int nMom = 0, nCov = 0, cellAuxStride = 0;
    long long launches = 0;
    double lastEvolveMs = 0;
    // per-region device time of evolve(), accumulated from CUDA events on `stream`
    cudaEvent_t evMark[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double profMs[6] = {0, 0, 0, 0, 0, 0};
    double profParticleSteps = 0;
    long long profSteps = 0;
    int numSMs = 148;

    // ---- host copies of the small mesh tables ----
    int nCells = 0, nFaces = 0, nInternalFaces = 0, nPoints = 0, nBnd = 0;
    long long nTets = 0;
    int maxTetsPerCell = 0, tetBits = 0, keyBits = 0;
    std::vector<HostPatch> patches;
    std::vector<int> hOwner, hFaceStart, hFacePoints;   // needed for inlet tables
    bool axisym = false;
    double axis[3] = {0, 0, 0}, centreNormal[3] = {0, 0, 0}, openingAngle = 0;
    bool hasOpen = false, hasInlet = false;

    // ---- device mesh ----
    DMesh dm{};
    DBuf<double> points, faceCentres, faceAreas, magSf, cellVolumes, volumeOrArea, bbMin, bbMax, courantCoeffs,
        linWeights, pointWeights, patchFaceT;
    DBuf<double4> cellCentre4, bfNormal;
    DBuf<int> faceStart, facePoints, owner, neighbour, cellFaceStart, cellFaces, cellTetStart, faceTetStartOwn,
        faceTetStartNei, pointCellStart, pointCells, bfInfo, openCellStart, openCellFaces, cellRank, cellOfRank;
    DBuf<TetA> tetA;
    DBuf<D4> tetB, tetC, tetD, cellCo4;
    DBuf<int> tetCell;

    // ---- fields ----
    // FV-owned
    DBuf<double> Ufv_c, Ufv_b, p_c, p_b, kfv_c, kfv_b, eps_c, eps_b, R_c, R_b;
    DBuf<uint8_t> UbFixed, kbFixed, rhoFixed;
    // second set of the FV-owned fields: target of pdfb200_stage_fv_fields, swapped in by pdfb200_commit_fv_fields
    DBuf<double> stg_U_c, stg_U_b, stg_p_c, stg_p_b, stg_k_c, stg_k_b, stg_eps_c, stg_eps_b, stg_R_c, stg_R_b;
    DBuf<uint8_t> stg_UbFixed, stg_kbFixed;
    cudaStream_t copyStream = nullptr;
    cudaEvent_t evStaged = nullptr;
    // asynchronous read-back (pdfb200_download_cell_fields_async): own stream, ordered behind the library's stream by
    // evStepDone; the step that overwrites the cell fields waits for evDownloaded
    cudaStream_t dlStream = nullptr;
    cudaEvent_t evStepDone = nullptr, evDownloaded = nullptr;
    bool downloadPending = false;
    // dynamic work distribution of the particle kernel: claim counter, per-chunk mass sums (evolve_kernel.cuh)
    DBuf<int> k1WorkCtr;
    DBuf<double> chunkSums, chunkPart;
    int k1Dynamic = K1_DYNAMIC;   // 0 static waves, 1 chunks claimed per warp
    int k1Chunk = K1_CHUNK;       // entries per row of chunkSums
    // pdfb200_profile_ctas: per-CTA start/end times of the last particle-kernel launch (allocated on first use)
    DBuf<long long> ctaClock;
    int ctaClockGrid = 0;
    bool stagedPending = false;
    // cloud fields (cell + boundary)
    DBuf<double> rho_c, rho_b, rhoInst_c, pnd_c, pndInst_c, Updf_c, Updf_b, Tau_c, Tau_b, kpdf_c, kpdf_b;
    DBuf<double> Phi_c, Phi_b, Cov_c, Cov_b;       // [nPhi][nCells] etc.
    DBuf<uint8_t> PhiFixed, CovFixed;              // [nPhi][nBnd]
    DBuf<double> PaNIC, lostMass, lostShare;
    DBuf<LostRec> lostList, lostSorted;           // particles lost in the current step (k_lost_reduce)
    DBuf<unsigned> lostCount;
    bool lostDirty = false;                        // lostMass holds non-zero entries from the last step
    // time-averaged moments
    DBuf<double> mMom, VMom, UMom, UUMom, PhiMom, PhiPhiMom, UPhiMom, UPhi_c;
    DBuf<double> momentBlock;                      // [nCells*nMom + BLOCK_TAIL]
    // step work arrays
    DBuf<double> phiPC, UPC_c, OmegaRaw_c, etaRaw_c;
    DBuf<double> nodeMid, nodePC, cellAux;
    // externally solved position correction: rows of 12 = (G0, l, G1, zeta, G2, 0) per cell / boundary face
    DBuf<double> pcExt_c, pcExt_b;
    double pcScale = 0;
    bool havePcExt = false;
    int pcPlanes = 1;                 // 32-byte planes per node of nodePC (3 with external fields)
    DBuf<double> Un, inletU, inletR, fwdTrans, probVec;
    bool inletCached = false;
    // flamelet tables
    DBuf<double> flChi, flZ, flFields, flCEq;     // Kulkarni: flChi = z, flZ = pv, tables [nZ][nPv], flCEq = cEq[nZ]
    int nChi = 0, nZ = 0, nFlFields = 0;
    double cEqMax = 0;
    int tableKind = 0;                // 0 none, 1 flamelet (chi, z), 2 auto-ignition (z, pv)

    // ---- particles ----
    PState S[2]{};
    std::vector<void *> stateAllocs;
    int cur = 0;
    DBuf<uint32_t> keysA, keysB, valsA, valsB, iota;
    DBuf<int> injPatchTail;
    DBuf<unsigned char> cubTemp;
    DBuf<int> cellStart, cellEnd, ncFlag, ncCount, ncOffset, scanTmp, scanTileSums, scanTileOffs;
    DBuf<int> injCount, injOffset, csCount, csCursor, csStart, ncCounts;
    DBuf<double> injSums;
    DBuf<double> k1PartSum, k1PartMax, cellPartSum, finalSums;   // deterministic reductions
    DBuf<double> ncDelta, cloneDelta;   // cloneDelta: [capacity][K1_NC1], axisymmetric cases only
    DBuf<StepScalars> scal;
    StepScalars *hScal = nullptr;     // pinned mirror
    DBuf<double> reportRing;          // REPORT_RING step reports of 64 doubles (k_report_out)
    double *hRing = nullptr;          // pinned copy of the ring
    std::vector<cudaEvent_t> evMarks; // 7 per step of a batch: the regions of pdfb200_profile
    void *stepGraph[2] = {nullptr, nullptr};        // cudaGraphExec_t of one step, per parity of the state double buffer
    long long stepGraphLaunches[2] = {0, 0};
    bool radixPath = false;
    int graphMode = -1;               // pdfb200_set_graph_mode
    double *hFinal = nullptr;         // pinned reduction results
    int64_t nParticlesHost = 0;       // alive particles after the last step
    int64_t nSlotsHost = 0;           // slots used in the current buffer (upper bound for launches)
    bool haveFv = false, haveCloudFields = false, haveMoments = false, sortedValid = false;
    uint32_t stepHost = 0;
    double deltaTHost = 1e-13;

    // ---- run-time specialised particle kernel (jit.cu) ----
    int jitState = 0;                 // 0 not tried, 1 loaded, -1 not available
    void *jitLibrary = nullptr, *jitKernel = nullptr;
    void *specialisedKernel();        // nullptr -> use the ahead-of-time generic kernel
    void dropSpecialisedKernel();     // parameters or mesh flags changed

    // ---- multi-GPU ----
    int nRanks = 1, rank = 0;
    NcclApi *nccl = nullptr;
    void *comm = nullptr;

    ~Cloud();
    void create(const pdfb200_mesh &m, const pdfb200_params &p, int device, int64_t capacity);
    void buildMesh(const pdfb200_mesh &m);            // mesh_build.cu
    void chooseSortKey();                             // mesh_build.cu
    void setParams(const pdfb200_params &p);
    void requireTables() const;
    void setAutoIgnitionTables(int nZ, int nPv, const double *z, const double *pv, const double *cEq, int nFields, const double *const *fields);
    void setTables(int nChi, int nZ, const double *chi, const double *z, int nFields, const double *const *fields);
    void uploadFv(const pdfb200_fv_fields &f);
    void stageFv(const pdfb200_fv_fields &f);
    void commitFv();
    void downloadCellFieldsAsync(const pdfb200_cell_fields &o);
    void syncDownloads();
    void orderAfterDownloads();   // makes the library's stream wait for pending asynchronous downloads
    void uploadCloudFields(const pdfb200_cloud_fields &f);
    void uploadPoscorr(const pdfb200_poscorr_fields &f);
    void initMoments();
    void uploadMoments(const pdfb200_cell_fields &m);
    void allocParticles();
    void seedParticles();
    void uploadParticles(int64_t n, const pdfb200_particles &p);
    void downloadParticles(int64_t maxn, pdfb200_particles &out, int64_t *n);
    void evolve(int nSteps, pdfb200_step_report *reports);
    void enqueueStep(void *k1Jit, size_t k1JitSmem, int k1Blocks, cudaEvent_t *ev);
    void dropGraphs();
    void downloadCellFields(pdfb200_cell_fields &o);
    void commInit(int nRanks, int rank, const char id[128]);

    // step pieces (fields.cu / particles.cu)
    void prepareFields();
    void stepParticles(pdfb200_step_report *rep);
    void sortAndMoments();
    void finaliseStep(pdfb200_step_report *rep);
    void numberControl();
    void inject();
    void syncScalars();
    DMesh &mesh() { return dm; }
};

// launch helper counting this library's kernel launches
#define LAUNCH(cloud, kernel, grid, block, smem, ...)                                      \
    do {                                                                                   \
        kernel<<<(grid), (block), (smem), (cloud)->stream>>>(__VA_ARGS__);                  \
        ++(cloud)->launches;                                                               \
        CUDA_CHECK(cudaGetLastError());                                                    \
    } while (0)

static inline int gridFor(long long n, int block, int cap = 1 << 30) {
    long long g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}
