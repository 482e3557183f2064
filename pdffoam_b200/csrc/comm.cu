// comm.cu — multi-GPU: particles are sharded by rank, mesh/fields/tables are
// replicated, and the per-cell instantaneous moment block is summed across
// ranks once per step over NVLink (ncclAllReduce). This replaces OpenFOAM's
// processor patches, Cloud<T>::move transfers and sendToOrigProc
// (mcParticle/mcParticleCloud/mcParticleCloud.C:103-231) and the reduce() calls
// of evolve (:1451-1458).
//
// NCCL is bound at run time with dlopen so that the library shares the copy
// already loaded in the process (e.g. the one PyTorch ships) instead of
// bringing a second one.
#include <dlfcn.h>

#include <string>

#include "cloud.hpp"

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void *ncclComm_p;
enum { NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2 };

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
    int (*CommInitRank)(ncclComm_p *, int, ncclUniqueId_t, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclComm_p) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
};

static NcclApi *loadNccl() {
    static NcclApi api;
    if (api.handle) return &api;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // already in the process?
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) throw std::runtime_error(std::string("NCCL not found: ") + dlerror());
    api.handle = h;
    auto sym = [&](const char *n) { void *p = dlsym(h, n); if (!p) throw std::runtime_error(std::string("NCCL symbol missing: ") + n); return p; };
    api.GetUniqueId = (int (*)(ncclUniqueId_t *))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(ncclComm_p *, int, ncclUniqueId_t, int))sym("ncclCommInitRank");
    api.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_p, cudaStream_t))sym("ncclAllReduce");
    api.CommDestroy = (int (*)(ncclComm_p))sym("ncclCommDestroy");
    api.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    return &api;
}

#define NCCL_CHECK(api, expr)                                                                       \
    do {                                                                                            \
        int _r = (expr);                                                                            \
        if (_r != 0) throw std::runtime_error(std::string("NCCL: ") + (api)->GetErrorString(_r));   \
    } while (0)

void Cloud::commInit(int nRanks_, int rank_, const char id[128]) {
    if (nRanks_ < 1 || rank_ < 0 || rank_ >= nRanks_) throw std::invalid_argument("comm_init: bad rank/size");
    if (nParticlesHost > 0) throw std::logic_error("comm_init: call before seeding/uploading particles");
    nRanks = nRanks_; rank = rank_;
    chooseSortKey();   // the particles per cell held by this rank decide the key
    dropSpecialisedKernel();
    if (nRanks == 1) return;
    nccl = loadNccl();
    ncclUniqueId_t uid;
    memcpy(uid.internal, id, 128);
    ncclComm_p cm = nullptr;
    NCCL_CHECK(nccl, nccl->CommInitRank(&cm, nRanks, uid, rank));
    comm = cm;
}

// sums the moment block and the particle sums, maxes the particle maxima
void commAllreduceMoments(Cloud &c, float *ms) {
    NcclApi *api = c.nccl;
    if (!api || !c.comm) throw std::logic_error("allreduce: communicator not initialised");
    double *rep = c.finalSums.p + 128;
    CUDA_CHECK(cudaEventRecord(c.evArA, c.stream));
    NCCL_CHECK(api, api->GroupStart());
    NCCL_CHECK(api, api->AllReduce(c.momentBlock.p, c.momentBlock.p, (size_t)c.nCells * c.nMom, NCCL_FLOAT64, NCCL_SUM, c.comm, c.stream));
    NCCL_CHECK(api, api->AllReduce(rep, rep, ACC_NSUM, NCCL_FLOAT64, NCCL_SUM, c.comm, c.stream));
    NCCL_CHECK(api, api->AllReduce(rep + ACC_NSUM, rep + ACC_NSUM, ACC_NMAX, NCCL_FLOAT64, NCCL_MAX, c.comm, c.stream));
    NCCL_CHECK(api, api->GroupEnd());
    CUDA_CHECK(cudaEventRecord(c.evArB, c.stream));
    *ms = -1.f;   // resolved by the caller after the stream is synchronised
}

void commDestroy(Cloud &c) {
    if (c.comm && c.nccl) c.nccl->CommDestroy(c.comm);
    c.comm = nullptr;
}

extern "C" int pdfb200_comm_unique_id(char id[128]) {
    try {
        NcclApi *api = loadNccl();
        ncclUniqueId_t uid;
        const int r = api->GetUniqueId(&uid);
        if (r != 0) return PDFB200_ERR_COMM;
        memcpy(id, uid.internal, 128);
        return PDFB200_OK;
    } catch (const std::exception &) {
        return PDFB200_ERR_COMM;
    }
}
