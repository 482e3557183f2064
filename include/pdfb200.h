/*
 * pdfb200.h — C ABI of the B200-native particle step (mcParticleCloud::evolve).
 *
 * This header is the drop-in boundary. pdfFoam has no FFI of its own: its
 * plugin boundary is C++ inside the OpenFOAM process. Every
 * entry point below therefore names the reference interface it replaces
 * (paths relative to the pdfFoam source tree), and INTEGRATION.md shows the
 * OpenFOAM-side adaptor that binds them.
 *
 * Conventions: plain pointers and sizes only, caller-owned HOST buffers,
 * int status returns (0 = ok) with pdfb200_last_error() for the message,
 * no exceptions cross the ABI, one handle per cloud, a handle is not
 * thread-safe. All floating point data is fp64 and all labels are int32,
 * as in the reference (scalar = double, label = int).
 *
 * The same declarations with the prefix `pdforacle_` are exported by the CPU
 * oracle (oracle/), which is test infrastructure only.
 */
#ifndef PDFB200_H
#define PDFB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDFB200_MAX_PHI 6      /* max particle scalars (scalarFields)          */
#define PDFB200_MAX_COV 21     /* MAX_PHI*(MAX_PHI+1)/2                         */
#define PDFB200_MAX_ADDED 4    /* max flamelet "added" scalars                 */

/* status codes */
enum {
    PDFB200_OK = 0,
    PDFB200_ERR_INVALID = 1,   /* bad argument / inconsistent input           */
    PDFB200_ERR_CUDA = 2,      /* CUDA runtime failure                        */
    PDFB200_ERR_CAPACITY = 3,  /* particle capacity exceeded                  */
    PDFB200_ERR_STATE = 4,     /* call order violated (e.g. evolve before fields) */
    PDFB200_ERR_COMM = 5       /* NCCL failure                                */
};

/* polyPatch geometry class (what OpenFOAM's isA<emptyPolyPatch>/<wedgePolyPatch>
 * tests see; mcParticle/mcParticle/mcParticle.C:211-235, mcParticleCloud.C:661-717) */
enum { PDFB200_PATCH_GENERIC = 0, PDFB200_PATCH_EMPTY = 1, PDFB200_PATCH_WEDGE = 2 };

/* mcBoundary run-time selection names (mcBoundary/mcBoundary/mcBoundary.C:42-82):
 * "slip", "empty", "inletOutlet" (mcInletOutletBoundary, injects),
 * open = mcOpenBoundary behaviour without injection. */
enum {
    PDFB200_BC_SLIP = 0,
    PDFB200_BC_EMPTY = 1,
    PDFB200_BC_OPEN = 2,
    PDFB200_BC_INLET_OUTLET = 3
};

/* model selectors; the words are the reference's run-time selection names */
enum { PDFB200_VELOCITY_NONE = 0, PDFB200_VELOCITY_SLMFULL = 1 };      /* "SLMFull" */
/* "IEM" (mcIEMMixingModel.C:68-83); "modifiedCurl" is new (BASELINE north_star, SURVEY
 * Appendix C; not in the reference): after the step's sort the particles of a cell are matched
 * in disjoint random pairs, a pair (p,q) mixes with probability
 * 3*Cmix*Omega_pq*eta_pq*deltaT*(w_p+w_q)/(2*wbar) (w = eta*m; several rounds when > 1) and
 * both move a fraction a ~ U(0,1) towards the pair's weighted mean. */
enum { PDFB200_MIXING_NONE = 0, PDFB200_MIXING_IEM = 1, PDFB200_MIXING_MODCURL = 2 };
enum { PDFB200_OMEGA_NONE = 0, PDFB200_OMEGA_RAS = 1 };                /* "RAS"     */
enum {
    PDFB200_REACTION_NONE = 0,
    PDFB200_REACTION_COLD = 1,            /* "cold"           */
    PDFB200_REACTION_BURKE_SCHUMANN = 2,  /* "BurkeSchumann"  */
    PDFB200_REACTION_STEADY_FLAMELET = 3, /* "SteadyFlamelet" */
    PDFB200_REACTION_NAIVE_FLAMELET = 4,  /* "naiveFlamelet" (mcNaiveFlameletModel.C:64-69): rho = 1 - 3.2 z (1 - z), T = 1e5 / rho */
    PDFB200_REACTION_KULKARNI = 5         /* "KulkarniAutoIgnition" (mcKulkarniAutoIgnition.C:339-380): progress variable c integrated
                                             with a tabulated source cdot(z, pv), rho and the added scalars looked up on (z, pv) */
};
/* mcPositionCorrection run-time selection (mcPositionCorrection/mcPositionCorrection.C): "limitedSimple" and
 * "simple" (mcSimplePositionCorrection.C:127-163) are algebraic and run entirely on the device. "external" is the
 * boundary for the three variants whose updateInternals() solves a PDE with OpenFOAM's linear solvers
 * ("ellipticRelaxation", "Muradoglu", "integrated"): the OpenFOAM side keeps solving, uploads the resulting cell
 * fields with pdfb200_upload_poscorr_fields and the device does the per-particle part (their correct(mcParticle&)). */
enum { PDFB200_POSCORR_NONE = 0, PDFB200_POSCORR_LIMITED_SIMPLE = 1, PDFB200_POSCORR_SIMPLE = 2, PDFB200_POSCORR_EXTERNAL = 3 };
enum { PDFB200_LTS_NONE = 0, PDFB200_LTS_CELL = 1, PDFB200_LTS_PARTICLE = 2 }; /* "cell","particle" */
enum { PDFB200_INTERP_CELL = 0, PDFB200_INTERP_CELL_POINT_FACE = 1 };  /* "cell","cellPointFace" */
/* mcInletRandom run-time selection (mcInletRandom.C:66-90): "inversion" (Newton inversion of the CDF,
 * mcInversionInletRandom.C:63-119) and "sampling" (acceptance-rejection, mcSamplingInletRandom.C:37-102: normal
 * deviates U+u'N accepted with probability proportional to their value). Both draw from x exp(-(x-U)^2 b^2)/Q. */
enum { PDFB200_INLET_RANDOM_INVERSION = 0, PDFB200_INLET_RANDOM_SAMPLING = 1 };

/* ---- mesh: the polyMesh the cloud borrows (mcParticleCloud.C:301-310) ---- */
typedef struct pdfb200_mesh {
    int32_t nPoints, nFaces, nInternalFaces, nCells, nPatches;
    const double  *points;       /* [nPoints*3]                               */
    const int32_t *faceStart;    /* [nFaces+1] CSR offsets into facePoints    */
    const int32_t *facePoints;   /* [faceStart[nFaces]] point labels per face */
    const int32_t *owner;        /* [nFaces]                                  */
    const int32_t *neighbour;    /* [nInternalFaces]                          */
    const int32_t *patchStart;   /* [nPatches] first face of each patch       */
    const int32_t *patchSize;    /* [nPatches]                                */
    const int32_t *patchGeom;    /* [nPatches] PDFB200_PATCH_*                */
    const int32_t *patchHandler; /* [nPatches] PDFB200_BC_* (boundaryHandlers{} dict) */
    const int32_t *patchReflecting; /* [nPatches] mcOpenBoundary "reflecting" switch */
    int32_t geometricD[3];       /* polyMesh::geometricD(): +1 valid, -1 empty/wedge */
    int32_t solutionD[3];        /* polyMesh::solutionD():  +1 valid, -1 empty       */
} pdfb200_mesh;

/* ---- numerics + model selection: mcSolution (mcSolution/mcSolution.C:64-240)
 *      and <cloud>Properties in constant/thermophysicalProperties           ---- */
typedef struct pdfb200_params {
    /* scalar bookkeeping (mcParticleCloud.C:1704-1899) */
    int32_t nPhi;
    int32_t nMixed;      int32_t mixed[PDFB200_MAX_PHI];
    int32_t nConserved;  int32_t conserved[PDFB200_MAX_PHI];
    /* mcSolution */
    double  CFL;
    double  averagingCoeff;
    int32_t particlesPerCell;
    int32_t particleNumberControl;
    double  cloneAt, eliminateAt;
    double  kMin;
    double  DNum;
    double  relaxTimeU, relaxTimek;  /* relaxationTimes{U,k} or the default   */
    double  minRelaxTime;            /* FV deltaT * minRelaxationFactor (mcSolution.C:213-230) */
    /* interpolationSchemes{} for the fields the models interpolate           */
    int32_t interpU, interpk, interpDiffU, interpOmega, interpUPosCorr, interpZVar, interpEta;
    /* model selectors + coefficients */
    int32_t velocityModel;  double C0, C1;        /* mcSLMFullVelocityModel.C:94-99 */
    int32_t mixingModel;    double Cmix;          /* mcIEMMixingModel.C:62-65 */
    int32_t omegaModel;     double Omega0;        /* mcRASOmegaModel.C:114    */
    int32_t reactionModel;
    double  coldDensity;                          /* mcColdReactionModel.C:57 */
    double  bsRfuel, bsRox, bsRstoi, bsTfuel, bsTox, bsTstoi, bsZstoi; /* mcBurkeSchumannReactionModel.C:63-69 */
    int32_t zIdx, TIdx, chiIdx;                   /* indices into the particle scalars */
    int32_t cIdx, pvIdx;                          /* KulkarniAutoIgnition: progress variable c and its normalised form pv */
    double  Cchi;                                 /* mcSteadyFlamelet: chi = Cchi*Omega*zVar */
    int32_t nAdded; int32_t addedIdx[PDFB200_MAX_ADDED]; /* SteadyFlamelet "scalars (T)" */
    int32_t positionCorrection; double pcC, pcCFL; /* mcLimitedSimplePositionCorrection.C:90-102 */
    int32_t localTimeStepping;  double ltsUpperBound; /* mcCellLocalTimeStepping.C:84-85 */
    double  k0;              /* RASModel::k0()/kMin(): floor for R diag / inlet u_rms */
    double  deltaT0;         /* initial particle time step (reference: 100*SMALL) */
    /* RNG */
    uint64_t seed;
    int32_t  rngMode;        /* oracle only: 0 = keyed Philox (same streams as the device),
                                1 = one sequential stream consumed in list order like
                                OpenFOAM's Random (the reference's behaviour)           */
    int32_t  inletRandom;    /* PDFB200_INLET_RANDOM_* ("randomGenerator { type ...; }" of an inletOutlet handler) */
} pdfb200_params;

/* ---- FV-owned fields the cloud borrows (mcParticleCloud.C:319-331, .H:206):
 *      cell values + values on the boundary faces in mesh face order
 *      (index = face - nInternalFaces). *Fixed flags: 1 = fixedValue patch face
 *      (cloud-derived fields keep the FV value there), 0 = zeroGradient-like
 *      (boundary takes the cell value after correctBoundaryConditions()).     */
typedef struct pdfb200_fv_fields {
    const double *U,   *Ub;   const uint8_t *UbFixed;   /* [nCells*3], [nBnd*3], [nBnd] */
    const double *p,   *pb;                              /* [nCells], [nBnd]             */
    const double *k,   *kb;   const uint8_t *kbFixed;
    const double *epsilon, *epsilonb;                    /* Omega = epsilon/k; pass omega*k for kOmegaSST */
    const double *R,   *Rb;                              /* symmTensor xx,xy,xz,yy,yz,zz */
} pdfb200_fv_fields;

/* ---- fields the cloud writes into / owns: rho (owned by mcThermo), the user
 *      scalar fields and their covariances (mcParticleCloud.C:415-420,1731-1797) ---- */
typedef struct pdfb200_cloud_fields {
    const double *rho, *rhob; const uint8_t *rhobFixed;
    const double *Phi[PDFB200_MAX_PHI];  const double *Phib[PDFB200_MAX_PHI];
    const uint8_t *PhibFixed[PDFB200_MAX_PHI];
    const double *Cov[PDFB200_MAX_COV];  const double *Covb[PDFB200_MAX_COV];
    const uint8_t *CovbFixed[PDFB200_MAX_COV];           /* order: (0,0),(0,1)..(1,1).. */
} pdfb200_cloud_fields;

/* ---- particle columns in mcParticleIO.C:167-234 order (positions, m,
 *      UParticle, Omega, rho, eta, Phi...) + cell label. tet/id may be NULL on
 *      upload (tet is located, id = running index).                         ---- */
typedef struct pdfb200_particles {
    double  *pos;      /* [n*3] */
    double  *U;        /* [n*3] UParticle */
    double  *m, *Omega, *rho, *eta;   /* [n] */
    double  *Phi[PDFB200_MAX_PHI];    /* [n] each */
    int32_t *cell;     /* [n] */
    int32_t *tet;      /* [n] global tet label (tetFacePointCellDecomposition order) */
    uint64_t *id;      /* [n] persistent particle id, 64 bits (keys the random streams; Particle<>::origId) */
} pdfb200_particles;

/* ---- per-step scalars printed/returned by evolve (mcParticleCloud.C:1376-1530) ---- */
typedef struct pdfb200_step_report {
    double deltaT;          /* time step used for this step                    */
    double deltaTNext;      /* CFL/CoMax (mcParticleCloud.C:1512)              */
    double rhoRes;          /* return value: mean |pnd - rho|/rho (:1376-1381) */
    double rhoErrMin, rhoErrMax;
    double UMax, UcorrMax, etaMin, etaMax, CoMax;
    double totalMass, totalParticleMass;               /* :1432-1433 */
    double deltaMassInst[1 + PDFB200_MAX_PHI];         /* :1238-1246,1440-1444 */
    double massInInst[1 + PDFB200_MAX_PHI];
    double massOutInst[1 + PDFB200_MAX_PHI];
    int64_t nParticles;     /* alive after the step                            */
    int64_t nInjected, nDeleted, nCloned, nEliminated, nLost;
    double momentsAllreduceMs; /* time spent in the cross-GPU moment allreduce  */
} pdfb200_step_report;

/* ---- cell fields produced by updateCloudPDF (mcParticleCloud.C:824-1009);
 *      any pointer may be NULL to skip it                                  ---- */
typedef struct pdfb200_cell_fields {
    double *rho;        /* rhocPdf_ (the only field fed back to the FV side)  */
    double *rhoInst, *pnd, *pndInst;
    double *U;          /* UCloudPDF [nCells*3]                               */
    double *Tau;        /* TauCloudPDF [nCells*6]                             */
    double *k;          /* kCloudPDF                                          */
    double *Phi[PDFB200_MAX_PHI];
    double *Cov[PDFB200_MAX_COV];
    /* velocity-scalar flux <u'' phi''> = UPhiMom/mMom - U Phi, [nCells*3] per scalar. Not in the reference
     * (its moments stop at mcParticleCloud.C:933-956); BASELINE north_star asks for "scalar fluxes".        */
    double *UPhi[PDFB200_MAX_PHI];
    double *PaNIC;      /* particles per cell                                 */
    /* time-averaged moments (mMoment, VMoment, UMoment, <Phi>Moment, ...)    */
    double *mMom, *VMom, *UMom, *UUMom;
    double *PhiMom[PDFB200_MAX_PHI]; double *PhiPhiMom[PDFB200_MAX_COV];
    double *UPhiMom[PDFB200_MAX_PHI];   /* time-averaged sum of eta m U phi, [nCells*3] per scalar */
} pdfb200_cell_fields;

/* ---- fields of an externally solved position correction (PDFB200_POSCORR_EXTERNAL). Per particle
 *        Ucorrection -= interp(G0) + scale * interp(l) * (interp(zeta) != 0 ? interp(G1) : interp(G2))
 *      with the interpolation scheme interpUPosCorr, which is
 *        mcMuradogluPositionCorrection::correct (.C:158-169): G0 = grad(phi), G1 = grad(Q), G2 = grad(QInst),
 *            l = L, zeta, scale = a*U0;
 *        mcEllipticRelaxationPositionCorrection::correct (.C:238-248): the same without G0 (NULL);
 *        mcIntegratedPositionCorrection::correct (.C:132-138, Ucorrection += UPosCorr): G0 = -UPosCorr, scale = 0.
 *      Cell values [nCells*3] / [nCells] and boundary-face values [nBnd*3] / [nBnd] after
 *      correctBoundaryConditions(); a NULL pair counts as zero.                                         ---- */
typedef struct pdfb200_poscorr_fields {
    const double *G0, *G0b;
    const double *G1, *G1b;
    const double *G2, *G2b;
    const double *l, *lb;
    const double *zeta, *zetab;
    double scale;
} pdfb200_poscorr_fields;

typedef struct pdfb200_cloud pdfb200_cloud;   /* opaque handle */

/* (the types above are also read by the kernel source when the library specialises it at run time with
 * NVRTC, which accepts no host function declarations: the entry points are hidden from that compiler) */
#ifndef __CUDACC_RTC__

/* Last error message of the calling thread ("" if none). Replaces
 * FatalErrorIn(...) << exit(FatalError) (process abort) in the reference.   */
const char *pdfb200_last_error(void);

/* Construct the cloud on `device` with room for `capacity` particles.
 * Replaces mcParticleCloud::mcParticleCloud (mcParticleCloud.C:301-735):
 * builds face/cell geometry, the tetFacePointCellDecomposition
 * (tetFacePointCellDecomposition.C:31-120) with richTetrahedron gradients
 * (richTetrahedronI.H:107-134), CourantCoeffs (:526-535,649), hNum (:1937-1971)
 * and detects axisymmetry from a wedge patch pair (:661-717).              */
int pdfb200_create(const pdfb200_mesh *mesh, const pdfb200_params *params,
                   int device, int64_t capacity, pdfb200_cloud **out);
void pdfb200_destroy(pdfb200_cloud *cloud);

/* Live re-read of the numerics (mcSolution::readIfModified, mcSolution.C:186-197). */
int pdfb200_set_params(pdfb200_cloud *cloud, const pdfb200_params *params);

/* Flamelet library of mcSteadyFlamelet (mcSteadyFlamelet.C:120-286): chi[nChi],
 * z[nZ], then nFields tables, each row-major [nChi][nZ]; field 0 is rho, fields
 * 1.. are the "added" scalars in addedIdx order.                            */
int pdfb200_set_flamelet_tables(pdfb200_cloud *cloud, int32_t nChi, int32_t nZ,
                                const double *chi, const double *z,
                                int32_t nFields, const double *const *fields);

/* Library of mcKulkarniAutoIgnitionReactionModel (constant/autoIgnition/{z,pv,cEq,cdot,rho,<scalars>},
 * mcKulkarniAutoIgnition.C:111-335): z[nZ], pv[nPv], cEq[nZ], then nFields tables, each row-major [nZ][nPv]; field 0 is
 * cdot, field 1 rho, fields 2.. the "added" scalars in addedIdx order.                                              */
int pdfb200_set_autoignition_tables(pdfb200_cloud *cloud, int32_t nZ, int32_t nPv, const double *z, const double *pv,
                                    const double *cEq, int32_t nFields, const double *const *fields);

/* Upload the FV-owned fields (once per FV cycle block). */
int pdfb200_upload_fv_fields(pdfb200_cloud *cloud, const pdfb200_fv_fields *f);
/* The same in two halves, so that the copy of the next FV state overlaps the particle steps of the current one:
 * stage starts an asynchronous copy (own stream, second set of device buffers) from the caller's buffers, which
 * must be page-locked for the copy to overlap and must stay untouched until commit returns; commit waits for the
 * copy and makes the staged fields the ones evolve() reads (a pointer swap). pdfSimpleFoam changes its FV fields
 * once per block of PDF steps (tutorials/SandiaD/system/controlDict:41-43: 10 FV : 200 PDF), time-varying inflow
 * data (timeVaryingMappedFixedValue) more often; either way the particle steps need not wait for the transfer. */
int pdfb200_stage_fv_fields(pdfb200_cloud *cloud, const pdfb200_fv_fields *f);
int pdfb200_commit_fv_fields(pdfb200_cloud *cloud);
/* Upload rho / scalar / covariance fields and their boundary conditions. */
int pdfb200_upload_cloud_fields(pdfb200_cloud *cloud, const pdfb200_cloud_fields *f);
/* Upload the fields of an externally solved position correction (see pdfb200_poscorr_fields); once per solve. */
int pdfb200_upload_poscorr_fields(pdfb200_cloud *cloud, const pdfb200_poscorr_fields *f);
/* mcParticleCloud::initMoments (mcParticleCloud.C:1633-1701). */
int pdfb200_init_moments(pdfb200_cloud *cloud);

/* mcParticleCloud::initReleaseParticles (mcParticleCloud.C:1534-1563):
 * particlesPerCell particles per cell, generated on the device. In a
 * multi-rank cloud each rank seeds its share of every cell.                */
int pdfb200_seed_particles(pdfb200_cloud *cloud);
/* mcParticle::readFields (mcParticleIO.C:94-164). */
int pdfb200_upload_particles(pdfb200_cloud *cloud, int64_t n, const pdfb200_particles *p);
int64_t pdfb200_num_particles(pdfb200_cloud *cloud);
/* mcParticle::writeFields (mcParticleIO.C:167-234). Columns in the cloud's
 * current storage order; out buffers must hold maxn entries.               */
int pdfb200_download_particles(pdfb200_cloud *cloud, int64_t maxn,
                               pdfb200_particles *out, int64_t *n);

/* mcParticleCloud::evolve (mcParticleCloud.C:1226-1530), nSteps times;
 * reports[nSteps] may be NULL.                                             */
int pdfb200_evolve(pdfb200_cloud *cloud, int32_t nSteps, pdfb200_step_report *reports);

/* Cell fields of updateCloudPDF (mcParticleCloud.C:824-1009) back to the host. */
int pdfb200_download_cell_fields(pdfb200_cloud *cloud, pdfb200_cell_fields *out);
/* The same without waiting: the copies are queued on a stream of their own behind everything issued so far, into host
 * buffers that should be page-locked; pdfb200_sync_downloads() returns when they have landed. The next evolve() may be
 * called at once - the step that overwrites the fields (updateCloudPDF) waits on the device for the copies - so the
 * read-back of step n (rho for the FV side, pdfSimpleFoam/createFields.H:18) runs under step n+1.                     */
int pdfb200_download_cell_fields_async(pdfb200_cloud *cloud, const pdfb200_cell_fields *out);
int pdfb200_sync_downloads(pdfb200_cloud *cloud);
/* Restart: the time-averaged moment fields of a previous run instead of init_moments — the mMoment, VMoment,
 * UMoment, UUMoment, <Phi>Moment and <Phi><Phi>Moment fields that mcParticleCloud's constructor reads with
 * READ_IF_PRESENT (mcParticleCloud.C:339-517) and checkMoments() accepts only as a complete set (:739-820:
 * "Moments are missing. Forced re-initialization."): any of them NULL -> PDFB200_ERR_INVALID and the caller
 * falls back to pdfb200_init_moments. UPhiMom and PaNIC may be NULL (flux moments restart from the means,
 * counts from zero). The mean fields (rho, U, Tau, k, Phi, Cov, pnd) are re-derived from the moments exactly
 * as updateCloudPDF derives them (:958-1008). Call after upload_fv_fields / upload_cloud_fields.            */
int pdfb200_upload_moments(pdfb200_cloud *cloud, const pdfb200_cell_fields *moments);
/* The step counter: third word of every random-number counter (the keyed generator replaces
 * OpenFOAM's Random, whose state the reference does not checkpoint). A restarted run that sets it to the value of
 * the run it continues draws the same numbers that run would have drawn.                                     */
int64_t pdfb200_step_counter(pdfb200_cloud *cloud);
int pdfb200_set_step_counter(pdfb200_cloud *cloud, int64_t step);
/* The id the next injected or cloned particle receives (ids key the random streams; Particle<>::origId in the
 * reference). upload_particles sets it just above the largest id handed in; a restarted run that wants the
 * particles it creates to carry the ids (and so the streams) the continued run would have given them restores
 * the counter of that run. It must stay congruent to the rank modulo the number of ranks.                   */
int64_t pdfb200_id_counter(pdfb200_cloud *cloud);
int pdfb200_set_id_counter(pdfb200_cloud *cloud, int64_t nextId);

/* current particle time step (mcParticleCloud::deltaT()) */
double pdfb200_delta_t(pdfb200_cloud *cloud);
int pdfb200_set_delta_t(pdfb200_cloud *cloud, double dt);

/* ---- multi-GPU: particles sharded by rank, mesh/fields replicated, one sum
 *      all-reduce of the per-cell moment block per step. Replaces OpenFOAM's
 *      processor patches + sendToOrigProc (mcParticleCloud.C:103-231) and the
 *      reduce() calls at :1451-1458.                                       ---- */
int pdfb200_comm_unique_id(char id[128]);
int pdfb200_comm_init(pdfb200_cloud *cloud, int32_t nRanks, int32_t rank, const char id[128]);

/* ---- introspection used by the parity tests ---- */
typedef struct pdfb200_mesh_info {
    int64_t nTets;
    int32_t maxTetsPerCell;
    int32_t isAxiSymmetric;
    double  axis[3], centrePlaneNormal[3], openingAngle;
} pdfb200_mesh_info;
int pdfb200_mesh_info_get(pdfb200_cloud *cloud, pdfb200_mesh_info *info);
/* tet tables: gradN [nTets*9] (vertices a,b,c: richTetrahedron::gradNi()[0..2]),
 * nbr [nTets*4], faceLabel/ptB/ptC/cell [nTets]; any pointer may be NULL.   */
int pdfb200_download_tets(pdfb200_cloud *cloud, double *gradN, int32_t *nbr,
                          int32_t *faceLabel, int32_t *ptB, int32_t *ptC, int32_t *cell);
/* mesh geometry: faceCentres[nFaces*3], faceAreas[nFaces*3], cellCentres[nCells*3],
 * cellVolumes[nCells], courantCoeffs[nFaces*3], hNum[nCells]; NULL to skip.  */
int pdfb200_download_geometry(pdfb200_cloud *cloud, double *faceCentres, double *faceAreas,
                              double *cellCentres, double *cellVolumes,
                              double *courantCoeffs, double *hNum);
/* The keyed counter-based generator (Philox4x32-10, counter = (entity, step, purpose << 24 | sub),
 * key = seed) evaluated on `device` for n entities: out6[6*i..] = two uniforms in [0,1) and four
 * N(0,1) deviates (single-precision Box-Muller). Replaces OpenFOAM's Random (seeded at
 * mcParticleCloud.C:332); lets the tests hold the device generator to the CPU restatement bit for bit. */
int pdfb200_rng_probe(int device, uint64_t seed, int64_t n, const uint64_t *entity, uint32_t step,
                      uint32_t purpose, uint32_t sub, double *out6);
/* number of kernels of this library launched since the handle was created  */
int64_t pdfb200_kernel_launches(pdfb200_cloud *cloud);
/* accumulated device time (ms, CUDA events on the library's stream) of the regions
 * of evolve(): out[0] field preparation + injection, [1] the fused particle kernel,
 * [2] radix sort, [3] segment bounds + moment reduction, [4] moment all-reduce,
 * [5] time-averaging + number control + step end; out[6] = steps, out[7] =
 * particle-steps processed. reset != 0 clears the accumulators.              */
int pdfb200_profile(pdfb200_cloud *cloud, double out[8], int reset);
/* Load balance of the particle kernel: the first call switches a recording hook on (nCtas = 0); later calls return,
 * for the last launch of the particle kernel, three values per CTA: start and end (%globaltimer, ns) and the SM it ran
 * on. out holds 3 * maxCtas values.                                                                                 */
int pdfb200_profile_ctas(pdfb200_cloud *cloud, int64_t maxCtas, int64_t *out, int64_t *nCtas);
/* How evolve() issues its steps. A step is about fifty kernels and never waits for the device in its middle (entry
 * counts, time step and flags stay on the device; the step reports are fetched once per batch of up to 64 steps).
 * For small clouds - the tutorial cases hold 1e5 particles, a step is then bound by launch overhead - the step is
 * captured once into a CUDA graph per state-buffer parity and replayed. mode: -1 automatic (graph when the capacity is
 * at most 4e6 particles, one rank, cell-only sort key; $PDFB200_GRAPH overrides), 0 never, 1 whenever possible.
 * pdfb200_profile has no per-region times for replayed steps.                                                    */
int pdfb200_set_graph_mode(pdfb200_cloud *cloud, int mode);
/* device time (ms) of the last evolve() call, CUDA events on the library's stream */
double pdfb200_last_evolve_ms(pdfb200_cloud *cloud);
/* Run-time selection, GPU style (mcModel::New / run-time selection tables, e.g.
 * mcVelocityModel.C:54-79): at the first evolve() the particle kernel is compiled once more
 * with the cloud's parameters and mesh flags as constants (jit.cu). This
 * returns the translation unit that would be compiled for the given parameters and flags
 * (host only, no device needed; the CPU test-suite builds it); length of the text or -1. */
long pdfb200_jit_source(const pdfb200_params *params, int axisym, int hasOpen, int tetBits, int cellFacesUniform,
                        const int32_t geometricD[3], const int32_t solutionD[3], char *buf, long bufLen);

/* Compiles that translation unit exactly as the first evolve() does — kernel sources embedded in the
 * library at build time, NVRTC in memory (libnvrtc is dlopen'ed), `nvcc -cubin` in a private temporary
 * directory when NVRTC is missing — without needing a device. compiler: 0 = first available, 1 = NVRTC,
 * 2 = nvcc. Returns the size of the cubin or -1; the compiler's messages go to log.                  */
long pdfb200_jit_compile(const pdfb200_params *params, int axisym, int hasOpen, int tetBits, int cellFacesUniform,
                         const int32_t geometricD[3], const int32_t solutionD[3], int compiler, char *log, long logLen);

#endif /* !__CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* PDFB200_H */
