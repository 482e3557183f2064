/*---------------------------------------------------------------------------*\
  mcDeviceCloud.C — see mcDeviceCloud.H. Dictionary keys and defaults follow the reference:
  mcSolution (mcSolution/mcSolution.C:44-57,64-240), the model selectors of <cloud>Properties
  (mcParticleCloud.C:654-659 and the New() functions, e.g. mcVelocityModel.C:54-79), the model coefficient
  sub-dictionaries named <Type><Family>Coeffs (mcModel.C:47-60) and boundaryHandlers (mcParticleCloud.C:1902-1934).
\*---------------------------------------------------------------------------*/
#include "mcDeviceCloud.H"

#include <cstring>

// * * * * * * * * * * * * * * * Local helpers * * * * * * * * * * * * * * * //

namespace
{
using namespace Foam;

void putScalar(scalar* d, const scalar& v) { d[0] = v; }
void putVector(scalar* d, const vector& v) { d[0] = v.x(); d[1] = v.y(); d[2] = v.z(); }
void putSymmTensor(scalar* d, const symmTensor& v) { d[0] = v.xx(); d[1] = v.xy(); d[2] = v.xz(); d[3] = v.yy(); d[4] = v.yz(); d[5] = v.zz(); }

label indexOf(const wordList& l, const word& w)
{
    forAll(l, i) if (l[i] == w) return i;
    return -1;
}
}

namespace { Foam::label selectFrom(const char* family, const Foam::word& w, const char* const* names, Foam::label n); }

// * * * * * * * * * * * * * Private member functions * * * * * * * * * * * * //

void Foam::mcDeviceCloud::check(int rc, const char* where) const
{
    if (rc)
    {
        FatalErrorIn(where) << pdfb200_last_error() << nl << exit(FatalError);
    }
}


Foam::label Foam::mcDeviceCloud::selectWord(const char* family, const word& w, const char* const* names, label n) const
{
    return selectFrom(family, w, names, n);
}


template<class GF, class Type>
void Foam::mcDeviceCloud::flatten
(
    const GF& f, List<scalar>& cells, List<scalar>& bnd, List<unsigned char>* fixed, label nCmpt, void (*put)(scalar*, const Type&)
) const
{
    const label nc = mesh_.nCells(), nI = mesh_.nInternalFaces();
    cells.setSize(nc*nCmpt);
    bnd.setSize(nBnd_*nCmpt);
    if (fixed) fixed->setSize(nBnd_);
    for (label c = 0; c < nc; ++c) put(&cells[c*nCmpt], f.internalField()[c]);
    const polyBoundaryMesh& pbm = mesh_.boundaryMesh();
    for (label patchI = 0; patchI < pbm.size(); ++patchI)
    {
        const label off = pbm[patchI].start() - nI;
        // empty patches carry no patch field values (size 0): the cell value stands in, as interpolation does
        const bool haveValues = f.boundaryField()[patchI].size() == pbm[patchI].size();
        for (label j = 0; j < pbm[patchI].size(); ++j)
        {
            if (haveValues) put(&bnd[(off + j)*nCmpt], f.boundaryField()[patchI][j]);
            else put(&bnd[(off + j)*nCmpt], f.internalField()[mesh_.faceOwner()[pbm[patchI].start() + j]]);
            if (fixed) (*fixed)[off + j] = f.boundaryField()[patchI].fixesValue() ? 1 : 0;
        }
    }
}


namespace
{
Foam::label selectFrom(const char* family, const Foam::word& w, const char* const* names, Foam::label n)
{
    for (Foam::label i = 0; i < n; ++i)
    {
        if (names[i] && w == names[i]) return i;
    }
    std::string valid;
    for (Foam::label i = 0; i < n; ++i) if (names[i]) { valid += names[i]; valid += ' '; }
    FatalErrorIn("mcDeviceCloud") << "Unknown " << family << " type " << w << Foam::nl << Foam::nl
        << "Valid " << family << " types are :" << Foam::nl << valid << Foam::exit(FatalError);
    return -1;
}
}


pdfb200_params Foam::mcDeviceCloud::readParams
(
    const fvMesh& mesh_,
    const dictionary& thermoDict_,
    const dictionary& solutionDict_,
    wordList& scalarNames_
)
{
    pdfb200_params P;
    std::memset(&P, 0, sizeof(P));

    // scalar bookkeeping (mcParticleCloud.C:1704-1899)
    scalarNames_ = readWordList(thermoDict_.lookup("scalarFields"));
    P.nPhi = scalarNames_.size();
    if (P.nPhi > PDFB200_MAX_PHI)
    {
        FatalErrorIn("mcDeviceCloud::readParams()") << "at most " << label(PDFB200_MAX_PHI) << " scalarFields" << nl << exit(FatalError);
    }
    const wordList mixed = thermoDict_.found("mixedScalars") ? readWordList(thermoDict_.lookup("mixedScalars")) : wordList();
    const wordList cons = thermoDict_.found("conservedScalars") ? readWordList(thermoDict_.lookup("conservedScalars")) : wordList();
    P.nMixed = mixed.size();
    forAll(mixed, i)
    {
        P.mixed[i] = indexOf(scalarNames_, mixed[i]);
        if (P.mixed[i] < 0) FatalErrorIn("mcDeviceCloud::readParams()") << "No such scalar field " << mixed[i] << " (mixedScalars)" << nl << exit(FatalError);
    }
    P.nConserved = cons.size();
    forAll(cons, i)
    {
        P.conserved[i] = indexOf(scalarNames_, cons[i]);
        if (P.conserved[i] < 0) FatalErrorIn("mcDeviceCloud::readParams()") << "No such scalar field " << cons[i] << " (conservedScalars)" << nl << exit(FatalError);
    }

    // mcSolution (defaults mcSolution.C:44-57)
    const dictionary& S = solutionDict_;
    P.CFL = S.lookupOrDefault<scalar>("CFL", 0.5);
    P.averagingCoeff = S.lookupOrDefault<scalar>("averagingCoeff", 1000.);
    P.particlesPerCell = S.lookupOrDefault<label>("particlesPerCell", 40);
    P.particleNumberControl = S.lookupOrDefault<Switch>("particleNumberControl", Switch(true)) ? 1 : 0;
    P.cloneAt = S.lookupOrDefault<scalar>("cloneAt", 0.8);
    P.eliminateAt = S.lookupOrDefault<scalar>("eliminateAt", 1.2);
    P.kMin = S.lookupOrDefault<scalar>("kMin", 100.*SMALL);
    P.DNum = S.lookupOrDefault<scalar>("DNum", 0.);
    const dictionary relax = S.subOrEmptyDict("relaxationTimes");
    const scalar relaxDefault = relax.lookupOrDefault<scalar>("default", GREAT);
    P.relaxTimeU = relax.lookupOrDefault<scalar>("U", relaxDefault);
    P.relaxTimek = relax.lookupOrDefault<scalar>("k", relaxDefault);
    // FV time step times minRelaxationFactor (mcSolution.C:213-230)
    P.minRelaxTime = S.lookupOrDefault<scalar>("minRelaxationFactor", 10.)*S.lookupOrDefault<scalar>("fvDeltaT", 1.);
    const dictionary schemes = S.subOrEmptyDict("interpolationSchemes");
    static const char* const schemeNames[] = {"cell", "cellPointFace"};
    struct { int32_t* dst; const char* key; } sch[] =
    {
        {&P.interpU, "U"}, {&P.interpk, "k"}, {&P.interpDiffU, "SLMFullVelocityModel::diffU"},
        {&P.interpOmega, "mcRASOmegaModel::Omega"}, {&P.interpUPosCorr, "UPosCorr"}, {&P.interpEta, "mcCellLocaltimeStepping::eta"}
    };
    for (unsigned i = 0; i < sizeof(sch)/sizeof(sch[0]); ++i)
    {
        *sch[i].dst = selectFrom("interpolation", schemes.lookupOrDefault<word>(sch[i].key, "cell"), schemeNames, 2);   // default mcSolution.C:49
    }

    // model selectors; the words are the reference's run-time selection names
    static const char* const velocityNames[] = {"none", "SLMFull"};
    static const char* const mixingNames[] = {"none", "IEM", "modifiedCurl"};
    static const char* const omegaNames[] = {"none", "RAS"};
    static const char* const reactionNames[] = {"none", "cold", "BurkeSchumann", "SteadyFlamelet", "naiveFlamelet"};
    static const char* const posCorrNames[] = {"none", "limitedSimple", "simple", "external"};
    static const char* const ltsNames[] = {"off", "cell", "particle"};
    P.velocityModel = selectFrom("mcVelocityModel", thermoDict_.lookupOrDefault<word>("velocityModel", "none"), velocityNames, 2);
    P.mixingModel = selectFrom("mcMixingModel", thermoDict_.lookupOrDefault<word>("mixingModel", "none"), mixingNames, 3);
    P.omegaModel = selectFrom("mcOmegaModel", thermoDict_.lookupOrDefault<word>("OmegaModel", "none"), omegaNames, 2);
    P.reactionModel = selectFrom("mcReactionModel", thermoDict_.lookupOrDefault<word>("reactionModel", "none"), reactionNames, 5);
    P.positionCorrection = selectFrom("mcPositionCorrection", thermoDict_.lookupOrDefault<word>("positionCorrection", "none"), posCorrNames, 4);
    P.localTimeStepping = selectFrom("mcLocalTimeStepping", thermoDict_.lookupOrDefault<word>("localTimeStepping", "off"), ltsNames, 3);

    // coefficients: sub-dictionaries <Type><Family>Coeffs of the solution dictionary (mcModel.C:47-60)
    const dictionary slm = S.subOrEmptyDict("SLMFullVelocityModelCoeffs");
    P.C0 = slm.lookupOrDefault<scalar>("C0", 2.1);           // mcSLMFullVelocityModel.C:94-99
    P.C1 = slm.lookupOrDefault<scalar>("C1", 1.0);
    const word mixCoeffs = P.mixingModel == PDFB200_MIXING_MODCURL ? "modifiedCurlMixingModelCoeffs" : "IEMMixingModelCoeffs";
    P.Cmix = S.subOrEmptyDict(mixCoeffs).lookupOrDefault<scalar>("Cmix", 2.0);   // mcIEMMixingModel.C:62-65
    P.Omega0 = S.subOrEmptyDict("RASOmegaModelCoeffs").lookupOrDefault<scalar>("Omega0", VGREAT);   // mcRASOmegaModel.C:114
    const word pcCoeffs = P.positionCorrection == PDFB200_POSCORR_SIMPLE ? "simplePositionCorrectionCoeffs" : "limitedSimplePositionCorrectionCoeffs";
    const dictionary pc = S.subOrEmptyDict(pcCoeffs);
    P.pcC = pc.lookupOrDefault<scalar>("C", 1e-2);           // mcLimitedSimplePositionCorrection.C:90-102
    P.pcCFL = pc.lookupOrDefault<scalar>("CFL", 0.5);
    const word ltsCoeffs = P.localTimeStepping == PDFB200_LTS_PARTICLE ? "particleLocalTimeSteppingCoeffs" : "cellLocalTimeSteppingCoeffs";
    P.ltsUpperBound = S.subOrEmptyDict(ltsCoeffs).lookupOrDefault<scalar>("upperBound", 10.);   // mcCellLocalTimeStepping.C:84-85

    // reaction models (thermoDict sub-dictionaries, mcReactionModel.C)
    const dictionary cold = thermoDict_.subOrEmptyDict("coldReactionModelCoeffs");
    P.coldDensity = cold.lookupOrDefault<scalar>("density", 1.0);   // mcColdReactionModel.C:57
    P.zIdx = indexOf(scalarNames_, thermoDict_.lookupOrDefault<word>("zName", "z"));
    P.TIdx = indexOf(scalarNames_, thermoDict_.lookupOrDefault<word>("TName", "T"));
    P.chiIdx = indexOf(scalarNames_, thermoDict_.lookupOrDefault<word>("chiName", "chi"));
    P.cIdx = indexOf(scalarNames_, thermoDict_.lookupOrDefault<word>("cName", "c"));
    P.pvIdx = indexOf(scalarNames_, thermoDict_.lookupOrDefault<word>("pvName", "pv"));
    if (P.reactionModel == PDFB200_REACTION_BURKE_SCHUMANN)
    {
        const dictionary& bs = thermoDict_.subDict("BurkeSchumannReactionModelCoeffs");   // mcBurkeSchumannReactionModel.C:63-69
        P.bsRfuel = readScalar(bs.lookup("Rfuel")); P.bsRox = readScalar(bs.lookup("Rox")); P.bsRstoi = readScalar(bs.lookup("Rstoi"));
        P.bsTfuel = readScalar(bs.lookup("Tfuel")); P.bsTox = readScalar(bs.lookup("Tox")); P.bsTstoi = readScalar(bs.lookup("Tstoi"));
        P.bsZstoi = readScalar(bs.lookup("zstoi"));
    }
    if (P.reactionModel == PDFB200_REACTION_STEADY_FLAMELET)
    {
        const dictionary fl = thermoDict_.subOrEmptyDict("SteadyFlameletReactionModelCoeffs");
        P.Cchi = fl.lookupOrDefault<scalar>("Cchi", 2.0);
        const wordList added = fl.found("scalars") ? readWordList(fl.lookup("scalars")) : wordList();   // mcSteadyFlamelet.C:154-178
        P.nAdded = added.size();
        forAll(added, i) P.addedIdx[i] = indexOf(scalarNames_, added[i]);
    }
    // zzCov drives chi: its scheme (interpolationSchemes "zzCov")
    P.interpZVar = selectFrom("interpolation", schemes.lookupOrDefault<word>(thermoDict_.lookupOrDefault<word>("zName", "z") + thermoDict_.lookupOrDefault<word>("zName", "z") + "Cov", "cell"), schemeNames, 2);

    // what the reference takes from elsewhere: RASModel::k0(), the initial cloud time step (100*SMALL), Random's seed
    const dictionary dev = S.subOrEmptyDict("deviceCoeffs");
    P.k0 = dev.lookupOrDefault<scalar>("k0", SMALL);
    P.deltaT0 = dev.lookupOrDefault<scalar>("deltaT0", 100.*SMALL);
    P.seed = (uint64_t)::strtoull(dev.lookupOrDefault<word>("seed", "0").c_str(), 0, 0);
    P.rngMode = 0;
    static const char* const inletRandomNames[] = {"inversion", "sampling"};
    P.inletRandom = 0;
    // one generator type for all inletOutlet handlers (boundaryHandlers{<patch>{randomGenerator{type ...;}}})
    const dictionary& bh = thermoDict_.subDict("boundaryHandlers");
    const polyBoundaryMesh& pbm = mesh_.boundaryMesh();
    for (label patchI = 0; patchI < pbm.size(); ++patchI)
    {
        const dictionary& d = bh.subDict(pbm[patchI].name());
        if (d.isDict("randomGenerator"))
        {
            P.inletRandom = selectFrom("mcInletRandom", d.subDict("randomGenerator").lookupOrDefault<word>("type", "inversion"), inletRandomNames, 2);
        }
    }
    return P;
}


void Foam::mcDeviceCloud::createHandle(int device, label capacity)
{
    const label nI = mesh_.nInternalFaces(), nF = mesh_.nFaces(), nP = mesh_.points().size();
    const polyBoundaryMesh& pbm = mesh_.boundaryMesh();
    nBnd_ = nF - nI;

    List<scalar> points(3*nP);
    for (label i = 0; i < nP; ++i) putVector(&points[3*i], mesh_.points()[i]);
    List<label> faceStart(nF + 1), facePoints;
    faceStart[0] = 0;
    for (label f = 0; f < nF; ++f)
    {
        forAll(mesh_.faces()[f], k) facePoints.append(mesh_.faces()[f][k]);
        faceStart[f + 1] = facePoints.size();
    }
    List<label> owner(nF), neighbour(nI), patchStart(pbm.size()), patchSize(pbm.size()), patchGeom(pbm.size()),
        patchHandler(pbm.size()), patchReflecting(pbm.size());
    for (label f = 0; f < nF; ++f) owner[f] = mesh_.faceOwner()[f];
    for (label f = 0; f < nI; ++f) neighbour[f] = mesh_.faceNeighbour()[f];

    // boundaryHandlers{} (mcParticleCloud.C:1902-1934, mcBoundary.C:42-82, mcOpenBoundary.C:52)
    static const char* const handlerNames[] = {"slip", "empty", "open", "inletOutlet"};
    const dictionary& bh = thermoDict_.subDict("boundaryHandlers");
    for (label patchI = 0; patchI < pbm.size(); ++patchI)
    {
        const polyPatch& pp = pbm[patchI];
        patchStart[patchI] = pp.start();
        patchSize[patchI] = pp.size();
        patchGeom[patchI] = isA<emptyPolyPatch>(pp) ? PDFB200_PATCH_EMPTY : isA<wedgePolyPatch>(pp) ? PDFB200_PATCH_WEDGE : PDFB200_PATCH_GENERIC;
        if (!bh.isDict(pp.name()))
        {
            FatalErrorIn("mcDeviceCloud::createHandle()") << "No boundary handler for patch " << pp.name() << nl << exit(FatalError);
        }
        const dictionary& d = bh.subDict(pp.name());
        patchHandler[patchI] = selectWord("mcBoundary", d.lookup("type"), handlerNames, 4);
        patchReflecting[patchI] = d.lookupOrDefault<Switch>("reflecting", Switch(true)) ? 1 : 0;
    }

    pdfb200_mesh m;
    std::memset(&m, 0, sizeof(m));
    m.nPoints = nP; m.nFaces = nF; m.nInternalFaces = nI; m.nCells = mesh_.nCells(); m.nPatches = pbm.size();
    m.points = &points[0]; m.faceStart = &faceStart[0]; m.facePoints = &facePoints[0];
    m.owner = &owner[0]; m.neighbour = nI ? &neighbour[0] : 0;
    m.patchStart = &patchStart[0]; m.patchSize = &patchSize[0]; m.patchGeom = &patchGeom[0];
    m.patchHandler = &patchHandler[0]; m.patchReflecting = &patchReflecting[0];
    for (label d = 0; d < 3; ++d) { m.geometricD[d] = mesh_.geometricD()[d]; m.solutionD[d] = mesh_.solutionD()[d]; }

    if (capacity <= 0) capacity = label(1.5*mesh_.nCells()*params_.particlesPerCell) + 4096;
    check(pdfb200_create(&m, &params_, device, capacity, &h_), "mcDeviceCloud::createHandle()");
}


void Foam::mcDeviceCloud::uploadFvFields()
{
    List<unsigned char> noFlags;
    flatten<volVectorField, vector>(Ufv_, U_, Ub_, &UbFixed_, 3, putVector);
    flatten<volScalarField, scalar>(pfv_, p_, pb_, 0, 1, putScalar);
    flatten<volScalarField, scalar>(turbModel_.k(), k_, kb_, &kbFixed_, 1, putScalar);
    flatten<volScalarField, scalar>(turbModel_.epsilon(), eps_, epsb_, 0, 1, putScalar);
    flatten<volSymmTensorField, symmTensor>(turbModel_.R(), R_, Rb_, 0, 6, putSymmTensor);
    pdfb200_fv_fields f;
    std::memset(&f, 0, sizeof(f));
    f.U = &U_[0]; f.Ub = &Ub_[0]; f.UbFixed = &UbFixed_[0];
    f.p = &p_[0]; f.pb = &pb_[0];
    f.k = &k_[0]; f.kb = &kb_[0]; f.kbFixed = &kbFixed_[0];
    f.epsilon = &eps_[0]; f.epsilonb = &epsb_[0];
    f.R = &R_[0]; f.Rb = &Rb_[0];
    check(pdfb200_upload_fv_fields(h_, &f), "mcDeviceCloud::uploadFvFields()");
}


void Foam::mcDeviceCloud::uploadCloudFields()
{
    const label nPhi = params_.nPhi, nCov = nPhi*(nPhi + 1)/2;
    List<List<scalar> > c(1 + nPhi + nCov), b(1 + nPhi + nCov);
    List<List<unsigned char> > fx(1 + nPhi + nCov);
    pdfb200_cloud_fields f;
    std::memset(&f, 0, sizeof(f));
    flatten<volScalarField, scalar>(rhocPdf_, c[0], b[0], &fx[0], 1, putScalar);
    f.rho = &c[0][0]; f.rhob = &b[0][0]; f.rhobFixed = &fx[0][0];
    for (label i = 0; i < nPhi; ++i)
    {
        flatten<volScalarField, scalar>(*PhicPdf_[i], c[1 + i], b[1 + i], &fx[1 + i], 1, putScalar);
        f.Phi[i] = &c[1 + i][0]; f.Phib[i] = &b[1 + i][0]; f.PhibFixed[i] = &fx[1 + i][0];
    }
    for (label i = 0; i < nCov; ++i)
    {
        if (!PhiPhicPdf_[i]) continue;
        flatten<volScalarField, scalar>(*PhiPhicPdf_[i], c[1 + nPhi + i], b[1 + nPhi + i], &fx[1 + nPhi + i], 1, putScalar);
        f.Cov[i] = &c[1 + nPhi + i][0]; f.Covb[i] = &b[1 + nPhi + i][0]; f.CovbFixed[i] = &fx[1 + nPhi + i][0];
    }
    check(pdfb200_upload_cloud_fields(h_, &f), "mcDeviceCloud::uploadCloudFields()");
}

// * * * * * * * * * * * * * * * * Constructors  * * * * * * * * * * * * * * //

Foam::mcDeviceCloud::mcDeviceCloud
(
    const fvMesh& mesh,
    const dictionary& thermoDict,
    const dictionary& solutionDict,
    const word& cloudName,
    const compressible::turbulenceModel& turbModel,
    const volVectorField& U,
    const volScalarField& p,
    volScalarField& rho,
    int device,
    label capacity
)
:
    mesh_(mesh),
    name_(cloudName),
    thermoDict_(thermoDict),
    solutionDict_(solutionDict),
    turbModel_(turbModel),
    Ufv_(U),
    pfv_(p),
    rhocPdf_(rho),
    h_(0),
    nBnd_(0)
{
    std::memset(&lastReport_, 0, sizeof(lastReport_));
    params_ = readParams(mesh_, thermoDict_, solutionDict_, scalarNames_);
    // the scalar fields and their covariances live in the registry under the names of scalarFields and
    // <a><b>Cov (mcParticleCloud.C:1731-1797); missing covariance fields start from zero
    const label nPhi = params_.nPhi;
    PhicPdf_.setSize(nPhi);
    forAll(scalarNames_, i) PhicPdf_[i] = &mesh_.lookupObject<volScalarField>(scalarNames_[i]);
    PhiPhicPdf_.setSize(nPhi*(nPhi + 1)/2);
    label pp = 0;
    for (label i = 0; i < nPhi; ++i)
        for (label j = i; j < nPhi; ++j, ++pp)
        {
            const word covName = scalarNames_[i] + scalarNames_[j] + "Cov";
            PhiPhicPdf_[pp] = mesh_.foundObject<volScalarField>(covName) ? &mesh_.lookupObject<volScalarField>(covName) : 0;
        }
    createHandle(device, capacity);
    uploadFvFields();
    uploadCloudFields();
}

// * * * * * * * * * * * * * * * * Destructor  * * * * * * * * * * * * * * * //

Foam::mcDeviceCloud::~mcDeviceCloud()
{
    pdfb200_destroy(h_);
}

// * * * * * * * * * * * * * * * Member Functions  * * * * * * * * * * * * * //

void Foam::mcDeviceCloud::setFlameletLibrary(const List<scalar>& chi, const List<scalar>& z, const List<List<scalar> >& tables)
{
    List<const double*> ptr(tables.size());
    forAll(tables, i)
    {
        if (tables[i].size() != chi.size()*z.size())
        {
            FatalErrorIn("mcDeviceCloud::setFlameletLibrary()") << "table " << label(i) << " does not hold chi.size()*z.size() values" << nl << exit(FatalError);
        }
        ptr[i] = &tables[i][0];
    }
    check(pdfb200_set_flamelet_tables(h_, chi.size(), z.size(), &chi[0], &z[0], tables.size(), &ptr[0]), "mcDeviceCloud::setFlameletLibrary()");
}


void Foam::mcDeviceCloud::initialise()
{
    check(pdfb200_init_moments(h_), "mcDeviceCloud::initialise()");
    check(pdfb200_seed_particles(h_), "mcDeviceCloud::initialise()");
}


Foam::scalar Foam::mcDeviceCloud::evolve()
{
    uploadFvFields();            // U, p, k, epsilon, R and their patch values as the FV side left them
    check(pdfb200_evolve(h_, 1, &lastReport_), "mcDeviceCloud::evolve()");
    downloadCellFields();
    const pdfb200_step_report& r = lastReport_;
    // the lines mcParticleCloud::evolve prints (mcParticleCloud.C:1382-1509)
    Info<< "Cloud " << name_ << nl
        << "    deltaT = " << r.deltaT << nl
        << "    number of particles: " << r.nParticles << " (injected " << r.nInjected << ", deleted " << r.nDeleted
        << ", cloned " << r.nCloned << ", eliminated " << r.nEliminated << ", lost " << r.nLost << ")" << nl
        << "    pmd relative error: averageMag = " << r.rhoRes << ", min = " << r.rhoErrMin << ", max = " << r.rhoErrMax << nl
        << "    max(mag(U)) = " << r.UMax << ", max(mag(Ucorrection)) = " << r.UcorrMax << ", CoMax = " << r.CoMax << endl;
    return r.rhoRes;
}


void Foam::mcDeviceCloud::downloadCellFields()
{
    const label nc = mesh_.nCells(), nPhi = params_.nPhi, nCov = nPhi*(nPhi + 1)/2;
    List<scalar> rho(nc), U(3*nc), k(nc), Tau(6*nc), pnd(nc);
    List<List<scalar> > phi(nPhi), cov(nCov);
    pdfb200_cell_fields o;
    std::memset(&o, 0, sizeof(o));
    o.rho = &rho[0]; o.U = &U[0]; o.k = &k[0]; o.Tau = &Tau[0]; o.pnd = &pnd[0];
    for (label i = 0; i < nPhi; ++i) { phi[i].setSize(nc); o.Phi[i] = &phi[i][0]; }
    for (label i = 0; i < nCov; ++i) { cov[i].setSize(nc); o.Cov[i] = &cov[i][0]; }
    check(pdfb200_download_cell_fields(h_, &o), "mcDeviceCloud::downloadCellFields()");
    for (label c = 0; c < nc; ++c) rhocPdf_.internalField()[c] = rho[c];
    for (label i = 0; i < nPhi; ++i) for (label c = 0; c < nc; ++c) PhicPdf_[i]->internalField()[c] = phi[i][c];
    for (label i = 0; i < nCov; ++i) if (PhiPhicPdf_[i]) for (label c = 0; c < nc; ++c) PhiPhicPdf_[i]->internalField()[c] = cov[i][c];
    if (mesh_.foundObject<volVectorField>("UCloudPDF"))
    {
        volVectorField& f = mesh_.lookupObject<volVectorField>("UCloudPDF");
        for (label c = 0; c < nc; ++c) f.internalField()[c] = vector(U[3*c], U[3*c + 1], U[3*c + 2]);
    }
    if (mesh_.foundObject<volScalarField>("kCloudPDF"))
    {
        volScalarField& f = mesh_.lookupObject<volScalarField>("kCloudPDF");
        for (label c = 0; c < nc; ++c) f.internalField()[c] = k[c];
    }
    if (mesh_.foundObject<volSymmTensorField>("TauCloudPDF"))
    {
        volSymmTensorField& f = mesh_.lookupObject<volSymmTensorField>("TauCloudPDF");
        for (label c = 0; c < nc; ++c) f.internalField()[c] = symmTensor(Tau[6*c], Tau[6*c + 1], Tau[6*c + 2], Tau[6*c + 3], Tau[6*c + 4], Tau[6*c + 5]);
    }
    if (mesh_.foundObject<volScalarField>("pndCloudPDF"))
    {
        volScalarField& f = mesh_.lookupObject<volScalarField>("pndCloudPDF");
        for (label c = 0; c < nc; ++c) f.internalField()[c] = pnd[c];
    }
}

// ************************************************************************* //
