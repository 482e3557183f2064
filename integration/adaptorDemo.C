// adaptorDemo.C — drives mcDeviceCloud the way mcThermo does (mcThermo.C:114-143): builds the OpenFOAM-side
// objects (fvMesh, FV fields, turbulence model, dictionaries) from a case directory written by
// tests/test_adaptor.py, constructs the cloud, evolves N steps and writes the density field fed back to the FV
// equations (pdfSimpleFoam/createFields.H:18) plus the step reports. Test infrastructure for the adaptor.
#include "mcDeviceCloud.H"

#include <cstring>
#include <fstream>
#include <sstream>

using namespace Foam;

namespace
{
std::string slurp(const std::string& p)
{
    std::ifstream f(p.c_str(), std::ios::binary);
    if (!f) throw std::runtime_error("cannot read " + p);
    std::ostringstream s; s << f.rdbuf();
    return s.str();
}
template<class T> std::vector<T> raw(const std::string& p)
{
    const std::string b = slurp(p);
    std::vector<T> v(b.size()/sizeof(T));
    std::memcpy(v.data(), b.data(), v.size()*sizeof(T));
    return v;
}
void fillScalar(volScalarField& f, const std::vector<double>& c, const std::vector<double>& b, const std::vector<unsigned char>* fixed, const fvMesh& m)
{
    for (label i = 0; i < m.nCells(); ++i) f.internalField()[i] = c[i];
    for (label p = 0; p < m.boundaryMesh().size(); ++p)
    {
        const label off = m.boundaryMesh()[p].start() - m.nInternalFaces();
        if (fixed && m.boundaryMesh()[p].size()) f.setPatchFixesValue(p, (*fixed)[off] != 0);
        for (label j = 0; j < m.boundaryMesh()[p].size(); ++j) f.boundaryField()[p][j] = b[off + j];
    }
}
}

int main(int argc, char* argv[])
{
    if (argc < 2) { fprintf(stderr, "usage: adaptorDemo <case directory>\n"); return 2; }
    const std::string d = std::string(argv[1]) + "/";
    try
    {
        // ---- polyMesh ----
        std::istringstream sz(slurp(d + "sizes.txt"));
        label nPoints, nFaces, nInternal, nCells, nPatches, nSteps, nChi, nZ;
        fvMesh mesh;
        sz >> nPoints >> nFaces >> nInternal >> nCells >> nPatches >> nSteps >> nChi >> nZ;
        for (int k = 0; k < 3; ++k) sz >> mesh.geometricD_[k];
        for (int k = 0; k < 3; ++k) sz >> mesh.solutionD_[k];
        const std::vector<double> pts = raw<double>(d + "points.f64");
        const std::vector<int> faceStart = raw<int>(d + "faceStart.i32"), facePoints = raw<int>(d + "facePoints.i32");
        const std::vector<int> owner = raw<int>(d + "owner.i32"), neighbour = raw<int>(d + "neighbour.i32");
        mesh.nCells_ = nCells;
        mesh.points_.setSize(nPoints);
        for (label i = 0; i < nPoints; ++i) mesh.points_[i] = point(pts[3*i], pts[3*i + 1], pts[3*i + 2]);
        mesh.faces_.setSize(nFaces); mesh.owner_.setSize(nFaces); mesh.neighbour_.setSize(nInternal);
        for (label f = 0; f < nFaces; ++f)
        {
            mesh.faces_[f].setSize(faceStart[f + 1] - faceStart[f]);
            for (int k = faceStart[f]; k < faceStart[f + 1]; ++k) mesh.faces_[f][k - faceStart[f]] = facePoints[k];
            mesh.owner_[f] = owner[f];
        }
        for (label f = 0; f < nInternal; ++f) mesh.neighbour_[f] = neighbour[f];
        std::istringstream ps(slurp(d + "patches.txt"));
        for (label p = 0; p < nPatches; ++p)
        {
            std::string name, cls; label start, size;
            ps >> name >> start >> size >> cls;
            if (cls == "empty") mesh.boundary_.append(new emptyPolyPatch(name, start, size));
            else if (cls == "wedge") mesh.boundary_.append(new wedgePolyPatch(name, start, size));
            else mesh.boundary_.append(new polyPatch(name, start, size));
        }

        // ---- FV fields ----
        const std::vector<unsigned char> UbFixed = raw<unsigned char>(d + "UbFixed.u8"), kbFixed = raw<unsigned char>(d + "kbFixed.u8"),
            rhobFixed = raw<unsigned char>(d + "rhobFixed.u8");
        volVectorField U("U", mesh);
        {
            const std::vector<double> c = raw<double>(d + "U.f64"), b = raw<double>(d + "Ub.f64");
            for (label i = 0; i < nCells; ++i) U.internalField()[i] = vector(c[3*i], c[3*i + 1], c[3*i + 2]);
            for (label p = 0; p < nPatches; ++p)
            {
                const label off = mesh.boundaryMesh()[p].start() - nInternal;
                if (mesh.boundaryMesh()[p].size()) U.setPatchFixesValue(p, UbFixed[off] != 0);
                for (label j = 0; j < mesh.boundaryMesh()[p].size(); ++j) U.boundaryField()[p][j] = vector(b[3*(off + j)], b[3*(off + j) + 1], b[3*(off + j) + 2]);
            }
        }
        volScalarField p("p", mesh), rho("rho", mesh);
        fillScalar(p, raw<double>(d + "p.f64"), raw<double>(d + "pb.f64"), 0, mesh);
        fillScalar(rho, raw<double>(d + "rho.f64"), raw<double>(d + "rhob.f64"), &rhobFixed, mesh);
        compressible::turbulenceModel turb(mesh);
        fillScalar(turb.k_, raw<double>(d + "k.f64"), raw<double>(d + "kb.f64"), &kbFixed, mesh);
        fillScalar(turb.epsilon_, raw<double>(d + "epsilon.f64"), raw<double>(d + "epsilonb.f64"), 0, mesh);
        {
            const std::vector<double> c = raw<double>(d + "R.f64"), b = raw<double>(d + "Rb.f64");
            for (label i = 0; i < nCells; ++i) turb.R_.internalField()[i] = symmTensor(c[6*i], c[6*i + 1], c[6*i + 2], c[6*i + 3], c[6*i + 4], c[6*i + 5]);
            for (label pa = 0; pa < nPatches; ++pa)
            {
                const label off = mesh.boundaryMesh()[pa].start() - nInternal;
                for (label j = 0; j < mesh.boundaryMesh()[pa].size(); ++j)
                {
                    const double* r = &b[6*(off + j)];
                    turb.R_.boundaryField()[pa][j] = symmTensor(r[0], r[1], r[2], r[3], r[4], r[5]);
                }
            }
        }

        // ---- dictionaries, scalar fields in the registry ----
        const dictionary thermoDict = dictionary::parse(slurp(d + "thermoCloudProperties"));
        const dictionary solutionDict = dictionary::parse(slurp(d + "thermoCloudSolution"));
        const wordList names = readWordList(thermoDict.lookup("scalarFields"));
        std::vector<std::shared_ptr<volScalarField> > scalars;
        forAll(names, i)
        {
            scalars.push_back(std::shared_ptr<volScalarField>(new volScalarField(names[i], mesh)));
            const std::vector<unsigned char> fx = raw<unsigned char>(d + "PhibFixed_" + names[i] + ".u8");
            fillScalar(*scalars.back(), raw<double>(d + "Phi_" + names[i] + ".f64"), raw<double>(d + "Phib_" + names[i] + ".f64"), &fx, mesh);
            mesh.store(names[i], scalars.back().get());
        }
        volVectorField UCloudPDF("UCloudPDF", mesh);
        mesh.store("UCloudPDF", &UCloudPDF);

        {
            // what the dictionaries were read as (compared with the Python-side parameters by tests/test_adaptor.py)
            wordList namesRead;
            const pdfb200_params P = mcDeviceCloud::readParams(mesh, thermoDict, solutionDict, namesRead);
            std::ofstream o((d + "params_read.txt").c_str());
            o.precision(17);
#define W(f) o << #f << ' ' << P.f << '\n';
            W(nPhi) W(nMixed) W(nConserved) W(CFL) W(averagingCoeff) W(particlesPerCell) W(particleNumberControl) W(cloneAt) W(eliminateAt)
            W(kMin) W(DNum) W(relaxTimeU) W(relaxTimek) W(minRelaxTime) W(interpU) W(interpk) W(interpDiffU) W(interpOmega) W(interpUPosCorr)
            W(interpZVar) W(interpEta) W(velocityModel) W(C0) W(C1) W(mixingModel) W(Cmix) W(omegaModel) W(Omega0) W(reactionModel)
            W(coldDensity) W(bsRfuel) W(bsRox) W(bsRstoi) W(bsTfuel) W(bsTox) W(bsTstoi) W(bsZstoi) W(zIdx) W(Cchi) W(nAdded)
            W(positionCorrection) W(pcC) W(pcCFL) W(localTimeStepping) W(ltsUpperBound) W(k0) W(deltaT0) W(seed) W(inletRandom)
#undef W
            for (int i = 0; i < P.nMixed; ++i) o << "mixed" << i << ' ' << P.mixed[i] << '\n';
            for (int i = 0; i < P.nConserved; ++i) o << "conserved" << i << ' ' << P.conserved[i] << '\n';
            for (int i = 0; i < P.nAdded; ++i) o << "addedIdx" << i << ' ' << P.addedIdx[i] << '\n';
        }
        // ---- the cloud, as mcThermo::createCloud does ----
        mcDeviceCloud cloud(mesh, thermoDict, solutionDict, "thermoCloud", turb, U, p, rho, 0, 0);
        if (nChi > 0)
        {
            const std::vector<double> chi = raw<double>(d + "chi.f64"), z = raw<double>(d + "z.f64");
            List<scalar> lchi(nChi), lz(nZ);
            for (label i = 0; i < nChi; ++i) lchi[i] = chi[i];
            for (label i = 0; i < nZ; ++i) lz[i] = z[i];
            std::istringstream ts(slurp(d + "tables.txt"));
            List<List<scalar> > tables;
            std::string tname;
            while (ts >> tname)
            {
                const std::vector<double> t = raw<double>(d + tname);
                List<scalar> l(label(t.size()));
                for (size_t i = 0; i < t.size(); ++i) l[label(i)] = t[i];
                tables.append(l);
            }
            cloud.setFlameletLibrary(lchi, lz, tables);
        }
        cloud.initialise();
        std::ofstream rep((d + "report.txt").c_str());
        rep.precision(17);
        for (label s = 0; s < nSteps; ++s)
        {
            const scalar res = cloud.evolve();
            const pdfb200_step_report& r = cloud.lastReport();
            rep << res << ' ' << r.deltaT << ' ' << r.deltaTNext << ' ' << r.nParticles << ' ' << r.nInjected << ' ' << r.nDeleted << ' '
                << r.nCloned << ' ' << r.nEliminated << ' ' << r.UMax << '\n';
        }
        std::vector<double> out(nCells), uout(3*nCells);
        for (label i = 0; i < nCells; ++i) { out[i] = rho.internalField()[i]; for (int k = 0; k < 3; ++k) uout[3*i + k] = UCloudPDF.internalField()[i][k]; }
        std::ofstream((d + "rho_out.f64").c_str(), std::ios::binary).write(reinterpret_cast<const char*>(out.data()), out.size()*sizeof(double));
        std::ofstream((d + "U_out.f64").c_str(), std::ios::binary).write(reinterpret_cast<const char*>(uout.data()), uout.size()*sizeof(double));
        forAll(names, i)
        {
            std::vector<double> o(nCells);
            for (label c = 0; c < nCells; ++c) o[c] = scalars[i]->internalField()[c];
            std::ofstream((d + "Phi_out_" + names[i] + ".f64").c_str(), std::ios::binary).write(reinterpret_cast<const char*>(o.data()), o.size()*sizeof(double));
        }
    }
    catch (const std::exception& e)
    {
        fprintf(stderr, "FOAM FATAL ERROR: %s\n", e.what());
        return 1;
    }
    return 0;
}
