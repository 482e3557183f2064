// capi.cpp — CPU ORACLE (test infrastructure): the pdfb200.h entry points with
// the prefix pdforacle_, plus a few probes used by the known-answer tests.
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>

#include "pdforacle.hpp"

using namespace pdforacle;

static thread_local std::string g_err;

struct pdforacle_cloud { Cloud c; };

#define ORACLE_TRY try {
#define ORACLE_CATCH                                                     \
    }                                                                    \
    catch (const std::exception &e) { g_err = e.what(); return PDFB200_ERR_INVALID; } \
    return PDFB200_OK;

static void copyScalar(ScalarField &f, const double *c, const double *b, const uint8_t *fx, int nc, int nb) {
    f.c.assign(c, c + nc);
    if (b) f.b.assign(b, b + nb); else f.b.assign(nb, 0.0);
    if (fx) f.fixed.assign(fx, fx + nb); else f.fixed.assign(nb, 0);
}

extern "C" {

const char *pdforacle_last_error(void) { return g_err.c_str(); }

int pdforacle_create(const pdfb200_mesh *mesh, const pdfb200_params *params, int, int64_t, pdforacle_cloud **out) {
    ORACLE_TRY
    auto *h = new pdforacle_cloud();
    std::string err = h->c.init(*mesh, *params);
    if (!err.empty()) { delete h; g_err = err; return PDFB200_ERR_INVALID; }
    *out = h;
    ORACLE_CATCH
}

void pdforacle_destroy(pdforacle_cloud *h) { delete h; }

int pdforacle_set_params(pdforacle_cloud *h, const pdfb200_params *p) {
    std::string err = h->c.setParams(*p);
    if (!err.empty()) { g_err = err; return PDFB200_ERR_INVALID; }
    return PDFB200_OK;
}

int pdforacle_set_rank(pdforacle_cloud *h, int nRanks, int rank) {
    h->c.nRanks = nRanks; h->c.rank = rank;
    return PDFB200_OK;
}

int pdforacle_set_flamelet_tables(pdforacle_cloud *h, int32_t nChi, int32_t nZ, const double *chi, const double *z,
                                  int32_t nFields, const double *const *fields) {
    Cloud &c = h->c;
    if (nZ < 2 || nChi < 1) { g_err = "flamelet tables: need nZ >= 2 and nChi >= 1"; return PDFB200_ERR_INVALID; }
    for (int i = 1; i < nZ; ++i) if (!(z[i] - z[i - 1] > 0)) { g_err = "z not strictly monotonically increasing"; return PDFB200_ERR_INVALID; }
    for (int i = 1; i < nChi; ++i) if (!(chi[i] - chi[i - 1] > 0)) { g_err = "chi not strictly monotonically increasing"; return PDFB200_ERR_INVALID; }
    c.nChi = nChi; c.nZ = nZ;
    c.flChi.assign(chi, chi + nChi); c.flZ.assign(z, z + nZ);
    c.flFields.resize(nFields);
    for (int f = 0; f < nFields; ++f) c.flFields[f].assign(fields[f], fields[f] + (size_t)nChi * nZ);
    return PDFB200_OK;
}

int pdforacle_set_autoignition_tables(pdforacle_cloud *h, int32_t nZ, int32_t nPv, const double *z, const double *pv, const double *cEq,
                                      int32_t nFields, const double *const *fields) {
    Cloud &c = h->c;
    if (nZ < 2 || nPv < 2 || nFields < 2) { g_err = "auto-ignition tables: need nZ >= 2, nPv >= 2, cdot and rho"; return PDFB200_ERR_INVALID; }
    for (int i = 1; i < nZ; ++i) if (!(z[i] - z[i - 1] > 0)) { g_err = "auto-ignition z not strictly monotonically increasing"; return PDFB200_ERR_INVALID; }
    for (int i = 1; i < nPv; ++i) if (!(pv[i] - pv[i - 1] > 0)) { g_err = "auto-ignition pv not strictly monotonically increasing"; return PDFB200_ERR_INVALID; }
    c.aiZ.assign(z, z + nZ); c.aiPv.assign(pv, pv + nPv); c.aiCEq.assign(cEq, cEq + nZ);
    c.aiCEqMax = *std::max_element(c.aiCEq.begin(), c.aiCEq.end());
    c.aiFields.resize(nFields);
    for (int f = 0; f < nFields; ++f) c.aiFields[f].assign(fields[f], fields[f] + (size_t)nZ * nPv);
    return PDFB200_OK;
}

int pdforacle_upload_fv_fields(pdforacle_cloud *h, const pdfb200_fv_fields *f) {
    ORACLE_TRY
    Cloud &c = h->c;
    const int nc = c.mesh.nCells, nb = c.nBnd();
    c.Ufv.c.resize(nc); c.Ufv.b.resize(nb);
    for (int i = 0; i < nc; ++i) c.Ufv.c[i] = Vec3(f->U[3 * i], f->U[3 * i + 1], f->U[3 * i + 2]);
    for (int i = 0; i < nb; ++i) c.Ufv.b[i] = Vec3(f->Ub[3 * i], f->Ub[3 * i + 1], f->Ub[3 * i + 2]);
    if (f->UbFixed) c.Ufv.fixed.assign(f->UbFixed, f->UbFixed + nb); else c.Ufv.fixed.assign(nb, 0);
    copyScalar(c.pfv, f->p, f->pb, nullptr, nc, nb);
    copyScalar(c.kfv, f->k, f->kb, f->kbFixed, nc, nb);
    copyScalar(c.epsfv, f->epsilon, f->epsilonb, nullptr, nc, nb);
    c.Rfv.c.resize(nc); c.Rfv.b.resize(nb);
    auto st = [](const double *r) { SymmTensor s; s.xx = r[0]; s.xy = r[1]; s.xz = r[2]; s.yy = r[3]; s.yz = r[4]; s.zz = r[5]; return s; };
    for (int i = 0; i < nc; ++i) c.Rfv.c[i] = st(f->R + 6 * i);
    for (int i = 0; i < nb; ++i) c.Rfv.b[i] = st(f->Rb + 6 * i);
    c.Rfv.fixed = c.kfv.fixed;
    c.haveFv = true;
    ORACLE_CATCH
}

// stage / commit (pdfb200_stage_fv_fields / pdfb200_commit_fv_fields): the CPU has nothing to overlap; the staged
// fields are held in a second cloud-side copy until the commit
struct StagedFv { std::vector<double> U, Ub, p, pb, k, kb, e, eb, R, Rb; std::vector<uint8_t> UbF, kbF; bool have = false; };
static thread_local StagedFv g_staged;
int pdforacle_stage_fv_fields(pdforacle_cloud *h, const pdfb200_fv_fields *f) {
    const int nc = h->c.mesh.nCells, nb = h->c.nBnd();
    StagedFv &s = g_staged;
    s.U.assign(f->U, f->U + 3 * nc); s.Ub.assign(f->Ub, f->Ub + 3 * nb); s.p.assign(f->p, f->p + nc); s.pb.assign(f->pb, f->pb + nb);
    s.k.assign(f->k, f->k + nc); s.kb.assign(f->kb, f->kb + nb); s.e.assign(f->epsilon, f->epsilon + nc); s.eb.assign(f->epsilonb, f->epsilonb + nb);
    s.R.assign(f->R, f->R + 6 * nc); s.Rb.assign(f->Rb, f->Rb + 6 * nb);
    s.UbF.assign(nb, 0); s.kbF.assign(nb, 0);
    if (f->UbFixed) s.UbF.assign(f->UbFixed, f->UbFixed + nb);
    if (f->kbFixed) s.kbF.assign(f->kbFixed, f->kbFixed + nb);
    s.have = true;
    return PDFB200_OK;
}
int pdforacle_upload_fv_fields(pdforacle_cloud *h, const pdfb200_fv_fields *f);
int pdforacle_commit_fv_fields(pdforacle_cloud *h) {
    StagedFv &s = g_staged;
    if (!s.have) { g_err = "commit_fv_fields: nothing staged (stage_fv_fields first)"; return PDFB200_ERR_STATE; }
    pdfb200_fv_fields f{};
    f.U = s.U.data(); f.Ub = s.Ub.data(); f.UbFixed = s.UbF.data(); f.p = s.p.data(); f.pb = s.pb.data();
    f.k = s.k.data(); f.kb = s.kb.data(); f.kbFixed = s.kbF.data(); f.epsilon = s.e.data(); f.epsilonb = s.eb.data();
    f.R = s.R.data(); f.Rb = s.Rb.data();
    s.have = false;
    return pdforacle_upload_fv_fields(h, &f);
}
int pdforacle_upload_cloud_fields(pdforacle_cloud *h, const pdfb200_cloud_fields *f) {
    ORACLE_TRY
    Cloud &c = h->c;
    const int nc = c.mesh.nCells, nb = c.nBnd();
    copyScalar(c.rhocPdf, f->rho, f->rhob, f->rhobFixed, nc, nb);
    c.PhicPdf.resize(c.prm.nPhi); c.PhiPhicPdf.resize(c.nCov());
    for (int i = 0; i < c.prm.nPhi; ++i) copyScalar(c.PhicPdf[i], f->Phi[i], f->Phib[i], f->PhibFixed[i], nc, nb);
    for (int i = 0; i < c.nCov(); ++i) {
        if (f->Cov[i]) copyScalar(c.PhiPhicPdf[i], f->Cov[i], f->Covb[i], f->CovbFixed[i], nc, nb);
        else { c.PhiPhicPdf[i].c.assign(nc, 0.0); c.PhiPhicPdf[i].b.assign(nb, 0.0); c.PhiPhicPdf[i].fixed.assign(nb, 0); }
    }
    c.haveCloudFields = true;
    ORACLE_CATCH
}

int pdforacle_upload_poscorr_fields(pdforacle_cloud *h, const pdfb200_poscorr_fields *f) {
    ORACLE_TRY
    Cloud &c = h->c;
    const int nc = c.mesh.nCells, nb = c.nBnd();
    auto vec = [&](VectorField &F, const double *cv, const double *bv) {
        F.fixed.assign(nb, 1);   // boundary values are taken as uploaded
        F.resize(nc, nb);
        if (!cv) return;
        if (!bv) throw std::invalid_argument("upload_poscorr_fields: cell values without boundary values");
        for (int i = 0; i < nc; ++i) F.c[i] = Vec3(cv[3 * i], cv[3 * i + 1], cv[3 * i + 2]);
        for (int i = 0; i < nb; ++i) F.b[i] = Vec3(bv[3 * i], bv[3 * i + 1], bv[3 * i + 2]);
    };
    auto sca = [&](ScalarField &F, const double *cv, const double *bv) {
        F.fixed.assign(nb, 1);
        F.resize(nc, nb);
        if (!cv) return;
        if (!bv) throw std::invalid_argument("upload_poscorr_fields: cell values without boundary values");
        F.c.assign(cv, cv + nc); F.b.assign(bv, bv + nb);
    };
    vec(c.pcG0, f->G0, f->G0b); vec(c.pcG1, f->G1, f->G1b); vec(c.pcG2, f->G2, f->G2b);
    sca(c.pcL, f->l, f->lb); sca(c.pcZeta, f->zeta, f->zetab);
    c.pcScale = f->scale;
    c.havePcExt = true;
    ORACLE_CATCH
}

int pdforacle_init_moments(pdforacle_cloud *h) {
    if (!h->c.haveFv || !h->c.haveCloudFields) { g_err = "init_moments: upload fields first"; return PDFB200_ERR_STATE; }
    h->c.initMoments();
    return PDFB200_OK;
}

int pdforacle_seed_particles(pdforacle_cloud *h) {
    ORACLE_TRY
    if (!h->c.haveMoments) throw std::runtime_error("seed_particles: init_moments first");
    h->c.seedParticles();
    ORACLE_CATCH
}

int pdforacle_upload_particles(pdforacle_cloud *h, int64_t n, const pdfb200_particles *p) {
    ORACLE_TRY
    Cloud &c = h->c;
    c.particles.clear(); c.particles.resize(n);
    for (int64_t i = 0; i < n; ++i) {
        Particle &q = c.particles[i];
        q.position = Vec3(p->pos[3 * i], p->pos[3 * i + 1], p->pos[3 * i + 2]);
        q.UParticle = Vec3(p->U[3 * i], p->U[3 * i + 1], p->U[3 * i + 2]);
        q.m = p->m[i]; q.Omega = p->Omega[i]; q.rho = p->rho[i]; q.eta = p->eta[i];
        for (int k = 0; k < c.prm.nPhi; ++k) q.Phi[k] = p->Phi[k][i];
        q.cell = p->cell[i];
        q.tet = p->tet ? p->tet[i] : c.mesh.locate(q.position, q.cell);
        q.id = p->id ? p->id[i] : (uint64_t)i;
    }
    // the next free id: above every id in use and congruent to the rank (ids advance in steps of nRanks); for the
    // default ids 0..n-1 of one rank this is n
    uint64_t idEnd = 0;
    for (const Particle &q : c.particles) idEnd = std::max<uint64_t>(idEnd, q.id + 1);
    const uint64_t nr = (uint64_t)std::max(c.nRanks, 1);
    c.nextId = (idEnd + nr - 1) / nr * nr + (uint64_t)c.rank;
    ORACLE_CATCH
}

// the CPU has nothing to overlap: the "asynchronous" download copies at once
int pdforacle_download_cell_fields(pdforacle_cloud *h, pdfb200_cell_fields *o);
int pdforacle_download_cell_fields_async(pdforacle_cloud *h, const pdfb200_cell_fields *o) {
    pdfb200_cell_fields tmp = *o;
    return pdforacle_download_cell_fields(h, &tmp);
}
int pdforacle_sync_downloads(pdforacle_cloud *) { return PDFB200_OK; }

// diagnostics: histogram of the tet faces crossed by every particle in the full-step walk of the last evolve
// (bin k: k crossings, last bin: >= nBins-1) and the largest count
int pdforacle_debug_crossings(pdforacle_cloud *h, int nBins, int64_t *hist, int *maxSteps) {
    int mx = 0;
    for (int k = 0; k < nBins; ++k) hist[k] = 0;
    for (const Particle &p : h->c.particles) { mx = std::max(mx, p.nSteps); ++hist[std::min(p.nSteps, nBins - 1)]; }
    *maxSteps = mx;
    return PDFB200_OK;
}

int64_t pdforacle_num_particles(pdforacle_cloud *h) { return (int64_t)h->c.particles.size(); }

int pdforacle_download_particles(pdforacle_cloud *h, int64_t maxn, pdfb200_particles *out, int64_t *n) {
    Cloud &c = h->c;
    const int64_t np = (int64_t)c.particles.size();
    if (np > maxn) { g_err = "download_particles: buffer too small"; return PDFB200_ERR_CAPACITY; }
    for (int64_t i = 0; i < np; ++i) {
        const Particle &q = c.particles[i];
        if (out->pos) { out->pos[3 * i] = q.position.x; out->pos[3 * i + 1] = q.position.y; out->pos[3 * i + 2] = q.position.z; }
        if (out->U) { out->U[3 * i] = q.UParticle.x; out->U[3 * i + 1] = q.UParticle.y; out->U[3 * i + 2] = q.UParticle.z; }
        if (out->m) out->m[i] = q.m;
        if (out->Omega) out->Omega[i] = q.Omega;
        if (out->rho) out->rho[i] = q.rho;
        if (out->eta) out->eta[i] = q.eta;
        for (int k = 0; k < c.prm.nPhi; ++k) if (out->Phi[k]) out->Phi[k][i] = q.Phi[k];
        if (out->cell) out->cell[i] = q.cell;
        if (out->tet) out->tet[i] = q.tet;
        if (out->id) out->id[i] = q.id;
    }
    *n = np;
    return PDFB200_OK;
}

int pdforacle_evolve(pdforacle_cloud *h, int32_t nSteps, pdfb200_step_report *reports) {
    ORACLE_TRY
    if (!h->c.haveMoments) throw std::runtime_error("evolve: fields/moments not initialised");
    for (int s = 0; s < nSteps; ++s) h->c.evolve(reports ? reports + s : nullptr);
    ORACLE_CATCH
}

// split step for particle-sharded runs: the caller sums `block` (and maxes
// rep->{UMax,UcorrMax,etaMax,CoMax}, mins etaMin) across ranks in between.
int64_t pdforacle_moment_block_size(pdforacle_cloud *h) { return h->c.momentBlockSize(); }
int pdforacle_evolve_begin(pdforacle_cloud *h, pdfb200_step_report *rep, double *block) {
    ORACLE_TRY
    std::vector<double> b;
    h->c.evolveBegin(rep, b);
    std::memcpy(block, b.data(), b.size() * sizeof(double));
    ORACLE_CATCH
}
int pdforacle_evolve_end(pdforacle_cloud *h, pdfb200_step_report *rep, const double *block) {
    ORACLE_TRY
    std::vector<double> b(block, block + h->c.momentBlockSize());
    h->c.evolveEnd(rep, b);
    ORACLE_CATCH
}

int pdforacle_download_cell_fields(pdforacle_cloud *h, pdfb200_cell_fields *o) {
    Cloud &c = h->c;
    const int nc = c.mesh.nCells;
    auto cp = [&](double *dst, const std::vector<double> &src) { if (dst) std::memcpy(dst, src.data(), nc * sizeof(double)); };
    cp(o->rho, c.rhocPdf.c); cp(o->rhoInst, c.rhocPdfInst.c); cp(o->pnd, c.pndcPdf.c); cp(o->pndInst, c.pndcPdfInst.c);
    cp(o->k, c.kcPdf.c); cp(o->PaNIC, c.PaNIC); cp(o->mMom, c.mMom); cp(o->VMom, c.VMom);
    for (int i = 0; i < nc; ++i) {
        if (o->U) { o->U[3 * i] = c.UcPdf.c[i].x; o->U[3 * i + 1] = c.UcPdf.c[i].y; o->U[3 * i + 2] = c.UcPdf.c[i].z; }
        if (o->UMom) { o->UMom[3 * i] = c.UMom[i].x; o->UMom[3 * i + 1] = c.UMom[i].y; o->UMom[3 * i + 2] = c.UMom[i].z; }
        if (o->Tau) { const SymmTensor &t = c.TaucPdf.c[i]; double *d = o->Tau + 6 * i; d[0] = t.xx; d[1] = t.xy; d[2] = t.xz; d[3] = t.yy; d[4] = t.yz; d[5] = t.zz; }
        if (o->UUMom) { const SymmTensor &t = c.UUMom[i]; double *d = o->UUMom + 6 * i; d[0] = t.xx; d[1] = t.xy; d[2] = t.xz; d[3] = t.yy; d[4] = t.yz; d[5] = t.zz; }
    }
    for (int k = 0; k < c.prm.nPhi; ++k) { cp(o->Phi[k], c.PhicPdf[k].c); cp(o->PhiMom[k], c.PhiMom[k]); }
    for (int k = 0; k < c.nCov(); ++k) { cp(o->Cov[k], c.PhiPhicPdf[k].c); cp(o->PhiPhiMom[k], c.PhiPhiMom[k]); }
    for (int k = 0; k < c.prm.nPhi; ++k)
        for (int i = 0; i < nc; ++i) {
            if (o->UPhi[k]) { const Vec3 &v = c.UPhiFlux[k][i]; o->UPhi[k][3 * i] = v.x; o->UPhi[k][3 * i + 1] = v.y; o->UPhi[k][3 * i + 2] = v.z; }
            if (o->UPhiMom[k]) { const Vec3 &v = c.UPhiMom[k][i]; o->UPhiMom[k][3 * i] = v.x; o->UPhiMom[k][3 * i + 1] = v.y; o->UPhiMom[k][3 * i + 2] = v.z; }
        }
    return PDFB200_OK;
}

// restart from the moment fields of a previous run (mcParticleCloud.C:339-517, checkMoments :739-820)
int pdforacle_upload_moments(pdforacle_cloud *h, const pdfb200_cell_fields *m) {
    ORACLE_TRY
    Cloud &c = h->c;
    const int nc = c.mesh.nCells;
    bool complete = m->mMom && m->VMom && m->UMom && m->UUMom;
    for (int k = 0; k < c.prm.nPhi; ++k) complete = complete && m->PhiMom[k];
    for (int k = 0; k < c.nCov(); ++k) complete = complete && m->PhiPhiMom[k];
    if (!complete) throw std::invalid_argument("upload_moments: Moments are missing (use init_moments)");
    c.initMoments();   // sizes, boundary types of the derived fields
    c.mMom.assign(m->mMom, m->mMom + nc); c.VMom.assign(m->VMom, m->VMom + nc);
    for (int i = 0; i < nc; ++i) {
        c.UMom[i] = Vec3(m->UMom[3 * i], m->UMom[3 * i + 1], m->UMom[3 * i + 2]);
        const double *d = m->UUMom + 6 * i;
        SymmTensor t; t.xx = d[0]; t.xy = d[1]; t.xz = d[2]; t.yy = d[3]; t.yz = d[4]; t.zz = d[5];
        c.UUMom[i] = t;
    }
    for (int k = 0; k < c.prm.nPhi; ++k) c.PhiMom[k].assign(m->PhiMom[k], m->PhiMom[k] + nc);
    for (int k = 0; k < c.nCov(); ++k) c.PhiPhiMom[k].assign(m->PhiPhiMom[k], m->PhiPhiMom[k] + nc);
    for (int k = 0; k < c.prm.nPhi; ++k)
        for (int i = 0; i < nc; ++i)
            c.UPhiMom[k][i] = m->UPhiMom[k] ? Vec3(m->UPhiMom[k][3 * i], m->UPhiMom[k][3 * i + 1], m->UPhiMom[k][3 * i + 2])
                                            : (c.PhiMom[k][i] / std::max(c.mMom[i], SMALL)) * c.UMom[i];
    // the mean fields follow from the moments as in updateCloudPDF (:958-1008): existing weight 1, nothing new
    const std::vector<double> none(c.momentBlockSize(), 0.0);
    c.finaliseMoments(none, 1.0);
    c.rhocPdfInst.c = c.rhocPdf.c; c.pndcPdfInst.c = c.pndcPdf.c;
    if (m->PaNIC) c.PaNIC.assign(m->PaNIC, m->PaNIC + nc);
    ORACLE_CATCH
}
int pdforacle_set_graph_mode(pdforacle_cloud *, int) { return PDFB200_OK; }   // nothing to replay on the CPU
int64_t pdforacle_id_counter(pdforacle_cloud *h) { return (int64_t)h->c.nextId; }
int pdforacle_set_id_counter(pdforacle_cloud *h, int64_t id) { h->c.nextId = (uint64_t)id; return PDFB200_OK; }
int64_t pdforacle_step_counter(pdforacle_cloud *h) { return (int64_t)h->c.step; }
int pdforacle_set_step_counter(pdforacle_cloud *h, int64_t step) { h->c.step = (uint32_t)step; return PDFB200_OK; }

double pdforacle_delta_t(pdforacle_cloud *h) { return h->c.deltaT; }
int pdforacle_set_delta_t(pdforacle_cloud *h, double dt) { h->c.deltaT = dt; return PDFB200_OK; }

int pdforacle_mesh_info_get(pdforacle_cloud *h, pdfb200_mesh_info *info) {
    const Mesh &m = h->c.mesh;
    info->nTets = (int64_t)m.tets.size(); info->maxTetsPerCell = m.maxTetsPerCell; info->isAxiSymmetric = m.axisym;
    for (int k = 0; k < 3; ++k) { info->axis[k] = m.axis[k]; info->centrePlaneNormal[k] = m.centreNormal[k]; }
    info->openingAngle = m.openingAngle;
    return PDFB200_OK;
}

int pdforacle_download_tets(pdforacle_cloud *h, double *gradN, int32_t *nbr, int32_t *faceLabel, int32_t *ptB, int32_t *ptC, int32_t *cell) {
    const Mesh &m = h->c.mesh;
    for (size_t t = 0; t < m.tets.size(); ++t) {
        const Tet &T = m.tets[t];
        if (gradN) for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) gradN[9 * t + 3 * i + k] = T.gradN[i][k];
        if (nbr) for (int i = 0; i < 4; ++i) nbr[4 * t + i] = T.nbr[i];
        if (faceLabel) faceLabel[t] = T.face;
        if (ptB) ptB[t] = T.ptB;
        if (ptC) ptC[t] = T.ptC;
        if (cell) cell[t] = T.cell;
    }
    return PDFB200_OK;
}

int pdforacle_download_geometry(pdforacle_cloud *h, double *fc, double *fa, double *cc, double *cv, double *co, double *hn) {
    const Mesh &m = h->c.mesh;
    for (int f = 0; f < m.nFaces; ++f) for (int k = 0; k < 3; ++k) {
        if (fc) fc[3 * f + k] = m.faceCentres[f][k];
        if (fa) fa[3 * f + k] = m.faceAreas[f][k];
        if (co) co[3 * f + k] = m.courantCoeffs[f][k];
    }
    for (int c = 0; c < m.nCells; ++c) {
        if (cc) for (int k = 0; k < 3; ++k) cc[3 * c + k] = m.cellCentres[c][k];
        if (cv) cv[c] = m.cellVolumes[c];
        if (hn) hn[c] = m.hNum[c];
    }
    return PDFB200_OK;
}

// ---- probes for the known-answer tests -------------------------------------

// richTetrahedron of tet t: S (4x3), mag, gradN (4x3), vertices (4x3), local point pair
int pdforacle_tet_detail(pdforacle_cloud *h, int64_t t, double *S, double *magV, double *gradN4, double *verts, int32_t *ptPair) {
    const Tet &T = h->c.mesh.tets[t];
    const Vec3 *s[4] = {&T.Sa, &T.Sb, &T.Sc, &T.Sd};
    const Vec3 *v[4] = {&T.a, &T.b, &T.c, &T.d};
    for (int i = 0; i < 4; ++i) for (int k = 0; k < 3; ++k) {
        S[3 * i + k] = (*s[i])[k]; gradN4[3 * i + k] = T.gradN[i][k]; verts[3 * i + k] = (*v[i])[k];
    }
    *magV = T.vol; ptPair[0] = T.ptFirst; ptPair[1] = T.ptSecond;
    return PDFB200_OK;
}
// tetFacePointCellDecomposition::find with reference semantics; and the shared locate rule
int pdforacle_find_reference(pdforacle_cloud *h, int64_t n, const double *pts, const int32_t *cellHint, int32_t *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = h->c.mesh.findReference(Vec3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), cellHint[i]);
    return PDFB200_OK;
}
int pdforacle_locate(pdforacle_cloud *h, int64_t n, const double *pts, const int32_t *cell, int32_t *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = h->c.mesh.locate(Vec3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), cell[i]);
    return PDFB200_OK;
}
int pdforacle_tet_inside(pdforacle_cloud *h, int64_t t, const double *pt) { return h->c.mesh.tets[t].inside(Vec3(pt[0], pt[1], pt[2])); }

// cellPointFace interpolation of a scalar cell field at points (for linear-field exactness tests)
int pdforacle_interpolate_scalar(pdforacle_cloud *h, const double *cellValues, const double *bndValues, int64_t n,
                                 const double *pts, const int32_t *cell, double *out, double *gradOut) {
    ORACLE_TRY
    Cloud &c = h->c;
    ScalarField f; f.c.assign(cellValues, cellValues + c.mesh.nCells); f.b.assign(bndValues, bndValues + c.nBnd()); f.fixed.assign(c.nBnd(), 1);
    NodeField<double> nf;
    c.makeNodes(f, nf, PDFB200_INTERP_CELL_POINT_FACE);
    for (int64_t i = 0; i < n; ++i) {
        Particle p; p.position = Vec3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]); p.cell = cell[i]; p.tet = c.mesh.locate(p.position, p.cell);
        out[i] = c.interp(nf, p);
        if (gradOut) { Vec3 g = c.gradInterp(nf, p); gradOut[3 * i] = g.x; gradOut[3 * i + 1] = g.y; gradOut[3 * i + 2] = g.z; }
    }
    ORACLE_CATCH
}

int pdforacle_find_coeffs(int32_t n, const double *lst, double v, int32_t *i, double *w) {
    ORACLE_TRY
    std::vector<double> l(lst, lst + n); int ii; double ww;
    findCoeffs(l, v, ii, ww); *i = ii; *w = ww;
    ORACLE_CATCH
}
double pdforacle_binterp(const double *v, int32_t nRows, int32_t nCols, int32_t ix, double wx, int32_t iy, double wy) {
    std::vector<double> vv(v, v + (size_t)nRows * nCols);
    return binterp(vv, nCols, ix, wx, iy, wy);
}

// mcInletRandom probes: out = {Q, m, PDF(x), CDF(x), value(u)}
int pdforacle_inlet_random(double UMean, double uRms, double x, double u, double *out) {
    InletRandom r; r.updateCoeffs(UMean, uRms);
    out[0] = r.Q; out[1] = r.m; out[2] = r.PDF(x); out[3] = r.CDF(x); out[4] = r.value(u);
    return PDFB200_OK;
}
// "sampling" generator probes: keyed (face bf, draws 0..n-1 of step `step`) and sequential
int pdforacle_inlet_sampling(double UMean, double uRms, uint64_t seed, uint32_t bf, uint32_t step, int64_t n, double *keyed, double *sequential) {
    InletRandom r; r.updateCoeffs(UMean, uRms);
    KeyedRng k; k.setSeed(seed);
    SeqRandom s; s.seed(seed);
    for (int64_t i = 0; i < n; ++i) {
        if (keyed) keyed[i] = r.sample(k, bf, step, (uint32_t)i);
        if (sequential) sequential[i] = r.sample(s);
    }
    return PDFB200_OK;
}
int pdforacle_inlet_random_sample(double UMean, double uRms, uint64_t seed, int64_t n, double *out) {
    InletRandom r; r.updateCoeffs(UMean, uRms);
    SeqRandom s; s.seed(seed);
    for (int64_t i = 0; i < n; ++i) out[i] = r.value(s.scalar01());
    return PDFB200_OK;
}

// RNG probe (keyed Philox): for each entity two uniforms and four normals of (entity, step, purpose, sub)
int pdforacle_rng_probe(uint64_t seed, int64_t n, const uint64_t *entity, uint32_t step, uint32_t purpose, uint32_t sub, double *out6) {
    KeyedRng k; k.setSeed(seed);
    for (int64_t i = 0; i < n; ++i) {
        k.uniform2(entity[i], step, purpose, sub, out6[6 * i], out6[6 * i + 1]);
        k.normal4(entity[i], step, purpose, sub, out6 + 6 * i + 2);
    }
    return PDFB200_OK;
}
int pdforacle_normal_pair(uint32_t a, uint32_t b, float *out2) { normalPairF32(a, b, out2[0], out2[1]); return PDFB200_OK; }
int pdforacle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); return PDFB200_OK; }

// rotationTensor probe
int pdforacle_rotation_tensor(const double *n1, const double *n2, double *T9) {
    Tensor t = rotationTensor(Vec3(n1[0], n1[1], n1[2]), Vec3(n2[0], n2[1], n2[2]));
    std::memcpy(T9, t.m, sizeof(t.m));
    return PDFB200_OK;
}

}  // extern "C"
