// TEST INFRASTRUCTURE -- not part of the product path.
//
// C-ABI harness around the UNMODIFIED reference ch4/v3 sources
// (/root/reference/ch4/v3/src/*.cpp minus main.cpp).  It is compiled together
// with those sources, where they lie, by oracle/Makefile into
// oracle/_ref/libref_v3.so.  Nothing here re-implements reference arithmetic:
// every entry point constructs reference objects and calls reference methods.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
// arm may load the resulting library.
//
// Array conventions (all little-endian, host):
//   node fields   : double[nv], index (i*nj+j)*nk+k  (Field<T> order, v3/Field.h:16,88)
//   vector fields : double[3*nv], interleaved xyz per node
//   particles     : double[7*n] AoS  x y z u v w mpw  (Particle, v3/Species.h:12-29)
#include <cstring>
#include <memory>
#include <vector>
#include <string>
#include <sstream>
#include <fstream>
#include "all.h"
#include "World.h"
#include "Species.h"
#include "PotentialSolver.h"
#include "Interactions.h"
#include "Source.h"
#include "Config.h"
#include "Rnd.h"
#include "Object.h"

#define API extern "C" __attribute__((visibility("default")))

namespace {
struct SpeciesAccess : public Species {      // reach the protected moment sums
    using Species::n_sum; using Species::nv_sum;
    using Species::nuu_sum; using Species::nvv_sum; using Species::nww_sum;
};
template <class F> void put_scalar(const F& f, double* out) {
    size_t u = 0;
    for (int i = 0; i < f.ni; i++) for (int j = 0; j < f.nj; j++) for (int k = 0; k < f.nk; k++) out[u++] = (double)f[i][j][k];
}
template <class F> void get_scalar(F& f, const double* in) {
    size_t u = 0;
    for (int i = 0; i < f.ni; i++) for (int j = 0; j < f.nj; j++) for (int k = 0; k < f.nk; k++) f[i][j][k] = in[u++];
}
void put_vec(const Field<type_calc3>& f, double* out) {
    size_t u = 0;
    for (int i = 0; i < f.ni; i++) for (int j = 0; j < f.nj; j++) for (int k = 0; k < f.nk; k++)
        for (int c = 0; c < 3; c++) out[u++] = f[i][j][k][c];
}
void get_vec(Field<type_calc3>& f, const double* in) {
    size_t u = 0;
    for (int i = 0; i < f.ni; i++) for (int j = 0; j < f.nj; j++) for (int k = 0; k < f.nk; k++)
        for (int c = 0; c < 3; c++) f[i][j][k][c] = in[u++];
}
}

// ---------------------------------------------------------------- global state
API void refv3_seed(unsigned seed) { rnd = Rnd(seed); }                       // v3/Rnd.cpp:7
API double refv3_rnd() { return rnd(); }
API void refv3_config(int subcycling, int multithreading, int num_threads, int merging, int sputtering) {
    Config& c = Config::getInstance();
    c.setSUBCYCLING(subcycling); c.setMULTITHREADING(multithreading);
    if (num_threads > 0) c.setNUM_THREADS(num_threads);
    c.setMERGING(merging); c.setSPUTTERING(sputtering);
}
API int refv3_num_threads() { return (int)Config::getInstance().getNUM_THREADS(); }

// ---------------------------------------------------------------------- World
API World* refv3_world_create(int ni, int nj, int nk, const double* x0, const double* xm) {
    return new World(ni, nj, nk, type_calc3(x0[0], x0[1], x0[2]), type_calc3(xm[0], xm[1], xm[2]));
}
API void refv3_world_destroy(World* w) { delete w; }
API void refv3_world_set_time(World* w, double dt, int num_ts) { w->setTime(dt, num_ts); }
API int  refv3_world_advance_time(World* w) { return w->advanceTime(); }
API void refv3_world_add_rectangle(World* w, const double* c, double phi, const double* sides) {
    w->addObject<Rectangle>(type_calc3(c[0], c[1], c[2]), phi, type_calc3(sides[0], sides[1], sides[2]));
}
API void refv3_world_add_sphere(World* w, const double* c, double phi, double r) {
    w->addObject<Sphere>(type_calc3(c[0], c[1], c[2]), phi, r);
}
API void refv3_world_compute_object_id(World* w) { w->computeObjectID(); }
API void refv3_world_add_inlet(World* w, const char* face) { w->addInlet(face); }
API int  refv3_world_in_object(World* w, const double* p) { return w->inObject(type_calc3(p[0], p[1], p[2])); }
API int  refv3_world_in_bounds(World* w, const double* p) { return w->inBounds(type_calc3(p[0], p[1], p[2])); }
API void refv3_world_line_intersect(World* w, const double* x1, const double* x2, int in_object, double* tp, double* pos, double* n) {
    type_calc t; type_calc3 P, N;
    w->lineIntersect(type_calc3(x1[0], x1[1], x1[2]), type_calc3(x2[0], x2[1], x2[2]), in_object, t, P, N);
    *tp = t; for (int c = 0; c < 3; c++) { pos[c] = P[c]; n[c] = N[c]; }
}
API double refv3_world_get_pe(World* w) { return w->getPE(); }
// field ids: 0 phi 1 rho 2 node_vol 3 ef(3) 4 object_id 5 node_type
API void refv3_world_get_field(World* w, int id, double* out) {
    switch (id) {
        case 0: put_scalar(w->phi, out); break;
        case 1: put_scalar(w->rho, out); break;
        case 2: put_scalar(w->node_vol, out); break;
        case 3: put_vec(w->ef, out); break;
        case 4: put_scalar(w->object_id, out); break;
        case 5: put_scalar(w->node_type, out); break;
    }
}
API void refv3_world_set_field(World* w, int id, const double* in) {
    switch (id) {
        case 0: get_scalar(w->phi, in); break;
        case 1: get_scalar(w->rho, in); break;
        case 3: get_vec(w->ef, in); break;
        case 4: get_scalar(w->object_id, in); break;
    }
}

// -------------------------------------------------------------------- Species
API Species* refv3_species_create(World* w, const char* name, double mass, double charge, double mpw0, double E_ion) {
    return new Species(name, mass, charge, *w, mpw0, E_ion);
}
API void refv3_species_destroy(Species* s) { delete s; }
API size_t refv3_species_count(Species* s) { return s->getNumParticles(); }
API void refv3_species_set_particles(Species* s, size_t n, const double* a) {   // raw store, no addParticle filtering
    std::vector<Particle>& p = s->getPartRef();
    p.clear(); p.reserve(n);
    for (size_t i = 0; i < n; i++, a += 7) p.emplace_back(a[0], a[1], a[2], a[3], a[4], a[5], a[6]);
    s->setSorted(false);
}
API void refv3_species_get_particles(Species* s, double* a) {
    for (const Particle& p : s->getConstPartRef()) {
        a[0] = p.pos[0]; a[1] = p.pos[1]; a[2] = p.pos[2];
        a[3] = p.vel[0]; a[4] = p.vel[1]; a[5] = p.vel[2]; a[6] = p.macro_weight; a += 7;
    }
}
API void refv3_species_add_particle(Species* s, const double* a) {              // v3/Species.cpp:420-434 (filter + half-step rewind)
    s->addParticle(type_calc3(a[0], a[1], a[2]), type_calc3(a[3], a[4], a[5]), a[6]);
}
API void refv3_species_load_box_thermal(Species* s, const double* x0, const double* sides, double den, double T) {
    s->loadParticleBoxThermal(type_calc3(x0[0], x0[1], x0[2]), type_calc3(sides[0], sides[1], sides[2]), den, T);
}
API void refv3_species_advance_electrons(Species* s, double dt) { s->advanceElectrons(dt); }          // :258-317
API void refv3_species_advance_non_electron(Species* s, Species* neutrals, Species* spherium, double dt) {
    s->advanceNonElectron(*neutrals, *spherium, dt);                                                   // :47-77
}
API void refv3_species_compute_number_density(Species* s) { s->computeNumberDensity(); }               // :401-416
API void refv3_species_sample_moments(Species* s) { s->sampleMoments(); }                              // :767-776
API void refv3_species_compute_gas_properties(Species* s) { s->computeGasProperties(); }
API void refv3_species_clear_samples(Species* s) { s->clearSamples(); }
API void refv3_species_update_averages(Species* s) { s->updateAverages(); }
API void refv3_species_compute_macro_count(Species* s) { s->computeMacroParticlesCount(); }            // :813-819
API void refv3_species_merge(Species* s) { s->merge(); }
API double refv3_species_ke(Species* s) { return s->getKE(); }
API double refv3_species_micro_count(Species* s) { return s->getMicroCount(); }
API void refv3_species_momentum(Species* s, double* m) { type_calc3 v = s->getMomentum(); m[0] = v[0]; m[1] = v[1]; m[2] = v[2]; }
API void refv3_species_sample_v3th(Species* s, double T, double* v) { type_calc3 r = s->sampleV3th(T); v[0] = r[0]; v[1] = r[1]; v[2] = r[2]; }
API void refv3_species_sample_reflected(Species* s, const double* pos, double vmag, const double* n, double* v) {
    type_calc3 r = s->sampleReflectedVelocity(type_calc3(pos[0], pos[1], pos[2]), vmag, type_calc3(n[0], n[1], n[2]));
    v[0] = r[0]; v[1] = r[1]; v[2] = r[2];
}
// per-cell index lists flattened: counts[num_cells]; returns total
API size_t refv3_species_sort_counts(Species* s, World* w, int* counts) {
    std::vector<std::vector<int>> v = s->sortIndexes();
    size_t tot = 0;
    for (int c = 0; c < w->num_cells; c++) { counts[c] = (int)v[c].size(); tot += v[c].size(); }
    return tot;
}
// field ids: 0 den 1 den_avg 2 T 3 vel(3) 4 macro_part_count(cells) 5 n_sum 6 nv_sum(3) 7 nuu 8 nvv 9 nww
API void refv3_species_get_field(Species* s, int id, double* out) {
    SpeciesAccess* a = static_cast<SpeciesAccess*>(s);
    switch (id) {
        case 0: put_scalar(s->den, out); break;
        case 1: put_scalar(s->den_avg, out); break;
        case 2: put_scalar(s->T, out); break;
        case 3: put_vec(s->vel, out); break;
        case 4: put_scalar(s->macro_part_count, out); break;
        case 5: put_scalar(a->n_sum, out); break;
        case 6: put_vec(a->nv_sum, out); break;
        case 7: put_scalar(a->nuu_sum, out); break;
        case 8: put_scalar(a->nvv_sum, out); break;
        case 9: put_scalar(a->nww_sum, out); break;
    }
}
API void refv3_world_compute_charge_density(World* w, Species** sp, int n) {   // v3/World.cpp:193-200
    // the reference takes std::vector<Species>&; Species is copy-constructible, but copying
    // 1e6-particle stores per call would distort timings, so build rho exactly as the body does.
    w->rho = 0;
    for (int i = 0; i < n; i++) { if (sp[i]->charge == 0) continue; w->rho += sp[i]->charge * sp[i]->den; }
}

// --------------------------------------------------------------------- Solver
API PotentialSolver* refv3_solver_create(World* w, unsigned max_it, double tol, int type) {
    return new PotentialSolver(*w, max_it, tol, (SolverType)type);
}
API void refv3_solver_destroy(PotentialSolver* s) { delete s; }
API void refv3_solver_set_reference(PotentialSolver* s, double phi0, double n0, double Te0) { s->setReferenceValues(phi0, n0, Te0); }
API int  refv3_solver_solve_gs(PotentialSolver* s) { return s->solveGS(); }       // v3/PotentialSolver.cpp:69-166
API int  refv3_solver_solve(PotentialSolver* s) { return s->solve(); }
API void refv3_solver_compute_ef(PotentialSolver* s) { s->computeEF(); }          // :354-408

// ------------------------------------------------------------ MC_MEX_Ionization
API MC_MEX_Ionization* refv3_mcc_create(Species* neutrals, Species* ions, Species* electrons, World* w, const char* table_path) {
    try { return new MC_MEX_Ionization(*neutrals, *ions, *electrons, *w, table_path); }
    catch (const std::exception& e) { std::cerr << "refv3_mcc_create: " << e.what() << "\n"; return nullptr; }
}
API void refv3_mcc_destroy(MC_MEX_Ionization* m) { delete m; }
API void refv3_mcc_apply(MC_MEX_Ionization* m, double dt) { m->apply(dt); }        // v3/Interactions.cpp:567-598
namespace {
struct MccAccess : public MC_MEX_Ionization {
    using MC_MEX_Ionization::evaluateSigmaColl; using MC_MEX_Ionization::evaluateSigmaIon;
    using MC_MEX_Ionization::W_sigma_v_rel_max; using MC_MEX_Ionization::collide;
};
}
API double refv3_mcc_sigma_coll(MC_MEX_Ionization* m, double E) { return static_cast<MccAccess*>(m)->evaluateSigmaColl(E); }
API double refv3_mcc_sigma_ion(MC_MEX_Ionization* m, double E) { return static_cast<MccAccess*>(m)->evaluateSigmaIon(E); }
API double refv3_mcc_get_wsv_max(MC_MEX_Ionization* m) { return static_cast<MccAccess*>(m)->W_sigma_v_rel_max; }
API void   refv3_mcc_set_wsv_max(MC_MEX_Ionization* m, double v) { static_cast<MccAccess*>(m)->W_sigma_v_rel_max = v; }
// one collide() call: in vel_neu[3], vel_ele[3] (updated), out ionised flag, vel_new[3]
API int refv3_mcc_collide(MC_MEX_Ionization* m, double* vn, double* ve, double* vnew, double sigma_coll) {
    type_calc3 a(vn[0], vn[1], vn[2]), b(ve[0], ve[1], ve[2]), c; bool ion = false;
    static_cast<MccAccess*>(m)->collide(a, b, ion, c, sigma_coll);
    for (int i = 0; i < 3; i++) { vn[i] = a[i]; ve[i] = b[i]; vnew[i] = c[i]; }
    return ion;
}

// --------------------------------------------------------------------- DSMC_MEX
namespace {
struct DsmcAccess : public DSMC_MEX {
    using DSMC_MEX::evaluateSigma; using DSMC_MEX::sigma_v_rel_max; using DSMC_MEX::collide;
};
}
API DSMC_MEX* refv3_dsmc_create(Species* s1, Species* s2, World* w) {              // v3/Interactions.cpp:143-176
    try { return s2 ? new DSMC_MEX(*s1, *s2, *w) : new DSMC_MEX(*s1, *w); }
    catch (const std::exception& e) { std::cerr << "refv3_dsmc_create: " << e.what() << "\n"; return nullptr; }
}
API void refv3_dsmc_destroy(DSMC_MEX* m) { delete m; }
API void refv3_dsmc_apply(DSMC_MEX* m, double dt) { m->apply(dt); }                // :183-265
API double refv3_dsmc_sigma(DSMC_MEX* m, double v_rel) { return static_cast<DsmcAccess*>(m)->evaluateSigma(v_rel); }
API double refv3_dsmc_get_sigma_v_max(DSMC_MEX* m) { return static_cast<DsmcAccess*>(m)->sigma_v_rel_max; }
API void   refv3_dsmc_set_sigma_v_max(DSMC_MEX* m, double v) { static_cast<DsmcAccess*>(m)->sigma_v_rel_max = v; }
API void refv3_dsmc_collide(DSMC_MEX* m, double* v1, double* v2) {                 // :267-285
    type_calc3 a(v1[0], v1[1], v1[2]), b(v2[0], v2[1], v2[2]);
    static_cast<DsmcAccess*>(m)->collide(a, b);
    for (int i = 0; i < 3; i++) { v1[i] = a[i]; v2[i] = b[i]; }
}

// -------------------------------------------------------------------- Sources
API Source* refv3_source_cold(Species* s, World* w, double v_drift, double den, const char* face) {
    return new ColdBeamSource(*s, *w, v_drift, den, face);
}
API Source* refv3_source_warm(Species* s, World* w, double v_drift, double den, double T, const char* face) {
    return new WarmBeamSource(*s, *w, v_drift, den, T, face);
}
API void refv3_source_destroy(Source* s) { delete s; }
API void refv3_source_sample(Source* s) { s->sample(); }
