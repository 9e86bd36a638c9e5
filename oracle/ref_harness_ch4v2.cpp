// TEST INFRASTRUCTURE -- not part of the product path.
//
// C-ABI harness around the UNMODIFIED reference ch4/v2 sources (/root/reference/ch4/v2/*.cpp minus main.cpp and
// Instantiator.cpp): the fixed-weight MC_MEX_Ionization of BASELINE config 3 (ch4/v2/Interactions.cpp:476-735).
// Compiled together with those sources, where they lie, by oracle/Makefile into oracle/_ref/libref_ch4v2.so.
// Nothing here re-implements reference arithmetic: every entry point constructs reference objects and calls
// reference methods.  Only tests/ may load the resulting library.
//
// Array conventions as in ref_harness_v3.cpp: node fields double[nv] in Field order, vector fields xyz-interleaved,
// particles double[7*n] AoS x y z u v w mpw.
#include <cstring>
#include <memory>
#include <vector>
#include <string>
#include <iostream>
#include "all.h"
#include "World.h"
#include "Species.h"
#include "Interactions.h"
#include "Rnd.h"
#include "Object.h"

#define API extern "C" __attribute__((visibility("default")))
extern Rnd rnd;                                                                   // ch4/v2/Rnd.cpp:18

API void refv2_seed(unsigned seed) { rnd = Rnd(seed); }
API double refv2_rnd() { return rnd(); }

API World* refv2_world_create(int ni, int nj, int nk, const double* x0, const double* xm) {
    return new World(ni, nj, nk, type_calc3(x0[0], x0[1], x0[2]), type_calc3(xm[0], xm[1], xm[2]));
}
API void refv2_world_destroy(World* w) { delete w; }
API void refv2_world_set_time(World* w, double dt, int num_ts) { w->setTime(dt, num_ts); }
API void refv2_world_add_rectangle(World* w, const double* c, double phi, const double* sides) {
    w->addObject<Rectangle>(type_calc3(c[0], c[1], c[2]), phi, type_calc3(sides[0], sides[1], sides[2]));
}
API void refv2_world_compute_object_id(World* w) { w->computeObjectID(); }
API void refv2_world_set_ef(World* w, const double* in) {
    size_t u = 0;
    for (int i = 0; i < w->ef.ni; i++) for (int j = 0; j < w->ef.nj; j++) for (int k = 0; k < w->ef.nk; k++)
        for (int c = 0; c < 3; c++) w->ef[i][j][k][c] = in[u++];
}

API Species* refv2_species_create(World* w, const char* name, double mass, double charge, double mpw0, double E_ion) {
    return new Species(name, mass, charge, *w, mpw0, E_ion);
}
API void refv2_species_destroy(Species* s) { delete s; }
API size_t refv2_species_count(Species* s) { return s->getNumParticles(); }
API void refv2_species_set_particles(Species* s, size_t n, const double* a) {     // raw store, no addParticle filtering
    std::vector<Particle>& p = s->getPartRef();
    p.clear(); p.reserve(n);
    for (size_t i = 0; i < n; i++, a += 7) p.emplace_back(a[0], a[1], a[2], a[3], a[4], a[5], a[6]);
}
API void refv2_species_get_particles(Species* s, double* a) {
    for (const Particle& p : s->getConstPartRef()) {
        a[0] = p.pos[0]; a[1] = p.pos[1]; a[2] = p.pos[2];
        a[3] = p.vel[0]; a[4] = p.vel[1]; a[5] = p.vel[2]; a[6] = p.macro_weight; a += 7;
    }
}
API void refv2_species_add_particle(Species* s, const double* a) {                // ch4/v2/Species.cpp:226-237
    s->addParticle(type_calc3(a[0], a[1], a[2]), type_calc3(a[3], a[4], a[5]), a[6]);
}
API void refv2_species_compute_macro_count(Species* s) { s->computeMacroParticlesCount(); }   // map_indexes reads it (:740-776)

API MC_MEX_Ionization* refv2_mcc_create(Species* neutrals, Species* ions, Species* electrons, World* w, const char* table_path) {
    try { return new MC_MEX_Ionization(*neutrals, *ions, *electrons, *w, table_path); }
    catch (const std::exception& e) { std::cerr << "refv2_mcc_create: " << e.what() << "\n"; return nullptr; }
}
API void refv2_mcc_destroy(MC_MEX_Ionization* m) { delete m; }
API void refv2_mcc_apply(MC_MEX_Ionization* m, double dt) { m->apply(dt); }       // ch4/v2/Interactions.cpp:566-641
namespace {
struct MccAccess : public MC_MEX_Ionization {
    using MC_MEX_Ionization::evaluateSigmaColl; using MC_MEX_Ionization::evaluateSigmaIon;
    using MC_MEX_Ionization::sigma_v_rel_max; using MC_MEX_Ionization::collide;
};
}
API double refv2_mcc_sigma_coll(MC_MEX_Ionization* m, double E) { return static_cast<MccAccess*>(m)->evaluateSigmaColl(E); }
API double refv2_mcc_sigma_ion(MC_MEX_Ionization* m, double E) { return static_cast<MccAccess*>(m)->evaluateSigmaIon(E); }
API double refv2_mcc_get_sv_max(MC_MEX_Ionization* m) { return static_cast<MccAccess*>(m)->sigma_v_rel_max; }
API void   refv2_mcc_set_sv_max(MC_MEX_Ionization* m, double v) { static_cast<MccAccess*>(m)->sigma_v_rel_max = v; }
// one collide() call (:678-735): in vel_neu[3], vel_ele[3] (updated), out ionised flag, vel_new[3]
API int refv2_mcc_collide(MC_MEX_Ionization* m, double* vn, double* ve, double* vnew, int n_atoms, double sigma_coll) {
    type_calc3 a(vn[0], vn[1], vn[2]), b(ve[0], ve[1], ve[2]), c; bool ion = false;
    static_cast<MccAccess*>(m)->collide(a, b, ion, c, n_atoms, sigma_coll);
    for (int i = 0; i < 3; i++) { vn[i] = a[i]; ve[i] = b[i]; vnew[i] = c[i]; }
    return ion;
}
