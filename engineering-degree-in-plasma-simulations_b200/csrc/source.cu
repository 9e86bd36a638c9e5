// Particle creation on the device: thermal box loader and face inlet sources.
//   k_gen_box_thermal : Species::loadParticleBoxThermal            ch4/v3/src/Species.cpp:560-598
//   k_gen_source      : ColdBeamSource / WarmBeamSource::sample    ch4/v3/src/Source.cpp:38-103,111-191
// Both generate the candidate particles into an AoS staging buffer with counter-based Philox streams and
// hand them to the addParticle kernel (filter NaN / out of bounds / in object, half-step rewind,
// stream-compacted append; species.cu).
#include "common.cuh"
#include "samplers.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

using namespace picg;

namespace picg {
int species_add_staged(picg_species_s* s, size_t n, const double* d_aos);     // species.cu
}

__global__ void __launch_bounds__(256) k_gen_box_thermal(size_t n, size_t first, uint64_t seed, uint32_t stream, uint32_t call,
                                                         double x0, double y0, double z0, double x1, double y1, double z1,
                                                         double T, double mass, double mpw, double* __restrict__ aos) {
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        PhiloxStream r; r.init(seed, stream, first + t, call);
        double* a = aos + t * 7;
        a[0] = x0 + r.next() * (x1 - x0);          // rnd(min,max) = min + rnd()*(max-min)   Rnd.cpp:14-16
        a[1] = y0 + r.next() * (y1 - y0);
        a[2] = z0 + r.next() * (z1 - z0);
        double v[3]; sample_v3th(r, T, mass, v);
        a[3] = v[0]; a[4] = v[1]; a[5] = v[2]; a[6] = mpw;
    }
}

struct SourceGeom { int face; double x0[3]; double L[3]; double v_drift, T, mass, mpw; };
__global__ void __launch_bounds__(256) k_gen_source(size_t n, uint64_t seed, uint32_t stream, uint32_t step, SourceGeom sg, double* __restrict__ aos) {
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        PhiloxStream r; r.init(seed, stream, t, step);
        double v[3] = {0, 0, 0};
        if (sg.T > 0) sample_v3th(r, sg.T, sg.mass, v);            // Source.cpp:117 (warm): velocity is sampled before the position
        int axis = sg.face >> 1; bool plus = sg.face & 1;
        v[axis] += plus ? -sg.v_drift : sg.v_drift;
        double p[3];
        for (int a = 0; a < 3; a++) {
            if (a == axis) p[a] = plus ? sg.x0[a] + sg.L[a] : sg.x0[a];      // L[axis] already shortened by one cell on "+" faces (:51,69,86)
            else p[a] = sg.x0[a] + r.next() * sg.L[a];
        }
        double* a = aos + t * 7;
        a[0] = p[0]; a[1] = p[1]; a[2] = p[2]; a[3] = v[0]; a[4] = v[1]; a[5] = v[2]; a[6] = sg.mpw;
    }
}

// World::addInlet (World.cpp:206-261): node_type = DIRICHLET and phi = 0 on the inlet face
__global__ void k_set_face(Grid g, int face, double phi_set, int type, double* __restrict__ phi, int* __restrict__ node_type) {
    int axis = face >> 1; bool plus = face & 1;
    int n1 = axis == 0 ? g.nj : g.ni, n2 = axis == 2 ? g.nj : g.nk;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n1 * n2; t += gridDim.x * blockDim.x) {
        int a = t / n2, b = t % n2, i, j, k;
        if (axis == 0) { i = plus ? g.ni - 1 : 0; j = a; k = b; }
        else if (axis == 1) { i = a; j = plus ? g.nj - 1 : 0; k = b; }
        else { i = a; j = b; k = plus ? g.nk - 1 : 0; }
        size_t u = ((size_t)i * g.nj + j) * g.nk + k;
        phi[u] = phi_set; node_type[u] = type;
    }
}

static const size_t kGenChunk = 1u << 22;

extern "C" {

int picg_species_load_box_thermal(picg_species_t s, const double centre[3], const double sides[3], double num_den, double T, size_t* loaded) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && centre && sides, "picg_species_load_box_thermal: null argument");
    double box_vol = sides[0] * sides[1] * sides[2];
    double num_micro = num_den * box_vol;
    size_t num_macro = (size_t)(num_micro / s->mpw0);
    REQUIRE_ARG(num_macro >= 1, "picg_species_load_box_thermal: number of macroparticles less than 1, change initial values");   // Species.cpp:571-573
    if (g_world_size > 1) {                           // particles are split evenly by index across ranks
        size_t per = num_macro / g_world_size, rem = num_macro % g_world_size;
        num_macro = per + ((size_t)g_rank < rem ? 1 : 0);
    }
    int rc = species_refresh_count(s); if (rc) return rc;
    size_t before = s->n_host;
    rc = species_ensure_capacity(s, before + num_macro); if (rc) return rc;
    double lo[3], hi[3];
    for (int a = 0; a < 3; a++) { lo[a] = centre[a] - sides[a] / 2; hi[a] = centre[a] + sides[a] / 2; }   // x0 -/+ sides/2 (:580-581)
    uint32_t call = ++s->n_load_calls;
    uint32_t stream = rng_stream_id(RNG_LOADER, s->id, g_rank);
    rc = ensure_scratch(s->w, std::min(num_macro, kGenChunk) * 56 + 64); if (rc) return rc;
    for (size_t off = 0; off < num_macro; off += kGenChunk) {
        size_t m = std::min(kGenChunk, num_macro - off);
        LAUNCH(K_SOURCE, k_gen_box_thermal, std::min(div_up(m, 256), g_sm_count * 8), 256, 0, m, off, g_seed, stream, call, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2],
               T, s->mass, s->mpw0, (double*)s->w->scratch);
        CHECK_LAUNCH();
        rc = species_add_staged(s, m, (const double*)s->w->scratch); if (rc) return rc;
    }
    rc = species_refresh_count(s); if (rc) return rc;
    if (loaded) *loaded = s->n_host - before;
    return PICG_OK;
}

int picg_source_create(picg_species_t s, picg_world_t w, double v_drift, double den, double T, int face, picg_source_t* out) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && w && out && s->w == w, "picg_source_create: bad argument");
    REQUIRE_ARG(face >= 0 && face < 6, "picg_source_create: face must be 0..5 (x- x+ y- y+ z- z+)");
    picg_source_s* src = new picg_source_s();
    src->sp = s; src->w = w; src->v_drift = v_drift; src->den = den; src->T = T; src->face = face;
    const Grid& g = w->g;
    src->L[0] = g.dx[0] * (g.ni - 1); src->L[1] = g.dx[1] * (g.nj - 1); src->L[2] = g.dx[2] * (g.nk - 1);      // Source.cpp:17-19
    int axis = face >> 1;
    src->A = axis == 0 ? src->L[1] * src->L[2] : (axis == 1 ? src->L[0] * src->L[2] : src->L[0] * src->L[1]);   // :20-26
    src->num_micro = den * v_drift * src->A * w->dt;                                                            // :27  N = n*v*A*dt
    if (face & 1) src->L[axis] -= g.dx[axis];                                                                   // :51,69,86
    int n1 = axis == 0 ? g.nj : g.ni, n2 = axis == 2 ? g.nj : g.nk;
    LAUNCH(K_MISC, k_set_face, div_up((size_t)n1 * n2, 256), 256, 0, g, face, 0.0, 2, w->phi, w->node_type); CHECK_LAUNCH();   // :28 world.addInlet
    *out = src;
    return PICG_OK;
}
int picg_source_destroy(picg_source_t src) { delete src; return PICG_OK; }

int picg_source_sample(picg_source_t src, size_t* injected) {
    REQUIRE_DEVICE(); REQUIRE_ARG(src, "picg_source_sample: null source");
    picg_species_s* s = src->sp;
    uint32_t stream = rng_stream_id(RNG_SOURCE, s->id, g_rank);
    src->step++;
    PhiloxStream r; r.init(g_seed, stream, ~0ull, (uint32_t)src->step);
    double num = src->num_micro / s->mpw0;
    if (g_world_size > 1) num /= g_world_size;                       // each rank injects its share
    int num_macro = (int)(num + r.next());                           // Source.cpp:41
    if (injected) *injected = 0;
    if (num_macro <= 0) return PICG_OK;
    int rc = species_refresh_count(s); if (rc) return rc;
    size_t before = s->n_host;
    rc = species_ensure_capacity(s, before + (size_t)num_macro); if (rc) return rc;
    rc = ensure_scratch(s->w, (size_t)num_macro * 56 + 64); if (rc) return rc;
    SourceGeom sg; sg.face = src->face; sg.v_drift = src->v_drift; sg.T = src->T; sg.mass = s->mass; sg.mpw = s->mpw0;
    for (int a = 0; a < 3; a++) { sg.x0[a] = s->w->g.x0[a]; sg.L[a] = src->L[a]; }
    LAUNCH(K_SOURCE, k_gen_source, std::min(div_up(num_macro, 256), g_sm_count * 8), 256, 0, (size_t)num_macro, g_seed, stream, (uint32_t)src->step, sg,
           (double*)s->w->scratch);
    CHECK_LAUNCH();
    rc = species_add_staged(s, (size_t)num_macro, (const double*)s->w->scratch); if (rc) return rc;
    rc = species_refresh_count(s); if (rc) return rc;
    if (injected) *injected = s->n_host - before;
    return PICG_OK;
}

}  // extern "C"
