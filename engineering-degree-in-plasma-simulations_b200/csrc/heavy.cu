// Ion / neutral push with surface interaction.
//   k_push_heavy : Species::advanceNoSputteringSerial  ch4/v3/src/Species.cpp:170-256
//                  Species::advanceSputteringSerial    :81-169   (adds the yield model :136-145)
// Kick as for electrons (neutrals have charge 0 so the kick is a no-op arithmetic), then up to 20
// sub-moves: a particle that ends a sub-move inside an object is put back on the surface
// (World::lineIntersect); neutrals are re-emitted diffusely (sampleReflectedVelocity) and continue with
// the remaining fraction of the step; ions are absorbed and inject int(mpw/neutrals.mpw0 + rnd()) neutrals
// through addParticle semantics (Species.cpp:225-232).  Particles leaving the box, or bouncing more than
// 20 times, are removed.  RNG: Philox stream (RNG_HEAVY, species) indexed by particle slot and call number.
#include "common.cuh"
#include "push.cuh"
#include "samplers.cuh"
#include <algorithm>

using namespace picg;

struct Emit {                    // target store for particles created at a surface (neutrals / sputtered material)
    double* a[7]; SpeciesCounters* ctr; u64 cap; double mpw0, q_over_m;
};

// Species::addParticle(pos, vel) with default weight (Species.cpp:420-437) for one particle created on the device
__device__ __forceinline__ void emit_particle(const Grid& g, const Emit& e, const double* __restrict__ ef, double half_dt,
                                              const double pos[3], double v[3]) {
    if (isnan(pos[0]) || isnan(pos[1]) || isnan(pos[2]) || isnan(v[0]) || isnan(v[1]) || isnan(v[2])) return;
    if (!in_bounds(g, pos[0], pos[1], pos[2]) || in_object(g, pos[0], pos[1], pos[2])) return;       // impact points are filtered
    double ex, ey, ez;
    gather_ef(g, ef, x_to_l(pos[0], g.x0[0], g.inv_dx[0]), x_to_l(pos[1], g.x0[1], g.inv_dx[1]), x_to_l(pos[2], g.x0[2], g.inv_dx[2]), ex, ey, ez);
    double u = __dsub_rn(v[0], __dmul_rn(__dmul_rn(ex, e.q_over_m), half_dt));
    double vv = __dsub_rn(v[1], __dmul_rn(__dmul_rn(ey, e.q_over_m), half_dt));
    double w = __dsub_rn(v[2], __dmul_rn(__dmul_rn(ez, e.q_over_m), half_dt));
    u64 dst = atomicAdd(&e.ctr->n, 1ull);
    if (dst >= e.cap) { atomicAdd(&e.ctr->overflow, 1ull); return; }
    e.a[0][dst] = pos[0]; e.a[1][dst] = pos[1]; e.a[2][dst] = pos[2]; e.a[3][dst] = u; e.a[4][dst] = vv; e.a[5][dst] = w; e.a[6][dst] = e.mpw0;
}

__global__ void __launch_bounds__(256) k_push_heavy(Grid g, PushArrays s, const double* __restrict__ pm, SpeciesCounters* ctr, u64 n,
                                                    const double* __restrict__ ef, double qm_dt, double dt, double charge, double mass,
                                                    unsigned* __restrict__ dead_list, Emit neutrals, Emit spherium, int sputtering, double half_world_dt,
                                                    uint64_t seed, uint32_t stream, uint32_t call) {
    const int lane = threadIdx.x & 31;
    for (u64 p0 = (blockIdx.x * (u64)blockDim.x + threadIdx.x) - lane; p0 < n; p0 += (u64)gridDim.x * blockDim.x) {
        u64 p = p0 + lane;
        bool dead = false;
        if (p < n) {
            double x[3] = {s.x[p], s.y[p], s.z[p]}, v[3] = {s.u[p], s.v[p], s.w[p]};
            double ex, ey, ez;
            gather_ef(g, ef, x_to_l(x[0], g.x0[0], g.inv_dx[0]), x_to_l(x[1], g.x0[1], g.inv_dx[1]), x_to_l(x[2], g.x0[2], g.inv_dx[2]), ex, ey, ez);
            v[0] = __dadd_rn(v[0], __dmul_rn(ex, qm_dt)); v[1] = __dadd_rn(v[1], __dmul_rn(ey, qm_dt)); v[2] = __dadd_rn(v[2], __dmul_rn(ez, qm_dt));
            PhiloxStream r; bool rng_ready = false;
            double t_rem = 1; int n_b = 0;
            while (t_rem > 0) {
                if (++n_b > 20) { dead = true; break; }                                  // :198-203
                double old[3] = {x[0], x[1], x[2]};
#pragma unroll
                for (int a = 0; a < 3; a++) x[a] = __dadd_rn(x[a], __dmul_rn(__dmul_rn(v[a], t_rem), dt));   // pos += vel*t_rem*dt
                int obj = in_object(g, x[0], x[1], x[2]);
                if (!in_bounds(g, x[0], x[1], x[2])) { dead = true; break; }
                if (obj) {
                    if (!rng_ready) { r.init(seed, stream, p, call); rng_ready = true; }
                    double tp, hit[3], nrm[3];
                    const ObjShape& o = g.obj[obj - 1];
                    if (o.type == 0) rect_line_intersect(o, old, x, &tp, hit, nrm); else sphere_line_intersect(o, old, x, &tp, hit, nrm);
                    x[0] = hit[0]; x[1] = hit[1]; x[2] = hit[2];
                    double v_mag = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
                    if (charge == 0) {                                                    // neutrals: diffuse re-emission
                        double nv[3]; sample_reflected(r, v_mag, nrm, mass, nv);
                        v[0] = nv[0]; v[1] = nv[1]; v[2] = nv[2];
                        t_rem *= (1 - tp);
                        continue;
                    }
                    double mpw = pm[p];                                                   // ions: neutralise on the surface
                    int mp_create = (int)(mpw / neutrals.mpw0 + r.next());
                    for (int c = 0; c < mp_create; c++) { double nv[3]; sample_reflected(r, v_mag, nrm, mass, nv); emit_particle(g, neutrals, ef, half_world_dt, x, nv); }
                    if (sputtering) {                                                     // :136-145
                        double yield = (v_mag > 5e3) ? 0.1 : 0;
                        int sp_create = (int)(yield * mpw / spherium.mpw0 + r.next());
                        for (int c = 0; c < sp_create; c++) { double nv[3]; sample_reflected(r, v_mag, nrm, mass, nv); emit_particle(g, spherium, ef, half_world_dt, x, nv); }
                    }
                    dead = true; break;
                }
                t_rem = 0;
            }
            if (!dead) { s.x[p] = x[0]; s.y[p] = x[1]; s.z[p] = x[2]; s.u[p] = v[0]; s.v[p] = v[1]; s.w[p] = v[2]; }
        }
        record_dead(dead, lane, p, ctr, dead_list);
    }
}

static Emit emit_of(picg_species_s* t) {
    Emit e; for (int c = 0; c < 7; c++) e.a[c] = t->a[c];
    e.ctr = t->ctr; e.cap = t->cap; e.mpw0 = t->mpw0; e.q_over_m = t->charge / t->mass;
    return e;
}

extern "C" int picg_species_push_heavy(picg_species_t s, picg_species_t neutrals, picg_species_t spherium, double dt, int sputtering) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && neutrals && spherium, "picg_species_push_heavy: null species");
    REQUIRE_ARG(neutrals->w == s->w && spherium->w == s->w, "picg_species_push_heavy: species belong to different worlds");
    // the kernel walks a fixed snapshot of the count: particles emitted into `neutrals` during this call
    // (possibly the same store) are not pushed in the same call, as in the reference (np is read once, :176).
    int rc = species_refresh_count(s); if (rc) return rc;
    size_t n = s->n_host;
    size_t cap = std::max<size_t>(n, 1);
    REQUIRE_ARG(cap < 0xffffffffull, "picg_species_push_heavy: more than 2^32-1 particles per GPU are not supported");
    rc = ensure_scratch(s->w, compact_scratch_bytes(cap)); if (rc) return rc;
    bool emits = s->charge != 0;
    if (emits) {                                   // room for injected neutrals (rarely needed: mpw_ion/mpw0_neutral is usually << 1)
        rc = species_refresh_count(neutrals); if (rc) return rc;
        if (neutrals->cap < neutrals->n_host + 1024) { rc = species_ensure_capacity(neutrals, neutrals->n_host + neutrals->n_host / 8 + 4096); if (rc) return rc; }
        if (sputtering) {
            rc = species_refresh_count(spherium); if (rc) return rc;
            if (spherium->cap < spherium->n_host + 1024) { rc = species_ensure_capacity(spherium, spherium->n_host + spherium->n_host / 8 + 4096); if (rc) return rc; }
        }
    }
    static uint32_t call = 0; call++;
    double qm_dt = dt * s->charge / s->mass;
    PushArrays a = {s->a[0], s->a[1], s->a[2], s->a[3], s->a[4], s->a[5]};
    LAUNCH(K_PUSH_HEAVY, k_push_heavy, push_grid(cap), 256, 0, s->w->g, a, s->a[6], s->ctr, (u64)n, s->w->ef, qm_dt, dt, s->charge, s->mass,
           (unsigned*)s->w->scratch, emit_of(neutrals), emit_of(spherium), (sputtering && emits) ? 1 : 0, 0.5 * s->w->dt, g_seed,
           rng_stream_id(RNG_HEAVY, s->id, g_rank), call);
    CHECK_LAUNCH();
    if (emits) { neutrals->n_host_valid = false; neutrals->sorted_valid = false; neutrals->n_upper = neutrals->cap;
                 if (sputtering) { spherium->n_host_valid = false; spherium->sorted_valid = false; spherium->n_upper = spherium->cap; } }
    return compact_dead(s, cap);
}
