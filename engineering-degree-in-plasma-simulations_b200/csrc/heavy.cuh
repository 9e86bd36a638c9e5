// Surface interaction of the heavy-species push (ions / neutrals hitting an object), shared by step.cu and cellstep.cu.
//   Species::advanceNoSputteringSerial / advanceSputteringSerial   ch4/v3/src/Species.cpp:213-238, :136-145
#pragma once
#include "common.cuh"
#include "samplers.cuh"

// write == 0: emit_particle only counts what it would append (made); write == 1: it writes to slot base + made.  The slots of an impacting
// particle are fixed by a prefix sum over the impacts in ascending slot order (k_heavy_impacts, step.cu): no cursor, no dependence on timing.
struct Emit { double* a[7]; SpeciesCounters* ctr; u64 cap; double mpw0, q_over_m; int write; u64 base; unsigned made; };
struct HeavyArgs { Emit neutrals, spherium; int sputtering; double charge, mass, half_world_dt; uint64_t seed; uint32_t stream, call; };

#ifdef __CUDACC__
// Species::addParticle(pos, vel) for one particle created on a surface (Species.cpp:420-437)
static __device__ __noinline__ void emit_particle(const Grid& g, Emit& e, const double* __restrict__ ef, double half_dt, const double pos[3], const double v[3]) {
    if (isnan(pos[0]) || isnan(pos[1]) || isnan(pos[2]) || isnan(v[0]) || isnan(v[1]) || isnan(v[2])) return;
    if (!in_bounds(g, pos[0], pos[1], pos[2]) || in_object(g, pos[0], pos[1], pos[2])) return;
    double ex, ey, ez;
    gather_ef(g, ef, x_to_l(pos[0], g.x0[0], g.inv_dx[0]), x_to_l(pos[1], g.x0[1], g.inv_dx[1]), x_to_l(pos[2], g.x0[2], g.inv_dx[2]), ex, ey, ez);
    double u = __dsub_rn(v[0], __dmul_rn(__dmul_rn(ex, e.q_over_m), half_dt));
    double vv = __dsub_rn(v[1], __dmul_rn(__dmul_rn(ey, e.q_over_m), half_dt));
    double w = __dsub_rn(v[2], __dmul_rn(__dmul_rn(ez, e.q_over_m), half_dt));
    const u64 dst = e.base + e.made++;
    if (!e.write || dst >= e.cap) return;                                   // counting pass / store full (the counter is clamped and the overflow recorded after the kernel)
    e.a[0][dst] = pos[0]; e.a[1][dst] = pos[1]; e.a[2][dst] = pos[2]; e.a[3][dst] = u; e.a[4][dst] = vv; e.a[5][dst] = w; e.a[6][dst] = e.mpw0;
}

// The part of the heavy push that follows an impact (Species.cpp:213-238): rare, kept out of line.
// Returns true when the particle is absorbed; otherwise x, v, t_rem are updated for the next sub-move.
static __device__ __noinline__ bool surface_interaction(const Grid& g, HeavyArgs& h, const double* __restrict__ ef, PhiloxStream& r, int obj,
                                                 const double old[3], double x[3], double v[3], double mpw, double& t_rem) {
    double tp, hit[3], nrm[3];
    const ObjShape& o = g.obj[obj - 1];
    if (o.type == 0) rect_line_intersect(o, old, x, &tp, hit, nrm); else sphere_line_intersect(o, old, x, &tp, hit, nrm);
    x[0] = hit[0]; x[1] = hit[1]; x[2] = hit[2];
    double v_mag = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (h.charge == 0) {                                                    // neutrals: diffuse re-emission
        double nv[3]; sample_reflected(r, v_mag, nrm, h.mass, nv);
        v[0] = nv[0]; v[1] = nv[1]; v[2] = nv[2];
        t_rem *= (1 - tp);
        return false;
    }
    int mp_create = (int)(mpw / h.neutrals.mpw0 + r.next());                // ions: neutralise on the surface (:225-232)
    for (int c = 0; c < mp_create; c++) { double nv[3]; sample_reflected(r, v_mag, nrm, h.mass, nv); emit_particle(g, h.neutrals, ef, h.half_world_dt, x, nv); }
    if (h.sputtering) {                                                     // :136-145
        double yield = (v_mag > 5e3) ? 0.1 : 0;
        int sp_create = (int)(yield * mpw / h.spherium.mpw0 + r.next());
        for (int c = 0; c < sp_create; c++) { double nv[3]; sample_reflected(r, v_mag, nrm, h.mass, nv); emit_particle(g, h.spherium, ef, h.half_world_dt, x, nv); }
    }
    return true;
}


// State of a heavy particle whose first sub-move ended inside an object; heavy_after_impact runs the rest of the
// reference's bounce loop (Species.cpp:194-249: up to 20 sub-moves in total).  Returns true when the particle is removed.
struct HeavyState { double ox, oy, oz, x, y, z, u, v, w; };
static __device__ __noinline__ bool heavy_after_impact(const Grid& g, HeavyArgs& h, const double* __restrict__ ef, double dt, unsigned long long p,
                                                       int obj, double mpw, HeavyState& st) {
    PhiloxStream rs; rs.init(h.seed, h.stream, p, h.call);
    double t_rem = 1; int n_b = 1;
    double old[3] = {st.ox, st.oy, st.oz}, x[3] = {st.x, st.y, st.z}, v[3] = {st.u, st.v, st.w};
    bool gone = surface_interaction(g, h, ef, rs, obj, old, x, v, mpw, t_rem);
    while (!gone && t_rem > 0) {
        if (++n_b > 20) { gone = true; break; }                                   // :198-203
        old[0] = x[0]; old[1] = x[1]; old[2] = x[2];
        for (int a = 0; a < 3; a++) x[a] = __dadd_rn(x[a], __dmul_rn(__dmul_rn(v[a], t_rem), dt));      // pos += vel*t_rem*dt
        int o2 = in_object(g, x[0], x[1], x[2]);
        if (!in_bounds(g, x[0], x[1], x[2])) { gone = true; break; }
        if (o2) { gone = surface_interaction(g, h, ef, rs, o2, old, x, v, mpw, t_rem); continue; }
        t_rem = 0;
    }
    st.x = x[0]; st.y = x[1]; st.z = x[2]; st.u = v[0]; st.v = v[1]; st.w = v[2];
    return gone;
}
#endif

static inline Emit emit_of(picg_species_s* t) {
    Emit e; for (int c = 0; c < 7; c++) e.a[c] = t->a[c];
    e.ctr = t->ctr; e.cap = t->cap; e.mpw0 = t->mpw0; e.q_over_m = t->charge / t->mass; e.write = 0; e.base = 0; e.made = 0;
    return e;
}
