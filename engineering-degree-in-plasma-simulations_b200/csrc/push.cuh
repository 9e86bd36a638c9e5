// Device pieces shared by the push kernels (push.cu, push_deposit.cu, heavy.cu).
#pragma once
#include "common.cuh"

struct PushArrays { double *x, *y, *z, *u, *v, *w; };

namespace picg {
int compact_dead(picg_species_s* s, size_t cap);
int compact_plan(picg_world_s* w, SpeciesCounters* ctr, size_t cap, size_t store_cap, unsigned** D_out);
int sort_slot_list(picg_world_s* w, const u64* count_ptr, const u64* n_ptr, size_t list_cap, size_t store_cap, const unsigned* list, unsigned* out, unsigned* counts);
size_t compact_scratch_bytes(size_t cap);
int push_grid(size_t n_upper);
}

#ifdef __CUDACC__
// Compaction plan (push.cu): D holds the dead slots in ascending order, the first h of them below n_alive; the tail [n_alive, n_alive + nd)
// holds nd - h dead slots (D[h..nd)) and h survivors.  Slot of the t-th survivor (0 <= t < h): t plus the number of dead tail slots
// before it, i.e. i - h for the first i in [h, nd] with (D[i] - n_alive) - (i - h) > t  (survivors ahead of D[i]; nd: all of them).
__device__ __forceinline__ unsigned compact_survivor(const unsigned* __restrict__ D, u64 h, u64 nd, u64 n_alive, u64 t) {
    u64 lo = h, hi = nd;
    while (lo < hi) { const u64 mid = (lo + hi) >> 1; if (((u64)D[mid] - n_alive) - (mid - h) > t) hi = mid; else lo = mid + 1; }
    return (unsigned)(n_alive + t + (lo - h));
}
// Species::advanceElectronsSerial body (Species.cpp:368-373):
//   lc = XtoL(pos); E = ef.gather(lc); vel += E*(dt*charge/mass); pos += vel*dt
// qm_dt is the scalar dt*charge/mass formed on the host exactly as the reference forms it.
__device__ __forceinline__ void push_kick_drift(const Grid& g, const double* __restrict__ ef, double qm_dt, double dt,
                                                double& x, double& y, double& z, double& u, double& v, double& w) {
    double ex, ey, ez;
    gather_ef(g, ef, x_to_l(x, g.x0[0], g.inv_dx[0]), x_to_l(y, g.x0[1], g.inv_dx[1]), x_to_l(z, g.x0[2], g.inv_dx[2]), ex, ey, ez);
    u = __dadd_rn(u, __dmul_rn(ex, qm_dt));
    v = __dadd_rn(v, __dmul_rn(ey, qm_dt));
    w = __dadd_rn(w, __dmul_rn(ez, qm_dt));
    x = __dadd_rn(x, __dmul_rn(u, dt));
    y = __dadd_rn(y, __dmul_rn(v, dt));
    z = __dadd_rn(z, __dmul_rn(w, dt));
}

// Warp-aggregated append of dead particle indices to the dead list (one atomic per warp).
// Must be called by all 32 lanes.
__device__ __forceinline__ void record_dead(bool dead, int lane, u64 p, SpeciesCounters* ctr, unsigned* __restrict__ dead_list) {
    unsigned mask = __ballot_sync(0xffffffffu, dead);
    if (mask) {
        int leader = __ffs(mask) - 1;
        u64 base = 0;
        if (lane == leader) base = atomicAdd(&ctr->n_dead, (u64)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (dead) dead_list[base + __popc(mask & ((1u << lane) - 1))] = (unsigned)p;
    }
}
// Same, for any index list with its own cursor (used for the heavy push's impact list).  All 32 lanes must call.
__device__ __forceinline__ void record_index(bool flag, int lane, u64 p, u64* cursor, unsigned* __restrict__ list) {
    unsigned mask = __ballot_sync(0xffffffffu, flag);
    if (mask) {
        int leader = __ffs(mask) - 1;
        u64 base = 0;
        if (lane == leader) base = atomicAdd(cursor, (u64)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (flag) list[base + __popc(mask & ((1u << lane) - 1))] = (unsigned)p;
    }
}
#endif
