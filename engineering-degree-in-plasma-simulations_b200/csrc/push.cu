// Particle push kernels + in-place removal of dead particles.
//   (the electron / heavy pushes are k_step in step.cu)
//   k_push_reflect   : ch2 Species::advance (specular walls)     ch2/v2/Species.cpp:18-55
//   compaction       : the swap-with-last removal :378-398 done as hole filling: survivors from the
//                      tail [n_alive,n) move into the holes left below n_alive, so the traffic is
//                      proportional to the number of dead particles, not to n.
// Algorithmic bytes per particle-step (fp64 SoA): 48 B read (pos, vel) + 48 B written = 96 B.
#include "common.cuh"
#include "push.cuh"
#include <algorithm>

using namespace picg;

// ---------------------------------------------------------------- ch2 reflective push
__global__ void __launch_bounds__(256) k_push_reflect(Grid g, PushArrays s, const SpeciesCounters* ctr, const double* __restrict__ ef,
                                                      double qm_dt, double dt) {
    const u64 n = ctr->n;
    for (u64 p = blockIdx.x * (u64)blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x) {
        double q[3] = {s.x[p], s.y[p], s.z[p]}, v[3] = {s.u[p], s.v[p], s.w[p]};
        push_kick_drift(g, ef, qm_dt, dt, q[0], q[1], q[2], v[0], v[1], v[2]);
#pragma unroll
        for (int a = 0; a < 3; a++) {                                              // ch2/v2/Species.cpp:44-53
            if (q[a] < g.x0[a]) { q[a] = __dsub_rn(__dmul_rn(2.0, g.x0[a]), q[a]); v[a] = -v[a]; }
            else if (q[a] >= g.xm[a]) { q[a] = __dsub_rn(__dmul_rn(2.0, g.xm[a]), q[a]); v[a] = -v[a]; }
        }
        s.x[p] = q[0]; s.y[p] = q[1]; s.z[p] = q[2]; s.u[p] = v[0]; s.v[p] = v[1]; s.w[p] = v[2];
    }
}

// ---------------------------------------------------------------- compaction (hole filling)
// scratch layout: dead_list[cap] | hole[cap] | surv[cap] | tailflag[cap] (bytes)
__global__ void k_compact_zero(const SpeciesCounters* ctr, unsigned char* __restrict__ tailflag) {
    u64 nd = ctr->n_dead;
    for (u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x; t < nd; t += (u64)gridDim.x * blockDim.x) tailflag[t] = 0;
}
__global__ void k_compact_mark(const SpeciesCounters* ctr, const unsigned* __restrict__ dead_list, unsigned char* __restrict__ tailflag) {
    u64 nd = ctr->n_dead, n_alive = ctr->n - nd;
    for (u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x; t < nd; t += (u64)gridDim.x * blockDim.x) {
        u64 idx = dead_list[t];
        if (idx >= n_alive) tailflag[idx - n_alive] = 1;
    }
}
__global__ void k_compact_collect(SpeciesCounters* ctr, const unsigned* __restrict__ dead_list, const unsigned char* __restrict__ tailflag,
                                  unsigned* __restrict__ hole, unsigned* __restrict__ surv) {
    u64 nd = ctr->n_dead, n_alive = ctr->n - nd;
    for (u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x; t < nd; t += (u64)gridDim.x * blockDim.x) {
        if (!tailflag[t]) surv[atomicAdd(&ctr->n_surv, 1ull)] = (unsigned)(n_alive + t);
        u64 idx = dead_list[t];
        if (idx < n_alive) hole[atomicAdd(&ctr->n_hole, 1ull)] = (unsigned)idx;
    }
}
__global__ void k_compact_move(const SpeciesCounters* ctr, PushArrays s, double* __restrict__ mpw, const unsigned* __restrict__ hole,
                               const unsigned* __restrict__ surv) {
    u64 nh = ctr->n_hole;     // == n_surv by construction
    for (u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x; t < nh; t += (u64)gridDim.x * blockDim.x) {
        unsigned d = hole[t], f = surv[t];
        s.x[d] = s.x[f]; s.y[d] = s.y[f]; s.z[d] = s.z[f]; s.u[d] = s.u[f]; s.v[d] = s.v[f]; s.w[d] = s.w[f]; mpw[d] = mpw[f];
    }
}
__global__ void k_compact_finish(SpeciesCounters* ctr) {
    ctr->n -= ctr->n_dead; ctr->n_dead = 0; ctr->n_hole = 0; ctr->n_surv = 0;
}

namespace picg {
// Removes the particles recorded in the dead list of `s` (scratch layout above).  All sizes are read on the device.
int compact_dead(picg_species_s* s, size_t cap) {
    unsigned* dead_list = (unsigned*)s->w->scratch;
    unsigned* hole = dead_list + cap;
    unsigned* surv = hole + cap;
    unsigned char* tailflag = (unsigned char*)(surv + cap);
    PushArrays a = {s->a[0], s->a[1], s->a[2], s->a[3], s->a[4], s->a[5]};
    int grid = std::max(1, std::min(div_up(std::max<size_t>(cap / 16, 1), 256), g_sm_count * 4));
    LAUNCH(K_COMPACT, k_compact_zero, grid, 256, 0, s->ctr, tailflag); CHECK_LAUNCH();
    LAUNCH(K_COMPACT, k_compact_mark, grid, 256, 0, s->ctr, dead_list, tailflag); CHECK_LAUNCH();
    LAUNCH(K_COMPACT, k_compact_collect, grid, 256, 0, s->ctr, dead_list, tailflag, hole, surv); CHECK_LAUNCH();
    LAUNCH(K_COMPACT, k_compact_move, grid, 256, 0, s->ctr, a, s->a[6], hole, surv); CHECK_LAUNCH();
    LAUNCH(K_COMPACT, k_compact_finish, 1, 1, 0, s->ctr); CHECK_LAUNCH();
    s->n_host_valid = false;       // count changed on the device; n_upper stays an upper bound
    s->sorted_valid = false; s->lists_valid = false; s->movers_fresh = false; s->count_valid = false;
    return PICG_OK;
}
size_t compact_scratch_bytes(size_t cap) { return ((cap * 13 + 64) + 255) & ~(size_t)255; }    // 256-byte multiple: what follows stays aligned
int push_grid(size_t n_upper) { return std::max(1, std::min(div_up(std::max<size_t>(n_upper, 1), 256), g_sm_count * 8)); }
}  // namespace picg

extern "C" {

int picg_species_push_reflect(picg_species_t s, double dt) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_push_reflect: null species");
    double qm_dt = dt * s->charge / s->mass;
    PushArrays a = {s->a[0], s->a[1], s->a[2], s->a[3], s->a[4], s->a[5]};
    LAUNCH(K_PUSH_REFLECT, k_push_reflect, push_grid(s->n_upper), 256, 0, s->w->g, a, s->ctr, s->w->ef, qm_dt, dt);
    CHECK_LAUNCH();
    s->sorted_valid = false; s->lists_valid = false; s->movers_fresh = false; s->count_valid = false;
    return PICG_OK;
}

}  // extern "C"
