// Fused electron push + number-density deposit (+ optional per-cell macro-particle count): one pass over
// the particle arrays instead of two.  Replaces the pair Species::advanceElectrons -> computeNumberDensity
// (ch4/v3/src/main.cpp:215-218, Species.cpp:356-416) with identical arithmetic per particle.
// Algorithmic bytes per particle-step: 48 B read + 8 B mpw read + 48 B written = 104 B.
#include "common.cuh"
#include "push.cuh"
#include "deposit.cuh"
#include <algorithm>
#include <cmath>

using namespace picg;

namespace picg {
int launch_finalize(picg_species_s* s);
int calibrate_scale(picg_species_s* s, bool count_cells);
int check_scale_after(picg_species_s* s);
int deposit_grid(size_t n_upper);
}

template <bool COUNT>
__global__ void __launch_bounds__(DEP_THREADS) k_push_deposit(Grid g, PushArrays s, const double* __restrict__ pm, SpeciesCounters* ctr,
                                                              const double* __restrict__ ef, double qm_dt, double dt, unsigned* __restrict__ dead_list,
                                                              u64* __restrict__ den_fixed, double scale, double* __restrict__ macro_count) {
    __shared__ i64 win[DEP_WINDOW * 8];
    __shared__ int s_c0;
    const u64 n = ctr->n;
    const int lane = threadIdx.x & 31;
    for (int t = threadIdx.x; t < DEP_WINDOW * 8; t += blockDim.x) win[t] = 0;
    for (u64 chunk = (u64)blockIdx.x * DEP_CHUNK; chunk < n; chunk += (u64)gridDim.x * DEP_CHUNK) {
        if (threadIdx.x == 0) {       // window placed at the chunk's first particle (pre-push position; drift per step << cell)
            int i = min(max((int)x_to_l(s.x[chunk], g.x0[0], g.inv_dx[0]), 0), g.ci - 1);
            int j = min(max((int)x_to_l(s.y[chunk], g.x0[1], g.inv_dx[1]), 0), g.cj - 1);
            int k = min(max((int)x_to_l(s.z[chunk], g.x0[2], g.inv_dx[2]), 0), g.ck - 1);
            s_c0 = cell_of(g, i, j, k) - 2;
        }
        __syncthreads();
        const int c0 = s_c0;
        const u64 end = min(chunk + DEP_CHUNK, n);
        for (u64 p0 = chunk + (threadIdx.x - lane); p0 < end; p0 += DEP_THREADS) {
            u64 p = p0 + lane;
            bool active = false, dead = false;
            int cell = -1; i64 q[8];
            if (p < end) {
                double x = s.x[p], y = s.y[p], z = s.z[p], u = s.u[p], v = s.v[p], w = s.w[p];
                push_kick_drift(g, ef, qm_dt, dt, x, y, z, u, v, w);
                dead = !in_bounds(g, x, y, z) || in_object(g, x, y, z) != 0;
                if (!dead) {
                    s.x[p] = x; s.y[p] = y; s.z[p] = z; s.u[p] = u; s.v[p] = v; s.w[p] = w;
                    int i, j, k;
                    scatter_weights_fixed(g, x_to_l(x, g.x0[0], g.inv_dx[0]), x_to_l(y, g.x0[1], g.inv_dx[1]), x_to_l(z, g.x0[2], g.inv_dx[2]),
                                          pm[p], scale, i, j, k, q);
                    cell = cell_of(g, i, j, k);
                    active = true;
                    if (COUNT) atomicAdd(&macro_count[cell], 1.0);
                }
            }
            record_dead(dead, lane, p, ctr, dead_list);
            warp_accumulate(g, active, cell, q, win, c0, den_fixed, lane);
        }
        __syncthreads();
        window_flush(g, win, c0, den_fixed);
        __syncthreads();
    }
}

extern "C" int picg_species_push_electrons_deposit(picg_species_t s, double dt, int count_cells) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_push_electrons_deposit: null species");
    if (!s->S_pinned && !s->S_calibrated) {
        // the scale needs one calibrating deposit on the current state; afterwards it is sticky
        int rc = calibrate_scale(s, false); if (rc < 0) return rc;
    }
    size_t cap = std::max<size_t>(s->n_upper, 1);
    REQUIRE_ARG(cap < 0xffffffffull, "picg_species_push_electrons_deposit: more than 2^32-1 particles per GPU are not supported");
    int rc = ensure_scratch(s->w, compact_scratch_bytes(cap)); if (rc) return rc;
    const Grid& g = s->w->g;
    double qm_dt = dt * s->charge / s->mass, scale = std::ldexp(1.0, s->S);
    PushArrays a = {s->a[0], s->a[1], s->a[2], s->a[3], s->a[4], s->a[5]};
    cudaMemsetAsync(s->den_fixed, 0, (size_t)g.nv * 8, g_stream);
    if (count_cells) cudaMemsetAsync(s->macro_count, 0, (size_t)g.nc * 8, g_stream);
    int grid = deposit_grid(cap);
    if (count_cells) LAUNCH(K_PUSH_DEPOSIT, k_push_deposit<true>, grid, DEP_THREADS, 0, g, a, s->a[6], s->ctr, s->w->ef, qm_dt, dt, (unsigned*)s->w->scratch,
                            (u64*)s->den_fixed, scale, s->macro_count);
    else LAUNCH(K_PUSH_DEPOSIT, k_push_deposit<false>, grid, DEP_THREADS, 0, g, a, s->a[6], s->ctr, s->w->ef, qm_dt, dt, (unsigned*)s->w->scratch,
                (u64*)s->den_fixed, scale, s->macro_count);
    CHECK_LAUNCH();
    rc = compact_dead(s, cap); if (rc) return rc;
    rc = launch_finalize(s); if (rc) return rc;
    if (s->S_pinned) return PICG_OK;
    rc = species_refresh_count(s); if (rc) return rc;
    return check_scale_after(s);
}
