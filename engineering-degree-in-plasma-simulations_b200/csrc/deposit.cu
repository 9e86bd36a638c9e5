// Number-density deposition (deterministic, fixed point) and the node passes that follow it.
//   (the particle pass itself -- computeNumberDensity's loop, Species.cpp:401-413 -- is k_step in step.cu)
//   k_finalize_den   : den /= node_vol (zero-divisor guard)   Species.cpp:415, Field.h:563-583
//   k_moments        : Species::sampleMoments                 Species.cpp:767-776
// Algorithmic bytes per particle: deposit 24 B (pos) + 8 B (mpw) = 32 B.
#include "common.cuh"
#include "deposit.cuh"
#include <algorithm>
#include <cmath>

using namespace picg;

// den = (fixed * 2^-S) / node_vol, 0 where node_vol == 0; tracks max and overflow (negative) nodes.
__global__ void __launch_bounds__(256) k_finalize_den(int u_begin, int u_end, const i64* __restrict__ fixed, const double* __restrict__ vol,
                                                      double inv_scale, double* __restrict__ den, SpeciesCounters* ctr) {
    i64 mx = 0; unsigned neg = 0;
    for (int u = u_begin + blockIdx.x * blockDim.x + threadIdx.x; u < u_end; u += gridDim.x * blockDim.x) {
        i64 f = fixed[u];
        mx = max(mx, f); neg += f < 0;
        double d = __dmul_rn((double)f, inv_scale);
        double v = vol[u];
        den[u] = (v != 0.0) ? __ddiv_rn(d, v) : 0.0;
    }
    for (int o = 16; o > 0; o >>= 1) { mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); neg += __shfl_xor_sync(0xffffffffu, neg, o); }
    if ((threadIdx.x & 31) == 0) {
        if (mx > 0) atomicMax((i64*)&ctr->den_max, mx);
        if (neg) atomicAdd(&ctr->den_neg, (u64)neg);
    }
}
__global__ void k_reset_den_stats(SpeciesCounters* ctr) { ctr->den_max = 0; ctr->den_neg = 0; }

// Species::sampleMoments (Species.cpp:767-776): five scatters per particle (n, n*v (3), n*u^2, n*v^2, n*w^2) with the
// reference's weights; these diagnostics accumulate in fp64 (order-dependent in the reference as well).
__global__ void __launch_bounds__(256) k_moments(Grid g, const double* __restrict__ px, const double* __restrict__ py, const double* __restrict__ pz,
                                                 const double* __restrict__ pu, const double* __restrict__ pv, const double* __restrict__ pw,
                                                 const double* __restrict__ pm, const SpeciesCounters* ctr, double* __restrict__ n_sum,
                                                 double* __restrict__ nv_sum, double* __restrict__ nuu, double* __restrict__ nvv, double* __restrict__ nww) {
    const u64 n = ctr->n;
    for (u64 p = blockIdx.x * (u64)blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x) {
        double lx = x_to_l(px[p], g.x0[0], g.inv_dx[0]), ly = x_to_l(py[p], g.x0[1], g.inv_dx[1]), lz = x_to_l(pz[p], g.x0[2], g.inv_dx[2]);
        int i = min((int)lx, g.ni - 2), j = min((int)ly, g.nj - 2), k = min((int)lz, g.nk - 2);
        double di = lx - i, dj = ly - j, dk = lz - k;
        double m = pm[p], u = pu[p], v = pv[p], w = pw[p];
        double val[7] = {m, m * u, m * v, m * w, m * u * u, m * v * v, m * w * w};
        double wt[8];
        wt[0] = (1 - di) * (1 - dj) * (1 - dk); wt[1] = (1 - di) * (1 - dj) * dk; wt[2] = (1 - di) * dj * (1 - dk); wt[3] = (1 - di) * dj * dk;
        wt[4] = di * (1 - dj) * (1 - dk); wt[5] = di * (1 - dj) * dk; wt[6] = di * dj * (1 - dk); wt[7] = di * dj * dk;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            size_t node = corner_node(g, i, j, k, c);
            atomicAdd(&n_sum[node], val[0] * wt[c]);
            atomicAdd(&nv_sum[node * 3], val[1] * wt[c]); atomicAdd(&nv_sum[node * 3 + 1], val[2] * wt[c]); atomicAdd(&nv_sum[node * 3 + 2], val[3] * wt[c]);
            atomicAdd(&nuu[node], val[4] * wt[c]); atomicAdd(&nvv[node], val[5] * wt[c]); atomicAdd(&nww[node], val[6] * wt[c]);
        }
    }
}

// Species::computeGasProperties (Species.cpp:777-804)
__global__ void k_gas_properties(int nv, const double* __restrict__ n_sum, const double* __restrict__ nv_sum, const double* __restrict__ nuu,
                                 const double* __restrict__ nvv, const double* __restrict__ nww, double mass_over_2k, double* __restrict__ vel, double* __restrict__ T) {
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < nv; u += gridDim.x * blockDim.x) {
        double c = n_sum[u];
        double vx = 0, vy = 0, vz = 0;
        if (c != 0) { vx = nv_sum[3 * (size_t)u] / c; vy = nv_sum[3 * (size_t)u + 1] / c; vz = nv_sum[3 * (size_t)u + 2] / c; }   // Field / Field zero guard
        vel[3 * (size_t)u] = vx; vel[3 * (size_t)u + 1] = vy; vel[3 * (size_t)u + 2] = vz;
        if (c <= 0) { T[u] = 0; continue; }
        double uu = nuu[u] / c - vx * vx, vv = nvv[u] / c - vy * vy, ww = nww[u] / c - vz * vz;
        T[u] = mass_over_2k * (uu + vv + ww);
    }
}

namespace picg {
int launch_step(picg_species_s* s, int mode, double dt, picg_species_s* neutrals, picg_species_s* spherium, int sputtering, size_t n_snapshot);   // step.cu
int launch_finalize(picg_species_s* s, size_t u_begin = 0, size_t u_end = (size_t)-1);
static int pow2_floor_log(i64 v) { int l = -1; while (v > 0) { v >>= 1; l++; } return l; }

int launch_finalize(picg_species_s* s, size_t u_begin, size_t u_end) {
    const Grid& g = s->w->g;
    u_end = std::min(u_end, (size_t)g.nv);
    if (u_begin >= u_end) return PICG_OK;
    LAUNCH(K_MISC, k_reset_den_stats, 1, 1, 0, s->ctr); CHECK_LAUNCH();
    LAUNCH(K_FINALIZE_DEN, k_finalize_den, std::min(div_up(u_end - u_begin, 256), g_sm_count * 8), 256, 0, (int)u_begin, (int)u_end, s->den_fixed, s->w->node_vol,
           std::ldexp(1.0, -s->S), s->den, s->ctr);
    CHECK_LAUNCH();
    return PICG_OK;
}

// Chooses the fixed-point scale S.  Contributions are non-negative, so partial sums never exceed the
// final node sums: it suffices that max_node_sum * 2^S < 2^63.  Target: max node sum ~ 2^54 (9 bits of
// headroom; quantum 2^-54 of the largest node).  S is sticky; it is re-chosen only when the measured
// maximum leaves [2^46, 2^59].
int calibrate_scale(picg_species_s* s, bool count_cells) {
    if (s->S_pinned) return PICG_OK;
    int rc;
    if (!s->S_calibrated) {
        // first deposit: rigorous bound max node sum <= total weight
        double total = 0;
        rc = picg_species_diagnostics(s, &total, nullptr, nullptr); if (rc) return rc;
        int e = 0; if (total > 0) std::frexp(total, &e);          // total < 2^e
        s->S = 61 - e;
        rc = launch_step(s, 4, 0.0, nullptr, nullptr, 0, 0); if (rc) return rc;
        rc = launch_finalize(s); if (rc) return rc;
        s->n_host_valid = false; rc = species_refresh_count(s); if (rc) return rc;   // reads den_max
        i64 mx = s->ctr_host->den_max;
        if (mx > 0) { s->S += 54 - pow2_floor_log(mx); s->S_calibrated = true; }      // an empty store stays uncalibrated
        return 1;   // caller must deposit again with the calibrated S
    }
    return PICG_OK;
}

int check_scale_after(picg_species_s* s) {      // called with fresh ctr_host
    if (s->ctr_host->den_neg) {
        s->S_calibrated = false;
        return set_error(PICG_ERR_OVERFLOW, "fixed-point density accumulator overflowed on %llu nodes (S=%d); deposit again", (u64)s->ctr_host->den_neg, s->S);
    }
    if (!s->S_pinned) {
        i64 mx = s->ctr_host->den_max;
        if (mx > 0) { int l = pow2_floor_log(mx); if (l < 46 || l > 59) s->S += 54 - l; }
    }
    return PICG_OK;
}
}  // namespace picg

extern "C" {

int picg_species_set_density_scale(picg_species_t s, int S) {
    REQUIRE_ARG(s, "picg_species_set_density_scale: null species");
    if (S < -1000) { s->S_pinned = false; s->S_calibrated = false; }
    else { REQUIRE_ARG(S >= -900 && S <= 900, "picg_species_set_density_scale: |S| too large"); s->S = S; s->S_pinned = true; }
    return PICG_OK;
}
int picg_species_density_scale(picg_species_t s, int* S) { REQUIRE_ARG(s && S, "picg_species_density_scale: null argument"); *S = s->S; return PICG_OK; }

int picg_species_finalize_density(picg_species_t s) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_finalize_density: null species");
    return launch_finalize(s);
}
int picg_species_finalize_density_range(picg_species_t s, size_t node_begin, size_t node_end) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && node_begin <= node_end, "picg_species_finalize_density_range: bad argument");
    return launch_finalize(s, node_begin, node_end);
}

int picg_species_sample_moments(picg_species_t s) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_sample_moments: null species");
    LAUNCH(K_MOMENTS, k_moments, std::max(1, std::min(div_up(std::max<size_t>(s->n_upper, 1), 256), g_sm_count * 8)), 256, 0, s->w->g,
           s->a[0], s->a[1], s->a[2], s->a[3], s->a[4], s->a[5], s->a[6], s->ctr, s->n_sum, s->nv_sum, s->nuu, s->nvv, s->nww);
    CHECK_LAUNCH();
    return PICG_OK;
}

int picg_species_compute_gas_properties(picg_species_t s) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_compute_gas_properties: null species");
    int nv = s->w->g.nv;
    LAUNCH(K_MISC, k_gas_properties, std::min(div_up(nv, 256), g_sm_count * 8), 256, 0, nv, s->n_sum, s->nv_sum, s->nuu, s->nvv, s->nww,
           s->mass / (2 * 1.380648e-23), s->vel, s->T);
    CHECK_LAUNCH();
    return PICG_OK;
}

int picg_species_clear_samples(picg_species_t s) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_clear_samples: null species");
    size_t nv = s->w->g.nv;
    cudaMemsetAsync(s->n_sum, 0, nv * 8, g_stream); cudaMemsetAsync(s->nv_sum, 0, nv * 24, g_stream);
    cudaMemsetAsync(s->nuu, 0, nv * 8, g_stream); cudaMemsetAsync(s->nvv, 0, nv * 8, g_stream); cudaMemsetAsync(s->nww, 0, nv * 8, g_stream);
    return PICG_OK;
}

}  // extern "C"
