// Bird no-time-counter DSMC collisions between macro-particles of equal weight (variable hard sphere cross-section).
//   k_dsmc<false> : DSMC_MEX::applyOneSpecies   ch4/v3/src/Interactions.cpp:185-223
//   k_dsmc<true>  : DSMC_MEX::applyTwoSpecies   ch4/v3/src/Interactions.cpp:225-265
//   evaluateSigma :178-181, collide :267-285, constructor constants :143-176, sigma_v_rel_max Interactions.h:58
// The species are cell-sorted on the device (sort.cu); cell c's particles are its exact per-cell list (celllists.cuh).
// One thread owns one cell and runs the reference's candidate loop sequentially: a collision rewrites the two velocities
// and a later candidate of the same cell may pick them again (:202-216), cells are independent.  Only velocities change,
// so the cell partition of both species stays valid.  RNG: Philox stream (RNG_DSMC) addressed by (cell, apply-call number).
#include "common.cuh"
#include "philox.cuh"
#include "celllists.cuh"
#include <algorithm>
#include <cmath>

using namespace picg;

namespace picg { int species_exact_lists(picg_species_s* s); }

struct DsmcParams {
    double mass1, mass2, sum_mass, dv, mpw0, rank_scale;
    int world, rank;
    double pi_c0_c0, c1_m_half, c2, c3;
};
// evaluateSigma (:178-181): pi * c0 * c0 * pow(c2 / (v_rel*v_rel), c1 - 0.5) / c3
__host__ __device__ __forceinline__ double dsmc_sigma(const DsmcParams& P, double v_rel) {
    return P.pi_c0_c0 * pow(P.c2 / (v_rel * v_rel), P.c1_m_half) / P.c3;
}
// collide (:267-285): isotropic scattering in the centre-of-mass frame
__device__ __forceinline__ void dsmc_collide(PhiloxStream& r, const DsmcParams& P, double v1[3], double v2[3]) {
    double cm[3], g[3];
    const double inv_sum = 1.0 / P.sum_mass;                                  // Vec3::operator/(scalar) multiplies by the reciprocal (Vec3.h:180-192)
    for (int c = 0; c < 3; c++) { cm[c] = (P.mass1 * v1[c] + P.mass2 * v2[c]) * inv_sum; g[c] = v1[c] - v2[c]; }
    double g_mag = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    double cos_ksi = 2 * r.next() - 1;
    double sin_ksi = sqrt(1 - cos_ksi * cos_ksi);
    double eps = 2 * 3.141592653 * r.next();                                  // Const::pi, all.h:20
    g[0] = g_mag * cos_ksi; g[1] = g_mag * sin_ksi * cos(eps); g[2] = g_mag * sin_ksi * sin(eps);
    double f2 = P.mass2 / P.sum_mass, f1 = P.mass1 / P.sum_mass;
    for (int c = 0; c < 3; c++) { v1[c] = cm[c] + f2 * g[c]; v2[c] = cm[c] - f1 * g[c]; }
}
__device__ __forceinline__ void atomic_max_pos_double(double* addr, double v) {     // valid for non-negative doubles
    atomicMax((unsigned long long*)addr, (unsigned long long)__double_as_longlong(v));
}

// svm: [0] sigma_v_rel_max in force, [1] largest sigma*v_rel sampled by this call.  stats: [0] candidates [1] collisions
template <bool TWO>
__global__ void __launch_bounds__(128, 6) k_dsmc(Grid g, DsmcParams P, Store s1, Store s2, CellLists L1, CellLists L2, double* __restrict__ svm,
                                                 u64* __restrict__ stats, double dt, uint64_t seed, uint32_t stream, uint32_t call) {
    const double sv_max = svm[0];
    u64 n_cand = 0, n_coll = 0; double step_max = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < g.nc; c += gridDim.x * blockDim.x) {
        CellView va = cell_view(L1, c); int np1 = va.np, np2 = np1;
        CellView vb = va;
        if (TWO) { vb = cell_view(L2, c); np2 = vb.np; if (np1 < 1 || np2 < 1) continue; }      // :238
        else if (np1 < 2) continue;                                                               // :194
        // :196 / :240.  Multi-GPU: the cell's candidate count is estimated from the local populations (x G^2), rounded once
        // like the reference's and dealt out to the ranks (n/G each, the n%G left over to a rotating subset), see mcc.cu
        double frac = 0.5 * np1 * np2 * P.mpw0 * sv_max * dt / P.dv * P.rank_scale;
        int n_groups = (int)(frac + 0.5);
        if (P.world > 1) n_groups = n_groups / P.world + ((unsigned)(c + (int)call + P.rank) % (unsigned)P.world < (unsigned)(n_groups % P.world) ? 1 : 0);
        if (n_groups <= 0) continue;
        PhiloxStream r; r.init(seed, stream, (u64)c, call);
        for (int t = 0; t < n_groups; t++) {
            int a = (int)(r.next() * np1);
            int b = (int)(r.next() * np2);
            if (!TWO) while (a == b) b = (int)(r.next() * np2);                                   // :201-203
            u64 p1 = (u64)cell_pick(L1, va, a);
            u64 p2 = (u64)cell_pick(TWO ? L2 : L1, vb, b);
            double v1[3] = {s1.a[3][p1], s1.a[4][p1], s1.a[5][p1]}, v2[3] = {s2.a[3][p2], s2.a[4][p2], s2.a[5][p2]};
            double d[3] = {v1[0] - v2[0], v1[1] - v2[1], v1[2] - v2[2]};
            double v_rel = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            double sv = dsmc_sigma(P, v_rel) * v_rel;
            if (sv > step_max) step_max = sv;
            n_cand++;
            if (sv / sv_max > r.next()) {                                                         // :212 (NaN for v_rel == 0: no collision, like the reference)
                n_coll++;
                dsmc_collide(r, P, v1, v2);
                s1.a[3][p1] = v1[0]; s1.a[4][p1] = v1[1]; s1.a[5][p1] = v1[2];
                s2.a[3][p2] = v2[0]; s2.a[4][p2] = v2[1]; s2.a[5][p2] = v2[2];
            }
        }
    }
    __shared__ u64 sh[2]; __shared__ double sh_max;
    if (threadIdx.x == 0) { sh[0] = sh[1] = 0; sh_max = 0; }
    __syncthreads();
    if (n_cand) { atomicAdd(&sh[0], n_cand); atomicAdd(&sh[1], n_coll); atomic_max_pos_double(&sh_max, step_max); }
    __syncthreads();
    if (threadIdx.x == 0 && sh[0]) { atomicAdd(&stats[0], sh[0]); atomicAdd(&stats[1], sh[1]); atomic_max_pos_double(&svm[1], sh_max); }
}
// sigma_v_rel_max <- largest value sampled by this call, only if a collision happened (:219-222, :261-264)
__global__ void k_dsmc_finish(double* svm, const u64* stats) { if (stats[1]) svm[0] = svm[1]; }
__global__ void k_dsmc_sigma(DsmcParams P, int n, const double* __restrict__ v, double* __restrict__ out) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) out[t] = dsmc_sigma(P, v[t]);
}

static DsmcParams make_params(const picg_dsmc_s* m) {
    DsmcParams P;
    P.mass1 = m->sp1->mass; P.mass2 = m->sp2->mass; P.sum_mass = P.mass1 + P.mass2;
    double m_reduced = P.mass1 * P.mass2 / (P.mass1 + P.mass2);                             // :146 / :164
    const Grid& g = m->w->g;
    P.dv = g.dx[0] * g.dx[1] * g.dx[2];                                                     // World::getCellVolume
    P.mpw0 = m->sp1->mpw0;
    P.rank_scale = (double)g_world_size * (double)g_world_size; P.world = g_world_size; P.rank = g_rank;
    const double c0 = 4.07e-10, c1 = 0.77;                                                  // :151-154 Bird's reference parameters at 273.15 K
    P.pi_c0_c0 = 3.141592653 * c0 * c0;
    P.c1_m_half = c1 - 0.5;
    P.c2 = 2 * 1.380648e-23 * 273.15 / m_reduced;
    P.c3 = std::tgamma(2.5 - c1);
    return P;
}

extern "C" {

int picg_dsmc_create(picg_species_t species1, picg_species_t species2, picg_world_t w, picg_dsmc_t* out) {
    REQUIRE_DEVICE(); REQUIRE_ARG(species1 && w && out, "picg_dsmc_create: null argument");
    if (!species2) species2 = species1;
    // two-species constructor precondition (:160-162) -> std::invalid_argument in the facade
    REQUIRE_ARG(species1->mpw0 == species2->mpw0, "species must have the same macroparticle weight for this algorithm to work properly.");
    REQUIRE_ARG(species1->w == w && species2->w == w, "picg_dsmc_create: species belong to another world");
    picg_dsmc_s* m = new picg_dsmc_s();
    m->sp1 = species1; m->sp2 = species2; m->w = w;
    cudaError_t e;
    if ((e = cudaMalloc(&m->svm, 2 * 8)) != cudaSuccess || (e = cudaMalloc(&m->stats, 8 * 8)) != cudaSuccess) {
        picg_dsmc_destroy(m); return cuda_fail(e, "cudaMalloc(dsmc)", __FILE__, __LINE__);
    }
    double svm[2] = {1e-14, 0.0};                                                           // Interactions.h:58
    CUDA_TRY(cudaMemcpyAsync(m->svm, svm, 16, cudaMemcpyHostToDevice, g_stream));
    CUDA_TRY(cudaMemsetAsync(m->stats, 0, 64, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    *out = m;
    return PICG_OK;
}

int picg_dsmc_destroy(picg_dsmc_t m) {
    if (!m) return PICG_OK;
    if (g_stream) cudaStreamSynchronize(g_stream);
    cudaFree(m->svm); cudaFree(m->stats);
    delete m; return PICG_OK;
}

int picg_dsmc_set_sigma_v_max(picg_dsmc_t m, double v) {
    REQUIRE_DEVICE(); REQUIRE_ARG(m && v > 0, "picg_dsmc_set_sigma_v_max: bad argument");
    CUDA_TRY(cudaMemcpyAsync(m->svm, &v, 8, cudaMemcpyHostToDevice, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return PICG_OK;
}

int picg_dsmc_sigma(picg_dsmc_t m, int n, const double* v_rel, double* sigma) {
    REQUIRE_DEVICE(); REQUIRE_ARG(m && v_rel && sigma && n >= 0, "picg_dsmc_sigma: bad argument");
    if (n == 0) return PICG_OK;
    int rc = ensure_scratch(m->w, (size_t)n * 16 + 64); if (rc) return rc;
    double* d = (double*)m->w->scratch;
    CUDA_TRY(cudaMemcpyAsync(d, v_rel, (size_t)n * 8, cudaMemcpyHostToDevice, g_stream));
    LAUNCH(K_MISC, k_dsmc_sigma, std::min(div_up(n, 256), 1024), 256, 0, make_params(m), n, d, d + n); CHECK_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(sigma, d + n, (size_t)n * 8, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return PICG_OK;
}

int picg_dsmc_apply(picg_dsmc_t m, double dt, picg_dsmc_stats* out) {
    REQUIRE_DEVICE(); REQUIRE_ARG(m, "picg_dsmc_apply: null handle");
    picg_species_s *a = m->sp1, *b = m->sp2;
    const bool two = a != b;
    int rc;
    rc = species_exact_lists(a); if (rc) return rc;            // sortPointers (:186, :226-227): no-op when sorted, mover pass on a stale partition
    if (two) { rc = species_exact_lists(b); if (rc) return rc; }
    DsmcParams P = make_params(m);
    const Grid& g = m->w->g;
    CUDA_TRY(cudaMemsetAsync(m->stats, 0, 64, g_stream));
    CUDA_TRY(cudaMemsetAsync(m->svm + 1, 0, 8, g_stream));
    m->step++;
    int grid = std::max(1, std::min(div_up(g.nc, 128), g_sm_count * 16));
    uint32_t stream = rng_stream_id(RNG_DSMC, a->id, g_rank);
    if (two) LAUNCH(K_DSMC, k_dsmc<true>, grid, 128, 0, g, P, store_of(a), store_of(b), lists_of(a), lists_of(b), m->svm, m->stats, dt, g_seed, stream, (uint32_t)m->step);
    else     LAUNCH(K_DSMC, k_dsmc<false>, grid, 128, 0, g, P, store_of(a), store_of(a), lists_of(a), lists_of(a), m->svm, m->stats, dt, g_seed, stream, (uint32_t)m->step);
    CHECK_LAUNCH();
    LAUNCH(K_DSMC, k_dsmc_finish, 1, 1, 0, m->svm, m->stats); CHECK_LAUNCH();
    if (out) {                                                  // the only synchronisation of the call: skipped when the caller does not want the numbers
        u64 hs[2]; double sv;
        CUDA_TRY(cudaMemcpyAsync(hs, m->stats, 16, cudaMemcpyDeviceToHost, g_stream));
        CUDA_TRY(cudaMemcpyAsync(&sv, m->svm, 8, cudaMemcpyDeviceToHost, g_stream));
        CUDA_TRY(cudaStreamSynchronize(g_stream));
        out->candidates = hs[0]; out->collisions = hs[1]; out->sigma_v_max = sv;
    }
    return PICG_OK;
}

}  // extern "C"
