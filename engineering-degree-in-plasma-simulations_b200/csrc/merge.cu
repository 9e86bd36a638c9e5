// Species::merge (ch4/v3/src/Species.cpp:1037-1145) with sortVelocitiesInCell (:981-1035): SURVEY.md 8(f) row 2.
// Every cell that holds at least 10 particles is cut into 15 x 15 x 15 velocity bins; the particles of a bin that holds more
// than two are replaced by two particles of half the bin's weight each, at mean velocity +- the per-axis standard deviation,
// placed where two (distinct) randomly chosen members were - weight, momentum and per-axis energy of the bin are kept.
//
// Work decomposition: ONE WARP PER CELL on the exact per-cell lists of the store (celllists.cuh).  The warp finds the velocity
// box of the cell, forms one 22-bit key per particle (bin << 10 | position in the cell list), sorts the keys in shared memory
// (bitonic, <= 1024 per cell), and the lane that meets the head of a run of equal bins walks the run: the sums are formed in list
// order with the reference's association (FMA-free), so weights and velocities of the merged pairs equal the reference's bit for
// bit on the same particle order; only the two randomly picked member positions differ (Philox instead of mt19937).
// The pair is written over the first two members of the bin (no append: the store never grows), the other members go to the dead
// list and the usual hole-filling compaction removes them (push.cu).
//
// Reference behaviour kept on purpose: the velocity box starts from the velocity of the species' FIRST particle (particles[0], not
// a particle of the cell, :985); addParticle() rewinds the new velocities by half a step of the local field (:420-433) and drops a
// particle whose velocity is NaN (a bin whose variance is below -1e-4 keeps its negative value, :1080-1087: both particles vanish).
// Deviation: cells with more than 1024 particles are left unmerged and counted (statistics word 3).
#include "common.cuh"
#include "philox.cuh"
#include "celllists.cuh"
#include "push.cuh"
#include <algorithm>
#include <cmath>

using namespace picg;

#define MG_THREADS 256
#define MG_WARPS (MG_THREADS / 32)
#define MG_CAP 1024                                // particles of one cell handled in shared memory
#define MG_GRID 15                                 // Species::m_vel_grid_n (Species.h:129)

namespace picg { int species_exact_lists(picg_species_s* s); }

struct MergeArgs {
    Store s; CellLists L; const double* ef; double q_over_m, half_dt;
    double v0[3];                                  // velocity of the particle in slot 0 (the reference's particles[0])
    unsigned* dead_list; u64* stats;               // stats: [0] merged bins [1] particles removed [2] pairs dropped as NaN / rejected [3] cells too large
    uint64_t seed; uint32_t stream, call;
};

__device__ __forceinline__ double warp_min(double v) { for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
__device__ __forceinline__ double warp_max(double v) { for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }

__global__ void __launch_bounds__(MG_THREADS) k_merge(Grid g, MergeArgs A) {
    __shared__ unsigned s_keys[MG_WARPS][MG_CAP];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned* keys = s_keys[wib];
    const int nwarps = gridDim.x * MG_WARPS;
    for (int cell = blockIdx.x * MG_WARPS + wib; cell < g.nc; cell += nwarps) {
        const CellView V = cell_view(A.L, cell);
        const int np = V.np;
        if (np < 10) continue;                                                    // :1049 (warp-uniform)
        if (np > MG_CAP) { if (lane == 0) atomicAdd(&A.stats[3], 1ull); continue; }
        // ---- velocity box of the cell, started from particles[0] (:985-1011)
        double lo[3] = {A.v0[0], A.v0[1], A.v0[2]}, hi[3] = {A.v0[0], A.v0[1], A.v0[2]};
        for (int a = lane; a < np; a += 32) {
            const unsigned slot = cell_pick(A.L, V, a);
#pragma unroll
            for (int c = 0; c < 3; c++) { const double v = A.s.a[3 + c][slot]; lo[c] = fmin(lo[c], v); hi[c] = fmax(hi[c], v); }
        }
        double dv[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            lo[c] = warp_min(lo[c]); hi[c] = warp_max(hi[c]);
            hi[c] = (hi[c] > 0) ? __dmul_rn(hi[c], 1.001) : __dmul_rn(hi[c], 0.999);   // :1006-1012
            dv[c] = __ddiv_rn(__dsub_rn(hi[c], lo[c]), (double)MG_GRID);           // :1014-1021
        }
        // ---- one key per particle: velocity bin (:1023-1031), then position in the list
        int npad = 32; while (npad < np) npad <<= 1;
        for (int a = lane; a < npad; a += 32) {
            unsigned key = 0xffffffffu;
            if (a < np) {
                const unsigned slot = cell_pick(A.L, V, a);
                int b[3];
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const double t = __ddiv_rn(__dsub_rn(A.s.a[3 + c][slot], lo[c]), dv[c]);
                    b[c] = (t >= 0.0 && t < (double)MG_GRID) ? (int)t : (t >= (double)MG_GRID ? MG_GRID - 1 : 0);   // the reference indexes out of range here
                }
                key = (unsigned)((b[0] * MG_GRID + b[1]) * MG_GRID + b[2]) << 10 | (unsigned)a;
            }
            keys[a] = key;
        }
        __syncwarp();
        for (int k = 2; k <= npad; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = lane; t < npad; t += 32) {
                    const int x = t ^ j;
                    if (x > t) {
                        const unsigned a = keys[t], b = keys[x];
                        if ((a > b) == ((t & k) == 0)) { keys[t] = b; keys[x] = a; }
                    }
                }
                __syncwarp();
            }
        // ---- runs of equal bins; a run of more than two particles is merged by the lane that holds its head (:1058-1127)
        for (int t = lane; t < np; t += 32) {
            const unsigned bin = keys[t] >> 10;
            if (t > 0 && (keys[t - 1] >> 10) == bin) continue;
            int len = 1; while (t + len < np && (keys[t + len] >> 10) == bin) len++;
            if (len <= 2) continue;
            double W = 0, M[3] = {0, 0, 0}, E[3] = {0, 0, 0};
            for (int m = 0; m < len; m++) {                                       // list order = the reference's index order
                const unsigned slot = cell_pick(A.L, V, (int)(keys[t + m] & 1023u));
                const double w = A.s.a[6][slot];
                W = __dadd_rn(W, w);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const double v = A.s.a[3 + c][slot];
                    M[c] = __dadd_rn(M[c], __dmul_rn(w, v));
                    E[c] = __dadd_rn(E[c], __dmul_rn(w, __dmul_rn(v, v)));
                }
            }
            const double wa = __dmul_rn(W, 0.5);
            double va[3], vb[3];
            bool broke = false;
            const double invW = __ddiv_rn(1.0, W);                                 // Vec3::operator/=(scalar) multiplies by the inverse (Vec3.h:212-222)
#pragma unroll
            for (int c = 0; c < 3; c++) { M[c] = __dmul_rn(M[c], invW); E[c] = __dsub_rn(__dmul_rn(E[c], invW), __dmul_rn(M[c], M[c])); }
            for (int c = 0; c < 3 && !broke; c++) {                               // :1080-1088
                if (E[c] < 0) { if (E[c] < -10e-5) broke = true; else E[c] = 0; }
            }
#pragma unroll
            for (int c = 0; c < 3; c++) { const double sd = sqrt(E[c]); va[c] = __dadd_rn(M[c], sd); vb[c] = __dsub_rn(M[c], sd); }
            // two distinct members give the positions (:1107-1113)
            PhiloxStream r; r.init(A.seed, A.stream, (uint64_t)cell * 4096ull + bin, A.call);
            const int ia = min((int)(r.next() * len), len - 1);
            int ib = min((int)(r.next() * len), len - 1);
            while (ib == ia) ib = min((int)(r.next() * len), len - 1);
            const unsigned sa = cell_pick(A.L, V, (int)(keys[t + ia] & 1023u)), sb = cell_pick(A.L, V, (int)(keys[t + ib] & 1023u));
            const double pa[3] = {A.s.a[0][sa], A.s.a[1][sa], A.s.a[2][sa]}, pb[3] = {A.s.a[0][sb], A.s.a[1][sb], A.s.a[2][sb]};
            // addParticle(pos, vel, w) (:420-433): NaN / bounds / object filter, half-step rewind in the local field
            int kept = 0;
            for (int which = 0; which < 2; which++) {
                const double* pos = which ? pb : pa; double* vel = which ? vb : va;
                const unsigned dst = cell_pick(A.L, V, (int)(keys[t + kept] & 1023u));
                bool ok = !(isnan(pos[0]) || isnan(pos[1]) || isnan(pos[2]) || isnan(vel[0]) || isnan(vel[1]) || isnan(vel[2]));
                ok = ok && in_bounds(g, pos[0], pos[1], pos[2]) && !in_object(g, pos[0], pos[1], pos[2]);
                if (!ok) { atomicAdd(&A.stats[2], 1ull); continue; }
                double ex, ey, ez;
                gather_ef(g, A.ef, x_to_l(pos[0], g.x0[0], g.inv_dx[0]), x_to_l(pos[1], g.x0[1], g.inv_dx[1]), x_to_l(pos[2], g.x0[2], g.inv_dx[2]), ex, ey, ez);
                vel[0] = __dsub_rn(vel[0], __dmul_rn(__dmul_rn(A.q_over_m, ex), A.half_dt));
                vel[1] = __dsub_rn(vel[1], __dmul_rn(__dmul_rn(A.q_over_m, ey), A.half_dt));
                vel[2] = __dsub_rn(vel[2], __dmul_rn(__dmul_rn(A.q_over_m, ez), A.half_dt));
                A.s.a[0][dst] = pos[0]; A.s.a[1][dst] = pos[1]; A.s.a[2][dst] = pos[2];
                A.s.a[3][dst] = vel[0]; A.s.a[4][dst] = vel[1]; A.s.a[5][dst] = vel[2]; A.s.a[6][dst] = wa;
                kept++;
            }
            for (int m = kept; m < len; m++)                                      // everyone else leaves the store
                A.dead_list[atomicAdd(&A.s.ctr->n_dead, 1ull)] = cell_pick(A.L, V, (int)(keys[t + m] & 1023u));
            atomicAdd(&A.stats[0], 1ull); atomicAdd(&A.stats[1], (u64)(len - kept));
        }
        __syncwarp();
    }
}

extern "C" {

int picg_species_merge(picg_species_t s, uint64_t* n_before, uint64_t* n_after, uint64_t stats[4]) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_merge: null species");
    int rc = species_refresh_count(s); if (rc) return rc;
    const size_t n0 = s->n_host;
    if (n_before) *n_before = n0;
    if (n_after) *n_after = n0;
    if (stats) for (int k = 0; k < 4; k++) stats[k] = 0;
    if (n0 == 0) return PICG_OK;
    // particles[0] (the reference's first particle) = slot 0 of the store as the caller last saw it: read before any re-sort
    MergeArgs A;
    for (int c = 0; c < 3; c++) CUDA_TRY(cudaMemcpyAsync(&A.v0[c], s->a[3 + c], 8, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    rc = species_exact_lists(s); if (rc) return rc;                                // :1044-1046 `if(!sorted) sortIndexes()`
    const size_t cap = std::max<size_t>(s->n_upper, 1);
    rc = ensure_scratch(s->w, compact_scratch_bytes(cap) + 64); if (rc) return rc;
    A.s = store_of(s); A.L = lists_of(s); A.ef = s->w->ef; A.q_over_m = s->charge / s->mass; A.half_dt = 0.5 * s->w->dt;
    A.dead_list = (unsigned*)s->w->scratch;
    A.stats = (u64*)((char*)s->w->scratch + compact_scratch_bytes(cap));
    A.seed = g_seed; A.stream = rng_stream_id(RNG_MERGE, s->id, g_rank); A.call = ++s->n_merge_calls;
    CUDA_TRY(cudaMemsetAsync(A.stats, 0, 32, g_stream));
    const Grid& g = s->w->g;
    LAUNCH(K_MISC, k_merge, std::max(1, std::min(div_up((size_t)g.nc, MG_WARPS), g_sm_count * 4)), MG_THREADS, 0, g, A); CHECK_LAUNCH();
    u64 host_stats[4];
    CUDA_TRY(cudaMemcpyAsync(host_stats, A.stats, 32, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    rc = compact_dead(s, cap); if (rc) return rc;                                  // :1129-1140 (removal of the zero-weight particles)
    rc = species_refresh_count(s); if (rc) return rc;
    if (n_after) *n_after = s->n_host;
    if (stats) for (int k = 0; k < 4; k++) stats[k] = host_stats[k];
    return PICG_OK;
}

}  // extern "C"
