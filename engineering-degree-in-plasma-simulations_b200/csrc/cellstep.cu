// Cell-centric particle kernel: the fast path of push / deposit / count for a species whose store carries a cell
// partition (cell_start[] from the last sort).  Same arithmetic per particle as k_step (step.cu) and the reference
// (ch4/v3/src/Species.cpp:356-416, Field.h:157-232); only the work decomposition differs.
//
//   one WARP owns one CELL:  particles [cell_start[c], cell_start[c+1]) are contiguous, so
//     * their loads and stores are coalesced streaming accesses (ld.global.cs / st.global.cs: they must not evict the
//       field from L2);
//     * the 8 corner nodes of E are the same for every particle that is still in cell c: the warp fetches the
//       24 values ONCE per cell into shared memory and the trilinear gather becomes broadcast LDS + arithmetic -
//       no dependent per-particle global gather is left on the critical path;
//     * the 8 fixed-point corner sums of the deposit are accumulated in registers over the whole cell and leave the
//       warp once per cell through a transposed butterfly and 8 global integer atomics (not 8 per particle).
//   Particles that drifted out of the cell since the last sort ("stragglers"), and particles moved into holes by the
//   compaction, are handled individually (global gather / global atomics): correctness never depends on how stale
//   the partition is, only speed does.  Particles appended after the sort lie beyond the partition and are covered
//   by a launch of the generic kernel on that tail.
// Algorithmic bytes per particle: push 96 B, deposit 32 B, fused 104 B.
#include "common.cuh"
#include "push.cuh"
#include "deposit.cuh"
#include "samplers.cuh"
#include "heavy.cuh"
#include <algorithm>
#include <cmath>

using namespace picg;

#define CS_THREADS 256
#define CS_WARPS (CS_THREADS / 32)

struct CellArgs {
    double* a[7]; SpeciesCounters* ctr; u64 n_limit;        // particles [0, n_limit) are covered by the partition
    const unsigned* cell_start; const double* ef; double qm_dt, dt;
    unsigned* dead_list; u64* den_fixed; double scale; double* macro_count;
};

// Field<Vec3>::gather (Field.h:201-232) from four row pointers (rows (i,j) (i,j+1) (i+1,j) (i+1,j+1), each holding the
// nodes k and k+1 as 6 consecutive doubles).  Identical association to gather_ef().
__device__ __forceinline__ void gather_rows(const double* r00, const double* r01, const double* r10, const double* r11,
                                            double di, double dj, double dk, double& ex, double& ey, double& ez) {
    double odi = __dsub_rn(1.0, di), odj = __dsub_rn(1.0, dj), odk = __dsub_rn(1.0, dk);
    double wa = __dmul_rn(odi, odj), wb = __dmul_rn(odi, dj), wc = __dmul_rn(di, odj), wd = __dmul_rn(di, dj);
    double acc[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double v;
        v = __dmul_rn(__dmul_rn(r00[c], wa), odk);
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(r00[3 + c], wa), dk));
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(r01[c], wb), odk));
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(r01[3 + c], wb), dk));
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(r10[c], wc), odk));
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(r10[3 + c], wc), dk));
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(r11[c], wd), odk));
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(r11[3 + c], wd), dk));
        acc[c] = v;
    }
    ex = acc[0]; ey = acc[1]; ez = acc[2];
}

template <bool PUSH, bool HEAVY, bool DEPOSIT, bool COUNT>
__global__ void __launch_bounds__(CS_THREADS, 3) k_cell_step(Grid g, CellArgs A, HeavyArgs H) {
    __shared__ double s_E[CS_WARPS][24];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = blockIdx.x * CS_WARPS + wib, nwarps = gridDim.x * CS_WARPS;
    double* E = s_E[wib];
    const u64 n_limit = (A.n_limit == ~0ull) ? A.ctr->n : A.n_limit;
    const size_t row_j = (size_t)g.nk * 3, row_i = (size_t)g.nj * g.nk * 3;

    for (int cell = warp; cell < g.nc; cell += nwarps) {
        const u64 s = A.cell_start[cell];
        u64 e = A.cell_start[cell + 1];
        if (e > n_limit) e = n_limit;
        if (s >= e) continue;                                             // warp-uniform
        int ci, cj, ck; cell_to_ijk(g, cell, ci, cj, ck);
        if (PUSH) {                                                       // the cell's 8 E nodes: 4 rows x (k, k+1) x 3 components
            __syncwarp();
            if (lane < 24) {
                int q = lane / 6, t = lane - q * 6;
                size_t base = ((size_t)((ci + (q >> 1)) * g.nj + (cj + (q & 1))) * g.nk + ck) * 3;
                E[lane] = __ldg(A.ef + base + t);
            }
            __syncwarp();
        }
        i64 acc[8]; double cnt = 0;
#pragma unroll
        for (int c = 0; c < 8; c++) acc[c] = 0;

        for (u64 base = s; base < e; base += 32) {
            const u64 p = base + lane;
            const bool ok = p < e;
            bool dead = false;
            double x = 0, y = 0, z = 0, u = 0, v = 0, w = 0, m = 0;
            if (ok) {
                x = __ldcs(A.a[0] + p); y = __ldcs(A.a[1] + p); z = __ldcs(A.a[2] + p);
                if (PUSH) { u = __ldcs(A.a[3] + p); v = __ldcs(A.a[4] + p); w = __ldcs(A.a[5] + p); }
                if (DEPOSIT || HEAVY) m = __ldcs(A.a[6] + p);
            }
            if (PUSH && ok) {
                double lx = x_to_l(x, g.x0[0], g.inv_dx[0]), ly = x_to_l(y, g.x0[1], g.inv_dx[1]), lz = x_to_l(z, g.x0[2], g.inv_dx[2]);
                int i = min((int)lx, g.ni - 2), j = min((int)ly, g.nj - 2), k = min((int)lz, g.nk - 2);
                double di = __dsub_rn(lx, (double)i), dj = __dsub_rn(ly, (double)j), dk = __dsub_rn(lz, (double)k);
                double ex, ey, ez;
                if (i == ci && j == cj && k == ck) gather_rows(E, E + 6, E + 12, E + 18, di, dj, dk, ex, ey, ez);
                else {                                                    // straggler: its own 8 nodes from global memory
                    const double* r00 = A.ef + ((size_t)(i * g.nj + j) * g.nk + k) * 3;
                    gather_rows(r00, r00 + row_j, r00 + row_i, r00 + row_i + row_j, di, dj, dk, ex, ey, ez);
                }
                u = __dadd_rn(u, __dmul_rn(ex, A.qm_dt)); v = __dadd_rn(v, __dmul_rn(ey, A.qm_dt)); w = __dadd_rn(w, __dmul_rn(ez, A.qm_dt));
                if (!HEAVY) {
                    x = __dadd_rn(x, __dmul_rn(u, A.dt)); y = __dadd_rn(y, __dmul_rn(v, A.dt)); z = __dadd_rn(z, __dmul_rn(w, A.dt));
                    dead = !in_bounds(g, x, y, z) || in_object(g, x, y, z) != 0;             // Species.cpp:375-388
                } else {
                    double t_rem = 1; int n_b = 0; bool rng_ready = false; PhiloxStream rs;
                    while (t_rem > 0) {
                        if (++n_b > 20) { dead = true; break; }                                // :198-203
                        const double ox = x, oy = y, oz = z;
                        x = __dadd_rn(x, __dmul_rn(__dmul_rn(u, t_rem), A.dt));
                        y = __dadd_rn(y, __dmul_rn(__dmul_rn(v, t_rem), A.dt));
                        z = __dadd_rn(z, __dmul_rn(__dmul_rn(w, t_rem), A.dt));
                        int obj = in_object(g, x, y, z);
                        if (!in_bounds(g, x, y, z)) { dead = true; break; }
                        if (obj) {
                            if (!rng_ready) { rs.init(H.seed, H.stream, p, H.call); rng_ready = true; }
                            double xx[3] = {x, y, z}, vv[3] = {u, v, w}, oo[3] = {ox, oy, oz};
                            bool absorbed = surface_interaction(g, H, A.ef, rs, obj, oo, xx, vv, m, t_rem);
                            x = xx[0]; y = xx[1]; z = xx[2]; u = vv[0]; v = vv[1]; w = vv[2];
                            if (absorbed) { dead = true; break; }
                            continue;
                        }
                        t_rem = 0;
                    }
                }
                if (!dead) {
                    __stcs(A.a[0] + p, x); __stcs(A.a[1] + p, y); __stcs(A.a[2] + p, z);
                    __stcs(A.a[3] + p, u); __stcs(A.a[4] + p, v); __stcs(A.a[5] + p, w);
                }
            }
            if (PUSH) record_dead(dead, lane, p, A.ctr, A.dead_list);
            if ((DEPOSIT || COUNT) && ok && !dead) {
                int i, j, k; i64 q[8];
                if (DEPOSIT) scatter_weights_fixed(g, x_to_l(x, g.x0[0], g.inv_dx[0]), x_to_l(y, g.x0[1], g.inv_dx[1]), x_to_l(z, g.x0[2], g.inv_dx[2]), m, A.scale, i, j, k, q);
                else {
                    i = min((int)x_to_l(x, g.x0[0], g.inv_dx[0]), g.ci - 1); j = min((int)x_to_l(y, g.x0[1], g.inv_dx[1]), g.cj - 1);
                    k = min((int)x_to_l(z, g.x0[2], g.inv_dx[2]), g.ck - 1);
                }
                if (i == ci && j == cj && k == ck) {
                    if (DEPOSIT) {
#pragma unroll
                        for (int c = 0; c < 8; c++) acc[c] += q[c];
                    }
                    cnt += 1.0;
                } else {                                                  // left the cell: deposit on its own
                    if (DEPOSIT) {
#pragma unroll
                        for (int c = 0; c < 8; c++) if (q[c]) atomicAdd(&A.den_fixed[corner_node(g, i, j, k, c)], (u64)q[c]);
                    }
                    if (COUNT) atomicAdd(&A.macro_count[cell_of(g, i, j, k)], 1.0);
                }
            }
        }
        // once per cell: combine the 32 lanes (transposed butterfly, deposit.cuh) and hand the 8 corner sums over
        if (DEPOSIT) {
            i64 t = butterfly8(acc, lane);
            if ((lane & 3) == 0 && t != 0) atomicAdd(&A.den_fixed[corner_node(g, ci, cj, ck, lane >> 2)], (u64)t);
        }
        if (COUNT) {
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);      // integer-valued: exact
            if (lane == 0 && cnt != 0) atomicAdd(&A.macro_count[cell], cnt);
        }
    }
}

namespace picg {
template <bool PUSH, bool HEAVY, bool DEPOSIT, bool COUNT>
static int launch_cell_variant(const Grid& g, const CellArgs& A, const HeavyArgs& H, int kid) {
    int grid = std::max(1, std::min(div_up((size_t)g.nc, CS_WARPS), g_sm_count * 3 * 4));
    LAUNCH(kid, (k_cell_step<PUSH, HEAVY, DEPOSIT, COUNT>), grid, CS_THREADS, 0, g, A, H);
    CHECK_LAUNCH();
    return PICG_OK;
}

// mode bits as in launch_step: 1 push, 2 heavy, 4 deposit, 8 count.  Covers particles [0, n_limit).
int launch_cell_step(picg_species_s* s, int mode, double dt, const HeavyArgs& H, size_t n_limit) {
    const Grid& g = s->w->g;
    CellArgs A;
    for (int c = 0; c < 7; c++) A.a[c] = s->a[c];
    A.ctr = s->ctr; A.n_limit = n_limit; A.cell_start = s->cell_start; A.ef = s->w->ef;
    A.qm_dt = dt * s->charge / s->mass; A.dt = dt;
    A.dead_list = (unsigned*)s->w->scratch; A.den_fixed = (u64*)s->den_fixed; A.scale = std::ldexp(1.0, s->S); A.macro_count = s->macro_count;
    switch (mode) {
        case 1:  return launch_cell_variant<true, false, false, false>(g, A, H, K_PUSH_ELECTRONS);
        case 3:  return launch_cell_variant<true, true, false, false>(g, A, H, K_PUSH_HEAVY);
        case 4:  return launch_cell_variant<false, false, true, false>(g, A, H, K_DEPOSIT);
        case 12: return launch_cell_variant<false, false, true, true>(g, A, H, K_DEPOSIT);
        case 8:  return launch_cell_variant<false, false, false, true>(g, A, H, K_COUNT_CELLS);
        case 5:  return launch_cell_variant<true, false, true, false>(g, A, H, K_PUSH_DEPOSIT);
        case 13: return launch_cell_variant<true, false, true, true>(g, A, H, K_PUSH_DEPOSIT);
        case 7:  return launch_cell_variant<true, true, true, false>(g, A, H, K_PUSH_HEAVY_DEPOSIT);
        case 15: return launch_cell_variant<true, true, true, true>(g, A, H, K_PUSH_HEAVY_DEPOSIT);
        default: return set_error(PICG_ERR_ARG, "launch_cell_step: unsupported mode %d", mode);
    }
}
}  // namespace picg
