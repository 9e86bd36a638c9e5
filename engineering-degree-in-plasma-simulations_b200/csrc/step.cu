// The particle-loop kernel: push (+ wall interaction) and/or fixed-point deposit (+ per-cell count) in one pass.
//
//   k_step<PUSH, HEAVY, DEPOSIT, COUNT>
//     PUSH            Species::advanceElectronsSerial              ch4/v3/src/Species.cpp:356-399
//     PUSH + HEAVY    Species::advanceNoSputteringSerial / ...SputteringSerial   :170-256 / :81-169
//     DEPOSIT         Species::computeNumberDensity                :401-413 (+ Field::scatter Field.h:157-199)
//     COUNT           Species::computeMacroParticlesCount          :813-819
//
// Data movement (the design point of this kernel; every stage is HBM-bound):
//   * a block owns a contiguous chunk of STEP_CHUNK particles of the (cell-sorted) SoA store;
//   * load phase: the chunk's arrays are copied global -> shared with fully coalesced accesses, all loads of a thread
//     issued before the first use (deep memory-level parallelism; no dependent gather sits between two particle loads);
//   * compute phase: thread t processes the R = STEP_RUN CONSECUTIVE particles t*R..t*R+R-1 from shared memory.  In a
//     cell-sorted store a run stays inside one cell almost always, so the eight fixed-point corner sums are accumulated
//     in registers and leave the thread once per run, not once per particle.  Run totals of the lanes of a warp that end
//     in the same cell are combined with a transposed butterfly (deposit.cuh) and go to a shared-memory window of
//     STEP_WINDOW cells x 8 corners with 64-bit integer atomics; cells outside the window (unsorted input, stragglers)
//     go straight to global memory.  Integer sums are associative: any order gives the same bits.
//   * store phase: updated positions / velocities go shared -> global, coalesced; the window is flushed with one
//     global atomic per touched (cell, corner).
// Algorithmic bytes per particle: push 96 B, deposit 32 B, fused 104 B (E-field and grid traffic are amortised over
// the ~60 particles of a cell and served by L1/L2).
#include "common.cuh"
#include "push.cuh"
#include "deposit.cuh"
#include "samplers.cuh"
#include "heavy.cuh"
#include <algorithm>
#include <cmath>

using namespace picg;

#define STEP_THREADS 256
#define STEP_RUN 4
#define STEP_CHUNK (STEP_THREADS * STEP_RUN)                 // 1024 particles per block iteration
#define STEP_PITCH (STEP_CHUNK + STEP_CHUNK / STEP_RUN)      // one pad double per run: conflict-free run-wise reads
#define STEP_WINDOW 256                                      // cells staged in shared memory (x 8 corners x 8 B = 16 KB)

struct StepArgs {
    double* a[7]; SpeciesCounters* ctr; u64 n_fixed; int use_fixed_n;      // heavy pushes walk a snapshot of the count (Species.cpp:176)
    const unsigned* tail_from;                                              // non-null: only particles [*tail_from, n) (the part beyond the cell partition)
    const double* ef; double qm_dt, dt;
    unsigned* dead_list; u64* den_fixed; double scale; double* macro_count;
};

__device__ __forceinline__ int spos(int i) { return i + (i / STEP_RUN); }

template <bool PUSH, bool HEAVY, bool DEPOSIT, bool COUNT>
__global__ void __launch_bounds__(STEP_THREADS, 2) k_step(Grid g, StepArgs A, HeavyArgs H) {
    extern __shared__ double smem[];
    constexpr int NARR = PUSH ? 7 : 4;                       // staged arrays: x y z [u v w] mpw
    double* sx = smem; double* sy = sx + STEP_PITCH; double* sz = sy + STEP_PITCH;
    double* su = PUSH ? sz + STEP_PITCH : nullptr; double* sv = PUSH ? su + STEP_PITCH : nullptr; double* sw = PUSH ? sv + STEP_PITCH : nullptr;
    double* sm = smem + (NARR - 1) * STEP_PITCH;
    i64* win = (i64*)(smem + NARR * STEP_PITCH);
    __shared__ int s_c0;
    const u64 n = A.use_fixed_n ? A.n_fixed : A.ctr->n;
    const int tid = threadIdx.x, lane = tid & 31;
    if (DEPOSIT) { for (int t = tid; t < STEP_WINDOW * 8; t += STEP_THREADS) win[t] = 0; }

    const u64 first = A.tail_from ? (u64)*A.tail_from : 0;
    for (u64 chunk = first + (u64)blockIdx.x * STEP_CHUNK; chunk < n; chunk += (u64)gridDim.x * STEP_CHUNK) {
        const int cnt = (int)min((u64)STEP_CHUNK, n - chunk);
        // ---- load phase (coalesced; every load of the thread is independent of every other)
#pragma unroll
        for (int r = 0; r < STEP_RUN; r++) {
            int i = r * STEP_THREADS + tid;
            bool ok = i < cnt; u64 p = chunk + i; int d = spos(i);
            sx[d] = ok ? A.a[0][p] : 0.0; sy[d] = ok ? A.a[1][p] : 0.0; sz[d] = ok ? A.a[2][p] : 0.0;
            if (PUSH) { su[d] = ok ? A.a[3][p] : 0.0; sv[d] = ok ? A.a[4][p] : 0.0; sw[d] = ok ? A.a[5][p] : 0.0; }
            if (DEPOSIT || HEAVY) sm[d] = ok ? A.a[6][p] : 0.0;
        }
        __syncthreads();
        if (DEPOSIT && tid == 0) {            // window placed at the chunk's first particle (2 cells of slack below)
            int i = min(max((int)x_to_l(sx[0], g.x0[0], g.inv_dx[0]), 0), g.ci - 1);
            int j = min(max((int)x_to_l(sy[0], g.x0[1], g.inv_dx[1]), 0), g.cj - 1);
            int k = min(max((int)x_to_l(sz[0], g.x0[2], g.inv_dx[2]), 0), g.ck - 1);
            s_c0 = cell_of(g, i, j, k) - 2;
        }
        if (DEPOSIT) __syncthreads();
        const int c0 = DEPOSIT ? s_c0 : 0;

        // ---- compute phase: a run of STEP_RUN consecutive particles per thread
        int cur = -1; i64 acc[8]; double cur_count = 0;
#pragma unroll
        for (int c = 0; c < 8; c++) acc[c] = 0;
#pragma unroll
        for (int r = 0; r < STEP_RUN; r++) {
            const int i = tid * STEP_RUN + r, d = spos(i);
            const u64 p = chunk + i;
            const bool ok = i < cnt;
            bool dead = false;
            double x = sx[d], y = sy[d], z = sz[d];
            if (PUSH && ok) {
                double u = su[d], v = sv[d], w = sw[d];
                if (!HEAVY) {
                    push_kick_drift(g, A.ef, A.qm_dt, A.dt, x, y, z, u, v, w);
                    dead = !in_bounds(g, x, y, z) || in_object(g, x, y, z) != 0;            // Species.cpp:375-388
                } else {
                    double ex, ey, ez;
                    gather_ef(g, A.ef, x_to_l(x, g.x0[0], g.inv_dx[0]), x_to_l(y, g.x0[1], g.inv_dx[1]), x_to_l(z, g.x0[2], g.inv_dx[2]), ex, ey, ez);
                    u = __dadd_rn(u, __dmul_rn(ex, A.qm_dt)); v = __dadd_rn(v, __dmul_rn(ey, A.qm_dt)); w = __dadd_rn(w, __dmul_rn(ez, A.qm_dt));
                    double t_rem = 1; int n_b = 0; bool rng_ready = false; PhiloxStream rs;
                    while (t_rem > 0) {
                        if (++n_b > 20) { dead = true; break; }                                // :198-203
                        double old[3] = {x, y, z};
                        x = __dadd_rn(x, __dmul_rn(__dmul_rn(u, t_rem), A.dt));               // pos += vel*t_rem*dt
                        y = __dadd_rn(y, __dmul_rn(__dmul_rn(v, t_rem), A.dt));
                        z = __dadd_rn(z, __dmul_rn(__dmul_rn(w, t_rem), A.dt));
                        int obj = in_object(g, x, y, z);
                        if (!in_bounds(g, x, y, z)) { dead = true; break; }
                        if (obj) {
                            if (!rng_ready) { rs.init(H.seed, H.stream, p, H.call); rng_ready = true; }
                            double xx[3] = {x, y, z}, vv[3] = {u, v, w};
                            bool absorbed = surface_interaction(g, H, A.ef, rs, obj, old, xx, vv, sm[d], t_rem);
                            x = xx[0]; y = xx[1]; z = xx[2]; u = vv[0]; v = vv[1]; w = vv[2];
                            if (absorbed) { dead = true; break; }
                            continue;
                        }
                        t_rem = 0;
                    }
                }
                if (!dead) { sx[d] = x; sy[d] = y; sz[d] = z; su[d] = u; sv[d] = v; sw[d] = w; }
            }
            if (PUSH) record_dead(dead, lane, p, A.ctr, A.dead_list);
            if (DEPOSIT || COUNT) {
                if (ok && !dead) {
                    int ci, cj, ck; i64 q[8];
                    if (DEPOSIT) scatter_weights_fixed(g, x_to_l(x, g.x0[0], g.inv_dx[0]), x_to_l(y, g.x0[1], g.inv_dx[1]), x_to_l(z, g.x0[2], g.inv_dx[2]),
                                                       sm[d], A.scale, ci, cj, ck, q);
                    else {
                        ci = min((int)x_to_l(x, g.x0[0], g.inv_dx[0]), g.ci - 1); cj = min((int)x_to_l(y, g.x0[1], g.inv_dx[1]), g.cj - 1);
                        ck = min((int)x_to_l(z, g.x0[2], g.inv_dx[2]), g.ck - 1);
                    }
                    int cell = cell_of(g, ci, cj, ck);
                    if (cell != cur) {
                        if (cur >= 0) {                                  // the run left its cell: hand the partial sums over
                            if (DEPOSIT) {
                                int rel = cur - c0;
                                if (rel >= 0 && rel < STEP_WINDOW) {
#pragma unroll
                                    for (int c = 0; c < 8; c++) if (acc[c]) atomicAdd((u64*)&win[rel * 8 + c], (u64)acc[c]);
                                } else {
                                    int i2, j2, k2; cell_to_ijk(g, cur, i2, j2, k2);
#pragma unroll
                                    for (int c = 0; c < 8; c++) if (acc[c]) atomicAdd(&A.den_fixed[corner_node(g, i2, j2, k2, c)], (u64)acc[c]);
                                }
                            }
                            if (COUNT) atomicAdd(&A.macro_count[cur], cur_count);
                        }
                        cur = cell; cur_count = 0;
#pragma unroll
                        for (int c = 0; c < 8; c++) acc[c] = 0;
                    }
                    if (DEPOSIT) {
#pragma unroll
                        for (int c = 0; c < 8; c++) acc[c] += q[c];
                    }
                    cur_count += 1.0;
                }
            }
        }
        // run totals: lanes ending in the same cell are combined before they touch shared / global memory
        if (DEPOSIT) warp_accumulate_w<STEP_WINDOW>(g, cur >= 0, cur, acc, win, c0, A.den_fixed, lane);
        if (COUNT) {
            unsigned peers = __match_any_sync(0xffffffffu, cur);
            double s = 0; unsigned m = peers;                            // integer-valued counts: exact in any order
            while (m) { int src = __ffs(m) - 1; m &= m - 1; s += __shfl_sync(peers, cur_count, src); }
            if (cur >= 0 && lane == __ffs(peers) - 1) atomicAdd(&A.macro_count[cur], s);
        }
        __syncthreads();
        // ---- store phase
        if (PUSH) {
#pragma unroll
            for (int r = 0; r < STEP_RUN; r++) {
                int i = r * STEP_THREADS + tid;
                if (i < cnt) {
                    u64 p = chunk + i; int d = spos(i);
                    A.a[0][p] = sx[d]; A.a[1][p] = sy[d]; A.a[2][p] = sz[d]; A.a[3][p] = su[d]; A.a[4][p] = sv[d]; A.a[5][p] = sw[d];
                }
            }
        }
        if (DEPOSIT) {
            for (int slot = tid; slot < STEP_WINDOW * 8; slot += STEP_THREADS) {
                i64 v = win[slot];
                if (v != 0) {
                    int i2, j2, k2; cell_to_ijk(g, c0 + (slot >> 3), i2, j2, k2);
                    atomicAdd(&A.den_fixed[corner_node(g, i2, j2, k2, slot & 7)], (u64)v);
                    win[slot] = 0;
                }
            }
        }
        __syncthreads();
    }
}

namespace picg {
int launch_finalize(picg_species_s* s);
int calibrate_scale(picg_species_s* s, bool count_cells);
int check_scale_after(picg_species_s* s);

template <bool PUSH, bool HEAVY, bool DEPOSIT, bool COUNT>
static int launch_variant(const Grid& g, const StepArgs& A, const HeavyArgs& H, size_t n_upper, int kid) {
    size_t smem = (size_t)((PUSH ? 7 : 4) * STEP_PITCH) * 8 + (DEPOSIT ? STEP_WINDOW * 64 : 0);
    static bool attr_set = false;
    if (!attr_set) { CUDA_TRY(cudaFuncSetAttribute(k_step<PUSH, HEAVY, DEPOSIT, COUNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set = true; }
    int per_sm = std::max(1, (int)((size_t)227 * 1024 / (smem + 1024)));
    per_sm = std::min(per_sm, 2048 / STEP_THREADS);
    int grid = std::max(1, std::min(div_up(std::max<size_t>(n_upper, 1), STEP_CHUNK), g_sm_count * per_sm));
    LAUNCH(kid, (k_step<PUSH, HEAVY, DEPOSIT, COUNT>), grid, STEP_THREADS, smem, g, A, H);
    CHECK_LAUNCH();
    return PICG_OK;
}

// mode bits: 1 push, 2 heavy, 4 deposit, 8 count
int launch_cell_step(picg_species_s* s, int mode, double dt, const HeavyArgs& H, size_t n_limit);     // cellstep.cu

int launch_step(picg_species_s* s, int mode, double dt, picg_species_s* neutrals, picg_species_s* spherium, int sputtering, size_t n_snapshot) {
    const Grid& g = s->w->g;
    StepArgs A;
    A.tail_from = nullptr;
    for (int c = 0; c < 7; c++) A.a[c] = s->a[c];
    A.ctr = s->ctr; A.use_fixed_n = (mode & 2) ? 1 : 0; A.n_fixed = n_snapshot;
    A.ef = s->w->ef; A.qm_dt = dt * s->charge / s->mass; A.dt = dt;                 // Species.cpp:372 `dt*charge/mass`
    A.dead_list = (unsigned*)s->w->scratch; A.den_fixed = (u64*)s->den_fixed; A.scale = std::ldexp(1.0, s->S); A.macro_count = s->macro_count;
    HeavyArgs H; memset(&H, 0, sizeof(H));
    if (mode & 2) {
        static uint32_t call = 0; call++;
        H.neutrals = emit_of(neutrals); H.spherium = emit_of(spherium); H.sputtering = (sputtering && s->charge != 0) ? 1 : 0;
        H.charge = s->charge; H.mass = s->mass; H.half_world_dt = 0.5 * s->w->dt; H.seed = g_seed; H.stream = rng_stream_id(RNG_HEAVY, s->id, g_rank); H.call = call;
    }
    if (mode & 4) cudaMemsetAsync(s->den_fixed, 0, (size_t)g.nv * 8, g_stream);
    if (mode & 8) cudaMemsetAsync(s->macro_count, 0, (size_t)g.nc * 8, g_stream);
    size_t nu = (mode & 2) ? n_snapshot : s->n_upper;
    // fast path: the store carries a cell partition (cell_start[] of the last sort) and is dense enough for a warp per cell
    if (s->part_valid && nu >= (size_t)4 * g.nc) {
        int rc = launch_cell_step(s, mode, dt, H, (mode & 2) ? n_snapshot : (size_t)-1); if (rc) return rc;
        if (nu <= s->part_n) return PICG_OK;              // nothing was appended since the sort
        A.tail_from = s->cell_start + g.nc;               // the appended tail goes through the generic kernel
        nu = nu - s->part_n;
    }
    switch (mode) {
        case 1:  return launch_variant<true, false, false, false>(g, A, H, nu, K_PUSH_ELECTRONS);
        case 3:  return launch_variant<true, true, false, false>(g, A, H, nu, K_PUSH_HEAVY);
        case 4:  return launch_variant<false, false, true, false>(g, A, H, nu, K_DEPOSIT);
        case 12: return launch_variant<false, false, true, true>(g, A, H, nu, K_DEPOSIT);
        case 8:  return launch_variant<false, false, false, true>(g, A, H, nu, K_COUNT_CELLS);
        case 5:  return launch_variant<true, false, true, false>(g, A, H, nu, K_PUSH_DEPOSIT);
        case 13: return launch_variant<true, false, true, true>(g, A, H, nu, K_PUSH_DEPOSIT);
        case 7:  return launch_variant<true, true, true, false>(g, A, H, nu, K_PUSH_HEAVY_DEPOSIT);
        case 15: return launch_variant<true, true, true, true>(g, A, H, nu, K_PUSH_HEAVY_DEPOSIT);
        default: return set_error(PICG_ERR_ARG, "launch_step: unsupported mode %d", mode);
    }
}

// common driver: optional push (electron or heavy), optional deposit (full or partial), optional count
int species_step(picg_species_s* s, bool push, bool heavy, bool deposit, bool finalize, bool count, double dt,
                 picg_species_s* neutrals, picg_species_s* spherium, int sputtering) {
    int rc;
    if (deposit && finalize && !s->S_pinned && !s->S_calibrated) { rc = calibrate_scale(s, false); if (rc < 0) return rc; }
    size_t cap = std::max<size_t>(s->n_upper, 1), n_snapshot = 0;
    if (heavy) {
        rc = species_refresh_count(s); if (rc) return rc;
        n_snapshot = s->n_host; cap = std::max<size_t>(n_snapshot, 1);
        if (s->charge != 0) {                          // room for injected neutrals / sputtered material
            for (picg_species_s* t : {neutrals, sputtering ? spherium : neutrals}) {
                rc = species_refresh_count(t); if (rc) return rc;
                if (t->cap < t->n_host + 1024) { rc = species_ensure_capacity(t, t->n_host + t->n_host / 8 + 4096); if (rc) return rc; }
            }
        }
    }
    if (push) {
        if (cap >= 0xffffffffull) return set_error(PICG_ERR_ARG, "more than 2^32-1 particles per GPU are not supported");
        rc = ensure_scratch(s->w, compact_scratch_bytes(cap)); if (rc) return rc;
    }
    int mode = (push ? 1 : 0) | (heavy ? 2 : 0) | (deposit ? 4 : 0) | (count ? 8 : 0);
    rc = launch_step(s, mode, dt, neutrals, spherium, sputtering, n_snapshot); if (rc) return rc;
    if (heavy && s->charge != 0) {
        for (picg_species_s* t : {neutrals, spherium}) { t->n_host_valid = false; t->sorted_valid = false; t->n_upper = t->cap; }
    }
    if (push) { rc = compact_dead(s, cap); if (rc) return rc; }
    if (deposit && finalize) {
        rc = launch_finalize(s); if (rc) return rc;
        if (!s->S_pinned) { s->n_host_valid = false; rc = species_refresh_count(s); if (rc) return rc; return check_scale_after(s); }
    }
    return PICG_OK;
}
}  // namespace picg

extern "C" {

int picg_species_push_electrons(picg_species_t s, double dt) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_push_electrons: null species");
    return species_step(s, true, false, false, false, false, dt, nullptr, nullptr, 0);
}
int picg_species_push_heavy(picg_species_t s, picg_species_t neutrals, picg_species_t spherium, double dt, int sputtering) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && neutrals && spherium, "picg_species_push_heavy: null species");
    REQUIRE_ARG(neutrals->w == s->w && spherium->w == s->w, "picg_species_push_heavy: species belong to different worlds");
    return species_step(s, true, true, false, false, false, dt, neutrals, spherium, sputtering);
}
int picg_species_push_electrons_deposit(picg_species_t s, double dt, int count_cells) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_push_electrons_deposit: null species");
    return species_step(s, true, false, true, true, count_cells != 0, dt, nullptr, nullptr, 0);
}
int picg_species_push_heavy_deposit(picg_species_t s, picg_species_t neutrals, picg_species_t spherium, double dt, int sputtering, int count_cells) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && neutrals && spherium, "picg_species_push_heavy_deposit: null species");
    REQUIRE_ARG(neutrals->w == s->w && spherium->w == s->w, "picg_species_push_heavy_deposit: species belong to different worlds");
    return species_step(s, true, true, true, true, count_cells != 0, dt, neutrals, spherium, sputtering);
}
// multi-GPU: fused push + deposit into the raw accumulator only (all ranks must share a pinned scale); finalize after the all-reduce
int picg_species_push_deposit_partial(picg_species_t s, picg_species_t neutrals, picg_species_t spherium, double dt, int heavy, int sputtering, int count_cells) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_push_deposit_partial: null species");
    REQUIRE_ARG(s->S_pinned, "picg_species_push_deposit_partial: pin a common scale with picg_species_set_density_scale first");
    REQUIRE_ARG(!heavy || (neutrals && spherium), "picg_species_push_deposit_partial: heavy push needs the neutral and sputtered species");
    return species_step(s, true, heavy != 0, true, false, count_cells != 0, dt, neutrals, spherium, sputtering);
}
int picg_species_deposit_density(picg_species_t s) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_deposit_density: null species");
    return species_step(s, false, false, true, true, false, 0.0, nullptr, nullptr, 0);
}
int picg_species_deposit_density_partial(picg_species_t s) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_deposit_density_partial: null species");
    REQUIRE_ARG(s->S_pinned, "picg_species_deposit_density_partial: pin a common scale with picg_species_set_density_scale first (all ranks must share S)");
    return species_step(s, false, false, true, false, false, 0.0, nullptr, nullptr, 0);
}
int picg_species_count_per_cell(picg_species_t s) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_count_per_cell: null species");
    return species_step(s, false, false, false, false, true, 0.0, nullptr, nullptr, 0);
}

}  // extern "C"
