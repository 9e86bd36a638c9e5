// The particle-loop kernel: push (+ wall interaction) and/or fixed-point deposit (+ per-cell count) in one pass.
//
//   k_run<PUSH, HEAVY, DEPOSIT, COUNT>
//     PUSH            Species::advanceElectronsSerial              ch4/v3/src/Species.cpp:356-399
//     PUSH + HEAVY    Species::advanceNoSputteringSerial / ...SputteringSerial   :170-256 / :81-169
//     DEPOSIT         Species::computeNumberDensity                :401-413 (+ Field::scatter Field.h:157-199)
//     COUNT           Species::computeMacroParticlesCount          :813-819
//
// Data movement (every stage is HBM-bound; the design point is bytes in flight and instructions per particle):
//   * a thread owns a RUN of 4 CONSECUTIVE particles.  Each SoA array is read with one 256-bit load per run
//     (ld.global.cs.v4.f64 -> LDG.E.256, evict-first so the streamed particles do not push the field out of L2): a warp
//     covers 128 consecutive particles with 7 fully coalesced, mutually independent loads, i.e. 7 KB in flight per warp
//     before the first dependent instruction.  Results go back with 256-bit stores (full sectors).
//   * the four particles of a run are independent, so their E-field gathers (24 L1/L2 loads each) overlap.
//   * in a cell-sorted store a run stays inside one cell almost always: the eight fixed-point corner sums are
//     accumulated in registers and leave the thread once per run, not once per particle.  Run totals of lanes that end
//     in the same cell are combined with a transposed butterfly (deposit.cuh) and go to a per-warp shared-memory window
//     of RUN_WINDOW cells x 8 corners with 64-bit integer atomics; cells outside the window (unsorted input,
//     stragglers) go straight to global memory.  Integer sums are associative: any order gives the same bits.
//   * warps are independent (no block-level barrier): a persistent grid strides over 128-particle warp chunks.
// Algorithmic bytes per particle: push 96 B, deposit 32 B, fused 104 B (E-field and grid traffic are amortised over
// the ~60 particles of a cell and served by L1/L2).
#include "common.cuh"
#include "push.cuh"
#include "deposit.cuh"
#include "samplers.cuh"
#include "heavy.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

using namespace picg;

#define RUN_THREADS 256
#define RUN_WARPS (RUN_THREADS / 32)
#define RUN_LEN_PUSH 4                                       // particles per thread run when the kernel pushes (7 arrays live in registers); 2 and 4 are supported
// The electron push (gather + kick + drift, no wall interaction) is bound by latency, not by registers in flight: runs of 2 particles at
// 4 blocks per SM (64 registers, 32 warps) beat runs of 4 at 2 blocks (128 registers, 16 warps) by 12 %; the heavy push and the drift-only
// push of the neutrals do not gain (profiles/r2_push_runlen.md).
#ifndef ELE_RL
#define ELE_RL 2
#endif
#ifndef ELE_BLOCKS
#define ELE_BLOCKS 4
#endif
#ifndef HEAVY_RL
#define HEAVY_RL 4
#endif
#ifndef HEAVY_BLOCKS
#define HEAVY_BLOCKS 2
#endif
__host__ __device__ constexpr int run_len(bool push, bool heavy, bool deposit, bool drift) {
    return !push ? 8 : (deposit || drift) ? RUN_LEN_PUSH : heavy ? HEAVY_RL : ELE_RL;
}
__host__ __device__ constexpr int run_blocks(bool push, bool heavy, bool deposit, bool drift) {
    return (!push || deposit || drift) ? 2 : heavy ? HEAVY_BLOCKS : ELE_BLOCKS;
}
#define RUN_LEN_SCAN 8                                       // deposit / count only (4 arrays): longer runs, fewer flushes per particle
#define RUN_WINDOW 64                                        // nodes along k in the per-warp window (4 rows x 64 x 8 B = 2 KB)

struct StepArgs {
    double* a[7]; SpeciesCounters* ctr; u64 n_fixed; int use_fixed_n;      // heavy pushes walk a snapshot of the count (Species.cpp:176)
    const unsigned* tail_from;                                              // non-null: only particles [*tail_from, n) (the part beyond the cell partition)
    const double* ef; double qm_dt, dt;
    unsigned* dead_list; unsigned* impact_list; u64* den_fixed; double scale; double* macro_count;
};

__device__ __forceinline__ void ld4_stream(const double* p, double v[4]) {
    asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st4_stream(double* p, const double v[4]) {
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
__device__ __forceinline__ void ld2_stream(const double* p, double v[2]) {
    asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
}
__device__ __forceinline__ void st2_stream(double* p, const double v[2]) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(v[0]), "d"(v[1]) : "memory");
}
template <int RL>
__device__ __forceinline__ void load_run(const double* base, u64 p0, bool full, u64 lo, u64 n, double* v) {
    if (full) {
        if (RL == 2) ld2_stream(base + p0, v);
#pragma unroll
        for (int q = 0; q < RL / 4; q++) ld4_stream(base + p0 + 4 * q, v + 4 * q);
    } else {
#pragma unroll
        for (int r = 0; r < RL; r++) v[r] = (p0 + r >= lo && p0 + r < n) ? __ldcs(base + p0 + r) : 0.0;
    }
}
template <int RL>
__device__ __forceinline__ void store_run(double* base, u64 p0, bool full, u64 lo, u64 n, const double* v) {
    if (full) {
        if (RL == 2) st2_stream(base + p0, v);
#pragma unroll
        for (int q = 0; q < RL / 4; q++) st4_stream(base + p0 + 4 * q, v + 4 * q);
    } else {
#pragma unroll
        for (int r = 0; r < RL; r++) if (p0 + r >= lo && p0 + r < n) __stcs(base + p0 + r, v[r]);
    }
}

// DRIFT: the species is neutral (charge == 0), so the kick `v += E * (dt*q/m)` adds E*0 = 0 (Species.cpp:187-188): the E gather
// and the velocity write-back are skipped (72 B per particle instead of 96).  Same values as the reference for every finite E
// (a -0.0 velocity component would become +0.0 there and stays -0.0 here: equal under ==).
template <bool PUSH, bool HEAVY, bool DEPOSIT, bool COUNT, bool DRIFT = false>
__global__ void __launch_bounds__(RUN_THREADS, run_blocks(PUSH, HEAVY, DEPOSIT, DRIFT)) k_run(Grid g, StepArgs A, HeavyArgs H) {
    __shared__ unsigned s_lo[DEPOSIT ? RUN_WARPS : 1][RUN_WINDOW * 4], s_hi[DEPOSIT ? RUN_WARPS : 1][RUN_WINDOW * 4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned* wlo = s_lo[DEPOSIT ? wib : 0]; unsigned* whi = s_hi[DEPOSIT ? wib : 0];
    if (DEPOSIT) { for (int t = lane; t < RUN_WINDOW * 4; t += 32) { wlo[t] = 0; whi[t] = 0; } __syncwarp(); }
    const u64 n = A.use_fixed_n ? A.n_fixed : A.ctr->n;
    const u64 lo = A.tail_from ? (u64)*A.tail_from : 0;
    const u64 warp = (u64)blockIdx.x * RUN_WARPS + wib, nwarps = (u64)gridDim.x * RUN_WARPS;

    constexpr int RL = run_len(PUSH, HEAVY, DEPOSIT, DRIFT);
    constexpr int WARP_CHUNK = 32 * RL;
    for (u64 chunk = (lo & ~(u64)3) + warp * WARP_CHUNK; chunk < n; chunk += nwarps * WARP_CHUNK) {
        const u64 p0 = chunk + (u64)lane * RL;
        const bool full = p0 >= lo && p0 + RL <= n;
        double x[RL], y[RL], z[RL], u[PUSH ? RL : 1], v[PUSH ? RL : 1], w[PUSH ? RL : 1], m[RL];
        load_run<RL>(A.a[0], p0, full, lo, n, x); load_run<RL>(A.a[1], p0, full, lo, n, y); load_run<RL>(A.a[2], p0, full, lo, n, z);
        if (PUSH) { load_run<RL>(A.a[3], p0, full, lo, n, u); load_run<RL>(A.a[4], p0, full, lo, n, v); load_run<RL>(A.a[5], p0, full, lo, n, w); }
        if (DEPOSIT) load_run<RL>(A.a[6], p0, full, lo, n, m);       // a push alone needs the weight only at an ion impact (read there)
        NodeWindow W = {0, 0, 0};
        if (DEPOSIT) {                  // window placed at the column of the warp's first particle (2 nodes of slack below)
            int i = min(max((int)x_to_l(x[0], g.x0[0], g.inv_dx[0]), 0), g.ci - 1);
            int j = min(max((int)x_to_l(y[0], g.x0[1], g.inv_dx[1]), 0), g.cj - 1);
            int k = min(max((int)x_to_l(z[0], g.x0[2], g.inv_dx[2]), 0), g.ck - 1);
            W.wi = __shfl_sync(0xffffffffu, i, 0); W.wj = __shfl_sync(0xffffffffu, j, 0);
            W.k0 = min(max(__shfl_sync(0xffffffffu, k, 0) - 2, 0), max(g.nk - RUN_WINDOW, 0));
        }
        int cur = -1, cur_i = 0, cur_j = 0, cur_k = 0; i64 acc[8]; double cur_count = 0;
#pragma unroll
        for (int c = 0; c < 8; c++) acc[c] = 0;
        constexpr int kUnroll = (DEPOSIT && PUSH) ? 1 : RL;     // the fused body is too large for the instruction cache when unrolled
#pragma unroll kUnroll
        for (int r = 0; r < RL; r++) {
            const u64 p = p0 + r;
            const bool ok = p >= lo && p < n;
            bool dead = false, impact = false;
            if (PUSH && ok) {
                double un = u[r], vn = v[r], wn = w[r];
                if (!DRIFT) {
                    double ex, ey, ez;
                    gather_ef(g, A.ef, x_to_l(x[r], g.x0[0], g.inv_dx[0]), x_to_l(y[r], g.x0[1], g.inv_dx[1]), x_to_l(z[r], g.x0[2], g.inv_dx[2]), ex, ey, ez);
                    un = __dadd_rn(un, __dmul_rn(ex, A.qm_dt)); vn = __dadd_rn(vn, __dmul_rn(ey, A.qm_dt)); wn = __dadd_rn(wn, __dmul_rn(ez, A.qm_dt));
                }
                double xn = x[r], yn = y[r], zn = z[r];
                if (!HEAVY) {
                    xn = __dadd_rn(xn, __dmul_rn(un, A.dt)); yn = __dadd_rn(yn, __dmul_rn(vn, A.dt)); zn = __dadd_rn(zn, __dmul_rn(wn, A.dt));
                    dead = !in_bounds(g, xn, yn, zn) || in_object(g, xn, yn, zn) != 0;        // Species.cpp:375-388
                } else {
                    // first sub-move (t_rem = 1, vel*1 is exact): identical to the electron drift.  Only particles that end it
                    // inside an object take the out-of-line path with the remaining sub-moves (Species.cpp:194-249).
                    const double ox = xn, oy = yn, oz = zn;
                    xn = __dadd_rn(xn, __dmul_rn(un, A.dt)); yn = __dadd_rn(yn, __dmul_rn(vn, A.dt)); zn = __dadd_rn(zn, __dmul_rn(wn, A.dt));
                    int obj = in_object(g, xn, yn, zn);
                    if (!in_bounds(g, xn, yn, zn)) dead = true;
                    else if (obj) impact = true;          // rare: the whole particle is re-done by k_heavy_impacts from its untouched state
                }
                if (!dead && !impact) { x[r] = xn; y[r] = yn; z[r] = zn; u[r] = un; v[r] = vn; w[r] = wn; }
            }
            if (PUSH) record_dead(dead, lane, p, A.ctr, A.dead_list);
            if (HEAVY) record_index(impact, lane, p, &A.ctr->n_impact, A.impact_list);
            int newcell = -1, ni_ = 0, nj_ = 0, nk_ = 0; bool have = false; i64 qq[8];
            if ((DEPOSIT || COUNT) && ok && !dead && !impact) {
                int ci, cj, ck; i64 q[8];
                if (DEPOSIT) scatter_weights_fixed(g, x_to_l(x[r], g.x0[0], g.inv_dx[0]), x_to_l(y[r], g.x0[1], g.inv_dx[1]), x_to_l(z[r], g.x0[2], g.inv_dx[2]),
                                                   m[r], A.scale, ci, cj, ck, q);
                else {
                    ci = min((int)x_to_l(x[r], g.x0[0], g.inv_dx[0]), g.ci - 1); cj = min((int)x_to_l(y[r], g.x0[1], g.inv_dx[1]), g.cj - 1);
                    ck = min((int)x_to_l(z[r], g.x0[2], g.inv_dx[2]), g.ck - 1);
                }
                int cell = cell_of(g, ci, cj, ck);
                newcell = cell; ni_ = ci; nj_ = cj; nk_ = ck; have = true;
                if (DEPOSIT) {
#pragma unroll
                    for (int c = 0; c < 8; c++) qq[c] = q[c];
                }
            }
            if (have) {
                if (newcell != cur) {                                   // the run leaves its cell: hand the partial sums over
                    if (cur >= 0) {
                        if (DEPOSIT) run_flush<RUN_WINDOW>(g, W, wlo, whi, cur_i, cur_j, cur_k, acc, A.den_fixed);
                        if (COUNT) atomicAdd(&A.macro_count[cur], cur_count);
                    }
                    cur = newcell; cur_i = ni_; cur_j = nj_; cur_k = nk_; cur_count = 0;
#pragma unroll
                    for (int c = 0; c < 8; c++) acc[c] = 0;
                }
                if (DEPOSIT) {
#pragma unroll
                    for (int c = 0; c < 8; c++) acc[c] += qq[c];
                }
                cur_count += 1.0;
            }
        }
        // results back to the store (dead slots keep their old contents; the compaction fills them)
        if (PUSH) {
            store_run<RL>(A.a[0], p0, full, lo, n, x); store_run<RL>(A.a[1], p0, full, lo, n, y); store_run<RL>(A.a[2], p0, full, lo, n, z);
            if (!DRIFT) { store_run<RL>(A.a[3], p0, full, lo, n, u); store_run<RL>(A.a[4], p0, full, lo, n, v); store_run<RL>(A.a[5], p0, full, lo, n, w); }
        }
        // end of the run: the register sums go to the warp's window, the window to the global grid
        if (DEPOSIT) {
            if (cur >= 0) run_flush<RUN_WINDOW>(g, W, wlo, whi, cur_i, cur_j, cur_k, acc, A.den_fixed);
            window_flush<RUN_WINDOW>(g, wlo, whi, W, A.den_fixed, lane);
        }
        if (COUNT) {
            unsigned peers = __match_any_sync(0xffffffffu, cur);
            double s = 0; unsigned mm = peers;                          // integer-valued counts: exact in any order
            while (mm) { int src = __ffs(mm) - 1; mm &= mm - 1; s += __shfl_sync(peers, cur_count, src); }
            if (cur >= 0 && lane == __ffs(peers) - 1) atomicAdd(&A.macro_count[cur], s);
        }
    }
}

// Second pass of the heavy push: the (few) particles whose first sub-move ended inside an object.  Each is re-done from
// its untouched state: kick, first sub-move, then the reference's bounce loop (surface hit, diffuse re-emission of
// neutrals, neutralisation of ions with injection of neutrals / sputtered material), Species.cpp:194-249.
// The impact list arrives sorted by slot (sort_slot_list).  Ions emit particles into other stores; where they land must not depend
// on which thread is first at a cursor, so a block takes 128 consecutive impacts, runs their bounce loops ONCE WITHOUT WRITING to count
// what each emits, gets the number emitted by all earlier impacts from a chain through global memory (blocks take tickets in starting
// order and pass the running totals on in ticket order), and runs the loops again, writing every emitted particle to its fixed slot.
struct EmitChain { unsigned long long ticket, turn, run_n, run_s; };
#define IMPACT_THREADS 128
__device__ __forceinline__ void block_exclusive_scan2(unsigned a, unsigned b, unsigned* sh /* 2 x 4 + 2 */, unsigned& ea, unsigned& eb, unsigned& ta, unsigned& tb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned xa = a, xb = b;
    for (int o = 1; o < 32; o <<= 1) { unsigned ya = __shfl_up_sync(0xffffffffu, xa, o), yb = __shfl_up_sync(0xffffffffu, xb, o); if (lane >= o) { xa += ya; xb += yb; } }
    if (lane == 31) { sh[warp] = xa; sh[4 + warp] = xb; }
    __syncthreads();
    unsigned ba = 0, bb = 0; ta = 0; tb = 0;
    for (int q = 0; q < IMPACT_THREADS / 32; q++) { if (q < warp) { ba += sh[q]; bb += sh[4 + q]; } ta += sh[q]; tb += sh[4 + q]; }
    ea = ba + xa - a; eb = bb + xb - b;
    __syncthreads();
}
template <bool DEPOSIT, bool COUNT>
__global__ void __launch_bounds__(IMPACT_THREADS) k_heavy_impacts(Grid g, StepArgs A, HeavyArgs H, EmitChain* chain) {
    __shared__ unsigned sh_scan[10]; __shared__ unsigned long long sh_v, sh_base[2];
    const u64 n_imp = A.ctr->n_impact;
    const bool emits = H.charge != 0;                                   // neutrals only bounce
    const u64 n0_n = emits ? H.neutrals.ctr->n : 0, n0_s = emits ? H.spherium.ctr->n : 0;      // nobody moves these counters while the kernel runs
    for (;;) {
        if (threadIdx.x == 0) sh_v = atomicAdd(&chain->ticket, 1ull);
        __syncthreads();
        const u64 vb = sh_v;
        if (vb * IMPACT_THREADS >= n_imp) break;
        const u64 t = vb * IMPACT_THREADS + threadIdx.x;
        const bool active = t < n_imp;
        const u64 p = active ? A.impact_list[t] : 0;
        double x = 0, y = 0, z = 0, u = 0, v = 0, w = 0, m = 0; int obj = 0;
        HeavyState st0 = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (active) {
            x = A.a[0][p]; y = A.a[1][p]; z = A.a[2][p]; u = A.a[3][p]; v = A.a[4][p]; w = A.a[5][p]; m = A.a[6][p];
            double ex, ey, ez;
            gather_ef(g, A.ef, x_to_l(x, g.x0[0], g.inv_dx[0]), x_to_l(y, g.x0[1], g.inv_dx[1]), x_to_l(z, g.x0[2], g.inv_dx[2]), ex, ey, ez);
            u = __dadd_rn(u, __dmul_rn(ex, A.qm_dt)); v = __dadd_rn(v, __dmul_rn(ey, A.qm_dt)); w = __dadd_rn(w, __dmul_rn(ez, A.qm_dt));
            HeavyState s0 = {x, y, z, __dadd_rn(x, __dmul_rn(u, A.dt)), __dadd_rn(y, __dmul_rn(v, A.dt)), __dadd_rn(z, __dmul_rn(w, A.dt)), u, v, w};
            st0 = s0;
            obj = in_object(g, st0.x, st0.y, st0.z);
        }
        HeavyArgs Hl = H;                                               // this thread's copy: emit_particle counts / numbers its particles in it
        if (emits) {
            if (active) { HeavyState st = st0; heavy_after_impact(g, Hl, A.ef, A.dt, p, obj, m, st); }      // counting pass: nothing is written
            unsigned en, es, tn, ts;
            block_exclusive_scan2(Hl.neutrals.made, Hl.spherium.made, sh_scan, en, es, tn, ts);
            if (threadIdx.x == 0) {
                while (atomicAdd(&chain->turn, 0ull) != vb) { }
                const unsigned long long rn = atomicAdd(&chain->run_n, (unsigned long long)tn), rs = atomicAdd(&chain->run_s, (unsigned long long)ts);
                __threadfence();
                atomicExch(&chain->turn, vb + 1);
                sh_base[0] = rn; sh_base[1] = rs;
            }
            __syncthreads();
            Hl.neutrals.write = 1; Hl.neutrals.made = 0; Hl.neutrals.base = n0_n + sh_base[0] + en;
            Hl.spherium.write = 1; Hl.spherium.made = 0; Hl.spherium.base = n0_s + sh_base[1] + es;
        }
        if (active) {
            HeavyState st = st0;
            bool gone = heavy_after_impact(g, Hl, A.ef, A.dt, p, obj, m, st);
            if (gone) A.dead_list[atomicAdd(&A.ctr->n_dead, 1ull)] = (unsigned)p;
            else {
                A.a[0][p] = st.x; A.a[1][p] = st.y; A.a[2][p] = st.z; A.a[3][p] = st.u; A.a[4][p] = st.v; A.a[5][p] = st.w;
                if (DEPOSIT || COUNT) {
                    int ci, cj, ck; i64 q[8];
                    scatter_weights_fixed(g, x_to_l(st.x, g.x0[0], g.inv_dx[0]), x_to_l(st.y, g.x0[1], g.inv_dx[1]), x_to_l(st.z, g.x0[2], g.inv_dx[2]), m, A.scale, ci, cj, ck, q);
                    if (DEPOSIT) { for (int c = 0; c < 8; c++) if (q[c]) atomicAdd(&A.den_fixed[corner_node(g, ci, cj, ck, c)], (u64)q[c]); }
                    if (COUNT) atomicAdd(&A.macro_count[cell_of(g, ci, cj, ck)], 1.0);
                }
            }
        }
        __syncthreads();                                                // sh_v / sh_base are reused by the next round
    }
}
// after the kernel: the emitted particles join their stores (a store that ran full keeps what fitted; the rest is counted as overflow)
__global__ void k_emit_finish(const EmitChain* chain, SpeciesCounters* cn, u64 cap_n, SpeciesCounters* cs, u64 cap_s, int same) {
    const u64 wn = cn->n + chain->run_n;
    if (wn > cap_n) { cn->overflow += wn - cap_n; cn->n = cap_n; } else cn->n = wn;
    if (!same) { const u64 wsn = cs->n + chain->run_s; if (wsn > cap_s) { cs->overflow += wsn - cap_s; cs->n = cap_s; } else cs->n = wsn; }
}

__global__ void k_clamp_count(SpeciesCounters* ctr, u64 cap);       // species.cu
namespace picg {
int launch_finalize(picg_species_s* s, size_t u_begin = 0, size_t u_end = (size_t)-1);
int calibrate_scale(picg_species_s* s, bool count_cells);
int check_scale_after(picg_species_s* s);

template <bool PUSH, bool HEAVY, bool DEPOSIT, bool COUNT, bool DRIFT = false>
static int launch_variant(const Grid& g, const StepArgs& A, const HeavyArgs& H, size_t n_upper, int kid) {
    constexpr int chunk = 32 * run_len(PUSH, HEAVY, DEPOSIT, DRIFT) * RUN_WARPS;
    int grid = std::max(1, std::min(div_up(std::max<size_t>(n_upper, 1), chunk), g_sm_count * run_blocks(PUSH, HEAVY, DEPOSIT, DRIFT) * 4));
    LAUNCH(kid, (k_run<PUSH, HEAVY, DEPOSIT, COUNT, DRIFT>), grid, RUN_THREADS, 0, g, A, H);
    CHECK_LAUNCH();
    return PICG_OK;
}

// mode bits: 1 push, 2 heavy, 4 deposit, 8 count
int launch_cell_step(picg_species_s* s, int mode, size_t n_limit, size_t n_est);     // cellstep.cu: deposit / count over the cell partition

int launch_step(picg_species_s* s, int mode, double dt, picg_species_s* neutrals, picg_species_s* spherium, int sputtering, size_t n_snapshot) {
    const Grid& g = s->w->g;
    StepArgs A;
    A.tail_from = nullptr;
    for (int c = 0; c < 7; c++) A.a[c] = s->a[c];
    A.ctr = s->ctr; A.use_fixed_n = 0; A.n_fixed = n_snapshot;                     // (the snapshot is an upper bound only: see species_step)
    A.ef = s->w->ef; A.qm_dt = dt * s->charge / s->mass; A.dt = dt;                 // Species.cpp:372 `dt*charge/mass`
    A.dead_list = (unsigned*)s->w->scratch; A.impact_list = (unsigned*)((char*)s->w->scratch + compact_scratch_bytes(std::max<size_t>(n_snapshot, 1)));
    A.den_fixed = (u64*)s->den_fixed; A.scale = std::ldexp(1.0, s->S); A.macro_count = s->macro_count;
    HeavyArgs H; memset(&H, 0, sizeof(H));
    if (mode & 2) {
        uint32_t call = ++s->n_heavy_calls;
        H.neutrals = emit_of(neutrals); H.spherium = emit_of(spherium); H.sputtering = (sputtering && s->charge != 0) ? 1 : 0;
        H.charge = s->charge; H.mass = s->mass; H.half_world_dt = 0.5 * s->w->dt; H.seed = g_seed; H.stream = rng_stream_id(RNG_HEAVY, s->id, g_rank); H.call = call;
    }
    if (mode & 4) cudaMemsetAsync(s->den_fixed, 0, (size_t)g.nv * 8, g_stream);
    if (mode & 8) cudaMemsetAsync(s->macro_count, 0, (size_t)g.nc * 8, g_stream);
    size_t nu = (mode & 2) ? n_snapshot : s->n_upper;
    // fast path of the deposit passes: the store carries a cell partition (cell_start[] of the last sort, possibly stale) and
    // is dense enough for a lane group per cell.  (A count alone is a pure stream: the thread-run kernel is at 85 % of HBM.)
    static const bool no_cell_path = getenv("PICG_NO_CELL_PATH") && atoi(getenv("PICG_NO_CELL_PATH")) != 0;      // A/B switch, see DESIGN.md
    if (!no_cell_path && (mode == 4 || mode == 12) && s->part_valid && nu >= (size_t)6 * g.nc) {
        int rc = launch_cell_step(s, mode, (size_t)-1, nu); if (rc) return rc;
        if (nu <= s->part_n) return PICG_OK;              // nothing was appended since the sort
        A.tail_from = s->cell_start + g.nc;               // the appended tail goes through the generic kernel
        nu = nu - s->part_n;
    }
    if (mode & 2) CUDA_TRY(cudaMemsetAsync(&s->ctr->n_impact, 0, 8, g_stream));
    static const bool no_drift_path = getenv("PICG_NO_DRIFT_PATH") && atoi(getenv("PICG_NO_DRIFT_PATH")) != 0;   // A/B switch: neutrals through the kick + gather
    int rc = PICG_OK;
    switch (mode) {
        case 1:  rc = launch_variant<true, false, false, false>(g, A, H, nu, K_PUSH_ELECTRONS); break;
        case 3:  rc = (A.qm_dt == 0.0 && !no_drift_path) ? launch_variant<true, true, false, false, true>(g, A, H, nu, K_PUSH_NEUTRAL)
                                                          : launch_variant<true, true, false, false>(g, A, H, nu, K_PUSH_HEAVY); break;
        // the thread-run kernel on the tail appended beyond the cell partition is timed on its own (the cell-group kernel did the rest)
        case 4:  rc = launch_variant<false, false, true, false>(g, A, H, nu, A.tail_from ? K_DEPOSIT_TAIL : K_DEPOSIT); break;
        case 12: rc = launch_variant<false, false, true, true>(g, A, H, nu, A.tail_from ? K_DEPOSIT_TAIL : K_DEPOSIT); break;
        case 8:  rc = launch_variant<false, false, false, true>(g, A, H, nu, K_COUNT_CELLS); break;
        case 5:  rc = launch_variant<true, false, true, false>(g, A, H, nu, K_PUSH_DEPOSIT); break;
        case 13: rc = launch_variant<true, false, true, true>(g, A, H, nu, K_PUSH_DEPOSIT); break;
        case 7:  rc = launch_variant<true, true, true, false>(g, A, H, nu, K_PUSH_HEAVY_DEPOSIT); break;
        case 15: rc = launch_variant<true, true, true, true>(g, A, H, nu, K_PUSH_HEAVY_DEPOSIT); break;
        default: return set_error(PICG_ERR_ARG, "launch_step: unsupported mode %d", mode);
    }
    if (rc) return rc;
    if (mode & 2) {                                   // surface interactions of the particles that hit an object (usually a handful)
        if (s->charge != 0 && H.sputtering && spherium == neutrals) return set_error(PICG_ERR_ARG, "heavy push: the sputtered material needs a store of its own");
        const size_t lcap = std::max<size_t>(n_snapshot, 1);
        unsigned* counts = (unsigned*)s->w->scratch + 2 * lcap;          // the compaction's block counts: free until compact_dead runs
        rc = sort_slot_list(s->w, &s->ctr->n_impact, &s->ctr->n, lcap, s->cap, A.impact_list, A.impact_list, counts); if (rc) return rc;      // ascending slots, in place
        EmitChain* chain = (EmitChain*)(s->w->reduce_buf + 4000);        // 4 words at the end of the reduction buffer
        CUDA_TRY(cudaMemsetAsync(chain, 0, sizeof(EmitChain), g_stream));
        int grid = g_sm_count * 2;
        if ((mode & 4) && (mode & 8)) LAUNCH(K_HEAVY_IMPACTS, (k_heavy_impacts<true, true>), grid, IMPACT_THREADS, 0, g, A, H, chain);
        else if (mode & 4) LAUNCH(K_HEAVY_IMPACTS, (k_heavy_impacts<true, false>), grid, IMPACT_THREADS, 0, g, A, H, chain);
        else if (mode & 8) LAUNCH(K_HEAVY_IMPACTS, (k_heavy_impacts<false, true>), grid, IMPACT_THREADS, 0, g, A, H, chain);
        else LAUNCH(K_HEAVY_IMPACTS, (k_heavy_impacts<false, false>), grid, IMPACT_THREADS, 0, g, A, H, chain);
        CHECK_LAUNCH();
        if (s->charge != 0) {                         // the emitted particles join their stores; a store that ran full keeps what fitted
            LAUNCH(K_HEAVY_IMPACTS, k_emit_finish, 1, 1, 0, (const EmitChain*)chain, neutrals->ctr, (u64)neutrals->cap, spherium->ctr, (u64)spherium->cap, spherium == neutrals ? 1 : 0); CHECK_LAUNCH();
        }
    }
    return PICG_OK;
}

// common driver: optional push (electron or heavy), optional deposit (full or partial), optional count
int species_step(picg_species_s* s, bool push, bool heavy, bool deposit, bool finalize, bool count, double dt,
                 picg_species_s* neutrals, picg_species_s* spherium, int sputtering) {
    int rc;
    if (deposit && finalize && !s->S_pinned && !s->S_calibrated) { rc = calibrate_scale(s, false); if (rc < 0) return rc; }
    size_t cap = std::max<size_t>(s->n_upper, 1), n_snapshot = 0;
    if (heavy) {
        // The reference walks a snapshot of the count (Species.cpp:176) because its loop appends to the store it walks.  Here nothing appends to the
        // store being pushed while its kernels run (emitted particles join their stores after the kernels, k_emit_finish), so the kernels read the
        // count on the device and the host only needs an upper bound: no synchronisation.
        n_snapshot = s->n_upper; cap = std::max<size_t>(n_snapshot, 1);
        if (s->charge != 0) {                          // room for injected neutrals / sputtered material
            for (picg_species_s* t : {neutrals, sputtering ? spherium : neutrals}) {
                const double per_ion_ub = std::min(64.0, s->mpw0 / t->mpw0 + 1.0);
                if (t->cap >= t->n_upper + (size_t)(per_ion_ub * (double)n_snapshot / 16.0) + 65536) continue;      // enough room by the upper bounds: no need for the exact count
                rc = species_refresh_count(t); if (rc) return rc;
                // every impacting ion emits int(mpw / mpw0_target + rnd()) particles (Species.cpp:225-232); at most a few per cent of a
                // species reach an electrode in one (possibly sub-cycled) push.  A store that still runs full is safe: the appends beyond
                // the capacity are dropped and counted, the counter is clamped (k_clamp_count below), the next refresh reports PICG_ERR_OOM.
                const double per_ion = std::min(64.0, s->mpw0 / t->mpw0 + 1.0);
                const size_t room = (size_t)(per_ion * (double)n_snapshot / 16.0) + 65536;
                if (t->cap < t->n_host + room) { rc = species_ensure_capacity(t, t->n_host + room + t->n_host / 8); if (rc) return rc; }
            }
        }
    }
    if (push) {
        if (cap >= 0xffffffffull) return set_error(PICG_ERR_ARG, "more than 2^32-1 particles per GPU are not supported");
        rc = ensure_scratch(s->w, compact_scratch_bytes(cap) + (heavy ? cap * 4 + 64 : 0)); if (rc) return rc;     // + the heavy push's impact list
    }
    if (deposit) count = true;                       // the per-cell count is a by-product of the deposit pass (computeMacroParticlesCount then costs nothing)
    int mode = (push ? 1 : 0) | (heavy ? 2 : 0) | (deposit ? 4 : 0) | (count ? 8 : 0);
    rc = launch_step(s, mode, dt, neutrals, spherium, sputtering, n_snapshot); if (rc) return rc;
    if (heavy && s->charge != 0) {
        for (picg_species_s* t : {neutrals, spherium}) { t->n_host_valid = false; t->sorted_valid = false; t->lists_valid = false; t->count_valid = false; t->n_upper = t->cap; }
    }
    if (push) { rc = compact_dead(s, cap); if (rc) return rc; }
    if (count) s->count_valid = true;                // counted at the post-push positions; the compaction only permutes survivors
    if (deposit && finalize) {
        rc = launch_finalize(s); if (rc) return rc;
        if (!s->S_pinned) {
            s->n_host_valid = false; rc = species_refresh_count(s); if (rc) return rc;
            if (s->ctr_host->den_neg) {                     // overflow (the population grew by orders of magnitude): re-calibrate, deposit once more
                s->S_calibrated = false;
                rc = calibrate_scale(s, false); if (rc < 0) return rc;
                rc = launch_step(s, 4, 0.0, nullptr, nullptr, 0, 0); if (rc) return rc;
                rc = launch_finalize(s); if (rc) return rc;
                s->n_host_valid = false; rc = species_refresh_count(s); if (rc) return rc;
            }
            return check_scale_after(s);
        }
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
    if (s->count_valid) return PICG_OK;              // already produced by the last deposit pass over the same particle state
    return species_step(s, false, false, false, false, true, 0.0, nullptr, nullptr, 0);
}

}  // extern "C"
