// Cell-group deposit: the fast path of Species::computeNumberDensity (ch4/v3/src/Species.cpp:401-413, Field::scatter
// Field.h:157-199) and Species::computeMacroParticlesCount (:813-819) for a species whose store carries a cell partition
// (cell_start[] of the last sort, possibly stale).  Same arithmetic per particle as k_run (step.cu); only the work
// decomposition and the data movement differ.
//
//   * STREAMS THROUGH TMA.  A warp owns a PASS of P = 32/G consecutive cells = one contiguous particle range.  Lane 0
//     issues one cp.async.bulk (global -> shared, mbarrier complete_tx, L2 evict-first) per particle array for the NEXT
//     chunk of the range while the warp computes on the current one (two stages per warp): the loads never occupy
//     registers and never stall the computing lanes.
//   * G LANES PER CELL (G = 4, 8, 16 or 32, chosen per launch from the mean cell population so that a lane sees ~5 particles
//     of its cell).  The particles of a cell are consecutive and all (but the stragglers) share the cell's eight nodes, so
//     the eight fixed-point corner sums stay in registers for the whole cell: no cell-change test, no shared-memory
//     atomics, no divergent flush.  "Still in its home cell" is 0 <= l - c < 1 on the already formed fractional
//     coordinates (exact, see below), tested on the high words - no float->int conversions.
//   * QUANTISATION ON THE FP64 PIPE.  q = llrint(w) is taken from the mantissa of w + 1.5*2^52 (round-to-nearest-even, the
//     same integer for 0 <= w < 2^51; larger weights take the generic path), the biased words are summed as integers and
//     the bias (trips x bits(1.5*2^52)) is removed once per cell.  No F2I on the slow conversion pipe.
//   * ONE EXIT PER CELL.  A transposed butterfly inside the lane group (the payload halves every round) leaves each
//     corner sum in one lane; the (.., k+1) sums are handed to the next group when its cell is the next cell of the same
//     column (they share the nodes), and what is left goes out as RED.64: about 4.5 global reductions per cell.
//   * Stragglers (particles that drifted out of the slot's home cell since the sort, or that the compaction moved into a
//     hole) deposit on their own with global reductions: correctness never depends on how stale the partition is.
//     Particles appended after the sort lie beyond the partition and go through k_run on that tail (step.cu).
// Algorithmic bytes per particle: deposit 32 B (pos + mpw), count 24 B.
#include "common.cuh"
#include "deposit.cuh"
#include "tma.cuh"
#include <algorithm>
#include <cmath>

using namespace picg;

#define CG_THREADS 256
#define CG_WARPS (CG_THREADS / 32)
#ifndef CG_CAP
#define CG_CAP 208                                  // particles per stage and array: the largest that lets two blocks share an SM (2 x 111 KB).  A pass longer than a
                                                    // stage is split and the lane groups whose cells ended in the first part idle through the second: 128 / 176 / 208 -> 14.2 / 10.2 / 9.7 ms
                                                    // per step for the three deposits of the bench (profiles/r2_deposit_stage.md; the last step 192 -> 208 is worth 2-3 %)
#endif
#ifndef CG_BLOCKS
#define CG_BLOCKS 2                                 // blocks per SM (registers per thread = 65536 / (256 * CG_BLOCKS))
#endif
#define CG_CHUNK (CG_CAP - 2)                       // the copied range is widened to even particle indices (16-byte alignment)
#define CG_MVB 64                                   // movers staged per warp before they go to the global list (one atomic per batch)

struct CellArgs {
    const double* a[4];                             // x y z mpw
    const SpeciesCounters* ctr; u64 n_limit;        // particles [0, n_limit) are covered by the partition (~0: the live count)
    const unsigned* cell_start; u64* den_fixed; double scale; double* macro_count;
    // optional by-product: (slot, current cell, home cell) of every particle found outside its slot's home cell (sort.cu: movers)
    unsigned *mv_slot, *mv_cell, *mv_home; u64* mv_count; u64* mv_listed; u64 mv_cap;
};

// ---------------------------------------------------------------- the kernel
// Uniform per warp (every lane holds the same values) except cs: lane l <= P holds cell_start[c0 + l] of the pass.
struct Cursor { int pass; unsigned cs, cs_next, pos, end; };
struct Item { int pass; unsigned cs, lo, hi; bool last; };

template <bool DEPOSIT, bool COUNT, int LG>
__global__ void __launch_bounds__(CG_THREADS, CG_BLOCKS) k_cell_deposit(Grid g, CellArgs A) {
    constexpr int NA = DEPOSIT ? 4 : 3;
    constexpr int G = 1 << LG, P = 32 >> LG;                    // lanes per cell, cells per pass
    extern __shared__ __align__(128) unsigned char cg_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* wbuf = reinterpret_cast<double*>(cg_smem) + (size_t)wib * 2 * NA * CG_CAP;
    u64* bars = reinterpret_cast<u64*>(reinterpret_cast<double*>(cg_smem) + (size_t)CG_WARPS * 2 * NA * CG_CAP) + wib * 2;
    unsigned* mvbuf = reinterpret_cast<unsigned*>(reinterpret_cast<u64*>(reinterpret_cast<double*>(cg_smem) + (size_t)CG_WARPS * 2 * NA * CG_CAP) + CG_WARPS * 2) + wib * 3 * CG_MVB;
    int mv_fill = 0;                                           // warp-uniform
    auto mv_flush = [&]() {                                    // staged (slot, cell, home) triples -> global list, one atomic per batch
        __syncwarp();
        u64 base = 0;
        if (lane == 0) base = atomicAdd(A.mv_count, (u64)mv_fill);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int e = lane; e < mv_fill; e += 32)
            if (base + e < A.mv_cap) { A.mv_slot[base + e] = mvbuf[e]; A.mv_cell[base + e] = mvbuf[CG_MVB + e]; A.mv_home[base + e] = mvbuf[2 * CG_MVB + e]; }
        __syncwarp();
        mv_fill = 0;
    };
    if (lane == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    const u64 policy = policy_evict_first();
    // the partition covers [0, n_lim): live particles that were in the store at the last sort
    const unsigned n_lim = (unsigned)min(min(A.n_limit == ~0ull ? A.ctr->n : A.n_limit, (u64)__ldg(A.cell_start + g.nc)), (u64)0xfffffff0u);
    if (A.mv_count && blockIdx.x == 0 && threadIdx.x == 0) *A.mv_listed = n_lim;     // slots beyond are left to the tail scan (sort.cu)
    const int npass = (g.nc + P - 1) / P;
    const int nw = gridDim.x * CG_WARPS;
    const int grp = lane >> LG, sub = lane & (G - 1);

    auto load_cs = [&](int pass) -> unsigned {
        if (pass >= npass) return 0u;
        return min(__ldg(A.cell_start + min(pass * P + min(lane, P), g.nc)), n_lim);
    };
    auto next_pass = [&](Cursor& F) {
        F.pass += nw; F.cs = F.cs_next; F.cs_next = load_cs(F.pass + nw);
        F.pos = __shfl_sync(0xffffffffu, F.cs, 0); F.end = __shfl_sync(0xffffffffu, F.cs, P);
    };
    auto skip_empty = [&](Cursor& F) { while (F.pass < npass && F.pos >= F.end) next_pass(F); };
    auto take = [&](Cursor& F) -> Item {                       // the chunk at the cursor; moves the cursor past it
        Item it; it.pass = F.pass; it.cs = F.cs; it.lo = F.pos; it.hi = min(F.pos + CG_CHUNK, F.end); it.last = it.hi == F.end;
        F.pos = it.hi;
        if (it.last) { next_pass(F); skip_empty(F); }
        return it;
    };
    auto issue = [&](const Item& it, int stage) {              // lane 0: one bulk copy per array into the stage
        if (lane == 0) {
            const unsigned a0 = it.lo & ~1u, a1 = (it.hi + 1) & ~1u, bytes = (a1 - a0) * 8;
            mbar_expect_tx(&bars[stage], NA * bytes);
#pragma unroll
            for (int c = 0; c < NA; c++) bulk_g2s(wbuf + (size_t)(stage * NA + c) * CG_CAP, A.a[c] + a0, bytes, &bars[stage], policy);
        }
    };

    Cursor F; F.pass = blockIdx.x * CG_WARPS + wib;
    F.cs = load_cs(F.pass); F.cs_next = load_cs(F.pass + nw);
    F.pos = __shfl_sync(0xffffffffu, F.cs, 0); F.end = __shfl_sync(0xffffffffu, F.cs, P);
    skip_empty(F);
    if (F.pass >= npass) return;                               // warp-uniform
    Item cur = take(F);
    issue(cur, 0);
    unsigned phase0 = 0, phase1 = 0; int stage = 0;

    constexpr double kMagic = 6755399441055744.0;              // 1.5 * 2^52: w + kMagic holds llrint(w) in its low mantissa bits for 0 <= w < 2^51
    constexpr int kMagicHi = 0x43380000;                       // high word of kMagic (its low word is 0)
    i64 acc[8]; int cnt = 0; int trips = 0;                    // acc carries trips * bits(kMagic) of bias until the flush
#pragma unroll
    for (int c = 0; c < 8; c++) acc[c] = 0;
    const double x0 = g.x0[0], y0 = g.x0[1], z0 = g.x0[2], idx = g.inv_dx[0], idy = g.inv_dx[1], idz = g.inv_dx[2];
    const int scale_hi = __double2hiint(A.scale);              // scale = 2^S: its low word is 0

    while (true) {
        const bool more = F.pass < npass;
        Item nxt = cur;
        if (more) { nxt = take(F); issue(nxt, stage ^ 1); }
        // ---- this lane's share of the chunk: cell c0+grp, particles [my_lo, my_hi), every G-th from sub
        const int cell = cur.pass * P + grp;
        const unsigned s_g = __shfl_sync(0xffffffffu, cur.cs, grp), e_g = __shfl_sync(0xffffffffu, cur.cs, grp + 1);
        const unsigned my_lo = max(s_g, cur.lo), my_hi = min(e_g, cur.hi);
        const int len = my_hi > my_lo ? (int)(my_hi - my_lo) : 0;
        const int ntrip = __reduce_max_sync(0xffffffffu, (len + G - 1) >> LG);
        int ci = 0, cj = 0, ck = 0;
        if (cell < g.nc) cell_to_ijk(g, cell, ci, cj, ck);
        const double dci = (double)ci, dcj = (double)cj, dck = (double)ck;
        const unsigned a0 = cur.lo & ~1u;
        const double* sx = wbuf + (size_t)(stage * NA) * CG_CAP; const double* sy = sx + CG_CAP; const double* sz = sy + CG_CAP; const double* sm = sz + CG_CAP;
        mbar_wait(&bars[stage], stage ? phase1 : phase0);
        if (stage) phase1 ^= 1; else phase0 ^= 1;
        trips += ntrip;
#pragma unroll 2
        for (int t = 0; t < ntrip; t++) {
            const unsigned p = my_lo + (t << LG) + sub;
            const bool ok = p < my_hi;
            const unsigned off = ok ? p - a0 : 0u;
            const double lx = x_to_l(sx[off], x0, idx), ly = x_to_l(sy[off], y0, idy), lz = x_to_l(sz[off], z0, idz);
            // (int)l == c  <=>  0 <= l - c < 1 (the subtraction is exact for l in [c, c+1)), and then l - c is the reference's
            // fractional weight l - (double)(int)l bit for bit (Field.h:161-169).  The reference's index clamp never acts here.
            // 0 <= d < 1  <=>  high word of d, as unsigned, below that of 1.0 (negative values and NaN have larger high words).
            const double di = __dsub_rn(lx, dci), dj = __dsub_rn(ly, dcj), dk = __dsub_rn(lz, dck);
            bool home = ok && (unsigned)__double2hiint(di) < 0x3ff00000u && (unsigned)__double2hiint(dj) < 0x3ff00000u && (unsigned)__double2hiint(dk) < 0x3ff00000u;
            const bool moved = ok && !home;
            if (DEPOSIT) {
                const double m = sm[off];
                home = home && (unsigned)__double2hiint(m) < (unsigned)(0x43200000 - (scale_hi - 0x3ff00000));   // 0 <= m * 2^S < 2^51
                const double odi = __dsub_rn(1.0, di), odj = __dsub_rn(1.0, dj), odk = __dsub_rn(1.0, dk);
                const double vs = __dmul_rn(m, __hiloint2double(home ? scale_hi : 0, 0));   // zero weight: every product below is +-0
                const double a_ = __dmul_rn(vs, odi), b_ = __dmul_rn(vs, di);
                const double w00 = __dmul_rn(a_, odj), w01 = __dmul_rn(a_, dj), w10 = __dmul_rn(b_, odj), w11 = __dmul_rn(b_, dj);
                acc[0] += __double_as_longlong(__dadd_rn(__dmul_rn(w00, odk), kMagic)); acc[1] += __double_as_longlong(__dadd_rn(__dmul_rn(w00, dk), kMagic));
                acc[2] += __double_as_longlong(__dadd_rn(__dmul_rn(w01, odk), kMagic)); acc[3] += __double_as_longlong(__dadd_rn(__dmul_rn(w01, dk), kMagic));
                acc[4] += __double_as_longlong(__dadd_rn(__dmul_rn(w10, odk), kMagic)); acc[5] += __double_as_longlong(__dadd_rn(__dmul_rn(w10, dk), kMagic));
                acc[6] += __double_as_longlong(__dadd_rn(__dmul_rn(w11, odk), kMagic)); acc[7] += __double_as_longlong(__dadd_rn(__dmul_rn(w11, dk), kMagic));
            }
            cnt += home ? 1 : 0;
            const bool stray = ok && !home;
            if (__any_sync(0xffffffffu, stray)) {
                int i = 0, j = 0, k = 0;
                if (stray) {                                                              // generic path, reference index rules
                    i64 q[8];
                    if (DEPOSIT) {
                        scatter_weights_fixed(g, lx, ly, lz, sm[off], A.scale, i, j, k, q);
#pragma unroll
                        for (int c = 0; c < 8; c++) if (q[c]) atomicAdd(&A.den_fixed[corner_node(g, i, j, k, c)], (u64)q[c]);
                    } else { i = min((int)lx, g.ci - 1); j = min((int)ly, g.cj - 1); k = min((int)lz, g.ck - 1); }
                    if (COUNT) atomicAdd(&A.macro_count[cell_of(g, i, j, k)], 1.0);
                }
                if (A.mv_count) {                                                         // warp-aggregated append to the staged mover list
                    const unsigned mask = __ballot_sync(0xffffffffu, moved);
                    if (mask) {
                        if (mv_fill + __popc(mask) > CG_MVB) mv_flush();
                        const int e = mv_fill + __popc(mask & ((1u << lane) - 1));
                        if (moved) { mvbuf[e] = p; mvbuf[CG_MVB + e] = (unsigned)cell_of(g, max(i, 0), max(j, 0), max(k, 0)); mvbuf[2 * CG_MVB + e] = (unsigned)cell; }
                        mv_fill += __popc(mask);
                    }
                }
            }
        }
        // ---- end of a pass: the lane groups' sums leave the registers.  Corner c = 4a + 2b + d (a, b, d = the i, j, k offsets);
        // the butterfly splits on a (lane bit G/2), then b (G/4), then d (G/8): the payload halves every round; the
        // remaining rounds (G > 8) are plain sums.
        if (cur.last) {
            const bool cell_ok = cell < g.nc;
            if (DEPOSIT) {
                const unsigned bias = (unsigned)trips * (unsigned)kMagicHi;              // the low word of bits(kMagic) is 0
#pragma unroll
                for (int c = 0; c < 8; c++) acc[c] -= (i64)((u64)bias << 32);
                const bool ba = sub & (G >> 1), bb = sub & (G >> 2);
                i64 v0 = (ba ? acc[4] : acc[0]) + __shfl_xor_sync(0xffffffffu, ba ? acc[0] : acc[4], G >> 1);
                i64 v1 = (ba ? acc[5] : acc[1]) + __shfl_xor_sync(0xffffffffu, ba ? acc[1] : acc[5], G >> 1);
                i64 v2 = (ba ? acc[6] : acc[2]) + __shfl_xor_sync(0xffffffffu, ba ? acc[2] : acc[6], G >> 1);
                i64 v3 = (ba ? acc[7] : acc[3]) + __shfl_xor_sync(0xffffffffu, ba ? acc[3] : acc[7], G >> 1);
                i64 t0 = (bb ? v2 : v0) + __shfl_xor_sync(0xffffffffu, bb ? v0 : v2, G >> 2);          // node (ci+a, cj+b, ck)
                i64 t1 = (bb ? v3 : v1) + __shfl_xor_sync(0xffffffffu, bb ? v1 : v3, G >> 2);          // node (ci+a, cj+b, ck+1)
                // the next group's cell is the next cell of the same column: its (.., ck') nodes are this group's (.., ck+1) nodes
                const bool chain = grp + 1 < P && cell + 1 < g.nc && ck + 1 < g.ck;
                const bool prev_chains = __shfl_up_sync(0xffffffffu, (int)chain, G) != 0 && grp > 0;
                u64* node = A.den_fixed + ((size_t)((ci + (ba ? 1 : 0)) * g.nj + (cj + (bb ? 1 : 0))) * g.nk + ck);
                if (LG == 2) {                                                            // lane holds (a, b; d): two nodes along k
                    const i64 from_prev = __shfl_up_sync(0xffffffffu, t1, G);
                    if (prev_chains) t0 += from_prev;
                    if (cell_ok) {
                        if (t0 != 0) atomicAdd(node, (u64)t0);
                        if (!chain && t1 != 0) atomicAdd(node + 1, (u64)t1);
                    }
                } else {                                                                  // one more split on d, then plain sums: one node per lane
                    const bool bd = sub & (G >> 3);
                    i64 u0 = (bd ? t1 : t0) + __shfl_xor_sync(0xffffffffu, bd ? t0 : t1, G >> 3);
#pragma unroll
                    for (int dist = G >> 4; dist >= 1; dist >>= 1) u0 += __shfl_xor_sync(0xffffffffu, u0, dist);
                    const i64 from_prev = __shfl_sync(0xffffffffu, u0, max(lane - G + (G >> 3), 0));   // same (a, b), d = 1, previous group
                    if (prev_chains && !bd) u0 += from_prev;
                    const bool writer = (sub & ((G >> 3) - 1)) == 0;
                    if (cell_ok && writer && u0 != 0 && !(bd && chain)) atomicAdd(node + (bd ? 1 : 0), (u64)u0);
                }
#pragma unroll
                for (int c = 0; c < 8; c++) acc[c] = 0;
            }
            if (COUNT) {
                int tot = cnt;
#pragma unroll
                for (int dist = G >> 1; dist >= 1; dist >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, dist);
                if (sub == 0 && cell_ok && tot != 0) atomicAdd(&A.macro_count[cell], (double)tot);         // integer-valued: exact in any order
            }
            cnt = 0; trips = 0;
        }
        if (!more) break;
        __syncwarp();                                          // every lane is done with this stage before it is refilled
        cur = nxt; stage ^= 1;
    }
    if (A.mv_count && mv_fill) mv_flush();
}

namespace picg {
int ensure_mover_triples(picg_species_s* s);                  // sort.cu
static size_t cell_smem_bytes(int na) { return (size_t)CG_WARPS * 2 * na * CG_CAP * 8 + CG_WARPS * 2 * 8 + (size_t)CG_WARPS * 3 * CG_MVB * 4; }

template <bool DEPOSIT, bool COUNT, int LG>
static int launch_cell_variant(const Grid& g, const CellArgs& A, int kid) {
    static bool attr_set = false;
    const size_t smem = cell_smem_bytes(DEPOSIT ? 4 : 3);
    if (!attr_set) { CUDA_TRY(cudaFuncSetAttribute(k_cell_deposit<DEPOSIT, COUNT, LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set = true; }
    const int npass = div_up((size_t)g.nc, 32 >> LG);
    int grid = std::max(1, std::min(div_up((size_t)npass, CG_WARPS), g_sm_count * CG_BLOCKS));
    LAUNCH(kid, (k_cell_deposit<DEPOSIT, COUNT, LG>), grid, CG_THREADS, smem, g, A);
    CHECK_LAUNCH();
    return PICG_OK;
}
template <int LG>
static int launch_cell_mode(const Grid& g, const CellArgs& A, int mode) {
    switch (mode) {
        case 4:  return launch_cell_variant<true, false, LG>(g, A, K_DEPOSIT);
        case 12: return launch_cell_variant<true, true, LG>(g, A, K_DEPOSIT);
        case 8:  return launch_cell_variant<false, true, LG>(g, A, K_COUNT_CELLS);
        default: return set_error(PICG_ERR_ARG, "launch_cell_step: unsupported mode %d", mode);
    }
}

// mode bits as in launch_step (step.cu): 4 deposit, 8 count (no push variants).  Covers particles [0, min(n_limit, partition)).
// n_est: an estimate of the particle count (sizes the lane groups: about 5 particles of a cell per lane).
int launch_cell_step(picg_species_s* s, int mode, size_t n_limit, size_t n_est) {
    const Grid& g = s->w->g;
    CellArgs A;
    A.a[0] = s->a[0]; A.a[1] = s->a[1]; A.a[2] = s->a[2]; A.a[3] = s->a[6];
    A.ctr = s->ctr; A.n_limit = n_limit; A.cell_start = s->cell_start;
    A.den_fixed = (u64*)s->den_fixed; A.scale = std::ldexp(1.0, s->S); A.macro_count = s->macro_count;
    A.mv_slot = A.mv_cell = A.mv_home = nullptr; A.mv_count = nullptr; A.mv_listed = nullptr; A.mv_cap = 0;
    const bool emit = s->wants_lists && n_limit == (size_t)-1;                    // per-cell lists are in use: list the movers on the fly
    if (emit) {
        int rc = ensure_mover_triples(s); if (rc) return rc;
        A.mv_slot = s->mv_trip; A.mv_cell = s->mv_trip + s->mv_trip_cap; A.mv_home = s->mv_trip + 2 * s->mv_trip_cap;
        A.mv_count = &s->ctr->n_movers; A.mv_listed = &s->ctr->n_listed; A.mv_cap = s->mv_trip_cap;
        CUDA_TRY(cudaMemsetAsync(&s->ctr->n_movers, 0, 8, g_stream));
    }
    s->movers_fresh = false;
    static const int force_lg = getenv("PICG_CELL_LG") ? atoi(getenv("PICG_CELL_LG")) : 0;             // tuning switch
    const double ppc = (double)n_est / std::max(1, g.nc);
    int lg = force_lg ? force_lg : (ppc <= 26 ? 2 : ppc <= 52 ? 3 : ppc <= 104 ? 4 : 5);
    int rc;
    switch (lg) {
        case 2:  rc = launch_cell_mode<2>(g, A, mode); break;
        case 3:  rc = launch_cell_mode<3>(g, A, mode); break;
        case 4:  rc = launch_cell_mode<4>(g, A, mode); break;
        default: rc = launch_cell_mode<5>(g, A, mode); break;
    }
    if (rc == PICG_OK && emit) { s->movers_fresh = true; s->movers_saved = false; }       // stays true until the particles move, die or are appended to
    return rc;
}
}  // namespace picg
