// Variable-weight Monte-Carlo electron-neutral collisions with electron-impact ionisation.
//   k_mcc : MC_MEX_Ionization::apply_vector_indexes   ch4/v3/src/Interactions.cpp:600-762
//           collide :885-947, newVelocityElecton :845-862, evaluateSigmaColl :541-558, evaluateSigmaIon :559-566
// Both species are cell-sorted on the device (sort.cu); cell c's particles are the index ranges
// [cell_start[c], cell_start[c+1]).  One thread owns one cell and runs the reference's candidate loop
// sequentially (weights and the neutral list mutate inside the loop, :687-725), cells are independent.
// RNG: Philox stream (RNG_MCC) addressed by (cell, apply-call number), so results are independent of the
// launch geometry.
//   k_mcc<1> : the fixed-weight ancestor of the same interaction, MC_MEX_Ionization::apply of ch4/v2
//           (ch4/v2/Interactions.cpp:566-641, collide :678-735, evaluateSigmaIon :560-563): Bird's NTC candidate count with
//           neutrals.mpw0, unweighted acceptance sigma*v_rel / (sigma*v_rel)_max, no ionisation threshold guard, products created
//           through Species::addParticle (bounds / object filter and half-step rewind, ch4/v2/Species.cpp:226-237), neutrals never
//           depleted.  BASELINE.json config 3.
// Appends.  Every product is a new particle at the end of a store.  One atomic per product on the store's counter serialises
// at the L2 (one address: ~0.45 ns each, 9 ms for the 2e7 split-off neutrals of a late step).  The candidate loop is therefore
// warp-synchronous: a warp owns 32 consecutive cells (lane = cell), iteration t handles the t-th candidate of every cell, and the
// products of one iteration take consecutive slots of ONE atomicAdd per store (warp_reserve).  Products of neighbouring cells end
// up next to each other in the appended tail, which the tail deposit likes.
// Staging.  The products do not go to the stores directly: every product is written as one 64-byte record (x y z u v w mpw) plus its
// cell into a staging area, indexed by the cursors above.  After the kernel the records of a store are ordered by (cell, order of
// creation inside the cell) with the counting sort of sort.cu - the key is the cell of the thread that made the product, no position
// is looked at - and appended to the store in that order (k_stage_commit).  Two things follow: the appended tail of a call is in cell
// order (its deposit runs on register sums and shared-memory windows instead of eight global reductions per particle), and it no longer
// depends on the order in which the warps got their turn at the cursors: the same input gives the same stores, bit for bit.
#include "common.cuh"
#include "philox.cuh"
#include "celllists.cuh"
#include "push.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

using namespace picg;

namespace picg {
int sort_species(picg_species_s* s); int species_exact_lists(picg_species_s* s); int species_prepare_lists(picg_species_s* s);
int species_grow_store(picg_species_s* s, size_t cap); int species_scratch_for_store(picg_species_s* s);      // species.cu
size_t counting_sort_words(const Grid& g);                  // sort.cu
int scan_cell_table(const Grid& g, unsigned* table, unsigned* work);
}

#define MCC_EXTRA 16          // split-off neutrals created in this call that remain selectable within the cell (:699-701)

struct MccParams {
    double m_n, m_e, sum_mass, E_rel_eV, E_ele_eV, two_qe_me, c0, c1, c2, B_inc, E_ion_eV, inv_dv, rank_scale;
    int world, rank;
    int n_tab; const double* tab_E; const double* tab_s;
    // fixed-weight variant (ch4/v2): weights of the created particles, Species::addParticle's rewind
    int fixed_weight; double neu_mpw0, ion_mpw0, ele_mpw0; int ions_to_create; const double* ef; double qm_ion, qm_ele, half_dt;
};
// evaluateSigmaColl (:541-558): std::map lower_bound + linear interpolation, clamped to the end values
__host__ __device__ __forceinline__ double sigma_coll(const MccParams& P, double E) {
    int lo = 0, hi = P.n_tab;                       // first index with tab_E >= E
    while (lo < hi) { int mid = (lo + hi) >> 1; if (P.tab_E[mid] < E) lo = mid + 1; else hi = mid; }
    if (lo == 0) return P.tab_s[0];
    if (lo == P.n_tab) return P.tab_s[P.n_tab - 1];
    double x1 = P.tab_E[lo - 1], x2 = P.tab_E[lo], y1 = P.tab_s[lo - 1], y2 = P.tab_s[lo];
    return y1 + (E - x1) * (y2 - y1) / (x2 - x1);
}
// evaluateSigmaIon (:559-566)
__host__ __device__ __forceinline__ double sigma_ion(const MccParams& P, double E) {
    if (!P.fixed_weight && E <= P.E_ion_eV) return 0;                                  // the guard is new in v3; ch4/v2 (:560-563) evaluates the fit everywhere
    return P.c0 * log(E / P.c1) / E * exp(-P.c2 / E);
}
// newVelocityElecton, IONIZE_1 / LAB frame (:845-862); note i x u is not normalised
__device__ __forceinline__ void new_velocity_electron(PhiloxStream& r, const MccParams& P, double E, const double u[3], double out[3]) {
    double cos_ksi = (2 + E - 2 * pow(1 + E, r.next())) / E;
    double sin_ksi = sqrt(1 - cos_ksi * cos_ksi);
    double phi = 2 * 3.141592653 * r.next();
    double v_mag = sqrt(E * P.two_qe_me);
    double ixu[3] = {0.0 * u[2] - 0.0 * u[1], 0.0 * u[0] - 1.0 * u[2], 1.0 * u[1] - 0.0 * u[0]};        // (1,0,0) x u
    double uxi[3] = {u[1] * ixu[2] - u[2] * ixu[1], u[2] * ixu[0] - u[0] * ixu[2], u[0] * ixu[1] - u[1] * ixu[0]};   // u x (i x u)
    double sp = sin(phi), cp = cos(phi);
    for (int c = 0; c < 3; c++) out[c] = (cos_ksi * u[c] + ixu[c] * sin_ksi * sp + uxi[c] * sin_ksi * cp) * v_mag;
}
// collide (:885-947).  vel_neu is never modified by the reference.  Returns ionised flag.
__device__ __forceinline__ bool collide(PhiloxStream& r, const MccParams& P, const double vn[3], double ve[3], double vnew[3], double s_coll) {
    double g[3] = {vn[0] - ve[0], vn[1] - ve[1], vn[2] - ve[2]};
    double g_mag = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    double m_r = P.m_n * P.m_e / P.sum_mass;
    double E_rel_J = 0.5 * m_r * g_mag * g_mag;
    double Pion = sigma_ion(P, E_rel_J / 1.602176565e-19) / s_coll;
    if (r.next() <= Pion) {
        double ve_mag = sqrt(ve[0] * ve[0] + ve[1] * ve[1] + ve[2] * ve[2]);
        double E_inc = ve_mag * ve_mag * P.E_ele_eV;
        if (!P.fixed_weight && E_inc < P.E_ion_eV) return false;                       // :900-905 (v3 only; ch4/v2 goes on and creates a NaN electron)
        double E_ej = 10.0 * tan(r.next() * atan((E_inc - P.E_ion_eV) / (2 * P.B_inc)));
        double E_sc = E_inc - P.E_ion_eV - E_ej;
        if (E_sc < 0) E_sc = 0.000001;
        double inv = 1.0 / ve_mag, u[3] = {ve[0] * inv, ve[1] * inv, ve[2] * inv};      // Vec3::unit
        new_velocity_electron(r, P, E_sc, u, ve);
        new_velocity_electron(r, P, E_ej, u, vnew);
        return true;
    }
    // elastic, isotropic in the centre of mass; only the electron changes (:933-947)
    double cm[3];
    for (int c = 0; c < 3; c++) cm[c] = (P.m_n * vn[c] + P.m_e * ve[c]) * (1.0 / P.sum_mass);
    double cos_ksi = 2 * r.next() - 1;
    double sin_ksi = sqrt(1 - cos_ksi * cos_ksi);
    double eps = 2 * 3.141592653 * r.next();
    g[0] = g_mag * cos_ksi; g[1] = g_mag * sin_ksi * cos(eps); g[2] = g_mag * sin_ksi * sin(eps);
    double f = P.m_n / P.sum_mass;
    for (int c = 0; c < 3; c++) ve[c] = cm[c] - f * g[c];
    return false;
}

// ---------------------------------------------------------------- appends: one atomic per warp iteration and store
#define MCC_THREADS 128
// A warp walks 32 consecutive cells in lockstep (lane = cell); iteration t handles candidate t of every cell that still has
// one.  The lanes whose candidate creates products take consecutive records of ONE atomicAdd on the staging cursor (a lane may want
// several).  All 32 lanes call.  Returns the first record of the lane, or -1 (nothing wanted, or the area is full - which the sizing
// by the candidate count rules out).
__device__ __forceinline__ long long warp_reserve(int want, int lane, u64* cursor, u64 cap) {
    int incl = want;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (!total) return -1;
    u64 base = 0;
    if (lane == 31) base = atomicAdd((unsigned long long*)cursor, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (!want) return -1;
    const u64 first = base + (u64)(incl - want);
    return first + (u64)want <= cap ? (long long)first : -1;
}
// Staging area shared by the three stores: 8 words per record - x y z u v w mpw and a tag word: cell | rank << 32 | store << 62, where
// rank numbers the records of one (cell, store) in order of creation.  A cell's records are made by one thread, one after the other, so
// the thread numbers them itself; made[k][c]: how many it made for store k (tables zeroed before the launch, written once per cell).
struct Stage { double* rec; unsigned* made[3]; u64* cursor; u64 cap; };
__device__ __forceinline__ void write_stage(const Stage& S, long long e, const double pos[3], const double v[3], double mpw, unsigned cell, int store, int& rank) {
    double4* r = reinterpret_cast<double4*>(S.rec + 8 * e);
    const u64 tag = (u64)cell | ((u64)(unsigned)rank++ << 32) | ((u64)store << 62);
    r[0] = make_double4(pos[0], pos[1], pos[2], v[0]); r[1] = make_double4(v[1], v[2], mpw, __longlong_as_double((long long)tag));
}
__device__ __forceinline__ void atomic_max_pos_double(double* addr, double v) {     // valid for non-negative doubles
    atomicMax((unsigned long long*)addr, (unsigned long long)__double_as_longlong(v));
}
// Species::addParticle of ch4/v2 (Species.cpp:226-237): bounds / object filter, then vel -= charge/mass * E(pos) * (0.5 * world.dt)
__device__ __forceinline__ bool add_particle_filter_rewind(const Grid& g, const double* __restrict__ ef, double q_over_m, double half_dt, const double pos[3], double v[3]) {
    if (!in_bounds(g, pos[0], pos[1], pos[2]) || in_object(g, pos[0], pos[1], pos[2])) return false;
    double ex, ey, ez;
    gather_ef(g, ef, x_to_l(pos[0], g.x0[0], g.inv_dx[0]), x_to_l(pos[1], g.x0[1], g.inv_dx[1]), x_to_l(pos[2], g.x0[2], g.inv_dx[2]), ex, ey, ez);
    v[0] = __dsub_rn(v[0], __dmul_rn(__dmul_rn(ex, q_over_m), half_dt));
    v[1] = __dsub_rn(v[1], __dmul_rn(__dmul_rn(ey, q_over_m), half_dt));
    v[2] = __dsub_rn(v[2], __dmul_rn(__dmul_rn(ez, q_over_m), half_dt));
    return true;
}

// Candidate pairs of cell c (v3 :646 / ch4/v2 :591).  Multi-GPU: the particles of a cell are spread over G ranks, so the
// cell's candidate count is estimated from the local populations (x G^2), rounded ONCE like the reference's and dealt out to the
// ranks: n/G each, the n%G left over to a rotating subset.  Rounding per rank instead would lose every cell whose share is below one half.
__device__ __forceinline__ int mcc_groups(const MccParams& P, int fixed, int c, int np_n, int np_e, double W_max, double dt, uint32_t call) {
    double frac = fixed ? 0.5 * np_n * np_e * P.neu_mpw0 * W_max * dt * P.inv_dv * P.rank_scale : np_n * np_e * W_max * dt * P.inv_dv * P.rank_scale;
    int n_groups = (int)(frac + 0.5);
    if (P.world > 1) n_groups = n_groups / P.world + ((unsigned)(c + (int)call + P.rank) % (unsigned)P.world < (unsigned)(n_groups % P.world) ? 1 : 0);
    if (n_groups > np_n) n_groups = np_n - 1;                                             // v3 :649-653 / v2 :598-600
    return n_groups < 0 ? 0 : n_groups;
}
// Sum of the candidate counts over the cells: the number of products of a call can not exceed it, which sizes the staging areas exactly
// (a collision is never dropped for lack of room, so the outcome of a call never depends on who reached a cursor first).
__global__ void __launch_bounds__(256) k_mcc_count(Grid g, MccParams P, CellLists Ln, CellLists Le, const double* __restrict__ wsv, double dt, uint32_t call, u64* __restrict__ total) {
    const double W_max = wsv[0];
    u64 sum = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < g.nc; c += gridDim.x * blockDim.x) {
        const int np_e = cell_view(Le, c).np;
        if (np_e <= 0) continue;
        const int np_n = cell_view(Ln, c).np;
        if (np_n > 0) sum += (u64)mcc_groups(P, P.fixed_weight, c, np_n, np_e, W_max, dt, call);
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd((unsigned long long*)total, (unsigned long long)sum);
}
// stats: [0] candidates [1] collisions [2] ionisations [3] skipped (electron heavier than neutral)
//        [5] dropped because a product store was full (collision skipped untouched)
//        [6] split-off neutrals beyond MCC_EXTRA per cell and call (created, but not selectable by later candidates of the same call)
//        [7] fixed-weight variant: created electrons with a NaN velocity (ionisation below the threshold, ch4/v2 only; appended as the reference does)
// Sn: the staging area of the products (stores 0 neutrals, 1 electrons, 2 ions)
template <int FIXED>
__global__ void __launch_bounds__(MCC_THREADS, 6) k_mcc(Grid g, MccParams P, Store neu, Store ele, Store ion, CellLists Ln, CellLists Le, Stage Sn,
                                                        double* __restrict__ wsv, u64* __restrict__ stats, double dt, uint64_t seed, uint32_t stream, uint32_t call) {
    const int lane = threadIdx.x & 31;
    const double W_max = wsv[0];
    u64 n_cand = 0, n_coll = 0, n_ion = 0, n_skip = 0, n_drop = 0, n_capped = 0, n_nan = 0; double step_max = 0;
    unsigned tot_made[3] = {0, 0, 0};                                                      // records staged per store by this thread
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int c0 = warp * 32; c0 < g.nc; c0 += nwarps * 32) {
        const int c = c0 + lane;
        CellView ve, vn; int np_e = 0, np_n0 = 0, n_groups = 0;
        if (c < g.nc) {
            ve = cell_view(Le, c); np_e = ve.np;
            if (np_e > 0) { vn = cell_view(Ln, c); np_n0 = vn.np; }
        }
        int np_n = np_n0;
        if (np_e > 0 && np_n0 > 0) n_groups = mcc_groups(P, FIXED, c, np_n, np_e, W_max, dt, call);
        const int max_groups = __reduce_max_sync(0xffffffffu, n_groups);
        if (max_groups == 0) continue;                                                    // warp-uniform
        PhiloxStream r; r.init(seed, stream, (u64)c, call);
        long long extra[MCC_EXTRA]; int n_extra = 0;
        int made_n = 0, made_e = 0, made_i = 0;                                            // records this cell has staged so far, per store
        for (int t = 0; t < max_groups; t++) {
            // kind of product this lane's candidate asks for: 0 none, 1 split-off neutral, 2 ion + electron
            int kind = 0; u64 pn = 0, pe = 0; bool staged = false;
            double vn_[3] = {0, 0, 0}, ve_[3] = {0, 0, 0}, vnew[3] = {0, 0, 0}, pos[3] = {0, 0, 0}, Wn = 0, We = 0, Wl = 0;
            bool ion_ok[2] = {false, false};                                               // FIXED: addParticle accepts the ion / the electron
            double vi[3] = {0, 0, 0};
            if (t < n_groups) {
                int a = (int)(r.next() * np_n);                                            // rnd(0,np) = 0 + rnd()*(np-0)
                int b = (int)(r.next() * np_e);
                staged = a >= np_n0;                                                       // a neutral split off earlier in this call: still in the staging area
                pn = staged ? (u64)extra[a - np_n0] : (u64)cell_pick(Ln, vn, a);
                pe = (u64)cell_pick(Le, ve, b);
                if (staged) { vn_[0] = Sn.rec[8 * pn + 3]; vn_[1] = Sn.rec[8 * pn + 4]; vn_[2] = Sn.rec[8 * pn + 5]; }
                else { vn_[0] = neu.a[3][pn]; vn_[1] = neu.a[4][pn]; vn_[2] = neu.a[5][pn]; }
                ve_[0] = ele.a[3][pe]; ve_[1] = ele.a[4][pe]; ve_[2] = ele.a[5][pe];
                double d[3] = {vn_[0] - ve_[0], vn_[1] - ve_[1], vn_[2] - ve_[2]};
                double v_rel = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                double E_rel = P.E_rel_eV * v_rel * v_rel;
                double s_coll = sigma_coll(P, E_rel);
                n_cand++;
                if (FIXED) {                                                               // ch4/v2 :608-632
                    double sv = s_coll * v_rel;
                    if (sv > step_max) step_max = sv;
                    if (sv / W_max > r.next()) {
                        n_coll++;
                        bool ionised = collide(r, P, vn_, ve_, vnew, s_coll);
                        ele.a[3][pe] = ve_[0]; ele.a[4][pe] = ve_[1]; ele.a[5][pe] = ve_[2];   // collide works on a reference to the electron's velocity
                        if (ionised) {
                            n_ion++; kind = 2;
                            pos[0] = neu.a[0][pn]; pos[1] = neu.a[1][pn]; pos[2] = neu.a[2][pn];          // (the fixed-weight variant never splits: pn is a store slot)
                            vi[0] = vn_[0]; vi[1] = vn_[1]; vi[2] = vn_[2];
                            ion_ok[0] = add_particle_filter_rewind(g, P.ef, P.qm_ion, P.half_dt, pos, vi);      // ions.addParticle(pos, vel_neutral, ions.mpw0) :624-626
                            // electrons.addParticle(pos, vel_new, electrons.mpw0) :628.  Below the ionisation threshold ch4/v2 computes the
                            // ejected electron from a negative energy: its velocity is NaN and the reference appends it all the same (addParticle
                            // only tests the position); between the electrodes it dies at the next electron push, where Rectangle::inObject
                            // holds for a NaN position (ch4/v2/Species.cpp:193-207) - here as there.  Reproduced, and counted.
                            if (isnan(vnew[0]) || isnan(vnew[1]) || isnan(vnew[2])) n_nan++;
                            ion_ok[1] = add_particle_filter_rewind(g, P.ef, P.qm_ele, P.half_dt, pos, vnew);
                        }
                    }
                } else {
                    Wn = staged ? Sn.rec[8 * pn + 6] : neu.a[6][pn]; We = ele.a[6][pe];
                    double Wg = Wn < We ? We : Wn; Wl = Wn < We ? Wn : We;                  // greaterLesser funkc.h:19-25
                    double Wsv = Wg * s_coll * v_rel;
                    if (Wsv > step_max) step_max = Wsv;
                    if (r.next() < Wsv / W_max) {
                        n_coll++;
                        if (Wn > We) {                                                     // split the neutral (:684-703)
                            bool ionised = collide(r, P, vn_, ve_, vnew, s_coll);          // pure: works on local copies
                            kind = ionised ? 2 : 1;
                            if (staged) { pos[0] = Sn.rec[8 * pn]; pos[1] = Sn.rec[8 * pn + 1]; pos[2] = Sn.rec[8 * pn + 2]; }
                            else { pos[0] = neu.a[0][pn]; pos[1] = neu.a[1][pn]; pos[2] = neu.a[2][pn]; }
                        } else if (Wn < We) {
                            n_skip++;      // the reference's electron-heavier branch is defective; not reproduced, counted
                        }                  // equal weights: accepted pair does nothing (:727-734)
                    }
                }
            }
            // ---- the warp's products of this iteration take their records (every lane is here: the loop bounds are warp-uniform)
            if (FIXED) {                                                                   // :623-629: ions_to_create ions, one electron
                const int n_i = (kind == 2 && ion_ok[0]) ? P.ions_to_create : 0, n_e = (kind == 2 && ion_ok[1]) ? 1 : 0;
                long long e0 = warp_reserve(n_i + n_e, lane, Sn.cursor, Sn.cap);
                if (n_i + n_e) {
                    if (e0 < 0) n_drop++;
                    else {
                        for (int q = 0; q < n_i; q++) write_stage(Sn, e0++, pos, vi, P.ion_mpw0, (unsigned)c, 2, made_i);
                        if (n_e) write_stage(Sn, e0, pos, vnew, P.ele_mpw0, (unsigned)c, 1, made_e);
                    }
                }
                continue;
            }
            // Record reservation BEFORE any state is changed: a collision whose products do not fit is skipped as a whole (and counted)
            long long e0 = warp_reserve(kind == 1 ? 1 : kind == 2 ? 2 : 0, lane, Sn.cursor, Sn.cap);
            if (kind > 0 && e0 < 0) { kind = -1; n_drop++; n_coll--; }
            if (kind > 0) {
                if (staged) Sn.rec[8 * pn + 6] = Wn - We; else neu.a[6][pn] = Wn - We;
                ele.a[3][pe] = ve_[0]; ele.a[4][pe] = ve_[1]; ele.a[5][pe] = ve_[2];
                if (kind == 2) {
                    n_ion++;
                    write_stage(Sn, e0, pos, vn_, Wl, (unsigned)c, 2, made_i);                 // no half-step rewind (:694-695)
                    write_stage(Sn, e0 + 1, pos, vnew, Wl, (unsigned)c, 1, made_e);
                } else {
                    write_stage(Sn, e0, pos, vn_, We, (unsigned)c, 0, made_n);                 // split-off neutral of the electron's weight
                    if (n_extra < MCC_EXTRA) { extra[n_extra++] = e0; np_n++; } else n_capped++;
                }
            }
        }
        if (made_n) { Sn.made[0][c] = (unsigned)made_n; tot_made[0] += made_n; }
        if (made_e) { Sn.made[1][c] = (unsigned)made_e; tot_made[1] += made_e; }
        if (made_i) { Sn.made[2][c] = (unsigned)made_i; tot_made[2] += made_i; }
    }
    // block-level reduction of the statistics
    __shared__ u64 sh[10]; __shared__ double sh_max;
    if (threadIdx.x == 0) { for (int q = 0; q < 10; q++) sh[q] = 0; sh_max = 0; }
    __syncthreads();
    if (n_cand) {
        atomicAdd(&sh[0], n_cand); atomicAdd(&sh[1], n_coll); atomicAdd(&sh[2], n_ion); atomicAdd(&sh[3], n_skip); atomicAdd(&sh[4], n_drop);
        atomicAdd(&sh[5], n_capped); atomicAdd(&sh[6], n_nan); atomic_max_pos_double(&sh_max, step_max);
        for (int k = 0; k < 3; k++) if (tot_made[k]) atomicAdd(&sh[7 + k], (u64)tot_made[k]);
    }
    __syncthreads();
    if (threadIdx.x == 0 && sh[0]) {
        atomicAdd(&stats[0], sh[0]); atomicAdd(&stats[1], sh[1]); atomicAdd(&stats[2], sh[2]); atomicAdd(&stats[3], sh[3]); atomicAdd(&stats[5], sh[4]);
        atomicAdd(&stats[6], sh[5]); atomicAdd(&stats[7], sh[6]);
        for (int k = 0; k < 3; k++) if (sh[7 + k]) atomicAdd(&stats[12 + k], sh[7 + k]);             // records of store k (0 neutrals, 1 electrons, 2 ions)
        atomic_max_pos_double(&wsv[1], sh_max);
    }
}
// After the kernel.  (1) W_sigma_v_rel_max <- max sampled value of this step, only if a collision happened (:751-756).  (2) A cursor
// that ran past the staging area (the reservations beyond it failed) goes back to the capacity.  stats[8]: records staged.
__global__ void k_mcc_finish(double* wsv, u64* stats, u64 cap) {
    if (stats[1]) wsv[0] = wsv[1];
    if (stats[8] > cap) stats[8] = cap;
}
// Staging index of the r-th record of every store in (cell, creation) order: start[k][] = exclusive scan of made[k][] (records of the
// cells below).  One pass over the tag words serves the three stores.
struct StageIndex { const unsigned* start[3]; unsigned* sorted[3]; };
__global__ void __launch_bounds__(256) k_stage_index(const u64* __restrict__ n_staged, const double* __restrict__ rec, StageIndex X) {
    const u64 n = *n_staged;
    for (u64 e = blockIdx.x * (u64)blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x) {
        const u64 tag = (u64)__double_as_longlong(rec[8 * e + 7]);
        const unsigned cell = (unsigned)(tag & 0xffffffffu), rank = (unsigned)((tag >> 32) & 0x3fffffffu); const int k = (int)(tag >> 62);
        X.sorted[k][X.start[k][cell] + rank] = (unsigned)e;
    }
}
// The staged records of one store in (cell, creation) order -> the end of the store.  sorted[r]: staging index of the r-th record,
// *n_keep: number of records of the store.  One thread per record: a 64-byte record in, seven coalesced 8-byte columns out.
__global__ void __launch_bounds__(256) k_stage_commit(Store st, const double* __restrict__ rec, const unsigned* __restrict__ sorted, const unsigned* __restrict__ n_keep) {
    const u64 n0 = st.ctr->n, m = *n_keep;
    for (u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x; r < m; r += (u64)gridDim.x * blockDim.x) {
        const double4* src = reinterpret_cast<const double4*>(rec + 8 * (u64)sorted[r]);
        const double4 lo = src[0], hi = src[1];
        const u64 d = n0 + r;
        st.a[0][d] = lo.x; st.a[1][d] = lo.y; st.a[2][d] = lo.z; st.a[3][d] = lo.w; st.a[4][d] = hi.x; st.a[5][d] = hi.y; st.a[6][d] = hi.z;
    }
}
__global__ void k_stage_count(SpeciesCounters* ctr, const unsigned* __restrict__ n_keep) { ctr->n += *n_keep; }
__global__ void k_sigma_eval(MccParams P, int n, const double* __restrict__ E, double* __restrict__ sc, double* __restrict__ si) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) { sc[t] = sigma_coll(P, E[t]); si[t] = sigma_ion(P, E[t]); }
}

static MccParams make_params(const picg_mcc_s* m) {
    MccParams P;
    P.m_n = m->neu->mass; P.m_e = m->ele->mass; P.sum_mass = P.m_n + P.m_e;
    double m_r = P.m_n * P.m_e / (P.m_n + P.m_e);                                         // :503
    P.E_rel_eV = 0.5 * m_r / 1.602176565e-19;                                             // :509
    P.E_ele_eV = 9.10938215e-31 * 0.5 / 1.602176565e-19;                                  // Interactions.h:123 (Const::m_e)
    P.two_qe_me = 2 * 1.602176565e-19 / 9.10938215e-31;                                   // Interactions.h:124
    P.c0 = 1.015e-18; P.c1 = 9.793e+00; P.c2 = 6.181e+01; P.B_inc = 10.0;                 // :513-518
    P.E_ion_eV = m->E_ion_J / 1.602176565e-19;                                            // :499
    const Grid& g = m->w->g;
    P.inv_dv = 1 / (g.dx[0] * g.dx[1] * g.dx[2]);                                         // :500-501
    P.rank_scale = (double)g_world_size * (double)g_world_size; P.world = g_world_size; P.rank = g_rank;
    P.n_tab = m->n_table; P.tab_E = m->tab_E; P.tab_s = m->tab_s;
    P.fixed_weight = m->fixed_weight; P.neu_mpw0 = m->neu->mpw0; P.ion_mpw0 = m->ion->mpw0; P.ele_mpw0 = m->ele->mpw0;
    P.ions_to_create = (int)(m->ele->mpw0 / m->ion->mpw0 + 0.5);                          // ch4/v2 :623
    P.ef = m->w->ef; P.qm_ion = m->ion->charge / m->ion->mass; P.qm_ele = m->ele->charge / m->ele->mass; P.half_dt = 0.5 * m->w->dt;
    return P;
}
// per-cell list lengths as the collision kernel sees them (debug / tests): must equal computeMacroParticlesCount
__global__ void k_list_counts(Grid g, CellLists L, double* __restrict__ out) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < g.nc; c += gridDim.x * blockDim.x) out[c] = (double)cell_view(L, c).np;
}

extern "C" {

int picg_mcc_create(picg_species_t neutrals, picg_species_t ions, picg_species_t electrons, picg_world_t w, const double* table_E,
                    const double* table_sigma, int n_table, double E_ion_J, picg_mcc_t* out) {
    REQUIRE_DEVICE(); REQUIRE_ARG(neutrals && ions && electrons && w && out && table_E && table_sigma, "picg_mcc_create: null argument");
    // constructor preconditions of the reference (:479-489) -> std::invalid_argument in the facade
    REQUIRE_ARG(!(neutrals->mpw0 < 1e2 * electrons->mpw0), "neutrals species should have mpw0 at least 100 times greater than electrons");
    REQUIRE_ARG(!(!E_ion_J || E_ion_J < 0), "neutral species must have proper ionization energy E_ion (in Joules)");
    REQUIRE_ARG(!(neutrals->mpw0 < ions->mpw0), "neutrals.mpw0 must be greater or equal ions.mpw0");
    REQUIRE_ARG(n_table >= 1, "picg_mcc_create: empty cross-section table (the reference throws when the file cannot be opened, :521)");
    picg_mcc_s* m = new picg_mcc_s();
    m->neu = neutrals; m->ion = ions; m->ele = electrons; m->w = w; m->E_ion_J = E_ion_J;
    // std::map semantics: sorted by energy, a repeated key keeps its last value (:530)
    std::vector<std::pair<double, double>> tab;
    for (int i = 0; i < n_table; i++) {
        bool dup = false;
        for (auto& t : tab) if (t.first == table_E[i]) { t.second = table_sigma[i]; dup = true; }
        if (!dup) tab.push_back({table_E[i], table_sigma[i]});
    }
    std::sort(tab.begin(), tab.end());
    m->n_table = (int)tab.size();
    std::vector<double> E(tab.size()), S(tab.size());
    for (size_t i = 0; i < tab.size(); i++) { E[i] = tab[i].first; S[i] = tab[i].second; }
    cudaError_t e;
    if ((e = cudaMalloc(&m->tab_E, E.size() * 8)) != cudaSuccess || (e = cudaMalloc(&m->tab_s, S.size() * 8)) != cudaSuccess ||
        (e = cudaMalloc(&m->wsv, 2 * 8)) != cudaSuccess || (e = cudaMalloc(&m->stats, 16 * 8)) != cudaSuccess) {
        picg_mcc_destroy(m); return cuda_fail(e, "cudaMalloc(mcc)", __FILE__, __LINE__);
    }
    CUDA_TRY(cudaMemcpyAsync(m->tab_E, E.data(), E.size() * 8, cudaMemcpyHostToDevice, g_stream));
    CUDA_TRY(cudaMemcpyAsync(m->tab_s, S.data(), S.size() * 8, cudaMemcpyHostToDevice, g_stream));
    double wsv[2] = {1e-14 * std::max(electrons->mpw0, neutrals->mpw0), 0.0};             // Interactions.h:129, .cpp:534
    CUDA_TRY(cudaMemcpyAsync(m->wsv, wsv, 16, cudaMemcpyHostToDevice, g_stream));
    CUDA_TRY(cudaMemsetAsync(m->stats, 0, 128, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    { int rc = species_prepare_lists(neutrals); if (rc) return rc; rc = species_prepare_lists(electrons); if (rc) return rc; }
    *out = m;
    return PICG_OK;
}

int picg_mcc_destroy(picg_mcc_t m) {
    if (!m) return PICG_OK;
    if (g_stream) cudaStreamSynchronize(g_stream);
    cudaFree(m->tab_E); cudaFree(m->tab_s); cudaFree(m->wsv); cudaFree(m->stats);
    delete m; return PICG_OK;
}

int picg_mcc_set_wsv_max(picg_mcc_t m, double v) {
    REQUIRE_DEVICE(); REQUIRE_ARG(m && v > 0, "picg_mcc_set_wsv_max: bad argument");
    CUDA_TRY(cudaMemcpyAsync(m->wsv, &v, 8, cudaMemcpyHostToDevice, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return PICG_OK;
}

// variant 0: variable weights (ch4/v3, the default); 1: fixed weights (ch4/v2/Interactions.cpp:566-641).  Resets the acceptance
// ceiling to the variant's initial value (v3: 1e-14 * max(mpw0), Interactions.cpp:534; v2: 1e-14, ch4/v2/Interactions.h:129).
int picg_mcc_set_variant(picg_mcc_t m, int variant) {
    REQUIRE_DEVICE(); REQUIRE_ARG(m && (variant == 0 || variant == 1), "picg_mcc_set_variant: variant must be 0 (ch4/v3, variable weights) or 1 (ch4/v2, fixed weights)");
    m->fixed_weight = variant;
    return picg_mcc_set_wsv_max(m, variant ? 1e-14 : 1e-14 * std::max(m->ele->mpw0, m->neu->mpw0));
}

int picg_mcc_sigma(picg_mcc_t m, int n, const double* E_eV, double* sc, double* si) {
    REQUIRE_DEVICE(); REQUIRE_ARG(m && E_eV && sc && si && n >= 0, "picg_mcc_sigma: bad argument");
    if (n == 0) return PICG_OK;
    int rc = ensure_scratch(m->w, (size_t)n * 24 + 64); if (rc) return rc;
    double* d = (double*)m->w->scratch;
    CUDA_TRY(cudaMemcpyAsync(d, E_eV, (size_t)n * 8, cudaMemcpyHostToDevice, g_stream));
    LAUNCH(K_MISC, k_sigma_eval, std::min(div_up(n, 256), 1024), 256, 0, make_params(m), n, d, d + n, d + 2 * (size_t)n); CHECK_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(sc, d + n, (size_t)n * 8, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaMemcpyAsync(si, d + 2 * (size_t)n, (size_t)n * 8, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return PICG_OK;
}

int picg_mcc_apply(picg_mcc_t m, double dt, picg_mcc_stats* out) {
    REQUIRE_DEVICE(); REQUIRE_ARG(m, "picg_mcc_apply: null handle");
    picg_species_s *neu = m->neu, *ele = m->ele, *ion = m->ion;
    int rc;
    // both collision partners must be exactly cell-sorted, as in the reference (:602-613)
    rc = species_exact_lists(neu); if (rc) return rc;          // no-op when sorted; mover pass on a stale partition; full sort otherwise
    rc = species_exact_lists(ele); if (rc) return rc;
    MccParams P = make_params(m);
    const Grid& g = m->w->g;
    // Room for the products: a call makes at most one product per store and candidate pair (ch4/v2: ions_to_create ions), and the number
    // of candidate pairs follows from the per-cell lists alone: counted first, so that the staging areas can never run full.
    CUDA_TRY(cudaMemsetAsync(m->stats, 0, 128, g_stream));
    m->step++;
    LAUNCH(K_MCC_APPEND, k_mcc_count, std::max(1, std::min(div_up(g.nc, 256), g_sm_count * 8)), 256, 0, g, P, lists_of(neu), lists_of(ele), (const double*)m->wsv, dt, (uint32_t)m->step, m->stats + 11);
    CHECK_LAUNCH();
    u64 n_pairs = 0;
    CUDA_TRY(cudaMemcpyAsync(&n_pairs, m->stats + 11, 8, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    // exact counts of the collision partners are known from the list builds above; the ion store only receives: its upper bound will do
    picg_species_s* sp3[3] = {neu, ele, ion};
    size_t n_before[3];
    for (int k = 0; k < 3; k++) n_before[k] = sp3[k]->n_host_valid ? sp3[k]->n_host : sp3[k]->n_upper;
    // a pair makes one split-off neutral, or one ion and one electron (ch4/v2: ions_to_create ions and one electron)
    const size_t per_pair = m->fixed_weight ? (size_t)std::max(P.ions_to_create, 1) + 1 : 2;
    const size_t scap = std::max<size_t>((size_t)n_pairs * per_pair, 64);
    double zero = 0; CUDA_TRY(cudaMemcpyAsync(m->wsv + 1, &zero, 8, cudaMemcpyHostToDevice, g_stream));
    // staging area in the scratch arena (free here: the list builds above are done with it): records | per-cell tables of the three stores |
    // index lists of the commits | work area of the scans
    const size_t nt = (size_t)g.nc + 1;
    size_t made_off[3], sorted_off[3], off = (scap * 64 + 255) & ~(size_t)255;
    for (int k = 0; k < 3; k++) { made_off[k] = off; off += ((nt * 4 + 255) & ~(size_t)255); }
    for (int k = 0; k < 3; k++) { sorted_off[k] = off; off += ((std::min(scap, (size_t)n_pairs * (k == 2 ? per_pair - 1 : 1) + 64) * 4 + 255) & ~(size_t)255); }
    const size_t work_off = off; off += counting_sort_words(g) * 4 + 256;
    rc = ensure_scratch(m->w, off); if (rc) return rc;
    char* base = (char*)m->w->scratch;
    Stage S; S.rec = (double*)base; S.cursor = m->stats + 8; S.cap = scap;
    for (int k = 0; k < 3; k++) { S.made[k] = (unsigned*)(base + made_off[k]); CUDA_TRY(cudaMemsetAsync(S.made[k], 0, nt * 4, g_stream)); }
    int grid = std::max(1, std::min(div_up(g.nc, MCC_THREADS), g_sm_count * 16));
    if (m->fixed_weight) LAUNCH(K_MCC, k_mcc<1>, grid, MCC_THREADS, 0, g, P, store_of(neu), store_of(ele), store_of(ion), lists_of(neu), lists_of(ele), S,
                                m->wsv, m->stats, dt, g_seed, rng_stream_id(RNG_MCC, neu->id, g_rank), (uint32_t)m->step);
    else LAUNCH(K_MCC, k_mcc<0>, grid, MCC_THREADS, 0, g, P, store_of(neu), store_of(ele), store_of(ion), lists_of(neu), lists_of(ele), S,
                m->wsv, m->stats, dt, g_seed, rng_stream_id(RNG_MCC, neu->id, g_rank), (uint32_t)m->step);
    CHECK_LAUNCH();
    LAUNCH(K_MCC_APPEND, k_mcc_finish, 1, 1, 0, m->wsv, m->stats, (u64)scap); CHECK_LAUNCH();
    u64 host_stats[16]; double host_wmax = 0;
    CUDA_TRY(cudaMemcpyAsync(host_stats, m->stats, 128, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaMemcpyAsync(&host_wmax, m->wsv, 8, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));                  // host_stats is valid from here on
    // the staged products, store by store, in (cell, creation) order behind the store's particles
    bool grown[3] = {false, false, false};
    size_t kept_all[3] = {0, 0, 0};
    const size_t staged = (size_t)host_stats[8];
    if (staged) {
        StageIndex X;
        size_t kept[3];
        for (int k = 0; k < 3; k++) {                           // made[k][] -> records in the cells below; made[k][nc] = records of store k
            kept[k] = kept_all[k] = (size_t)host_stats[12 + k];
            X.start[k] = S.made[k]; X.sorted[k] = (unsigned*)(base + sorted_off[k]);
            if (kept[k]) { rc = scan_cell_table(g, S.made[k], (unsigned*)(base + work_off)); if (rc) return rc; }
        }
        const int egrid = std::max(1, std::min(div_up(staged, 256), g_sm_count * 8));
        LAUNCH(K_MCC_APPEND, k_stage_index, egrid, 256, 0, (const u64*)S.cursor, (const double*)S.rec, X); CHECK_LAUNCH();
        for (int k = 0; k < 3; k++) {
            if (!kept[k]) continue;
            picg_species_s* sp = sp3[k];
            if (sp->cap < n_before[k] + kept[k]) {              // grow by what is needed plus headroom; the arena holds the staged records and is re-sized after the commits
                if (!sp->n_host_valid) { rc = species_refresh_count(sp); if (rc) return rc; n_before[k] = sp->n_host; }      // growing copies n_host particles: it must be exact
                rc = species_grow_store(sp, n_before[k] + kept[k] + std::max<size_t>(2 * kept[k], n_before[k] / 100)); if (rc) return rc;
                grown[k] = true;
            }
            const int cgrid = std::max(1, std::min(div_up(kept[k], 256), g_sm_count * 8));
            LAUNCH(K_MCC_APPEND, k_stage_commit, cgrid, 256, 0, store_of(sp), (const double*)S.rec, (const unsigned*)X.sorted[k], (const unsigned*)(S.made[k] + g.nc)); CHECK_LAUNCH();
            LAUNCH(K_MCC_APPEND, k_stage_count, 1, 1, 0, sp->ctr, (const unsigned*)(S.made[k] + g.nc)); CHECK_LAUNCH();
            // the store grew by exactly kept[k]: the host-side count follows without a read-back
            if (sp->n_host_valid) { sp->n_host += kept[k]; sp->n_upper = sp->n_host; } else sp->n_upper = std::min(sp->cap, sp->n_upper + kept[k]);
        }
    }
    for (int k = 0; k < 3; k++) if (grown[k]) { rc = species_scratch_for_store(sp3[k]); if (rc) return rc; }
    for (int k = 0; k < 3; k++) m->last_appends[k] = staged ? kept_all[k] : 0;
    if (host_stats[1]) { neu->sorted_valid = false; neu->lists_valid = false; neu->count_valid = false; ele->sorted_valid = false; ele->lists_valid = false; ele->count_valid = false; ion->sorted_valid = false; ion->lists_valid = false; ion->count_valid = false; }   // :751-754
    if (host_stats[5]) {                                        // grow so that the next call has room, and tell the caller
        for (int k = 0; k < 3; k++) { rc = species_ensure_capacity(sp3[k], sp3[k]->n_host + std::max<size_t>(4 * (size_t)host_stats[5], sp3[k]->n_host / 10)); if (rc) return rc; }
        set_error(PICG_OK, "picg_mcc_apply: %llu collisions were skipped because a particle store was full; stores were grown (reserve more up front)", (unsigned long long)host_stats[5]);
    }
    if (out) {
        out->candidates = host_stats[0]; out->collisions = host_stats[1]; out->ionizations = host_stats[2]; out->dropped = host_stats[5];
        out->extras_capped = host_stats[6]; out->nan_products = host_stats[7];
        out->w_sigma_v_max = host_wmax;
    }
    return PICG_OK;
}

// debug / tests: lengths of the exact per-cell lists of `which` (0 neutrals, 1 electrons), cells in Field order
int picg_mcc_list_counts(picg_mcc_t m, int which, double* host) {
    REQUIRE_DEVICE(); REQUIRE_ARG(m && host && (which == 0 || which == 1), "picg_mcc_list_counts: bad argument");
    picg_species_s* s = which ? m->ele : m->neu;
    int rc = species_exact_lists(s); if (rc) return rc;
    const Grid& g = m->w->g;
    rc = ensure_scratch(m->w, (size_t)g.nc * 8 + 64); if (rc) return rc;
    // the mover arrays live in species-owned buffers, not in the scratch arena, so the arena is free here
    LAUNCH(K_MISC, k_list_counts, std::min(div_up(g.nc, 256), 2048), 256, 0, g, lists_of(s), (double*)m->w->scratch); CHECK_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(host, m->w->scratch, (size_t)g.nc * 8, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return PICG_OK;
}

}  // extern "C"
