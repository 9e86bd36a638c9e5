// fp32 secondary path of the particle loop (north star: "1e-4 relative for fp32", "coalesced float4 loads"; SURVEY 8d "fp32 secondary";
// the reference is written against one scalar type, ch4/v3/src/all.h:11 `using type_calc = double`, and rebuilds with float).
//
// A single-precision ABSOLUTE position cannot carry the motion of the slow species (a neutral moves 4e-10 m per step in a 2.5 cm
// domain: below the fp32 spacing of 1.9e-9 m), so the fp32 store keeps CELL-RELATIVE positions: the cell index (u32, k fastest - it is
// also the sort key, so a sort needs no key pass) and the fractional coordinates fx, fy, fz in [0, 1) as floats (spacing 6e-8 of a cell
// = 6e-12 m).  The fractional coordinates ARE the trilinear weights of gather and scatter (Field.h:157-232: di = lc - (int)lc), so
// no subtraction of large numbers is left anywhere.  32 bytes per particle instead of 56:
//     fx fy fz u v w mpw : float[cap]       cell : u32[cap]
//   k_push32     Species::advanceElectronsSerial (Species.cpp:356-399) in fp32: gather E, kick, drift in cell units with carry
//                into the cell index, absorb outside the box / inside an object.         56 B per particle (28 R + 28 W)
//   k_deposit32  Species::computeNumberDensity (:401-416) + computeMacroParticlesCount (:813-819): fp32 weights, the eight contributions
//                quantised to the same int64 fixed-point grid as the fp64 path (deterministic).   20 B per particle
//   sort         radix sort on the stored cell index + permutation of the eight arrays.
// Node fields (E, phi, rho, node volumes, the density outputs) stay fp64 grids: they are 0.1 % of the memory traffic and phi
// needs the range.  The heavy species' wall interaction and the collision kernels have no fp32 variant (fp64 path only).
#include "common.cuh"
#include "push.cuh"
#include "deposit.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

using namespace picg;

struct picg_species32_s {
    picg_world_s* w;
    double mass, charge, mpw0;
    size_t cap = 0, n_host = 0; bool n_host_valid = true; size_t n_upper = 0;
    float* f[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // fx fy fz u v w mpw
    unsigned* cell = nullptr;
    float* fspare = nullptr; unsigned* cspare = nullptr;                               // out-of-place targets of the sort permutation
    SpeciesCounters* ctr = nullptr; SpeciesCounters* ctr_host = nullptr;
    i64* den_fixed = nullptr; double* den = nullptr; double* macro_count = nullptr;
    unsigned* cell_start = nullptr; bool sorted_valid = false;
    int S = 0; bool S_set = false;
};

__global__ void k_compact_finish(SpeciesCounters* ctr);
__global__ void k_finalize_den(int u_begin, int u_end, const i64* __restrict__ fixed, const double* __restrict__ vol, double inv_scale, double* __restrict__ den, SpeciesCounters* ctr);
__global__ void k_reset_den_stats(SpeciesCounters* ctr);
__global__ void k_cell_start(const u64* __restrict__ n_ptr, const unsigned* __restrict__ keys, int nc, unsigned* __restrict__ cell_start);
__global__ void k_iota_u32(const u64* __restrict__ n_ptr, unsigned* __restrict__ out);
namespace picg {
int radix_sort_pairs(const u64* n_ptr, size_t n_upper, int key_bits, unsigned*& keysA, unsigned*& valsA, unsigned*& keysB, unsigned*& valsB, unsigned* counts, int nblocks);   // sort.cu
}

struct Arr32 { float* f[7]; unsigned* cell; };
static Arr32 arr_of(picg_species32_s* s) { Arr32 a; for (int c = 0; c < 7; c++) a.f[c] = s->f[c]; a.cell = s->cell; return a; }

// ---------------------------------------------------------------- conversions (host AoS of doubles <-> cell-relative fp32 SoA)
__device__ __forceinline__ void to_cell_frac(double l, int cells, int& c, float& fr) {
    c = min(max((int)l, 0), cells - 1);
    fr = (float)(l - (double)c);
    if (fr >= 1.0f) fr = 0.99999994f;                       // rounding up to the next cell's origin: stay in the cell
    if (fr < 0.0f) fr = 0.0f;
}
__global__ void __launch_bounds__(256) k_aos_to_f32(Grid g, size_t n, const double* __restrict__ aos, Arr32 a, size_t base) {
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const double* q = aos + p * 7;
        int ci, cj, ck; float fx, fy, fz;
        to_cell_frac(x_to_l(q[0], g.x0[0], g.inv_dx[0]), g.ci, ci, fx); to_cell_frac(x_to_l(q[1], g.x0[1], g.inv_dx[1]), g.cj, cj, fy);
        to_cell_frac(x_to_l(q[2], g.x0[2], g.inv_dx[2]), g.ck, ck, fz);
        const size_t d = base + p;
        a.f[0][d] = fx; a.f[1][d] = fy; a.f[2][d] = fz; a.f[3][d] = (float)q[3]; a.f[4][d] = (float)q[4]; a.f[5][d] = (float)q[5]; a.f[6][d] = (float)q[6];
        a.cell[d] = (unsigned)cell_of(g, ci, cj, ck);
    }
}
__global__ void __launch_bounds__(256) k_soa64_to_f32(Grid g, const SpeciesCounters* ctr, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                                                      const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ w, const double* __restrict__ m, Arr32 a) {
    const u64 n = ctr->n;
    for (u64 p = blockIdx.x * (u64)blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x) {
        int ci, cj, ck; float fx, fy, fz;
        to_cell_frac(x_to_l(x[p], g.x0[0], g.inv_dx[0]), g.ci, ci, fx); to_cell_frac(x_to_l(y[p], g.x0[1], g.inv_dx[1]), g.cj, cj, fy);
        to_cell_frac(x_to_l(z[p], g.x0[2], g.inv_dx[2]), g.ck, ck, fz);
        a.f[0][p] = fx; a.f[1][p] = fy; a.f[2][p] = fz; a.f[3][p] = (float)u[p]; a.f[4][p] = (float)v[p]; a.f[5][p] = (float)w[p]; a.f[6][p] = (float)m[p];
        a.cell[p] = (unsigned)cell_of(g, ci, cj, ck);
    }
}
__global__ void __launch_bounds__(256) k_f32_to_aos(Grid g, size_t n, Arr32 a, size_t base, double* __restrict__ aos) {
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const size_t s = base + p;
        int ci, cj, ck; cell_to_ijk(g, (int)a.cell[s], ci, cj, ck);
        double* q = aos + p * 7;
        q[0] = g.x0[0] + ((double)ci + (double)a.f[0][s]) * g.dx[0]; q[1] = g.x0[1] + ((double)cj + (double)a.f[1][s]) * g.dx[1];
        q[2] = g.x0[2] + ((double)ck + (double)a.f[2][s]) * g.dx[2];
        q[3] = a.f[3][s]; q[4] = a.f[4][s]; q[5] = a.f[5][s]; q[6] = a.f[6][s];
    }
}

// ---------------------------------------------------------------- push
struct Push32Args { Arr32 a; SpeciesCounters* ctr; const double* ef; float qm_dt, dt_idx, dt_idy, dt_idz; unsigned* dead_list; };

__device__ __forceinline__ void ld4(const float* p, float v[4]) { float4 t = __ldcs(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
__device__ __forceinline__ void ld4(const unsigned* p, unsigned v[4]) { uint4 t = __ldcs(reinterpret_cast<const uint4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
__device__ __forceinline__ void st4(float* p, const float v[4]) { __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3])); }
__device__ __forceinline__ void st4(unsigned* p, const unsigned v[4]) { __stcs(reinterpret_cast<uint4*>(p), make_uint4(v[0], v[1], v[2], v[3])); }

// Field<Vec3>::gather (Field.h:201-232) in fp32: the eight terms (F*w2)*w1 summed left to right, weights from the stored fractions
__device__ __forceinline__ void gather_ef32(const Grid& g, const double* __restrict__ ef, int i, int j, int k, float di, float dj, float dk, float& ex, float& ey, float& ez) {
    const float odi = __fsub_rn(1.0f, di), odj = __fsub_rn(1.0f, dj), odk = __fsub_rn(1.0f, dk);
    const float wa = __fmul_rn(odi, odj), wb = __fmul_rn(odi, dj), wc = __fmul_rn(di, odj), wd = __fmul_rn(di, dj);
    const size_t r00 = ((size_t)(i * g.nj + j) * g.nk + k) * 3, r01 = r00 + (size_t)g.nk * 3, r10 = r00 + (size_t)g.nj * g.nk * 3, r11 = r10 + (size_t)g.nk * 3;
    float acc[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float v;
        v = __fmul_rn(__fmul_rn((float)__ldg(ef + r00 + c), wa), odk);
        v = __fadd_rn(v, __fmul_rn(__fmul_rn((float)__ldg(ef + r00 + 3 + c), wa), dk));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn((float)__ldg(ef + r01 + c), wb), odk));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn((float)__ldg(ef + r01 + 3 + c), wb), dk));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn((float)__ldg(ef + r10 + c), wc), odk));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn((float)__ldg(ef + r10 + 3 + c), wc), dk));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn((float)__ldg(ef + r11 + c), wd), odk));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn((float)__ldg(ef + r11 + 3 + c), wd), dk));
        acc[c] = v;
    }
    ex = acc[0]; ey = acc[1]; ez = acc[2];
}
// drift of one coordinate in cell units: f += v * dt / dx, whole cells carried into the index; false when the particle leaves the box
__device__ __forceinline__ bool drift32(float& fr, int& c, float v, float dt_id, int cells) {
    float t = __fadd_rn(fr, __fmul_rn(v, dt_id));
    const float fl = floorf(t);
    c += (int)fl;                                            // |v dt / dx| stays far below 2^31 for any particle worth keeping
    t = __fsub_rn(t, fl);
    if (t >= 1.0f) { t = 0.0f; c += 1; }                     // t - floor(t) rounded up to 1 (t slightly below an integer)
    fr = t;
    return c >= 0 && c < cells;                              // World::inBounds (World.cpp:201-205): x0 <= p < xm
}

__global__ void __launch_bounds__(256, 2) k_push32(Grid g, Push32Args A) {
    const u64 n = A.ctr->n;
    const int lane = threadIdx.x & 31;
    const u64 nthreads = (u64)gridDim.x * blockDim.x;
    for (u64 p0 = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * 4; p0 - (u64)lane * 4 < n; p0 += nthreads * 4) {      // warp-uniform trip count
        const bool full = p0 + 4 <= n;
        float fx[4], fy[4], fz[4], u[4], v[4], w[4]; unsigned cl[4];
        if (full) { ld4(A.a.f[0] + p0, fx); ld4(A.a.f[1] + p0, fy); ld4(A.a.f[2] + p0, fz); ld4(A.a.f[3] + p0, u); ld4(A.a.f[4] + p0, v); ld4(A.a.f[5] + p0, w); ld4(A.a.cell + p0, cl); }
        else {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const bool ok = p0 + r < n; const u64 p = ok ? p0 + r : 0;
                fx[r] = ok ? A.a.f[0][p] : 0; fy[r] = ok ? A.a.f[1][p] : 0; fz[r] = ok ? A.a.f[2][p] : 0; u[r] = ok ? A.a.f[3][p] : 0; v[r] = ok ? A.a.f[4][p] : 0; w[r] = ok ? A.a.f[5][p] : 0;
                cl[r] = ok ? A.a.cell[p] : 0;
            }
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const bool ok = p0 + r < n;
            bool dead = false;
            if (ok) {
                int ci, cj, ck; cell_to_ijk(g, (int)cl[r], ci, cj, ck);
                float ex, ey, ez;
                gather_ef32(g, A.ef, ci, cj, ck, fx[r], fy[r], fz[r], ex, ey, ez);
                const float un = __fadd_rn(u[r], __fmul_rn(ex, A.qm_dt)), vn = __fadd_rn(v[r], __fmul_rn(ey, A.qm_dt)), wn = __fadd_rn(w[r], __fmul_rn(ez, A.qm_dt));   // Species.cpp:372
                float nx = fx[r], ny = fy[r], nz = fz[r];
                bool in = drift32(nx, ci, un, A.dt_idx, g.ci); in = drift32(ny, cj, vn, A.dt_idy, g.cj) && in; in = drift32(nz, ck, wn, A.dt_idz, g.ck) && in;    // :373
                if (in) {                                    // World::inObject on the absolute position (:375-388)
                    const double x = g.x0[0] + ((double)ci + (double)nx) * g.dx[0], y = g.x0[1] + ((double)cj + (double)ny) * g.dx[1], z = g.x0[2] + ((double)ck + (double)nz) * g.dx[2];
                    dead = in_object(g, x, y, z) != 0;
                } else dead = true;
                if (!dead) { fx[r] = nx; fy[r] = ny; fz[r] = nz; u[r] = un; v[r] = vn; w[r] = wn; cl[r] = (unsigned)cell_of(g, ci, cj, ck); }
            }
            record_dead(dead, lane, p0 + r, A.ctr, A.dead_list);
        }
        if (full) { st4(A.a.f[0] + p0, fx); st4(A.a.f[1] + p0, fy); st4(A.a.f[2] + p0, fz); st4(A.a.f[3] + p0, u); st4(A.a.f[4] + p0, v); st4(A.a.f[5] + p0, w); st4(A.a.cell + p0, cl); }
        else {
#pragma unroll
            for (int r = 0; r < 4; r++) if (p0 + r < n) {
                const u64 p = p0 + r;
                A.a.f[0][p] = fx[r]; A.a.f[1][p] = fy[r]; A.a.f[2][p] = fz[r]; A.a.f[3][p] = u[r]; A.a.f[4][p] = v[r]; A.a.f[5][p] = w[r]; A.a.cell[p] = cl[r];
            }
        }
    }
}
__global__ void k_compact_move32(const SpeciesCounters* ctr, Arr32 a, const unsigned* __restrict__ D) {      // plan: push.cu / push.cuh
    const u64 nh = ctr->n_hole, nd = ctr->n_dead, n_alive = ctr->n - nd;
    for (u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x; t < nh; t += (u64)gridDim.x * blockDim.x) {
        const unsigned d = D[t], f = compact_survivor(D, nh, nd, n_alive, t);
#pragma unroll
        for (int c = 0; c < 7; c++) a.f[c][d] = a.f[c][f];
        a.cell[d] = a.cell[f];
    }
}

// ---------------------------------------------------------------- deposit + per-cell count
// Field<double>::scatter (Field.h:157-199) with fp32 weights: ((val*wi)*wj)*wk per corner, val pre-multiplied by 2^S (exact), rounded to the
// int64 fixed-point grid.  A thread owns 8 consecutive particles; in a cell-sorted store they share a cell, so the eight corner sums stay in
// registers and leave once per (thread, cell) as 64-bit integer reductions (associative: any order gives the same bits).
struct Dep32Args { Arr32 a; const SpeciesCounters* ctr; u64* den_fixed; double* macro_count; float scale; };
template <bool DEPOSIT, bool COUNT>
__global__ void __launch_bounds__(256, 2) k_deposit32(Grid g, Dep32Args A) {
    const u64 n = A.ctr->n;
    const u64 nthreads = (u64)gridDim.x * blockDim.x;
    for (u64 p0 = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * 8; p0 < n; p0 += nthreads * 8) {
        const bool full = p0 + 8 <= n;
        float fx[8], fy[8], fz[8], m[8]; unsigned cl[8];
        if (full) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                ld4(A.a.cell + p0 + 4 * h, cl + 4 * h);
                if (DEPOSIT) { ld4(A.a.f[0] + p0 + 4 * h, fx + 4 * h); ld4(A.a.f[1] + p0 + 4 * h, fy + 4 * h); ld4(A.a.f[2] + p0 + 4 * h, fz + 4 * h); ld4(A.a.f[6] + p0 + 4 * h, m + 4 * h); }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const bool ok = p0 + r < n; const u64 p = ok ? p0 + r : 0;
                cl[r] = ok ? A.a.cell[p] : 0xffffffffu;
                if (DEPOSIT) { fx[r] = ok ? A.a.f[0][p] : 0; fy[r] = ok ? A.a.f[1][p] : 0; fz[r] = ok ? A.a.f[2][p] : 0; m[r] = ok ? A.a.f[6][p] : 0; }
            }
        }
        unsigned cur = 0xffffffffu; i64 acc[8]; int cnt = 0;
#pragma unroll
        for (int c = 0; c < 8; c++) acc[c] = 0;
        auto flush = [&]() {
            if (cur == 0xffffffffu) return;
            if (DEPOSIT) {
                int i, j, k; cell_to_ijk(g, (int)cur, i, j, k);
#pragma unroll
                for (int c = 0; c < 8; c++) if (acc[c]) atomicAdd(&A.den_fixed[corner_node(g, i, j, k, c)], (u64)acc[c]);
            }
            if (COUNT) atomicAdd(&A.macro_count[cur], (double)cnt);
        };
#pragma unroll
        for (int r = 0; r < 8; r++) {
            if (cl[r] == 0xffffffffu) continue;
            if (cl[r] != cur) {
                flush();
                cur = cl[r]; cnt = 0;
#pragma unroll
                for (int c = 0; c < 8; c++) acc[c] = 0;
            }
            cnt++;
            if (DEPOSIT) {
                const float di = fx[r], dj = fy[r], dk = fz[r];
                const float odi = __fsub_rn(1.0f, di), odj = __fsub_rn(1.0f, dj), odk = __fsub_rn(1.0f, dk);
                const float vs = __fmul_rn(m[r], A.scale);
                const float w00 = __fmul_rn(__fmul_rn(vs, odi), odj), w01 = __fmul_rn(__fmul_rn(vs, odi), dj), w10 = __fmul_rn(__fmul_rn(vs, di), odj), w11 = __fmul_rn(__fmul_rn(vs, di), dj);
                acc[0] += __float2ll_rn(__fmul_rn(w00, odk)); acc[1] += __float2ll_rn(__fmul_rn(w00, dk));
                acc[2] += __float2ll_rn(__fmul_rn(w01, odk)); acc[3] += __float2ll_rn(__fmul_rn(w01, dk));
                acc[4] += __float2ll_rn(__fmul_rn(w10, odk)); acc[5] += __float2ll_rn(__fmul_rn(w10, dk));
                acc[6] += __float2ll_rn(__fmul_rn(w11, odk)); acc[7] += __float2ll_rn(__fmul_rn(w11, dk));
            }
        }
        flush();
    }
}

// ---------------------------------------------------------------- sort, diagnostics, charge density
template <typename V>
__global__ void __launch_bounds__(256) k_permute32(const SpeciesCounters* ctr, const unsigned* __restrict__ idx, const V* __restrict__ in, V* __restrict__ out) {
    const u64 n = ctr->n;
    for (u64 p = blockIdx.x * (u64)blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x) out[p] = in[idx[p]];
}
__global__ void __launch_bounds__(256) k_copy_cells(const SpeciesCounters* ctr, const unsigned* __restrict__ in, unsigned* __restrict__ out) {
    const u64 n = ctr->n;
    for (u64 p = blockIdx.x * (u64)blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x) out[p] = in[p];
}
// getMicroCount / getMomentum / getKE (Species.cpp:731-752), accumulated in fp64
__global__ void __launch_bounds__(256) k_diag32(const SpeciesCounters* ctr, Arr32 a, double* __restrict__ out) {
    __shared__ double sm[5][256];
    const u64 n = ctr->n;
    double acc[5] = {0, 0, 0, 0, 0};
    for (u64 p = blockIdx.x * (u64)blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x) {
        const double u = a.f[3][p], v = a.f[4][p], w = a.f[5][p], m = a.f[6][p];
        acc[0] += m; acc[1] += m * u; acc[2] += m * v; acc[3] += m * w; acc[4] += m * (u * u + v * v + w * w);
    }
    for (int c = 0; c < 5; c++) sm[c][threadIdx.x] = acc[c];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) for (int c = 0; c < 5; c++) sm[c][threadIdx.x] += sm[c][threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) for (int c = 0; c < 5; c++) atomicAdd(&out[c], sm[c][0]);
}
__global__ void __launch_bounds__(256) k_rho_add(int nv, double q, const double* __restrict__ den, double* __restrict__ rho, int first) {
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < nv; u += gridDim.x * blockDim.x) rho[u] = first ? __dmul_rn(q, den[u]) : __dadd_rn(rho[u], __dmul_rn(q, den[u]));
}

namespace {
int refresh_count32(picg_species32_s* s) {
    if (s->n_host_valid) return PICG_OK;
    CUDA_TRY(cudaMemcpyAsync(s->ctr_host, s->ctr, sizeof(SpeciesCounters), cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    s->n_host = (size_t)s->ctr_host->n; s->n_host_valid = true; s->n_upper = s->n_host;
    return PICG_OK;
}
int ensure_capacity32(picg_species32_s* s, size_t cap) {
    if (cap <= s->cap) return PICG_OK;
    int rc = refresh_count32(s); if (rc) return rc;
    const size_t newcap = (std::max(cap, s->cap + s->cap / 2) + 255) & ~(size_t)255;
    cudaStreamSynchronize(g_stream);
    cudaFree(s->fspare); cudaFree(s->cspare); s->fspare = nullptr; s->cspare = nullptr;      // the sort's out-of-place targets are allocated when a sort needs them
    for (int c = 0; c < 8; c++) {
        void** slot = c < 7 ? (void**)&s->f[c] : (void**)&s->cell;
        void* fresh = nullptr;
        cudaError_t e = cudaMalloc(&fresh, newcap * 4);
        if (e != cudaSuccess) {
            size_t fr = 0, tot = 0; cudaMemGetInfo(&fr, &tot);
            return set_error(PICG_ERR_OOM, "fp32 species store cannot grow to %zu particles: %s (%.1f GB free of %.1f)", newcap, cudaGetErrorString(e), fr / 1e9, tot / 1e9);
        }
        if (s->n_host && *slot) { CUDA_TRY(cudaMemcpyAsync(fresh, *slot, s->n_host * 4, cudaMemcpyDeviceToDevice, g_stream)); CUDA_TRY(cudaStreamSynchronize(g_stream)); }
        cudaFree(*slot); *slot = fresh;
    }
    note_realloc("fp32 particle store", newcap * 32);
    s->cap = newcap;
    return ensure_scratch(s->w, std::max(newcap * 16 + (1u << 20), compact_scratch_bytes(newcap) + 64));
}
int compact_dead32(picg_species32_s* s, size_t cap) {
    unsigned* D = nullptr;
    int rc = compact_plan(s->w, s->ctr, cap, s->cap, &D); if (rc) return rc;
    int grid = std::max(1, std::min(div_up(std::max<size_t>(cap / 16, 1), 256), g_sm_count * 4));
    LAUNCH(K_COMPACT, k_compact_move32, grid, 256, 0, s->ctr, arr_of(s), D); CHECK_LAUNCH();
    LAUNCH(K_COMPACT, k_compact_finish, 1, 1, 0, s->ctr); CHECK_LAUNCH();
    s->n_host_valid = false; s->sorted_valid = false;
    return PICG_OK;
}
const size_t kChunk32 = 1u << 22;
}  // namespace

extern "C" {

int picg_species32_create(picg_world_t w, double mass, double charge, double mpw0, picg_species32_t* out) {
    REQUIRE_DEVICE(); REQUIRE_ARG(w && out, "picg_species32_create: null argument");
    picg_species32_s* s = new picg_species32_s();
    s->w = w; s->mass = mass; s->charge = charge; s->mpw0 = mpw0;
    const Grid& g = w->g;
    cudaError_t e;
    if ((e = cudaMalloc(&s->ctr, sizeof(SpeciesCounters))) != cudaSuccess || (e = cudaMallocHost(&s->ctr_host, sizeof(SpeciesCounters))) != cudaSuccess ||
        (e = cudaMalloc(&s->den_fixed, (size_t)g.nv * 8)) != cudaSuccess || (e = cudaMalloc(&s->den, (size_t)g.nv * 8)) != cudaSuccess ||
        (e = cudaMalloc(&s->macro_count, (size_t)g.nc * 8)) != cudaSuccess || (e = cudaMalloc(&s->cell_start, ((size_t)g.nc + 1) * 4)) != cudaSuccess) {
        picg_species32_destroy(s); return cuda_fail(e, "cudaMalloc(species32)", __FILE__, __LINE__);
    }
    memset(s->ctr_host, 0, sizeof(SpeciesCounters));
    CUDA_TRY(cudaMemsetAsync(s->ctr, 0, sizeof(SpeciesCounters), g_stream));
    CUDA_TRY(cudaMemsetAsync(s->den_fixed, 0, (size_t)g.nv * 8, g_stream)); CUDA_TRY(cudaMemsetAsync(s->den, 0, (size_t)g.nv * 8, g_stream));
    CUDA_TRY(cudaMemsetAsync(s->macro_count, 0, (size_t)g.nc * 8, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    *out = s; return PICG_OK;
}
int picg_species32_destroy(picg_species32_t s) {
    if (!s) return PICG_OK;
    if (g_stream) cudaStreamSynchronize(g_stream);
    for (int c = 0; c < 7; c++) cudaFree(s->f[c]);
    cudaFree(s->cell); cudaFree(s->fspare); cudaFree(s->cspare); cudaFree(s->ctr); cudaFreeHost(s->ctr_host);
    cudaFree(s->den_fixed); cudaFree(s->den); cudaFree(s->macro_count); cudaFree(s->cell_start);
    delete s; return PICG_OK;
}
int picg_species32_reserve(picg_species32_t s, size_t capacity) { REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species32_reserve: null species"); return ensure_capacity32(s, capacity); }
int picg_species32_count(picg_species32_t s, size_t* n) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && n, "picg_species32_count: null argument");
    int rc = refresh_count32(s); *n = s->n_host; return rc;
}
int picg_species32_upload(picg_species32_t s, size_t n, const double* aos7) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && (aos7 || n == 0), "picg_species32_upload: null argument");
    s->n_host = 0; s->n_host_valid = true;
    int rc = ensure_capacity32(s, std::max<size_t>(n, 256)); if (rc) return rc;
    rc = ensure_scratch(s->w, std::min(std::max<size_t>(n, 1), kChunk32) * 56 + 64); if (rc) return rc;
    for (size_t off = 0; off < n; off += kChunk32) {
        const size_t m = std::min(kChunk32, n - off);
        CUDA_TRY(cudaMemcpyAsync(s->w->scratch, aos7 + off * 7, m * 56, cudaMemcpyHostToDevice, g_stream));
        LAUNCH(K_TRANSPOSE, k_aos_to_f32, std::min(div_up(m, 256), g_sm_count * 8), 256, 0, s->w->g, m, (const double*)s->w->scratch, arr_of(s), off); CHECK_LAUNCH();
        CUDA_TRY(cudaStreamSynchronize(g_stream));
    }
    SpeciesCounters z; memset(&z, 0, sizeof(z)); z.n = n; *s->ctr_host = z;
    CUDA_TRY(cudaMemcpyAsync(s->ctr, s->ctr_host, sizeof(z), cudaMemcpyHostToDevice, g_stream)); CUDA_TRY(cudaStreamSynchronize(g_stream));
    s->n_host = n; s->n_host_valid = true; s->n_upper = n; s->sorted_valid = false;
    return PICG_OK;
}
// device-side conversion of an fp64 store (same world): every particle keeps its order
int picg_species32_from_species(picg_species32_t s, picg_species_t src) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && src && s->w == src->w, "picg_species32_from_species: null argument or different worlds");
    int rc = species_refresh_count(src); if (rc) return rc;
    const size_t n = src->n_host;
    s->n_host = 0; s->n_host_valid = true;
    rc = ensure_capacity32(s, std::max<size_t>(n, 256)); if (rc) return rc;
    LAUNCH(K_TRANSPOSE, k_soa64_to_f32, std::max(1, std::min(div_up(std::max<size_t>(n, 1), 256), g_sm_count * 8)), 256, 0, s->w->g, src->ctr, src->a[0], src->a[1], src->a[2], src->a[3], src->a[4],
           src->a[5], src->a[6], arr_of(s)); CHECK_LAUNCH();
    SpeciesCounters z; memset(&z, 0, sizeof(z)); z.n = n; *s->ctr_host = z;
    CUDA_TRY(cudaMemcpyAsync(s->ctr, s->ctr_host, sizeof(z), cudaMemcpyHostToDevice, g_stream)); CUDA_TRY(cudaStreamSynchronize(g_stream));
    s->n_host = n; s->n_upper = n; s->sorted_valid = false;
    return PICG_OK;
}
int picg_species32_download(picg_species32_t s, size_t capacity, double* aos7, size_t* n_out) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && n_out, "picg_species32_download: null argument");
    int rc = refresh_count32(s); if (rc) return rc;
    const size_t n = s->n_host; *n_out = n;
    if (!aos7) return PICG_OK;
    REQUIRE_ARG(capacity >= n, "picg_species32_download: host buffer too small");
    rc = ensure_scratch(s->w, std::min(std::max<size_t>(n, 1), kChunk32) * 56 + 64); if (rc) return rc;
    for (size_t off = 0; off < n; off += kChunk32) {
        const size_t m = std::min(kChunk32, n - off);
        LAUNCH(K_TRANSPOSE, k_f32_to_aos, std::min(div_up(m, 256), g_sm_count * 8), 256, 0, s->w->g, m, arr_of(s), off, (double*)s->w->scratch); CHECK_LAUNCH();
        CUDA_TRY(cudaMemcpyAsync(aos7 + off * 7, s->w->scratch, m * 56, cudaMemcpyDeviceToHost, g_stream));
        CUDA_TRY(cudaStreamSynchronize(g_stream));
    }
    return PICG_OK;
}
// Species::advanceElectrons (Species.cpp:258-399) on the fp32 store
int picg_species32_push_electrons(picg_species32_t s, double dt) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species32_push_electrons: null species");
    const Grid& g = s->w->g;
    const size_t cap = std::max<size_t>(s->n_upper, 1);
    if (cap >= 0xfffffff0ull) return set_error(PICG_ERR_ARG, "more than 2^32-1 particles per GPU are not supported");
    int rc = ensure_scratch(s->w, compact_scratch_bytes(cap)); if (rc) return rc;
    Push32Args A; A.a = arr_of(s); A.ctr = s->ctr; A.ef = s->w->ef; A.qm_dt = (float)(dt * s->charge / s->mass);
    A.dt_idx = (float)(dt * g.inv_dx[0]); A.dt_idy = (float)(dt * g.inv_dx[1]); A.dt_idz = (float)(dt * g.inv_dx[2]); A.dead_list = (unsigned*)s->w->scratch;
    const int grid = std::max(1, std::min(div_up(cap, 256 * 4), g_sm_count * 2 * 4));
    LAUNCH(K_PUSH_ELECTRONS, k_push32, grid, 256, 0, g, A); CHECK_LAUNCH();
    s->sorted_valid = false;
    return compact_dead32(s, cap);
}
int picg_species32_set_density_scale(picg_species32_t s, int S) { REQUIRE_ARG(s && S > -1000 && S < 1000, "picg_species32_set_density_scale: bad argument"); s->S = S; s->S_set = true; return PICG_OK; }
int picg_species32_density_scale(picg_species32_t s, int* S) { REQUIRE_ARG(s && S, "picg_species32_density_scale: null argument"); *S = s->S; return PICG_OK; }
int picg_species32_diagnostics(picg_species32_t s, double* micro_count, double momentum[3], double* ke) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species32_diagnostics: null species");
    double* d = s->w->reduce_buf;
    CUDA_TRY(cudaMemsetAsync(d, 0, 5 * 8, g_stream));
    LAUNCH(K_DIAG, k_diag32, std::max(1, std::min(div_up(std::max<size_t>(s->n_upper, 1), 256), g_sm_count * 4)), 256, 0, s->ctr, arr_of(s), d); CHECK_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(s->w->reduce_host, d, 5 * 8, cudaMemcpyDeviceToHost, g_stream)); CUDA_TRY(cudaStreamSynchronize(g_stream));
    const double* h = s->w->reduce_host;
    if (micro_count) *micro_count = h[0];
    if (momentum) for (int c = 0; c < 3; c++) momentum[c] = s->mass * h[1 + c];
    if (ke) *ke = 0.5 * s->mass * h[4];
    return PICG_OK;
}
// Species::computeNumberDensity + computeMacroParticlesCount on the fp32 store (the count is a by-product)
int picg_species32_deposit_density(picg_species32_t s) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species32_deposit_density: null species");
    const Grid& g = s->w->g;
    if (!s->S_set) {                                   // rigorous bound: no node sum exceeds the total weight
        double total = 0; int rc = picg_species32_diagnostics(s, &total, nullptr, nullptr); if (rc) return rc;
        int e = 0; if (total > 0) std::frexp(total, &e);
        s->S = 61 - e; s->S_set = true;
    }
    CUDA_TRY(cudaMemsetAsync(s->den_fixed, 0, (size_t)g.nv * 8, g_stream));
    CUDA_TRY(cudaMemsetAsync(s->macro_count, 0, (size_t)g.nc * 8, g_stream));
    Dep32Args A; A.a = arr_of(s); A.ctr = s->ctr; A.den_fixed = (u64*)s->den_fixed; A.macro_count = s->macro_count; A.scale = (float)std::ldexp(1.0, s->S);
    const int grid = std::max(1, std::min(div_up(std::max<size_t>(s->n_upper, 1), 256 * 8), g_sm_count * 2 * 4));
    LAUNCH(K_DEPOSIT, (k_deposit32<true, true>), grid, 256, 0, g, A); CHECK_LAUNCH();
    LAUNCH(K_MISC, k_reset_den_stats, 1, 1, 0, s->ctr); CHECK_LAUNCH();
    LAUNCH(K_FINALIZE_DEN, k_finalize_den, std::min(div_up((size_t)g.nv, 256), g_sm_count * 8), 256, 0, 0, g.nv, s->den_fixed, s->w->node_vol, std::ldexp(1.0, -s->S), s->den, s->ctr);
    CHECK_LAUNCH();
    return PICG_OK;
}
int picg_species32_sort(picg_species32_t s) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species32_sort: null species");
    const Grid& g = s->w->g;
    int rc = refresh_count32(s); if (rc) return rc;
    const size_t cap = std::max<size_t>(s->n_host, 1);
    int nblocks = std::max(1, std::min(std::min(div_up(cap, 2048), g_sm_count * 4), 1024));
    const size_t capa = (cap + 63) & ~(size_t)63;
    rc = ensure_scratch(s->w, capa * 16 + (size_t)256 * nblocks * 4 + 256 * 4 + 256); if (rc) return rc;
    if (!s->fspare) { cudaError_t e = cudaMalloc(&s->fspare, s->cap * 4); if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(fp32 sort target)", __FILE__, __LINE__); }
    if (!s->cspare) { cudaError_t e = cudaMalloc(&s->cspare, s->cap * 4); if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(fp32 sort target)", __FILE__, __LINE__); }
    unsigned* keysA = (unsigned*)s->w->scratch; unsigned* keysB = keysA + capa; unsigned* idxA = keysB + capa; unsigned* idxB = idxA + capa; unsigned* counts = idxB + capa;
    const u64* n_ptr = &s->ctr->n;
    const int pgrid = std::max(1, std::min(div_up(cap, 256), g_sm_count * 8));
    LAUNCH(K_SORT_KEYS, k_copy_cells, pgrid, 256, 0, s->ctr, s->cell, keysA); CHECK_LAUNCH();      // the stored cell index is the key
    LAUNCH(K_SORT_KEYS, k_iota_u32, pgrid, 256, 0, n_ptr, idxA); CHECK_LAUNCH();
    int bits = 1; while ((1ull << bits) < (u64)g.nc) bits++;
    rc = radix_sort_pairs(n_ptr, cap, bits, keysA, idxA, keysB, idxB, counts, nblocks); if (rc) return rc;
    for (int c = 0; c < 7; c++) { LAUNCH(K_SORT_PERMUTE, (k_permute32<float>), pgrid, 256, 0, s->ctr, idxA, s->f[c], s->fspare); CHECK_LAUNCH(); std::swap(s->f[c], s->fspare); }
    LAUNCH(K_SORT_PERMUTE, k_copy_cells, pgrid, 256, 0, s->ctr, keysA, s->cspare); CHECK_LAUNCH(); std::swap(s->cell, s->cspare);
    LAUNCH(K_CELL_START, k_cell_start, pgrid, 256, 0, n_ptr, keysA, g.nc, s->cell_start); CHECK_LAUNCH();
    s->sorted_valid = true;
    return PICG_OK;
}
int picg_species32_download_field(picg_species32_t s, int field, void* host) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && host, "picg_species32_download_field: null argument");
    const Grid& g = s->w->g;
    const void* src = nullptr; size_t bytes = (size_t)g.nv * 8;
    switch (field) {
        case PICG_SF_DEN: src = s->den; break;
        case PICG_SF_DEN_FIXED: src = s->den_fixed; break;
        case PICG_SF_MACRO_COUNT: src = s->macro_count; bytes = (size_t)g.nc * 8; break;
        default: return set_error(PICG_ERR_ARG, "picg_species32_download_field: the fp32 store keeps den, den_fixed and macro_part_count");
    }
    CUDA_TRY(cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost, g_stream)); CUDA_TRY(cudaStreamSynchronize(g_stream));
    return PICG_OK;
}
// World::computeChargeDensity (World.cpp:193-200) over fp32 species: rho = sum of charge * den (neutral species skipped)
int picg_world_charge_density32(picg_world_t w, picg_species32_t* sp, int n) {
    REQUIRE_DEVICE(); REQUIRE_ARG(w && (sp || n == 0), "picg_world_charge_density32: null argument");
    const Grid& g = w->g;
    CUDA_TRY(cudaMemsetAsync(w->rho, 0, (size_t)g.nv * 8, g_stream));
    int first = 1;
    for (int k = 0; k < n; k++) {
        if (!sp[k] || sp[k]->w != w) return set_error(PICG_ERR_ARG, "picg_world_charge_density32: species of another world");
        if (sp[k]->charge == 0) continue;
        LAUNCH(K_CHARGE_DENSITY, k_rho_add, std::min(div_up((size_t)g.nv, 256), g_sm_count * 8), 256, 0, g.nv, sp[k]->charge, sp[k]->den, w->rho, first); CHECK_LAUNCH();
        first = 0;
    }
    return PICG_OK;
}

}  // extern "C"
