// Species: structure-of-arrays particle store on the device, host<->device transposes of the
// reference's AoS Particle record, addParticle semantics, diagnostics, node-field bookkeeping.
// Replaces the storage side of ch4/v3/src/Species.{h,cpp}.
#include "common.cuh"
#include "push.cuh"
#include <cstring>
#include <algorithm>

using namespace picg;

static const size_t kChunk = 1u << 22;     // particles per staging chunk for host<->device transposes

// ---------------------------------------------------------------- kernels
struct SoA { double* a[7]; };

// AoS (7 doubles / particle, staged in scratch) -> SoA at offset `base`
__global__ void __launch_bounds__(256) k_aos_to_soa(size_t n, const double* __restrict__ aos, SoA s, size_t base) {
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 7; c++) s.a[c][base + p] = aos[p * 7 + c];
    }
}
__global__ void __launch_bounds__(256) k_soa_to_aos(size_t n, SoA s, size_t base, double* __restrict__ aos) {
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 7; c++) aos[p * 7 + c] = s.a[c][base + p];
    }
}

// Species::addParticle (Species.cpp:420-434): reject NaN pos/vel, out of bounds, in object; gather E at pos;
// vel -= charge/mass * E * (0.5*world.dt); append.  The accepted candidates are appended IN CANDIDATE ORDER: block b owns a contiguous
// range of candidates, k_add_count counts what each range keeps, one block scans the 1024 counts, k_add_write places every accepted
// candidate at n + (kept in the blocks before) + (kept before it in its block).  No atomics: the same candidates give the same store.
#define ADD_BLOCKS 1024
__device__ __forceinline__ bool add_keep(const Grid& g, const double* __restrict__ a) {
    const double x = a[0], y = a[1], z = a[2];
    bool keep = !(isnan(x) || isnan(y) || isnan(z) || isnan(a[3]) || isnan(a[4]) || isnan(a[5]));
    return keep && in_bounds(g, x, y, z) && !in_object(g, x, y, z);
}
__device__ __forceinline__ void add_range(size_t n, size_t& p0, size_t& p1) {      // candidates of block blockIdx.x (multiples of the block size)
    const size_t per = ((n + ADD_BLOCKS - 1) / ADD_BLOCKS + 255) & ~(size_t)255;
    p0 = min(n, (size_t)blockIdx.x * per); p1 = min(n, p0 + per);
}
__global__ void __launch_bounds__(256) k_add_count(Grid g, size_t n, const double* __restrict__ aos, unsigned* __restrict__ counts) {
    __shared__ unsigned ws[8];
    size_t p0, p1; add_range(n, p0, p1);
    unsigned c = 0;
    for (size_t p = p0 + threadIdx.x; p < p1; p += 256) c += add_keep(g, aos + p * 7) ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned t = 0; for (int q = 0; q < 8; q++) t += ws[q]; counts[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(ADD_BLOCKS) k_add_scan(unsigned* __restrict__ counts) {      // exclusive scan in place, total in counts[ADD_BLOCKS]
    __shared__ unsigned ws[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned v = counts[threadIdx.x]; unsigned x = v;
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    if (warp == 0) { unsigned wv = ws[lane], wx = wv; for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, wx, o); if (lane >= o) wx += y; } ws[lane] = wx - wv; }
    __syncthreads();
    counts[threadIdx.x] = ws[warp] + x - v;
    if (threadIdx.x == ADD_BLOCKS - 1) counts[ADD_BLOCKS] = ws[warp] + x;
}
__global__ void __launch_bounds__(256) k_add_write(Grid g, size_t n, const double* __restrict__ aos, SoA s, size_t cap, const SpeciesCounters* ctr,
                                                   const unsigned* __restrict__ offsets, const double* __restrict__ ef, double q_over_m, double half_dt) {
    __shared__ unsigned ws[8];
    size_t p0, p1; add_range(n, p0, p1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 base = ctr->n + offsets[blockIdx.x];
    for (size_t q0 = p0; q0 < p1; q0 += 256) {
        const size_t p = q0 + threadIdx.x;
        const bool keep = p < p1 && add_keep(g, aos + p * 7);
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) ws[warp] = __popc(mask);
        __syncthreads();
        unsigned before = 0, total = 0;
        for (int q = 0; q < 8; q++) { const unsigned t = ws[q]; if (q < warp) before += t; total += t; }
        if (keep) {
            const u64 dst = base + before + __popc(mask & ((1u << lane) - 1));
            if (dst < cap) {
                const double x = aos[p * 7], y = aos[p * 7 + 1], z = aos[p * 7 + 2];
                double u = aos[p * 7 + 3], v = aos[p * 7 + 4], w = aos[p * 7 + 5];
                double ex, ey, ez;
                gather_ef(g, ef, x_to_l(x, g.x0[0], g.inv_dx[0]), x_to_l(y, g.x0[1], g.inv_dx[1]), x_to_l(z, g.x0[2], g.inv_dx[2]), ex, ey, ez);
                // vel -= charge/mass*ef_part*(0.5*dt):  ((q/m)*E)*(0.5 dt)   (scalar*Vec3 then Vec3*scalar)
                u = __dsub_rn(u, __dmul_rn(__dmul_rn(ex, q_over_m), half_dt));
                v = __dsub_rn(v, __dmul_rn(__dmul_rn(ey, q_over_m), half_dt));
                w = __dsub_rn(w, __dmul_rn(__dmul_rn(ez, q_over_m), half_dt));
                s.a[0][dst] = x; s.a[1][dst] = y; s.a[2][dst] = z; s.a[3][dst] = u; s.a[4][dst] = v; s.a[5][dst] = w; s.a[6][dst] = aos[p * 7 + 6];
            }
        }
        base += total;
        __syncthreads();
    }
}
__global__ void k_add_finish(SpeciesCounters* ctr, const unsigned* __restrict__ counts, u64 cap) {
    const u64 want = ctr->n + counts[ADD_BLOCKS];
    if (want > cap) { ctr->overflow += want - cap; ctr->n = cap; } else ctr->n = want;
}
__global__ void k_clamp_count(SpeciesCounters* ctr, u64 cap) { if (ctr->n > cap) ctr->n = cap; }

// getMicroCount / getMomentum / getKE (Species.cpp:731-752): sums of mpw, mpw*v, mpw*(v.v)
__global__ void __launch_bounds__(256) k_diag(const SpeciesCounters* ctr, SoA s, double* __restrict__ out) {
    __shared__ double sm[5][256];
    u64 n = ctr->n;
    double acc[5] = {0, 0, 0, 0, 0};
    for (u64 p = blockIdx.x * (u64)blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x) {
        double u = s.a[3][p], v = s.a[4][p], w = s.a[5][p], m = s.a[6][p];
        acc[0] += m; acc[1] += m * u; acc[2] += m * v; acc[3] += m * w; acc[4] += m * (u * u + v * v + w * w);
    }
    for (int c = 0; c < 5; c++) sm[c][threadIdx.x] = acc[c];
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) for (int c = 0; c < 5; c++) sm[c][threadIdx.x] += sm[c][threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x < 5) out[blockIdx.x * 5 + threadIdx.x] = sm[threadIdx.x][0];
}

// Field::updateMovingAverage (Field.h:246-261)
__global__ void k_moving_avg(int nv, const double* __restrict__ den, double* __restrict__ avg, double samples, int clear) {
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < nv; u += gridDim.x * blockDim.x) {
        double a = clear ? 0.0 : avg[u];
        avg[u] = (den[u] + a * (samples - 1.0)) / samples;
    }
}

// ---------------------------------------------------------------- helpers
namespace picg {
int species_refresh_count(picg_species_s* s) {
    if (s->n_host_valid) return PICG_OK;
    CUDA_TRY(cudaMemcpyAsync(s->ctr_host, s->ctr, sizeof(SpeciesCounters), cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    s->n_host = (size_t)s->ctr_host->n;
    s->n_host_valid = true; s->n_upper = s->n_host;
    if (s->ctr_host->overflow) {
        u64 lost = s->ctr_host->overflow;
        cudaMemsetAsync(&s->ctr->overflow, 0, 8, g_stream);
        return set_error(PICG_ERR_OOM, "species store full: %llu appended particles were dropped (capacity %zu); call picg_species_reserve", lost, s->cap);
    }
    if (s->S_pinned && s->ctr_host->den_neg) {        // pinned scale (multi-GPU): nobody re-calibrates, so the wrapped accumulator must not pass silently
        u64 bad = s->ctr_host->den_neg;
        cudaMemsetAsync(&s->ctr->den_neg, 0, 8, g_stream);
        return set_error(PICG_ERR_OVERFLOW, "fixed-point density accumulator overflowed on %llu nodes with the pinned scale S=%d: agree on a smaller S (picg_species_set_density_scale) and deposit again", bad, s->S);
    }
    return PICG_OK;
}

// Grows the particle arrays only; the caller sizes the scratch arena afterwards (species_scratch_for_store) - used where the arena holds
// live data while a store has to grow (mcc.cu: staged products).
int species_grow_store(picg_species_s* s, size_t cap) {
    if (cap <= s->cap) return PICG_OK;
    int rc = species_refresh_count(s); if (rc) return rc;
    size_t exact = (cap + 255) & ~(size_t)255;
    size_t newcap = (std::max(cap, s->cap + s->cap / 2) + 255) & ~(size_t)255;
    // array by array (allocate, copy, free the old one): the peak is the old store plus ONE new array, not two whole stores
    cudaStreamSynchronize(g_stream);
    for (int attempt = 0; attempt < 2; attempt++) {
        cudaError_t e = cudaSuccess; int c = 0;
        for (; c < 8 && e == cudaSuccess; c++) {
            double*& old = c < 7 ? s->a[c] : s->spare;
            double* fresh = nullptr;
            e = cudaMalloc(&fresh, newcap * 8);
            if (e != cudaSuccess) break;
            if (c < 7 && s->n_host && old) e = cudaMemcpyAsync(fresh, old, s->n_host * 8, cudaMemcpyDeviceToDevice, g_stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
            if (e != cudaSuccess) { cudaFree(fresh); break; }
            cudaFree(old); old = fresh;
        }
        if (e == cudaSuccess) break;
        cudaGetLastError();
        // arrays [0, c) already have the larger size, the rest the old one: the store stays valid at its old capacity
        if (attempt == 1 || newcap == exact || c > 0)
            return set_error(PICG_ERR_OOM, "species store cannot grow from %zu to %zu particles (%zu live): %s; reserve the capacity up front (picg_species_reserve)", s->cap, newcap, s->n_host, cudaGetErrorString(e));
        newcap = exact;                                           // retry with the exact request
    }
    note_realloc("particle store", newcap * 64);
    s->cap = newcap;
    return PICG_OK;
}
// the shared scratch arena is sized for a full store as well (radix sort: 16 B per particle; push: dead / hole lists +
// the heavy push's impact list), so that it does not have to grow - free + allocate of GBs, tens of ms - while the
// population grows into the reserved capacity
int species_scratch_for_store(picg_species_s* s) { return ensure_scratch(s->w, std::max(s->cap * 16 + (1u << 20), compact_scratch_bytes(s->cap) + s->cap * 4 + 64)); }
int species_ensure_capacity(picg_species_s* s, size_t cap) {
    if (cap <= s->cap) return PICG_OK;
    int rc = species_grow_store(s, cap); if (rc) return rc;
    return species_scratch_for_store(s);
}
}  // namespace picg

static SoA soa_of(picg_species_s* s) { SoA r; for (int c = 0; c < 7; c++) r.a[c] = s->a[c]; return r; }

namespace picg {
static int add_candidates(picg_species_s* s, size_t n, const double* d_aos, double q_over_m, double half_dt) {
    unsigned* counts = (unsigned*)s->w->reduce_buf;              // ADD_BLOCKS + 1 words
    LAUNCH(K_ADD_PARTICLES, k_add_count, ADD_BLOCKS, 256, 0, s->w->g, n, d_aos, counts); CHECK_LAUNCH();
    LAUNCH(K_ADD_PARTICLES, k_add_scan, 1, ADD_BLOCKS, 0, counts); CHECK_LAUNCH();
    LAUNCH(K_ADD_PARTICLES, k_add_write, ADD_BLOCKS, 256, 0, s->w->g, n, d_aos, soa_of(s), s->cap, (const SpeciesCounters*)s->ctr, (const unsigned*)counts, s->w->ef, q_over_m, half_dt); CHECK_LAUNCH();
    LAUNCH(K_ADD_PARTICLES, k_add_finish, 1, 1, 0, s->ctr, (const unsigned*)counts, (u64)s->cap); CHECK_LAUNCH();
    return PICG_OK;
}
// addParticle for n candidates already staged on the device (AoS).  Capacity must have been ensured.
int species_add_staged(picg_species_s* s, size_t n, const double* d_aos) {
    double q_over_m = s->charge / s->mass, half_dt = 0.5 * s->w->dt;
    int rc = add_candidates(s, n, d_aos, q_over_m, half_dt); if (rc) return rc;
    s->n_host_valid = false; s->n_upper = std::min(s->cap, s->n_upper + n); s->sorted_valid = false; s->lists_valid = false; s->count_valid = false;
    return PICG_OK;
}
}

extern "C" {

int picg_species_create(picg_world_t w, double mass, double charge, double mpw0, picg_species_t* out) {
    REQUIRE_DEVICE();
    REQUIRE_ARG(w && out, "picg_species_create: null argument");
    REQUIRE_ARG(mass > 0 && mpw0 > 0, "picg_species_create: mass and mpw0 must be positive");
    picg_species_s* s = new picg_species_s();
    s->w = w; s->mass = mass; s->charge = charge; s->mpw0 = mpw0;
    s->id = w->n_species++; w->live_species++;
    size_t nv = w->g.nv, nc = w->g.nc;
    cudaError_t e;
    double** nodef[] = {&s->den, &s->den_avg, &s->T, &s->n_sum, &s->nuu, &s->nvv, &s->nww};
    for (double** f : nodef) {
        if ((e = cudaMalloc(f, nv * 8)) != cudaSuccess) { picg_species_destroy(s); return cuda_fail(e, "cudaMalloc(species field)", __FILE__, __LINE__); }
        cudaMemsetAsync(*f, 0, nv * 8, g_stream);
    }
    if ((e = cudaMalloc(&s->vel, nv * 24)) != cudaSuccess || (e = cudaMalloc(&s->nv_sum, nv * 24)) != cudaSuccess ||
        (e = cudaMalloc(&s->den_fixed, nv * 8)) != cudaSuccess || (e = cudaMalloc(&s->macro_count, nc * 8)) != cudaSuccess ||
        (e = cudaMalloc(&s->cell_start, (nc + 1) * 4)) != cudaSuccess || (e = cudaMalloc(&s->ctr, sizeof(SpeciesCounters))) != cudaSuccess ||
        (e = cudaMallocHost(&s->ctr_host, sizeof(SpeciesCounters))) != cudaSuccess) {
        picg_species_destroy(s); return cuda_fail(e, "cudaMalloc(species)", __FILE__, __LINE__);
    }
    cudaMemsetAsync(s->vel, 0, nv * 24, g_stream); cudaMemsetAsync(s->nv_sum, 0, nv * 24, g_stream);
    cudaMemsetAsync(s->den_fixed, 0, nv * 8, g_stream); cudaMemsetAsync(s->macro_count, 0, nc * 8, g_stream);
    cudaMemsetAsync(s->ctr, 0, sizeof(SpeciesCounters), g_stream);
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    *out = s;
    return PICG_OK;
}

int picg_species_destroy(picg_species_t s) {
    if (!s) return PICG_OK;
    if (g_stream) cudaStreamSynchronize(g_stream);
    for (int c = 0; c < 7; c++) cudaFree(s->a[c]);
    cudaFree(s->spare); cudaFree(s->home); cudaFree(s->home_alt); cudaFree(s->in_start); cudaFree(s->out_start); cudaFree(s->mv_in); cudaFree(s->mv_trip);
    cudaFree(s->den_fixed); cudaFree(s->den); cudaFree(s->den_avg); cudaFree(s->T); cudaFree(s->vel); cudaFree(s->n_sum);
    cudaFree(s->nv_sum); cudaFree(s->nuu); cudaFree(s->nvv); cudaFree(s->nww); cudaFree(s->macro_count); cudaFree(s->cell_start);
    cudaFree(s->ctr); if (s->ctr_host) cudaFreeHost(s->ctr_host);
    if (s->w && s->w->live_species && --s->w->live_species == 0) s->w->n_species = 0;
    delete s;
    return PICG_OK;
}

int picg_species_reserve(picg_species_t s, size_t capacity) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_reserve: null species");
    return species_ensure_capacity(s, capacity);
}

int picg_species_partition_size(picg_species_t s, size_t* n) {
    REQUIRE_ARG(s && n, "picg_species_partition_size: null argument");
    *n = s->part_valid ? s->part_n : 0;
    return PICG_OK;
}

int picg_species_count(picg_species_t s, size_t* n) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && n, "picg_species_count: null argument");
    int rc = species_refresh_count(s);
    *n = s->n_host;
    return rc;
}

int picg_species_upload(picg_species_t s, size_t n, const double* aos7) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && (aos7 || n == 0), "picg_species_upload: null argument");
    s->n_host = 0; s->n_host_valid = true;                        // contents are replaced
    int rc = species_ensure_capacity(s, std::max<size_t>(n, 256)); if (rc) return rc;
    rc = ensure_scratch(s->w, std::min(n, kChunk) * 56 + 64); if (rc) return rc;
    for (size_t off = 0; off < n; off += kChunk) {
        size_t m = std::min(kChunk, n - off);
        CUDA_TRY(cudaMemcpyAsync(s->w->scratch, aos7 + off * 7, m * 56, cudaMemcpyHostToDevice, g_stream));
        LAUNCH(K_TRANSPOSE, k_aos_to_soa, std::min(div_up(m, 256), g_sm_count * 8), 256, 0, m, (const double*)s->w->scratch, soa_of(s), off);
        CHECK_LAUNCH();
        CUDA_TRY(cudaStreamSynchronize(g_stream));                // staging buffer is reused
    }
    SpeciesCounters z; memset(&z, 0, sizeof(z)); z.n = n;
    *s->ctr_host = z;
    CUDA_TRY(cudaMemcpyAsync(s->ctr, s->ctr_host, sizeof(z), cudaMemcpyHostToDevice, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    s->n_host = n; s->n_host_valid = true; s->n_upper = n; s->sorted_valid = false; s->lists_valid = false; s->movers_fresh = false; s->count_valid = false; s->part_valid = false;
    return PICG_OK;
}

// Device-resident hand-over of particles (no host staging): the caller gets the seven SoA arrays of the store, writes n particles
// into them on the library's stream (a peer copy, a collective, its own kernel) and adopts them.  The pointers stay valid until the
// store grows (reserve first).  Used by bench.py's multi-GPU self-check: every rank's particles gathered onto one GPU over NCCL.
int picg_species_particle_arrays(picg_species_t s, size_t capacity, void* arrays7[7], size_t* cap_out) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && arrays7, "picg_species_particle_arrays: null argument");
    int rc = species_ensure_capacity(s, std::max<size_t>(capacity, 256)); if (rc) return rc;
    for (int c = 0; c < 7; c++) arrays7[c] = s->a[c];
    if (cap_out) *cap_out = s->cap;
    return PICG_OK;
}
int picg_species_adopt(picg_species_t s, size_t n) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && n <= s->cap, "picg_species_adopt: more particles than the store holds (picg_species_particle_arrays sizes it)");
    SpeciesCounters z; memset(&z, 0, sizeof(z)); z.n = n;
    *s->ctr_host = z;
    CUDA_TRY(cudaMemcpyAsync(s->ctr, s->ctr_host, sizeof(z), cudaMemcpyHostToDevice, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    s->n_host = n; s->n_host_valid = true; s->n_upper = n; s->sorted_valid = false; s->lists_valid = false; s->movers_fresh = false; s->count_valid = false; s->part_valid = false;
    return PICG_OK;
}

int picg_species_download(picg_species_t s, size_t capacity, double* aos7, size_t* n_out) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && n_out, "picg_species_download: null argument");
    int rc = species_refresh_count(s); if (rc) return rc;
    size_t n = s->n_host; *n_out = n;
    if (!aos7) return PICG_OK;
    REQUIRE_ARG(capacity >= n, "picg_species_download: host buffer too small");
    rc = ensure_scratch(s->w, std::min(std::max<size_t>(n, 1), kChunk) * 56 + 64); if (rc) return rc;
    for (size_t off = 0; off < n; off += kChunk) {
        size_t m = std::min(kChunk, n - off);
        LAUNCH(K_TRANSPOSE, k_soa_to_aos, std::min(div_up(m, 256), g_sm_count * 8), 256, 0, m, soa_of(s), off, (double*)s->w->scratch);
        CHECK_LAUNCH();
        CUDA_TRY(cudaMemcpyAsync(aos7 + off * 7, s->w->scratch, m * 56, cudaMemcpyDeviceToHost, g_stream));
        CUDA_TRY(cudaStreamSynchronize(g_stream));
    }
    return PICG_OK;
}

int picg_species_add_particles(picg_species_t s, size_t n, const double* aos7, size_t* accepted) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && (aos7 || n == 0), "picg_species_add_particles: null argument");
    int rc = species_refresh_count(s); if (rc) return rc;
    size_t before = s->n_host;
    rc = species_ensure_capacity(s, before + n); if (rc) return rc;
    rc = ensure_scratch(s->w, std::min(std::max<size_t>(n, 1), kChunk) * 56 + 64); if (rc) return rc;
    double q_over_m = s->charge / s->mass, half_dt = 0.5 * s->w->dt;
    for (size_t off = 0; off < n; off += kChunk) {
        size_t m = std::min(kChunk, n - off);
        CUDA_TRY(cudaMemcpyAsync(s->w->scratch, aos7 + off * 7, m * 56, cudaMemcpyHostToDevice, g_stream));
        rc = add_candidates(s, m, (const double*)s->w->scratch, q_over_m, half_dt); if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(g_stream));
    }
    s->n_host_valid = false; s->n_upper = before + n; s->sorted_valid = false; s->lists_valid = false; s->count_valid = false;
    rc = species_refresh_count(s);
    if (accepted) *accepted = s->n_host - before;
    return rc;
}

int picg_species_diagnostics(picg_species_t s, double* micro_count, double momentum[3], double* ke) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_diagnostics: null species");
    picg_world_s* w = s->w;
    const int grid = 512;
    LAUNCH(K_DIAG, k_diag, grid, 256, 0, s->ctr, soa_of(s), w->reduce_buf); CHECK_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(w->reduce_host, w->reduce_buf, grid * 5 * 8, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    double acc[5] = {0, 0, 0, 0, 0};
    for (int b = 0; b < grid; b++) for (int c = 0; c < 5; c++) acc[c] += w->reduce_host[b * 5 + c];
    if (micro_count) *micro_count = acc[0];
    if (momentum) { momentum[0] = acc[1] * s->mass; momentum[1] = acc[2] * s->mass; momentum[2] = acc[3] * s->mass; }
    if (ke) *ke = acc[4] * s->mass * 0.5;
    return PICG_OK;
}

int picg_species_update_averages(picg_species_t s) {            // Field::updateMovingAverage Field.h:246-261
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_update_averages: null species");
    s->avg_samples++;
    int clear = 0;
    if (s->avg_samples > 20) { clear = 1; s->avg_samples = 1; }
    int nv = s->w->g.nv;
    LAUNCH(K_MISC, k_moving_avg, std::min(div_up(nv, 256), g_sm_count * 8), 256, 0, nv, s->den, s->den_avg, (double)s->avg_samples, clear);
    CHECK_LAUNCH();
    return PICG_OK;
}

static int species_field(picg_species_t s, int field, void** p, size_t* bytes) {
    size_t nv = s->w->g.nv, nc = s->w->g.nc;
    switch (field) {
        case PICG_SF_DEN: *p = s->den; *bytes = nv * 8; break;
        case PICG_SF_DEN_AVG: *p = s->den_avg; *bytes = nv * 8; break;
        case PICG_SF_T: *p = s->T; *bytes = nv * 8; break;
        case PICG_SF_VEL: *p = s->vel; *bytes = nv * 24; break;
        case PICG_SF_MACRO_COUNT: *p = s->macro_count; *bytes = nc * 8; break;
        case PICG_SF_N_SUM: *p = s->n_sum; *bytes = nv * 8; break;
        case PICG_SF_NV_SUM: *p = s->nv_sum; *bytes = nv * 24; break;
        case PICG_SF_NUU_SUM: *p = s->nuu; *bytes = nv * 8; break;
        case PICG_SF_NVV_SUM: *p = s->nvv; *bytes = nv * 8; break;
        case PICG_SF_NWW_SUM: *p = s->nww; *bytes = nv * 8; break;
        case PICG_SF_DEN_FIXED: *p = s->den_fixed; *bytes = nv * 8; break;
        default: return set_error(PICG_ERR_ARG, "unknown species field %d", field);
    }
    return PICG_OK;
}

int picg_species_download_field(picg_species_t s, int field, void* host) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && host, "picg_species_download_field: null argument");
    void* p; size_t bytes; int rc = species_field(s, field, &p, &bytes); if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(host, p, bytes, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return PICG_OK;
}

int picg_species_device_ptr(picg_species_t s, int field, void** dptr, size_t* bytes) {
    REQUIRE_ARG(s && dptr, "picg_species_device_ptr: null argument");
    size_t b; int rc;
    if (field >= 100 && field < 107) { *dptr = s->a[field - 100]; b = s->cap * 8; rc = PICG_OK; }   // 100..106: particle arrays x y z u v w mpw
    else if (field == 107) { *dptr = s->ctr; b = sizeof(SpeciesCounters); rc = PICG_OK; }
    else rc = species_field(s, field, dptr, &b);
    if (bytes) *bytes = b;
    return rc;
}

}  // extern "C"
