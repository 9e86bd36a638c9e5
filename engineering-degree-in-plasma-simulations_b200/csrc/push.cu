// Particle push kernels + in-place removal of dead particles.
//   (the electron / heavy pushes are k_step in step.cu)
//   k_push_reflect   : ch2 Species::advance (specular walls)     ch2/v2/Species.cpp:18-55
//   compaction       : the swap-with-last removal :378-398 done as hole filling: survivors from the
//                      tail [n_alive,n) move into the holes left below n_alive, so the traffic is
//                      proportional to the number of dead particles, not to n.
// Algorithmic bytes per particle-step (fp64 SoA): 48 B read (pos, vel) + 48 B written = 96 B.
#include "common.cuh"
#include "push.cuh"
#include <algorithm>

using namespace picg;

// ---------------------------------------------------------------- ch2 reflective push
__global__ void __launch_bounds__(256) k_push_reflect(Grid g, PushArrays s, const SpeciesCounters* ctr, const double* __restrict__ ef,
                                                      double qm_dt, double dt) {
    const u64 n = ctr->n;
    for (u64 p = blockIdx.x * (u64)blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x) {
        double q[3] = {s.x[p], s.y[p], s.z[p]}, v[3] = {s.u[p], s.v[p], s.w[p]};
        push_kick_drift(g, ef, qm_dt, dt, q[0], q[1], q[2], v[0], v[1], v[2]);
#pragma unroll
        for (int a = 0; a < 3; a++) {                                              // ch2/v2/Species.cpp:44-53
            if (q[a] < g.x0[a]) { q[a] = __dsub_rn(__dmul_rn(2.0, g.x0[a]), q[a]); v[a] = -v[a]; }
            else if (q[a] >= g.xm[a]) { q[a] = __dsub_rn(__dmul_rn(2.0, g.xm[a]), q[a]); v[a] = -v[a]; }
        }
        s.x[p] = q[0]; s.y[p] = q[1]; s.z[p] = q[2]; s.u[p] = v[0]; s.v[p] = v[1]; s.w[p] = v[2];
    }
}

// ---------------------------------------------------------------- compaction (hole filling)
// The dead list arrives in the order the warps reached its cursor.  The plan below does not depend on that order: the dead slots are
// marked in a bitmap over the store and read back in ascending order (D); the first h of them lie below n_alive = n - n_dead and are
// the holes, the rest are dead slots of the tail [n_alive, n), and the t-th surviving tail slot (found by bisection on D, see
// compact_survivor in push.cuh) moves into the t-th hole.  Same input, same store - bit for bit, whatever the atomics did.
// scratch layout: dead_list[cap] | D[cap] | block counts (1024 words);  the bitmap lives in the world (one bit per store slot)
#define BM_BLOCKS 1024
#define BM_THREADS 256
__global__ void __launch_bounds__(256) k_bm_mark(const u64* __restrict__ count_ptr, const unsigned* __restrict__ dead_list, unsigned* __restrict__ bitmap) {
    const u64 nd = *count_ptr;
    for (u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x; t < nd; t += (u64)gridDim.x * blockDim.x) { const unsigned idx = dead_list[t]; atomicOr(&bitmap[idx >> 5], 1u << (idx & 31)); }
}
__device__ __forceinline__ void bm_range(u64 n, u64& w0, u64& w1) {             // words of block blockIdx.x (contiguous, ascending with the block index)
    const u64 W = (n + 31) >> 5, per = (W + gridDim.x - 1) / gridDim.x;
    w0 = min(W, (u64)blockIdx.x * per); w1 = min(W, w0 + per);
}
__global__ void __launch_bounds__(BM_THREADS) k_bm_count(const u64* __restrict__ count_ptr, const u64* __restrict__ n_ptr, const unsigned* __restrict__ bitmap, unsigned* __restrict__ counts) {
    __shared__ unsigned ws[BM_THREADS / 32];
    u64 w0, w1; bm_range(*count_ptr ? *n_ptr : 0, w0, w1);
    unsigned c = 0;
    for (u64 w = w0 + threadIdx.x; w < w1; w += BM_THREADS) c += __popc(bitmap[w]);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned t = 0; for (int q = 0; q < BM_THREADS / 32; q++) t += ws[q]; counts[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(BM_BLOCKS) k_bm_scan(unsigned* __restrict__ counts) {          // exclusive scan of BM_BLOCKS counts, one block
    __shared__ unsigned ws[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned v = counts[threadIdx.x]; unsigned x = v;
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    if (warp == 0) { unsigned wv = ws[lane], wx = wv; for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, wx, o); if (lane >= o) wx += y; } ws[lane] = wx - wv; }
    __syncthreads();
    counts[threadIdx.x] = ws[warp] + x - v;
}
// set bits of the block's words in ascending order -> D; the words are cleared on the way (the bitmap is all zero between compactions)
__global__ void __launch_bounds__(BM_THREADS) k_bm_emit(const u64* __restrict__ count_ptr, const u64* __restrict__ n_ptr, unsigned* __restrict__ bitmap, const unsigned* __restrict__ offsets, unsigned* __restrict__ D) {
    __shared__ unsigned ws[BM_THREADS / 32]; __shared__ unsigned running;
    u64 w0, w1; bm_range(*count_ptr ? *n_ptr : 0, w0, w1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = offsets[blockIdx.x];
    __syncthreads();
    for (u64 base = w0; base < w1; base += BM_THREADS) {
        const u64 w = base + threadIdx.x;
        unsigned bits = w < w1 ? bitmap[w] : 0u;
        if (bits) bitmap[w] = 0u;
        const unsigned c = __popc(bits); unsigned x = c;
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) ws[warp] = x;
        __syncthreads();
        unsigned before = 0, total = 0;
        for (int q = 0; q < BM_THREADS / 32; q++) { const unsigned t = ws[q]; if (q < warp) before += t; total += t; }
        unsigned pos = running + before + x - c;
        while (bits) { const int b = __ffs(bits) - 1; bits &= bits - 1; D[pos++] = (unsigned)(w * 32 + b); }
        __syncthreads();
        if (threadIdx.x == 0) running += total;
        __syncthreads();
    }
}
// h = number of dead slots below n_alive (D is ascending)
__global__ void k_compact_split(SpeciesCounters* ctr, const unsigned* __restrict__ D) {
    const u64 nd = ctr->n_dead, n_alive = ctr->n - nd;
    u64 lo = 0, hi = nd;
    while (lo < hi) { const u64 mid = (lo + hi) >> 1; if ((u64)D[mid] < n_alive) lo = mid + 1; else hi = mid; }
    ctr->n_hole = lo;
}
__global__ void k_compact_move(const SpeciesCounters* ctr, PushArrays s, double* __restrict__ mpw, const unsigned* __restrict__ D) {
    const u64 nh = ctr->n_hole, nd = ctr->n_dead, n_alive = ctr->n - nd;
    for (u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x; t < nh; t += (u64)gridDim.x * blockDim.x) {
        const unsigned d = D[t], f = compact_survivor(D, nh, nd, n_alive, t);
        s.x[d] = s.x[f]; s.y[d] = s.y[f]; s.z[d] = s.z[f]; s.u[d] = s.u[f]; s.v[d] = s.v[f]; s.w[d] = s.w[f]; mpw[d] = mpw[f];
    }
}
__global__ void k_compact_finish(SpeciesCounters* ctr) {
    ctr->n -= ctr->n_dead; ctr->n_dead = 0; ctr->n_hole = 0; ctr->n_surv = 0;
}

namespace picg {
// `list` (count on the device, distinct slots of a store with *n_ptr slots) in ascending order -> `out` (may be `list` itself: the bitmap
// holds the set in between).  counts: BM_BLOCKS words of scratch.
static int ensure_bitmap(picg_world_s* w, size_t store_cap) {
    const size_t words = (store_cap + 31) / 32 + 32;
    if (w->cbm_words >= words) return PICG_OK;
    if (w->cbm) { cudaStreamSynchronize(g_stream); cudaFree(w->cbm); w->cbm = nullptr; w->cbm_words = 0; }
    const size_t want = words + words / 2;
    cudaError_t e = cudaMalloc(&w->cbm, want * 4);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(slot bitmap)", __FILE__, __LINE__);
    CUDA_TRY(cudaMemsetAsync(w->cbm, 0, want * 4, g_stream));
    w->cbm_words = want; note_realloc("slot bitmap", want * 4);
    return PICG_OK;
}
int sort_slot_list(picg_world_s* w, const u64* count_ptr, const u64* n_ptr, size_t list_cap, size_t store_cap, const unsigned* list, unsigned* out, unsigned* counts) {
    int rc = ensure_bitmap(w, store_cap); if (rc) return rc;
    int grid = std::max(1, std::min(div_up(std::max<size_t>(list_cap / 16, 1), 256), g_sm_count * 4));
    LAUNCH(K_COMPACT, k_bm_mark, grid, 256, 0, count_ptr, list, w->cbm); CHECK_LAUNCH();
    LAUNCH(K_COMPACT, k_bm_count, BM_BLOCKS, BM_THREADS, 0, count_ptr, n_ptr, w->cbm, counts); CHECK_LAUNCH();
    LAUNCH(K_COMPACT, k_bm_scan, 1, BM_BLOCKS, 0, counts); CHECK_LAUNCH();
    LAUNCH(K_COMPACT, k_bm_emit, BM_BLOCKS, BM_THREADS, 0, count_ptr, n_ptr, w->cbm, counts, out); CHECK_LAUNCH();
    return PICG_OK;
}
// The move plan of a compaction: D (ascending dead slots, in the scratch arena behind the dead list) and ctr->n_hole.  store_cap: slots
// the store can hold (sizes the bitmap).  All counts are read on the device.
int compact_plan(picg_world_s* w, SpeciesCounters* ctr, size_t cap, size_t store_cap, unsigned** D_out) {
    unsigned* dead_list = (unsigned*)w->scratch;
    unsigned* D = dead_list + cap;
    unsigned* counts = D + cap;
    int rc = sort_slot_list(w, &ctr->n_dead, &ctr->n, cap, store_cap, dead_list, D, counts); if (rc) return rc;
    LAUNCH(K_COMPACT, k_compact_split, 1, 1, 0, ctr, D); CHECK_LAUNCH();
    *D_out = D;
    return PICG_OK;
}
// Removes the particles recorded in the dead list of `s` (scratch layout above).  All sizes are read on the device.
int compact_dead(picg_species_s* s, size_t cap) {
    unsigned* D = nullptr;
    int rc = compact_plan(s->w, s->ctr, cap, s->cap, &D); if (rc) return rc;
    PushArrays a = {s->a[0], s->a[1], s->a[2], s->a[3], s->a[4], s->a[5]};
    int grid = std::max(1, std::min(div_up(std::max<size_t>(cap / 16, 1), 256), g_sm_count * 4));
    LAUNCH(K_COMPACT, k_compact_move, grid, 256, 0, s->ctr, a, s->a[6], D); CHECK_LAUNCH();
    LAUNCH(K_COMPACT, k_compact_finish, 1, 1, 0, s->ctr); CHECK_LAUNCH();
    s->n_host_valid = false;       // count changed on the device; n_upper stays an upper bound
    s->sorted_valid = false; s->lists_valid = false; s->movers_fresh = false; s->count_valid = false;
    return PICG_OK;
}
size_t compact_scratch_bytes(size_t cap) { return ((cap * 8 + BM_BLOCKS * 4 + 64) + 255) & ~(size_t)255; }    // 256-byte multiple: what follows stays aligned
int push_grid(size_t n_upper) { return std::max(1, std::min(div_up(std::max<size_t>(n_upper, 1), 256), g_sm_count * 8)); }
}  // namespace picg

extern "C" {

int picg_species_push_reflect(picg_species_t s, double dt) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_push_reflect: null species");
    double qm_dt = dt * s->charge / s->mass;
    PushArrays a = {s->a[0], s->a[1], s->a[2], s->a[3], s->a[4], s->a[5]};
    LAUNCH(K_PUSH_REFLECT, k_push_reflect, push_grid(s->n_upper), 256, 0, s->w->g, a, s->ctr, s->w->ef, qm_dt, dt);
    CHECK_LAUNCH();
    s->sorted_valid = false; s->lists_valid = false; s->movers_fresh = false; s->count_valid = false;
    return PICG_OK;
}

}  // extern "C"
