// Cell sort of the SoA particle store: stable LSD radix sort (8-bit digits) on 32-bit cell keys.
// Replaces Species::sortIndexes (ch4/v3/src/Species.cpp:905-929): instead of per-cell index vectors
// the particles themselves are permuted into cell order and cell_start[] (nc+1 offsets) is produced.
//
// Pass structure (classic three-kernel radix pass with a fixed number of blocks):
//   k_sort_upsweep   : per-block digit histogram over the block's contiguous key range
//   k_sort_scan      : exclusive scan of the [256][blocks] count table (digit-major)
//   k_sort_downsweep : each block walks its range tile by tile, ranks keys stably (warp match-any
//                      multisplit + per-warp counters) and scatters (key, index) pairs
// followed by k_sort_permute (gather each particle array through the final index list) and
// k_cell_start.  ceil(log2(num_cells)/8) passes: 3 for the 256^3 mesh.
#include "common.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>

using namespace picg;

#define SORT_THREADS 256
#define SORT_WARPS (SORT_THREADS / 32)
#define SORT_ITEMS 8                               // keys per thread per tile
#define SORT_TILE (SORT_THREADS * SORT_ITEMS)      // 2048 keys per tile

__global__ void __launch_bounds__(256) k_sort_keys(Grid g, const double* __restrict__ px, const double* __restrict__ py, const double* __restrict__ pz,
                                                   const SpeciesCounters* ctr, unsigned* __restrict__ keys, unsigned* __restrict__ idx) {
    const u64 n = ctr->n;
    for (u64 p = blockIdx.x * (u64)blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x) {
        int i = min(max((int)x_to_l(px[p], g.x0[0], g.inv_dx[0]), 0), g.ci - 1);
        int j = min(max((int)x_to_l(py[p], g.x0[1], g.inv_dx[1]), 0), g.cj - 1);
        int k = min(max((int)x_to_l(pz[p], g.x0[2], g.inv_dx[2]), 0), g.ck - 1);
        keys[p] = (unsigned)cell_of(g, i, j, k);
        idx[p] = (unsigned)p;
    }
}

// block b owns tiles [b*tiles_per_block, (b+1)*tiles_per_block)
__device__ __forceinline__ void block_range(u64 n, int nblocks, u64& begin, u64& end) {
    u64 tiles = (n + SORT_TILE - 1) / SORT_TILE;
    u64 per = (tiles + nblocks - 1) / nblocks;
    begin = min(n, (u64)blockIdx.x * per * SORT_TILE);
    end = min(n, begin + per * SORT_TILE);
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_upsweep(const u64* __restrict__ n_ptr, const unsigned* __restrict__ keys, int shift,
                                                               unsigned* __restrict__ counts /*[256][gridDim.x]*/) {
    __shared__ unsigned hist[SORT_WARPS][256];
    for (int t = threadIdx.x; t < SORT_WARPS * 256; t += SORT_THREADS) (&hist[0][0])[t] = 0;
    __syncthreads();
    u64 begin, end; block_range(*n_ptr, gridDim.x, begin, end);
    const int warp = threadIdx.x >> 5;
    for (u64 p = begin + threadIdx.x; p < end; p += SORT_THREADS) atomicAdd(&hist[warp][(keys[p] >> shift) & 255u], 1u);
    __syncthreads();
    for (int d = threadIdx.x; d < 256; d += SORT_THREADS) {
        unsigned s = 0;
        for (int w = 0; w < SORT_WARPS; w++) s += hist[w][d];
        counts[(size_t)d * gridDim.x + blockIdx.x] = s;
    }
}

// Exclusive scan of the [256][blocks] count table (digit-major), blocks <= 1024.  One thread block per digit scans its row and
// publishes the row total; the second kernel adds the exclusive prefix of the totals of the lower digits.
__global__ void __launch_bounds__(1024) k_sort_scan_rows(unsigned* __restrict__ counts, int nblocks, unsigned* __restrict__ totals) {
    __shared__ unsigned warp_tot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned* row = counts + (size_t)blockIdx.x * nblocks;
    const unsigned v = (int)threadIdx.x < nblocks ? row[threadIdx.x] : 0;
    unsigned x = v;
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned wv = warp_tot[lane], wx = wv;
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, wx, o); if (lane >= o) wx += y; }
        warp_tot[lane] = wx - wv;
        if (lane == 31) totals[blockIdx.x] = wx;
    }
    __syncthreads();
    if ((int)threadIdx.x < nblocks) row[threadIdx.x] = warp_tot[warp] + x - v;
}
__global__ void __launch_bounds__(1024) k_sort_scan_add(unsigned* __restrict__ counts, int nblocks, const unsigned* __restrict__ totals) {
    __shared__ unsigned base;
    if (threadIdx.x < 32) {                                   // sum of the totals of the digits below this one
        unsigned s = 0;
        for (int d = threadIdx.x; d < (int)blockIdx.x; d += 32) s += totals[d];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) base = s;
    }
    __syncthreads();
    if ((int)threadIdx.x < nblocks) counts[(size_t)blockIdx.x * nblocks + threadIdx.x] += base;
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_downsweep(const u64* __restrict__ n_ptr, const unsigned* __restrict__ keys_in,
                                                                 const unsigned* __restrict__ idx_in, unsigned* __restrict__ keys_out,
                                                                 unsigned* __restrict__ idx_out, int shift, const unsigned* __restrict__ counts) {
    __shared__ unsigned wcount[SORT_WARPS][256];     // per-warp digit counts of the current tile, then exclusive warp offsets
    __shared__ unsigned running[256];                // block's running global offset per digit
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = threadIdx.x; d < 256; d += SORT_THREADS) running[d] = counts[(size_t)d * gridDim.x + blockIdx.x];
    u64 begin, end; block_range(*n_ptr, gridDim.x, begin, end);
    for (u64 tile = begin; tile < end; tile += SORT_TILE) {
        for (int t = threadIdx.x; t < SORT_WARPS * 256; t += SORT_THREADS) (&wcount[0][0])[t] = 0;
        __syncthreads();
        // warp w owns the contiguous segment [tile + w*32*ITEMS, +32*ITEMS); rounds are visited in order => stable
        unsigned key[SORT_ITEMS], val[SORT_ITEMS], rank[SORT_ITEMS];
        u64 seg = tile + (u64)warp * 32 * SORT_ITEMS;
#pragma unroll
        for (int r = 0; r < SORT_ITEMS; r++) {
            u64 p = seg + r * 32 + lane;
            bool ok = p < end;
            key[r] = ok ? keys_in[p] : 0xffffffffu; val[r] = ok ? idx_in[p] : 0;
            unsigned d = ok ? ((key[r] >> shift) & 255u) : 256u;          // 256 = inactive lanes group together
            unsigned peers = __match_any_sync(0xffffffffu, d);
            unsigned before = 0;
            if (ok) before = wcount[warp][d];
            __syncwarp();
            if (ok && lane == __ffs(peers) - 1) wcount[warp][d] = before + __popc(peers);
            __syncwarp();
            rank[r] = before + __popc(peers & ((1u << lane) - 1));
        }
        __syncthreads();
        // exclusive scan over warps for every digit; add the tile totals to the running offsets afterwards
        for (int d = threadIdx.x; d < 256; d += SORT_THREADS) {
            unsigned s = 0;
            for (int w = 0; w < SORT_WARPS; w++) { unsigned c = wcount[w][d]; wcount[w][d] = s; s += c; }
            unsigned base = running[d];
            for (int w = 0; w < SORT_WARPS; w++) wcount[w][d] += base;
            running[d] = base + s;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < SORT_ITEMS; r++) {
            u64 p = seg + r * 32 + lane;
            if (p < end) {
                unsigned d = (key[r] >> shift) & 255u;
                unsigned dst = wcount[warp][d] + rank[r];
                keys_out[dst] = key[r]; idx_out[dst] = val[r];
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_sort_permute(const SpeciesCounters* ctr, const unsigned* __restrict__ idx, const double* __restrict__ in,
                                                      double* __restrict__ out) {
    const u64 n = ctr->n;
    for (u64 p = blockIdx.x * (u64)blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x) out[p] = in[idx[p]];
}

// cell_start[c] = first sorted position whose key >= c ; cell_start[nc] = n
__global__ void __launch_bounds__(256) k_cell_start(const u64* __restrict__ n_ptr, const unsigned* __restrict__ keys, int nc, unsigned* __restrict__ cell_start) {
    const u64 n = *n_ptr;
    for (u64 p = blockIdx.x * (u64)blockDim.x + threadIdx.x; p <= n; p += (u64)gridDim.x * blockDim.x) {
        int lo = (p == 0) ? 0 : (int)keys[p - 1] + 1;
        int hi = (p == n) ? nc : (int)keys[p];
        for (int c = lo; c <= hi; c++) cell_start[c] = (unsigned)p;
    }
}

// ---------------------------------------------------------------- movers (exact cell lists from a stale partition)
// After a sort every slot has a home cell (home[p]).  Later pushes / appends / hole filling make some slots hold a
// particle whose current cell differs from the slot's home: the movers.  k_find_movers lists them; sorted once by their
// current cell and once by their home cell they turn the stale partition into exact per-cell lists for MC collisions:
//   list(c) = home range of c  minus  movers whose home is c  plus  movers whose current cell is c.
__global__ void __launch_bounds__(256) k_find_movers(Grid g, const double* __restrict__ px, const double* __restrict__ py, const double* __restrict__ pz,
                                                     const unsigned* __restrict__ home, const u64* __restrict__ n_ptr, const unsigned* __restrict__ part_n_ptr,
                                                     const u64* __restrict__ listed_ptr, u64* __restrict__ count, u64 cap, unsigned* __restrict__ m_slot,
                                                     unsigned* __restrict__ m_cell, unsigned* __restrict__ m_home) {
    const u64 n = *n_ptr, part_n = *part_n_ptr;
    const u64 first = listed_ptr ? min(*listed_ptr, n) : 0;                  // slots below are already listed (by the last deposit pass)
    const int lane = threadIdx.x & 31;
    for (u64 p0 = (first & ~(u64)31) + (blockIdx.x * (u64)blockDim.x + threadIdx.x) - lane; p0 < n; p0 += (u64)gridDim.x * blockDim.x) {
        u64 p = p0 + lane;
        bool mover = false; unsigned cell = 0, hm = (unsigned)g.nc;          // home == nc: no home (appended after the sort)
        if (p >= first && p < n) {
            int i = min(max((int)x_to_l(px[p], g.x0[0], g.inv_dx[0]), 0), g.ci - 1);
            int j = min(max((int)x_to_l(py[p], g.x0[1], g.inv_dx[1]), 0), g.cj - 1);
            int k = min(max((int)x_to_l(pz[p], g.x0[2], g.inv_dx[2]), 0), g.ck - 1);
            cell = (unsigned)cell_of(g, i, j, k);
            if (p < part_n) { hm = home[p]; mover = hm != cell; } else mover = true;       // appended after the sort: no home
        }
        unsigned mask = __ballot_sync(0xffffffffu, mover);
        if (mask) {
            u64 base = 0; int leader = __ffs(mask) - 1;
            if (lane == leader) base = atomicAdd(count, (u64)__popc(mask));
            base = __shfl_sync(0xffffffffu, base, leader);
            u64 dst = base + __popc(mask & ((1u << lane) - 1));
            if (mover && dst < cap) { m_slot[dst] = (unsigned)p; m_cell[dst] = cell; m_home[dst] = hm; }
        }
    }
}
// slots of the partition that lie beyond the live count (particles died since the sort) also leave their home lists
__global__ void __launch_bounds__(256) k_find_vacated(const unsigned* __restrict__ home, const u64* __restrict__ n_ptr, const unsigned* __restrict__ part_n_ptr,
                                                      u64* __restrict__ count, u64 cap, unsigned* __restrict__ m_slot, unsigned* __restrict__ m_home) {
    const u64 n = *n_ptr, part_n = *part_n_ptr;
    for (u64 p = n + blockIdx.x * (u64)blockDim.x + threadIdx.x; p < part_n; p += (u64)gridDim.x * blockDim.x) {
        u64 dst = atomicAdd(count, 1ull);
        if (dst < cap) { m_slot[dst] = (unsigned)p; m_home[dst] = home[p]; }
    }
}
__global__ void k_copy_u32(const u64* __restrict__ n_ptr, const unsigned* __restrict__ in, unsigned* __restrict__ out) {
    const u64 n = *n_ptr;
    for (u64 p = blockIdx.x * (u64)blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x) out[p] = in[p];
}
__global__ void k_iota_u32(const u64* __restrict__ n_ptr, unsigned* __restrict__ out) {
    const u64 n = *n_ptr;
    for (u64 p = blockIdx.x * (u64)blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x) out[p] = (unsigned)p;
}
__global__ void k_gather_u32(const u64* __restrict__ n_ptr, const unsigned* __restrict__ idx, const unsigned* __restrict__ in, unsigned* __restrict__ out) {
    const u64 n = *n_ptr;
    for (u64 p = blockIdx.x * (u64)blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x) out[p] = in[idx[p]];
}
__global__ void k_clamp_u64(u64* v, u64 cap) { if (*v > cap) *v = cap; }
// The deposit pass left n_movers = the movers of the partition.  A list build appends the tail and the vacated slots behind them;
// the next build (no push or deposit in between: neutrals of the subcycled loop, products appended by a collision call) must start
// again from the deposit's count, not from the inflated one.
__global__ void k_movers_checkpoint(SpeciesCounters* ctr, int restore) { if (restore) ctr->n_movers = ctr->n_movers_dep; else ctr->n_movers_dep = ctr->n_movers; }

// The same table for a SPARSE sorted key list (movers).  A key at position p owns the cells (previous key, its key]: short gaps are
// filled by the key's thread, long ones (a few: at most nc/32) are queued and filled by a warp each.  nc + 1 writes in total.
__global__ void __launch_bounds__(256) k_cell_start_sparse(const u64* __restrict__ n_ptr, const unsigned* __restrict__ keys, int nc, unsigned* __restrict__ cell_start,
                                                           unsigned* __restrict__ queue /* [0]: count, then (lo, hi, p) triples */) {
    const u64 n = *n_ptr;
    for (u64 p = blockIdx.x * (u64)blockDim.x + threadIdx.x; p <= n; p += (u64)gridDim.x * blockDim.x) {
        const unsigned prev = (p == 0) ? 0u : keys[p - 1];
        if (p > 0 && prev >= (unsigned)nc) continue;                                      // everything up to nc is owned by an earlier position
        const int lo = (p == 0) ? 0 : (int)prev + 1;
        const int hi = (p == n) ? nc : (int)min(keys[p], (unsigned)nc);                   // keys >= nc (movers without a home) sort last
        if (hi - lo < 32) { for (int c = lo; c <= hi; c++) cell_start[c] = (unsigned)p; }
        else { unsigned q = atomicAdd(queue, 1u); queue[1 + 3 * q] = (unsigned)lo; queue[2 + 3 * q] = (unsigned)hi; queue[3 + 3 * q] = (unsigned)p; }
    }
}
__global__ void __launch_bounds__(256) k_cell_start_long(const unsigned* __restrict__ queue, unsigned* __restrict__ cell_start) {
    const unsigned nq = queue[0];
    const int lane = threadIdx.x & 31;
    for (unsigned q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nq; q += (gridDim.x * blockDim.x) >> 5) {
        const unsigned lo = queue[1 + 3 * q], hi = queue[2 + 3 * q], p = queue[3 + 3 * q];
        for (unsigned c = lo + lane; c <= hi; c += 32) cell_start[c] = p;
    }
}

// ---------------------------------------------------------------- counting sort by cell (short lists: movers, appended tails)
// The lists that patch the stale partition hold a few entries per cell.  Their keys are cells, and a table of one counter per cell
// (66 MB at 256^3) lives in the L2: count the keys with atomics, scan the table, hand out positions with a second round of atomics,
// then order each cell's (tiny) segment by value so that the result does not depend on the order of the atomics.  Keys >= nc are
// dropped.  The scanned table IS the per-cell offset table the list users need (nc + 1 entries, last = total).
#define SCAN_T 1024
#define SCAN_I 4
__global__ void __launch_bounds__(256) k_count_keys(const u64* __restrict__ n_ptr, const unsigned* __restrict__ keys, unsigned nc, unsigned* __restrict__ cnt) {
    const u64 n = *n_ptr;
    for (u64 e = blockIdx.x * (u64)blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x) { const unsigned k = keys[e]; if (k < nc) atomicAdd(&cnt[k], 1u); }
}
__global__ void __launch_bounds__(SCAN_T) k_scan_reduce(const unsigned* __restrict__ in, size_t n, unsigned* __restrict__ sums) {
    __shared__ unsigned ws[32];
    const size_t base = ((size_t)blockIdx.x * SCAN_T + threadIdx.x) * SCAN_I;
    unsigned v = 0;
#pragma unroll
    for (int q = 0; q < SCAN_I; q++) if (base + q < n) v += in[base + q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned t = ws[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) sums[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(1024) k_scan_sums(unsigned* __restrict__ sums, int nb) {      // exclusive scan in place, one block
    __shared__ unsigned ws[32]; __shared__ unsigned carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned v = i < nb ? sums[i] : 0;
        unsigned x = v;
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) ws[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned wv = ws[lane], wx = wv;
            for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, wx, o); if (lane >= o) wx += y; }
            ws[lane] = wx - wv;
        }
        __syncthreads();
        const unsigned excl = carry + ws[warp] + x - v;
        if (i < nb) sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(SCAN_T) k_scan_apply(unsigned* __restrict__ data, size_t n, const unsigned* __restrict__ sums, unsigned* __restrict__ copy) {
    __shared__ unsigned ws[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t base = ((size_t)blockIdx.x * SCAN_T + threadIdx.x) * SCAN_I;
    unsigned v[SCAN_I], tot = 0;
#pragma unroll
    for (int q = 0; q < SCAN_I; q++) { v[q] = base + q < n ? data[base + q] : 0; tot += v[q]; }
    unsigned x = tot;
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned wv = ws[lane], wx = wv;
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, wx, o); if (lane >= o) wx += y; }
        ws[lane] = wx - wv;
    }
    __syncthreads();
    unsigned run = sums[blockIdx.x] + ws[warp] + x - tot;
#pragma unroll
    for (int q = 0; q < SCAN_I; q++) { if (base + q < n) { data[base + q] = run; if (copy) copy[base + q] = run; } run += v[q]; }
}
__global__ void __launch_bounds__(256) k_fill_by_key(const u64* __restrict__ n_ptr, const unsigned* __restrict__ keys, const unsigned* __restrict__ vals, unsigned nc,
                                                     unsigned* __restrict__ cursor, unsigned* __restrict__ out_vals, unsigned* __restrict__ out_keys) {
    const u64 n = *n_ptr;
    for (u64 e = blockIdx.x * (u64)blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x) {
        const unsigned k = keys[e];
        if (k >= nc) continue;
        const unsigned pos = atomicAdd(&cursor[k], 1u);
        out_vals[pos] = vals[e];
        if (out_keys) out_keys[pos] = k;
    }
}
// each cell's segment in ascending order of its values (slots): the order no longer depends on the atomics above
__global__ void __launch_bounds__(256) k_sort_segments(unsigned nc, const unsigned* __restrict__ start, unsigned* __restrict__ vals) {
    for (unsigned c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
        const unsigned a = start[c], b = start[c + 1];
        for (unsigned i = a + 1; i < b; i++) {
            const unsigned v = vals[i]; unsigned j = i;
            while (j > a && vals[j - 1] > v) { vals[j] = vals[j - 1]; j--; }
            vals[j] = v;
        }
    }
}

// ---------------------------------------------------------------- merge of the appended tail into the cell partition
// The store is [partition | tail]: slots [0, part_n) ordered by home cell (cell_start of the last sort), the particles appended
// since then behind them in arrival order.  Sorting only the tail (T << n keys) and opening gaps in the partition puts every tail
// particle at the end of its current cell's range: slot p of the partition moves to p + tstart[home[p]], the r-th tail particle
// in cell order (cell c) to cell_start[c+1] + r, with tstart[c] = tail particles in cells below c.  Sources and destinations
// are two monotone streams, so the pass runs at copy speed - unlike the gather of a full re-sort - and needs no keys of the
// partition at all.  Home cells of the partition slots do not change (drifted particles stay movers).
__global__ void __launch_bounds__(256) k_tail_keys(Grid g, const double* __restrict__ px, const double* __restrict__ py, const double* __restrict__ pz,
                                                   const u64* __restrict__ n_ptr, const unsigned* __restrict__ part_n_ptr, u64* __restrict__ t_count,
                                                   unsigned* __restrict__ keys, unsigned* __restrict__ idx) {
    const u64 n = *n_ptr, first = *part_n_ptr;
    const u64 T = n > first ? n - first : 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) *t_count = T;
    for (u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x; r < T; r += (u64)gridDim.x * blockDim.x) {
        const u64 p = first + r;
        int i = min(max((int)x_to_l(px[p], g.x0[0], g.inv_dx[0]), 0), g.ci - 1);
        int j = min(max((int)x_to_l(py[p], g.x0[1], g.inv_dx[1]), 0), g.cj - 1);
        int k = min(max((int)x_to_l(pz[p], g.x0[2], g.inv_dx[2]), 0), g.ck - 1);
        keys[r] = (unsigned)cell_of(g, i, j, k);
        idx[r] = (unsigned)p;
    }
}
// source slot of every slot of the merged store (one scattered 4-byte pass; the seven particle arrays and the home array then
// follow with aligned, coalesced writes through k_sort_permute / k_gather_u32: scattered 8-byte writes with a gap per cell ran at half speed)
__global__ void __launch_bounds__(256) k_merge_sources(const unsigned* __restrict__ part_n_ptr, const u64* __restrict__ t_count, const unsigned* __restrict__ home,
                                                       const unsigned* __restrict__ cs, const unsigned* __restrict__ tstart, const unsigned* __restrict__ tkeys,
                                                       const unsigned* __restrict__ tidx, unsigned* __restrict__ src) {
    const u64 P = *part_n_ptr, T = *t_count;
    for (u64 p = blockIdx.x * (u64)blockDim.x + threadIdx.x; p < P + T; p += (u64)gridDim.x * blockDim.x) {
        if (p < P) src[p + tstart[home[p]]] = (unsigned)p;
        else { const u64 r = p - P; src[(u64)cs[tkeys[r] + 1] + r] = tidx[r]; }
    }
}
// home cell of the merged slots: unchanged for partition slots, the current cell for the merged tail particles
__global__ void __launch_bounds__(256) k_merge_home(const u64* __restrict__ n_ptr, const unsigned* __restrict__ part_n_ptr, const unsigned* __restrict__ src,
                                                    const unsigned* __restrict__ home, const unsigned* __restrict__ tail_cell /* indexed by slot - part_n */, unsigned* __restrict__ out) {
    const u64 n = *n_ptr, P = *part_n_ptr;
    for (u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x; q < n; q += (u64)gridDim.x * blockDim.x) { const unsigned sidx = src[q]; out[q] = sidx < P ? home[sidx] : tail_cell[sidx - P]; }
}
__global__ void __launch_bounds__(256) k_merge_cell_start(int nc, unsigned* __restrict__ cs, const unsigned* __restrict__ tstart) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= nc; c += gridDim.x * blockDim.x) cs[c] += tstart[c];
}
// the deposit pass listed the partition's movers by slot: the slots moved with the merge
__global__ void __launch_bounds__(256) k_merge_remap_movers(SpeciesCounters* ctr, int nc, const unsigned* __restrict__ tstart, unsigned* __restrict__ m_slot,
                                                            const unsigned* __restrict__ m_home, u64 n_new, u64 cap) {
    const u64 nm = min(ctr->n_movers, cap);                                               // (a list that overflowed its buffer leads to a full sort later)
    for (u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x; t < nm; t += (u64)gridDim.x * blockDim.x) {
        const unsigned h = m_home[t];
        if (h < (unsigned)nc) m_slot[t] += tstart[h];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ctr->n_listed = n_new;                       // merged tail particles sit in their current cell: none is a mover
}

namespace picg {
double g_mover_fraction = 0.15;
double g_merge_fraction = 0.12;      // a tail above this fraction of the store is merged into the partition (merge_tail)     // above this fraction of movers the store is re-sorted instead of patched
static uint64_t g_movers_from_deposit = 0, g_mover_scans = 0, g_mover_resorts = 0;
static bool trace_sort() { static const bool t = getenv("PICG_TRACE_SORT") && atoi(getenv("PICG_TRACE_SORT")) != 0; return t; }
// Stable LSD radix sort of (keys, vals) pairs; n is read on the device.  Returns the buffers that hold the result.
int radix_sort_pairs(const u64* n_ptr, size_t n_upper, int key_bits, unsigned*& keysA, unsigned*& valsA, unsigned*& keysB, unsigned*& valsB,
                     unsigned* counts, int nblocks) {
    int passes = (key_bits + 7) / 8;
    for (int pass = 0; pass < passes; pass++) {
        int shift = pass * 8;
        LAUNCH(K_SORT_HIST, k_sort_upsweep, nblocks, SORT_THREADS, 0, n_ptr, keysA, shift, counts); CHECK_LAUNCH();
        LAUNCH(K_SORT_SCAN, k_sort_scan_rows, 256, 1024, 0, counts, nblocks, counts + (size_t)256 * nblocks); CHECK_LAUNCH();
        LAUNCH(K_SORT_SCAN, k_sort_scan_add, 256, 1024, 0, counts, nblocks, counts + (size_t)256 * nblocks); CHECK_LAUNCH();
        LAUNCH(K_SORT_SCATTER, k_sort_downsweep, nblocks, SORT_THREADS, 0, n_ptr, keysA, valsA, keysB, valsB, shift, counts); CHECK_LAUNCH();
        std::swap(keysA, keysB); std::swap(valsA, valsB);
    }
    (void)n_upper;
    return PICG_OK;
}
static int key_bits_of(const Grid& g) { int bits = 1; while ((1ull << bits) < (u64)g.nc) bits++; return bits; }

// Counting sort of (keys, vals) by key in [0, nc): start[nc + 1] (zeroed here) becomes the per-cell offset table, out_vals the values in
// cell order (each cell's segment ascending), out_keys (optional) their keys.  work: (nc + 1) + 8192 words of scratch.
size_t counting_sort_words(const Grid& g) { return (((size_t)g.nc + 1 + 63) & ~(size_t)63) + 8192; }
// Exclusive scan in place of a per-cell table of nc + 1 entries (the last one, zero on entry, becomes the total).  work: 8192 words.
int scan_cell_table(const Grid& g, unsigned* table, unsigned* work) {
    const size_t nt = (size_t)g.nc + 1;
    const int nb = div_up(nt, (size_t)SCAN_T * SCAN_I);
    if (nb > 8192) return set_error(PICG_ERR_ARG, "scan_cell_table: more than 2^25 cells are not supported");
    LAUNCH(K_SORT_SCAN, k_scan_reduce, nb, SCAN_T, 0, table, nt, work); CHECK_LAUNCH();
    LAUNCH(K_SORT_SCAN, k_scan_sums, 1, 1024, 0, work, nb); CHECK_LAUNCH();
    LAUNCH(K_SORT_SCAN, k_scan_apply, nb, SCAN_T, 0, table, nt, work, (unsigned*)nullptr); CHECK_LAUNCH();
    return PICG_OK;
}
int counting_sort_by_cell(const Grid& g, const u64* n_ptr, size_t n_upper, const unsigned* keys, const unsigned* vals, unsigned* start, unsigned* out_vals,
                                 unsigned* out_keys, unsigned* work, bool ordered) {
    const size_t nt = (size_t)g.nc + 1;
    unsigned* cursor = work; unsigned* sums = work + ((nt + 63) & ~(size_t)63);
    const int nb = div_up(nt, (size_t)SCAN_T * SCAN_I);
    if (nb > 8192) return set_error(PICG_ERR_ARG, "counting_sort_by_cell: more than 2^25 cells are not supported");
    const int egrid = std::max(1, std::min(div_up(std::max<size_t>(n_upper, 1), 256), g_sm_count * 8));
    CUDA_TRY(cudaMemsetAsync(start, 0, nt * 4, g_stream));
    LAUNCH(K_SORT_HIST, k_count_keys, egrid, 256, 0, n_ptr, keys, (unsigned)g.nc, start); CHECK_LAUNCH();
    LAUNCH(K_SORT_SCAN, k_scan_reduce, nb, SCAN_T, 0, start, nt, sums); CHECK_LAUNCH();
    LAUNCH(K_SORT_SCAN, k_scan_sums, 1, 1024, 0, sums, nb); CHECK_LAUNCH();
    LAUNCH(K_SORT_SCAN, k_scan_apply, nb, SCAN_T, 0, start, nt, sums, cursor); CHECK_LAUNCH();
    LAUNCH(K_SORT_SCATTER, k_fill_by_key, egrid, 256, 0, n_ptr, keys, vals, (unsigned)g.nc, cursor, out_vals, out_keys); CHECK_LAUNCH();
    if (ordered) { LAUNCH(K_SORT_SCATTER, k_sort_segments, std::max(1, std::min(div_up((size_t)g.nc, 256), g_sm_count * 8)), 256, 0, (unsigned)g.nc, start, out_vals); CHECK_LAUNCH(); }
    return PICG_OK;
}

static int ensure_u32(unsigned*& p, size_t& cap, size_t want) {
    if (cap >= want) return PICG_OK;
    if (p) { cudaStreamSynchronize(g_stream); cudaFree(p); p = nullptr; cap = 0; }
    cudaError_t e = cudaMalloc(&p, want * 4);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(u32)", __FILE__, __LINE__);
    cap = want; note_realloc("u32 list buffer", want * 4); return PICG_OK;
}

// Sorts species s by cell.  Scratch: keysA | keysB | idxA | idxB | counts.
int sort_species(picg_species_s* s) {
    const Grid& g = s->w->g;
    int rc0 = species_refresh_count(s); if (rc0) return rc0;               // a sort is rare: the exact count makes part_n exact
    s->n_upper = s->n_host;
    if (trace_sort()) fprintf(stderr, "[picgpu] full sort of species %u: n = %zu (partition of the last sort: %zu)\n", s->id, s->n_host, s->part_valid ? s->part_n : (size_t)0);
    size_t cap = std::max<size_t>(s->n_upper, 1);
    if (cap >= 0xffffffffull) return set_error(PICG_ERR_ARG, "picg_species_sort: more than 2^32-1 particles per GPU are not supported");
    int nblocks = std::max(1, std::min(std::min(div_up(cap, SORT_TILE), g_sm_count * 4), 1024));
    size_t capa = (cap + 63) & ~(size_t)63;
    size_t bytes = capa * 16 + (size_t)256 * nblocks * 4 + 256 * 4 + 256;
    int rc = ensure_scratch(s->w, bytes); if (rc) return rc;
    rc = ensure_u32(s->home, s->home_cap, s->cap); if (rc) return rc;
    rc = ensure_u32(s->in_start, s->lists_cap, (size_t)g.nc + 1); if (rc) return rc;
    { size_t c2 = 0; if (!s->out_start) { rc = ensure_u32(s->out_start, c2, (size_t)g.nc + 1); if (rc) return rc; } }
    unsigned* keysA = (unsigned*)s->w->scratch; unsigned* keysB = keysA + capa;
    unsigned* idxA = keysB + capa; unsigned* idxB = idxA + capa;
    unsigned* counts = idxB + capa;
    const u64* n_ptr = &s->ctr->n;
    int pgrid = std::max(1, std::min(div_up(cap, 256), g_sm_count * 8));
    LAUNCH(K_SORT_KEYS, k_sort_keys, pgrid, 256, 0, g, s->a[0], s->a[1], s->a[2], s->ctr, keysA, idxA); CHECK_LAUNCH();
    rc = radix_sort_pairs(n_ptr, cap, key_bits_of(g), keysA, idxA, keysB, idxB, counts, nblocks); if (rc) return rc;
    // gather every particle array through the index list into the spare array, then rotate pointers
    for (int c = 0; c < 7; c++) {
        LAUNCH(K_SORT_PERMUTE, k_sort_permute, pgrid, 256, 0, s->ctr, idxA, s->a[c], s->spare); CHECK_LAUNCH();
        std::swap(s->a[c], s->spare);
    }
    LAUNCH(K_CELL_START, k_cell_start, pgrid, 256, 0, n_ptr, keysA, g.nc, s->cell_start); CHECK_LAUNCH();
    LAUNCH(K_CELL_START, k_copy_u32, pgrid, 256, 0, n_ptr, keysA, s->home); CHECK_LAUNCH();                 // home cell of every slot
    CUDA_TRY(cudaMemsetAsync(s->in_start, 0, ((size_t)g.nc + 1) * 4, g_stream));                           // no movers right after a sort
    CUDA_TRY(cudaMemsetAsync(s->out_start, 0, ((size_t)g.nc + 1) * 4, g_stream));
    s->sorted_valid = true; s->part_valid = true; s->part_n = cap; s->lists_valid = true; s->movers_fresh = false;
    return PICG_OK;
}

// Merges the appended tail [part_n, n) of s into its cell partition (see the kernels above).  n_host must be exact.
static uint64_t g_tail_merges = 0;
int merge_tail(picg_species_s* s) {
    const Grid& g = s->w->g;
    const size_t n = s->n_host;
    if (!s->part_valid || n <= s->part_n) return PICG_OK;
    const size_t T = n - s->part_n;
    if (trace_sort()) fprintf(stderr, "[picgpu] merge of the tail of species %u: n = %zu, partition %zu, tail %zu\n", s->id, n, s->part_n, T);
    const size_t Ta = (T + 63) & ~(size_t)63, nca = ((size_t)g.nc + 1 + 63) & ~(size_t)63;
    const size_t na = (n + 63) & ~(size_t)63;
    size_t bytes = Ta * 16 + nca * 4 + counting_sort_words(g) * 4 + 256 + 64 + na * 4;
    int rc = ensure_scratch(s->w, bytes); if (rc) return rc;
    rc = ensure_u32(s->home_alt, s->home_alt_cap, s->cap); if (rc) return rc;
    unsigned* keysU = (unsigned*)s->w->scratch; unsigned* idxU = keysU + Ta; unsigned* keysA = idxU + Ta; unsigned* idxA = keysA + Ta;
    unsigned* tstart = idxA + Ta; unsigned* work = tstart + nca;
    u64* t_count = (u64*)(work + counting_sort_words(g));
    unsigned* src = (unsigned*)(t_count + 8);
    const u64* n_ptr = &s->ctr->n; const unsigned* part_n_ptr = s->cell_start + g.nc;
    int tgrid = std::max(1, std::min(div_up(T, 256), g_sm_count * 8));
    LAUNCH(K_SORT_KEYS, k_tail_keys, tgrid, 256, 0, g, s->a[0], s->a[1], s->a[2], n_ptr, part_n_ptr, t_count, keysU, idxU); CHECK_LAUNCH();
    // tail in cell order: tstart[c] = tail particles in cells below c, (keysA, idxA) = cell and slot of the r-th tail particle
    rc = counting_sort_by_cell(g, t_count, T, keysU, idxU, tstart, idxA, keysA, work, true); if (rc) return rc;
    int pgrid = std::max(1, std::min(div_up(n, 256), g_sm_count * 8));
    LAUNCH(K_SORT_PERMUTE, k_merge_sources, pgrid, 256, 0, part_n_ptr, t_count, s->home, s->cell_start, tstart, keysA, idxA, src); CHECK_LAUNCH();
    for (int c = 0; c < 7; c++) {
        LAUNCH(K_SORT_PERMUTE, k_sort_permute, pgrid, 256, 0, s->ctr, src, s->a[c], s->spare); CHECK_LAUNCH();
        std::swap(s->a[c], s->spare);
    }
    LAUNCH(K_SORT_PERMUTE, k_merge_home, pgrid, 256, 0, n_ptr, part_n_ptr, src, s->home, keysU, s->home_alt); CHECK_LAUNCH();       // keysU[slot - part_n]: cell of tail slot
    if (s->movers_fresh) {
        unsigned* m_slot = s->mv_trip; unsigned* m_home = m_slot + 2 * s->mv_trip_cap;
        LAUNCH(K_SORT_KEYS, k_merge_remap_movers, g_sm_count * 2, 256, 0, s->ctr, g.nc, tstart, m_slot, m_home, (u64)n, (u64)s->mv_trip_cap); CHECK_LAUNCH();
    }
    std::swap(s->home, s->home_alt); std::swap(s->home_cap, s->home_alt_cap);
    LAUNCH(K_CELL_START, k_merge_cell_start, std::max(1, std::min(div_up((size_t)g.nc + 1, 256), g_sm_count * 8)), 256, 0, g.nc, s->cell_start, tstart); CHECK_LAUNCH();
    s->part_n = n; s->sorted_valid = false; s->lists_valid = false; s->count_valid = false;
    g_tail_merges++;
    return PICG_OK;
}

// Makes the per-cell lists of s exact for its current particle positions: nothing to do if the store is exactly sorted,
// a mover pass over a stale partition when few particles changed cell, a full sort otherwise.
// The species-owned (slot, cell, home) arrays sized by the store capacity (no regrowth while the population grows).
int ensure_mover_triples(picg_species_s* s) {
    size_t mcapa = ((size_t)(g_mover_fraction * (double)s->cap) + 1024 + 63) & ~(size_t)63;
    if (s->mv_trip_cap >= mcapa) return PICG_OK;
    size_t c3 = s->mv_trip_cap * 3;
    int rc = ensure_u32(s->mv_trip, c3, mcapa * 3); if (rc) return rc;
    s->mv_trip_cap = mcapa; s->movers_fresh = false;
    return PICG_OK;
}

int species_exact_lists(picg_species_s* s) {
    s->wants_lists = true;                                   // from now on deposit passes over the partition list the movers on the fly
    if (s->sorted_valid || s->lists_valid) return PICG_OK;
    int rc = species_refresh_count(s); if (rc) return rc;
    size_t n = s->n_host;
    const double max_frac = g_mover_fraction;
    if (!s->part_valid || n < 4096 || max_frac <= 0) return sort_species(s);
    const Grid& g = s->w->g;
    // a long tail of appended particles (MC products) is merged into the partition: a streaming pass, no keys of the partition needed
    if (g_merge_fraction > 0 && n > s->part_n && (double)(n - s->part_n) > g_merge_fraction * (double)n) {
        rc = merge_tail(s); if (rc) return rc;
    }
    size_t mcap = (size_t)(max_frac * (double)n) + 1024;
    size_t mcap_alloc = (size_t)(max_frac * (double)s->cap) + 1024;      // sized by the store capacity: no regrowth while the population grows
    // mover triples (slot / current cell / home cell) live in the species (a deposit pass may have listed them already);
    // radix ping-pong buffers in the scratch arena
    size_t mcapa = (mcap_alloc + 63) & ~(size_t)63;
    rc = ensure_scratch(s->w, counting_sort_words(g) * 4 + 256); if (rc) return rc;
    rc = ensure_u32(s->mv_in, s->mv_cap, mcapa * 2); if (rc) return rc;     // [0,mcapa): slots ordered by current cell, [mcapa, 2 mcapa): slots ordered by home cell
    s->mv_stride = mcapa;
    rc = ensure_mover_triples(s); if (rc) return rc;
    unsigned* m_slot = s->mv_trip; unsigned* m_cell = m_slot + s->mv_trip_cap; unsigned* m_home = m_cell + s->mv_trip_cap;
    unsigned* work = (unsigned*)s->w->scratch;
    u64* cnt = &s->ctr->n_movers;                                            // device-side mover count
    const int tail_only = s->movers_fresh ? 1 : 0;                           // the last deposit pass listed the partition's movers: only the appended tail is left
    if (!tail_only) CUDA_TRY(cudaMemsetAsync(cnt, 0, 8, g_stream));
    else { k_movers_checkpoint<<<1, 1, 0, g_stream>>>(s->ctr, s->movers_saved ? 1 : 0); CHECK_LAUNCH(); s->movers_saved = true; }
    if (tail_only) g_movers_from_deposit++; else g_mover_scans++;
    size_t n_scan = tail_only ? std::max<size_t>(n > s->part_n ? n - s->part_n : 0, (size_t)1 << 18) : n;   // grid size only (grid-stride loop)
    int pgrid = std::max(1, std::min(div_up(std::max<size_t>(n_scan, 1), 256), g_sm_count * 8));
    LAUNCH(K_SORT_KEYS, k_find_movers, pgrid, 256, 0, g, s->a[0], s->a[1], s->a[2], s->home, &s->ctr->n, s->cell_start + g.nc, tail_only ? &s->ctr->n_listed : nullptr, cnt, (u64)mcap, m_slot, m_cell, m_home);
    CHECK_LAUNCH();
    u64 n_live_movers = 0;
    CUDA_TRY(cudaMemcpyAsync(&n_live_movers, cnt, 8, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    if (trace_sort()) fprintf(stderr, "[picgpu] lists of species %u: n = %zu, partition %zu, movers + tail = %llu (cap %zu, %s)\n", s->id, n, s->part_n, (unsigned long long)n_live_movers, mcap, tail_only ? "movers from the deposit pass" : "full scan");
    if (n_live_movers > mcap) { g_mover_resorts++; return sort_species(s); }  // too stale: periodic full radix sort
    // (1) in-lists: live movers by their current cell (counting sort: the scanned table is in_start, each cell's slots ascending)
    rc = counting_sort_by_cell(g, cnt, (size_t)n_live_movers, m_cell, m_slot, s->in_start, s->mv_in, nullptr, work, true); if (rc) return rc;
    // (2) out-lists: every slot whose particle left its home cell (live movers with a home + slots vacated beyond n), by home cell;
    // movers without a home (appended after the sort) carry home = nc and are dropped
    {
        size_t vac_upper = s->part_n > n ? s->part_n - n : 0;
        if (n_live_movers + vac_upper > mcap) { g_mover_resorts++; return sort_species(s); }
        if (vac_upper) { LAUNCH(K_SORT_KEYS, k_find_vacated, std::max(1, std::min(div_up(vac_upper, 256), g_sm_count * 4)), 256, 0, s->home, &s->ctr->n, s->cell_start + g.nc, cnt, (u64)mcap, m_slot, m_home); CHECK_LAUNCH(); }
        rc = counting_sort_by_cell(g, cnt, (size_t)n_live_movers + vac_upper, m_home, m_slot, s->out_start, s->mv_in + mcapa, nullptr, work, false); if (rc) return rc;
    }
    s->lists_valid = true;
    return PICG_OK;
}
}  // namespace picg

extern "C" {
} extern "C++" { namespace picg {
// Species whose per-cell lists are in use (MC collisions): the second home array of the tail merge is allocated up front, so that no
// allocation happens inside a time step.
int species_prepare_lists(picg_species_s* s) { s->wants_lists = true; return ensure_u32(s->home_alt, s->home_alt_cap, s->cap); }
} } extern "C" {
int picg_set_mover_fraction(double f) { REQUIRE_ARG(f >= 0 && f <= 0.5, "picg_set_mover_fraction: 0 <= f <= 0.5"); g_mover_fraction = f; return PICG_OK; }
int picg_set_merge_fraction(double f) { REQUIRE_ARG(f >= 0 && f <= 0.5, "picg_set_merge_fraction: 0 <= f <= 0.5 (0: never merge)"); g_merge_fraction = f; return PICG_OK; }
int picg_tail_merge_count(uint64_t* merges) { if (merges) *merges = g_tail_merges; return PICG_OK; }
int picg_mover_stats(uint64_t* from_deposit, uint64_t* full_scans, uint64_t* resorts) {
    if (from_deposit) *from_deposit = g_movers_from_deposit;
    if (full_scans) *full_scans = g_mover_scans;
    if (resorts) *resorts = g_mover_resorts;
    return PICG_OK;
}
int picg_species_sort(picg_species_t s) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_species_sort: null species");
    return sort_species(s);
}
}
