// Device pieces of the deterministic fixed-point deposition, shared by step.cu and cellstep.cu.
//
// Every particle contributes eight products formed with the reference's association
// (Field.h:172-197), each quantised q = llrint(c * 2^S) and summed in int64.  Integer addition is
// associative, so ANY aggregation order (register runs, warp shuffles, shared-memory staging, global
// atomics, NCCL all-reduce across GPUs) yields the same bits.
//
// Staging (step.cu): every warp owns a private shared-memory window of NODES: the 4 node rows (i,j) (i,j+1) (i+1,j)
// (i+1,j+1) of one cell column, WK nodes along k.  A thread accumulates its run of consecutive particles in registers
// and adds the eight corner sums of the run to the window when the run ends or leaves its cell.  64-bit shared-memory
// atomics do not exist in hardware (they compile to compare-and-swap spin loops, ATOMS.CAST.SPIN.64), so every window
// node is a (lo, hi) pair of 32-bit words updated with two NATIVE 32-bit atomics and an explicit carry:
//     old = atomicAdd(&lo, x_lo);  carry = (old + x_lo) wrapped;  atomicAdd(&hi, x_hi + carry)
// which is exact for non-negative contributions in any interleaving.  The window is flushed with one RED.64 per
// touched node; cells outside the window (unsorted input, stragglers) go straight to global memory.
#pragma once
#include "common.cuh"

#ifdef __CUDACC__
// node index of corner c (0..7: dk = c&1, dj = (c>>1)&1, di = c>>2) of the cell whose low node is (i,j,k)
__device__ __forceinline__ size_t corner_node(const Grid& g, int i, int j, int k, int c) {
    return ((size_t)((i + (c >> 2)) * g.nj + (j + ((c >> 1) & 1))) * g.nk) + (k + (c & 1));
}
// cell = (i*cj + j)*ck + k  ->  (i,j,k) with multiply-shift division (Grid::div_ck / div_cj, exact for cell < 2^31)
__device__ __forceinline__ void cell_to_ijk(const Grid& g, int cell, int& i, int& j, int& k) {
    unsigned t = (unsigned)(((u64)(unsigned)cell * g.div_ck_mul) >> g.div_ck_shift);
    k = cell - (int)t * g.ck;
    unsigned ii = (unsigned)(((u64)t * g.div_cj_mul) >> g.div_cj_shift);
    j = (int)t - (int)ii * g.cj; i = (int)ii;
}

// Transposed butterfly: every lane passes its eight values r[0..7] (zeros for lanes that do not take part);
// on return the lanes with (lane & 3) == 0 hold in the result the sum over all 32 lanes of corner (lane >> 2).
__device__ __forceinline__ i64 butterfly8(const i64 r[8], int lane) {
    bool h = lane & 16;
    i64 a0 = (h ? r[4] : r[0]) + __shfl_xor_sync(0xffffffffu, h ? r[0] : r[4], 16);
    i64 a1 = (h ? r[5] : r[1]) + __shfl_xor_sync(0xffffffffu, h ? r[1] : r[5], 16);
    i64 a2 = (h ? r[6] : r[2]) + __shfl_xor_sync(0xffffffffu, h ? r[2] : r[6], 16);
    i64 a3 = (h ? r[7] : r[3]) + __shfl_xor_sync(0xffffffffu, h ? r[3] : r[7], 16);
    bool b = lane & 8;
    i64 b0 = (b ? a2 : a0) + __shfl_xor_sync(0xffffffffu, b ? a0 : a2, 8);
    i64 b1 = (b ? a3 : a1) + __shfl_xor_sync(0xffffffffu, b ? a1 : a3, 8);
    bool c = lane & 4;
    i64 t = (c ? b1 : b0) + __shfl_xor_sync(0xffffffffu, c ? b0 : b1, 4);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    return t;
}

// The warp's private window is NODE-indexed: 4 node rows (i,j) (i,j+1) (i+1,j) (i+1,j+1) of one cell column, WK nodes
// along k starting at k0.  In the k-fastest sorted order a warp chunk stays in one column, and consecutive cells share
// four of their eight nodes, so the window merges them before anything reaches global memory.
struct NodeWindow { int wi, wj, k0; };
template <int WK>
__device__ __forceinline__ bool window_has(const NodeWindow& W, int i, int j, int k) { return i == W.wi && j == W.wj && k >= W.k0 && k + 1 < W.k0 + WK; }

// v >= 0 is added to window node `slot` ((lo,hi) pair of native 32-bit shared atomics, exact 64-bit sum)
__device__ __forceinline__ void window_add(unsigned* lo, unsigned* hi, int slot, i64 v) {
    unsigned xl = (unsigned)(u64)v, xh = (unsigned)((u64)v >> 32);
    unsigned old = atomicAdd(&lo[slot], xl);
    unsigned carry = (unsigned)(old + xl < old);
    atomicAdd(&hi[slot], xh + carry);
}
// Hands the eight corner sums of a finished run (cell (i,j,k)) to the window or, outside it, to the global grid.
template <int WK>
__device__ __forceinline__ void run_flush(const Grid& g, const NodeWindow& W, unsigned* lo, unsigned* hi, int i, int j, int k, const i64 acc[8],
                                          u64* __restrict__ den_fixed) {
    if (window_has<WK>(W, i, j, k)) {
        const int base = k - W.k0;
#pragma unroll
        for (int c = 0; c < 8; c++) window_add(lo, hi, (c >> 1) * WK + base + (c & 1), acc[c]);
    } else {
#pragma unroll
        for (int c = 0; c < 8; c++) if (acc[c]) atomicAdd(&den_fixed[corner_node(g, i, j, k, c)], (u64)acc[c]);
    }
}
// Hands the non-zero window nodes over to the global grid (one RED.64 each) and clears them.  Warp-collective.
template <int WK>
__device__ __forceinline__ void window_flush(const Grid& g, unsigned* lo, unsigned* hi, const NodeWindow& W, u64* __restrict__ den_fixed, int lane) {
    __syncwarp();
    for (int slot = lane; slot < 4 * WK; slot += 32) {
        u64 t = ((u64)hi[slot] << 32) | lo[slot];
        if (t != 0) {
            int row = slot / WK, kk = slot - row * WK;
            size_t node = ((size_t)((W.wi + (row >> 1)) * g.nj + (W.wj + (row & 1))) * g.nk) + (W.k0 + kk);
            atomicAdd(&den_fixed[node], t);
            lo[slot] = 0; hi[slot] = 0;
        }
    }
    __syncwarp();
}
#endif
