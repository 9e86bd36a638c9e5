// Device pieces of the deterministic fixed-point deposition, shared by step.cu and cellstep.cu.
//
// Every particle contributes eight products formed with the reference's association
// (Field.h:172-197), each quantised q = llrint(c * 2^S) and summed in int64.  Integer addition is
// associative, so ANY aggregation order (register runs, warp shuffles, shared-memory staging, global
// atomics, NCCL all-reduce across GPUs) yields the same bits.
//
// Staging (step.cu): every warp owns a private shared-memory window of WINDOW consecutive cells x 8
// corner accumulators.  Lanes whose run totals fall in the same cell are combined with a transposed
// butterfly (9 64-bit shuffles for all eight corners instead of 40); afterwards eight lanes hold the eight
// corner sums and add them to eight DISTINCT window slots.  The window is private to the warp and the
// slots are distinct, so plain read-modify-write is enough - 64-bit shared-memory atomics would compile to
// compare-and-swap spin loops (ATOMS.CAST.SPIN.64).  Cells outside the window (unsorted input, stragglers)
// go straight to global memory with RED.64.
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

// Adds the contributions of the `active` lanes (cell, q[8]) to the warp's window / the global grid.
// Warp-collective: all 32 lanes must call.  win: this warp's private window [WINDOW*8], c0: its first cell.
template <int WINDOW>
__device__ __forceinline__ void warp_accumulate_w(const Grid& g, bool active, int cell, const i64 q[8], i64* win, int c0,
                                                  u64* __restrict__ den_fixed, int lane) {
    unsigned todo = __ballot_sync(0xffffffffu, active);
    while (todo) {
        int leader = __ffs(todo) - 1;
        int lcell = __shfl_sync(0xffffffffu, cell, leader);
        bool mine = active && cell == lcell;
        todo &= ~__ballot_sync(0xffffffffu, mine);
        i64 r[8];
#pragma unroll
        for (int c = 0; c < 8; c++) r[c] = mine ? q[c] : 0;
        i64 t = butterfly8(r, lane);
        int rel = lcell - c0;
        if ((lane & 3) == 0 && t != 0) {
            int corner = lane >> 2;
            if (rel >= 0 && rel < WINDOW) win[rel * 8 + corner] += t;            // private window, distinct slots: no atomic needed
            else { int i, j, k; cell_to_ijk(g, lcell, i, j, k); atomicAdd(&den_fixed[corner_node(g, i, j, k, corner)], (u64)t); }
        }
        __syncwarp();
    }
}
#endif
