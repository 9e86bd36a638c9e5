// Device pieces of the deterministic fixed-point deposition, shared by deposit.cu and push_deposit.cu.
//
// Every particle contributes eight products formed with the reference's association
// (Field.h:172-197), each quantised q = llrint(c * 2^S) and summed in int64.  Integer addition is
// associative, so ANY aggregation order (warp shuffles, shared-memory staging, global atomics,
// NCCL all-reduce across GPUs) yields the same bits.
//
// Staging (used by step.cu): a block owns a contiguous chunk of the (cell-sorted) particle array and a
// shared-memory window of WINDOW consecutive cells x 8 corner accumulators.  Lanes of a warp whose run
// totals fall in the same cell are first combined with a transposed butterfly (9 64-bit shuffles for all
// eight corners instead of 40), then one lane per corner issues the atomic: to shared memory when the
// cell is inside the window, straight to global memory otherwise (unsorted / straggler particles).
#pragma once
#include "common.cuh"


#ifdef __CUDACC__
// node index of corner c (0..7: dk = c&1, dj = (c>>1)&1, di = c>>2) of the cell whose low node is (i,j,k)
__device__ __forceinline__ size_t corner_node(const Grid& g, int i, int j, int k, int c) {
    return ((size_t)((i + (c >> 2)) * g.nj + (j + ((c >> 1) & 1))) * g.nk) + (k + (c & 1));
}
__device__ __forceinline__ void cell_to_ijk(const Grid& g, int cell, int& i, int& j, int& k) {
    k = cell % g.ck; int t = cell / g.ck; j = t % g.cj; i = t / g.cj;
}

// Adds the warp's contributions.  `active` lanes carry (cell, q[8]); all 32 lanes must call.
// win: shared window accumulators [DEP_WINDOW*8], c0: first cell of the window.
template <int WINDOW>
__device__ __forceinline__ void warp_accumulate_w(const Grid& g, bool active, int cell, const i64 q[8], i64* win, int c0,
                                                  u64* __restrict__ den_fixed, int lane) {
    unsigned todo = __ballot_sync(0xffffffffu, active);
    while (todo) {
        int leader = __ffs(todo) - 1;
        int lcell = __shfl_sync(0xffffffffu, cell, leader);
        bool mine = active && cell == lcell;
        unsigned m = __ballot_sync(0xffffffffu, mine);
        todo &= ~m;
        int rel = lcell - c0;
        bool in_win = rel >= 0 && rel < WINDOW;
        if (__popc(m) >= 3) {
            // transposed butterfly: after the three halving steps lane L holds corner (L>>2)'s partial
            // sum over 8 lanes; two more plain steps finish it.
            i64 r0 = mine ? q[0] : 0, r1 = mine ? q[1] : 0, r2 = mine ? q[2] : 0, r3 = mine ? q[3] : 0;
            i64 r4 = mine ? q[4] : 0, r5 = mine ? q[5] : 0, r6 = mine ? q[6] : 0, r7 = mine ? q[7] : 0;
            bool h = lane & 16;
            i64 a0 = (h ? r4 : r0) + __shfl_xor_sync(0xffffffffu, h ? r0 : r4, 16);
            i64 a1 = (h ? r5 : r1) + __shfl_xor_sync(0xffffffffu, h ? r1 : r5, 16);
            i64 a2 = (h ? r6 : r2) + __shfl_xor_sync(0xffffffffu, h ? r2 : r6, 16);
            i64 a3 = (h ? r7 : r3) + __shfl_xor_sync(0xffffffffu, h ? r3 : r7, 16);
            bool b = lane & 8;
            i64 b0 = (b ? a2 : a0) + __shfl_xor_sync(0xffffffffu, b ? a0 : a2, 8);
            i64 b1 = (b ? a3 : a1) + __shfl_xor_sync(0xffffffffu, b ? a1 : a3, 8);
            bool c = lane & 4;
            i64 t = (c ? b1 : b0) + __shfl_xor_sync(0xffffffffu, c ? b0 : b1, 4);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            t += __shfl_xor_sync(0xffffffffu, t, 1);
            if ((lane & 3) == 0 && t != 0) {
                int corner = lane >> 2;
                if (in_win) atomicAdd((u64*)&win[rel * 8 + corner], (u64)t);
                else { int i, j, k; cell_to_ijk(g, lcell, i, j, k); atomicAdd(&den_fixed[corner_node(g, i, j, k, corner)], (u64)t); }
            }
        } else if (mine) {
            if (in_win) {
#pragma unroll
                for (int c = 0; c < 8; c++) if (q[c] != 0) atomicAdd((u64*)&win[rel * 8 + c], (u64)q[c]);
            } else {
                int i, j, k; cell_to_ijk(g, lcell, i, j, k);
#pragma unroll
                for (int c = 0; c < 8; c++) if (q[c] != 0) atomicAdd(&den_fixed[corner_node(g, i, j, k, c)], (u64)q[c]);
            }
        }
    }
}

#endif
