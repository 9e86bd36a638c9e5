// Exact per-cell particle lists on top of a (possibly stale) cell partition, shared by the per-cell kernels (mcc.cu, merge.cu).
#pragma once
#include "common.cuh"

struct Store { double* a[7]; SpeciesCounters* ctr; u64 cap; };
// Exact per-cell particle lists on top of a (possibly stale) partition, see sort.cu:
//   list(c) = slots [cs[c], cs[c+1]) minus the out-movers of c, followed by the in-movers of c.
struct CellLists { const unsigned* cs; const unsigned* in_start; const unsigned* out_start; const unsigned* mv_in; const unsigned* mv_out; };
struct CellView { unsigned h0, n_stay, o0, o1, i0; int np; };
__device__ __forceinline__ CellView cell_view(const CellLists& L, int c) {
    CellView v;
    v.h0 = L.cs[c]; unsigned h1 = L.cs[c + 1];
    v.o0 = L.out_start[c]; v.o1 = L.out_start[c + 1];
    v.i0 = L.in_start[c]; unsigned i1 = L.in_start[c + 1];
    v.n_stay = (h1 - v.h0) - (v.o1 - v.o0);
    v.np = (int)(v.n_stay + (i1 - v.i0));
    return v;
}
// slot of the a-th particle of the list (0 <= a < np)
__device__ __forceinline__ unsigned cell_pick(const CellLists& L, const CellView& v, int a) {
    if ((unsigned)a >= v.n_stay) return L.mv_in[v.i0 + ((unsigned)a - v.n_stay)];
    unsigned slot = v.h0 + (unsigned)a;
    if (v.o1 == v.o0) return slot;
    for (;;) {                                        // a-th slot of the home range that is not an out-mover (the out list is tiny and unordered)
        unsigned k = 0;
        for (unsigned o = v.o0; o < v.o1; o++) k += (L.mv_out[o] <= slot);
        unsigned nxt = v.h0 + (unsigned)a + k;
        if (nxt == slot) return slot;
        slot = nxt;
    }
}


static inline CellLists lists_of(picg_species_s* s) {
    CellLists L; L.cs = s->cell_start; L.in_start = s->in_start; L.out_start = s->out_start; L.mv_in = s->mv_in; L.mv_out = s->mv_in ? s->mv_in + s->mv_stride : nullptr;
    return L;
}
static inline Store store_of(picg_species_s* s) { Store st; for (int c = 0; c < 7; c++) st.a[c] = s->a[c]; st.ctr = s->ctr; st.cap = s->cap; return st; }
