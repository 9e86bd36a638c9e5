// PotentialSolver on the device: red-black SOR with the reference's node classes, residual, E = -grad(phi).
//   k_sor_row / k_sor_slab : one colour half-sweep of PotentialSolver::solveGS     ch4/v3/src/PotentialSolver.cpp:86-121
//   k_residual    : the convergence check (every 25 iterations)           :124-159
//   k_compute_ef  : PotentialSolver::computeEF                             :354-408
// The reference sweeps lexicographically (Gauss-Seidel); red-black ordering visits the same node
// classes with the same update formula, so the two iterations share their fixed point (parity is
// checked on converged solutions, SURVEY.md section 7 "hard parts").
// Algorithmic bytes: 25 B/node/iteration (phi R+W, rho R, mask R); computeEF 32 B/node.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

using namespace picg;

struct SorParams {
    double inv_d2x, inv_d2y, inv_d2z, inv_eps0, twos, inv_twos;   // PotentialSolver::precalculate :473-491
    double phi0, n0, Te0, qe, w;
    int bc_mode;
};

// node update shared by sweep and residual
//  class 0: skipped (object node; in ch2 mode also the six faces)
//  class 1..6: zero-gradient face, first matching rule i0,iN,j0,jN,k0,kN (:96-107)
//  class 7: interior
__device__ __forceinline__ int node_class(const Grid& g, int bc_mode, int oid, int i, int j, int k) {
    if (oid > 0) return 0;
    bool face = (i == 0 || i == g.ni - 1 || j == 0 || j == g.nj - 1 || k == 0 || k == g.nk - 1);
    if (!face) return 7;
    if (bc_mode == 1) return 0;
    if (i == 0) return 1; if (i == g.ni - 1) return 2;
    if (j == 0) return 3; if (j == g.nj - 1) return 4;
    if (k == 0) return 5; return 6;
}
__device__ __forceinline__ size_t face_neighbor(const Grid& g, int cls, size_t u) {
    size_t si = (size_t)g.nj * g.nk, sj = g.nk;
    switch (cls) { case 1: return u + si; case 2: return u - si; case 3: return u + sj; case 4: return u - sj; case 5: return u + 1; default: return u - 1; }
}

// Colour-compact copies made once per solve.  In row (i,j) the nodes of colour c are k = (i+j+c)&1, +2, ...; node k is element
// row*hk + (k>>1) of colour c's half (hk = ceil(nk/2)).  A half-sweep then reads the node class (1 byte: no index arithmetic or
// geometry branches in the sweeps) and rho of its own colour with unit stride instead of every other value of whole sectors.
__global__ void __launch_bounds__(256) k_node_classes(Grid g, int bc_mode, const int* __restrict__ object_id, const double* __restrict__ rho,
                                                      unsigned char* __restrict__ cls, double* __restrict__ rho_split, unsigned char* __restrict__ cls_nat) {
    const int hk = (g.nk + 1) >> 1;
    const size_t half = (size_t)g.ni * g.nj * hk;
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < (size_t)g.nv; u += (size_t)gridDim.x * blockDim.x) {
        int k = (int)(u % g.nk); size_t row = u / g.nk; int j = (int)(row % g.nj), i = (int)(row / g.nj);
        const int color = (i + j + k) & 1;                                      // k == (i+j+color)&1 mod 2
        const size_t c = (size_t)color * half + row * hk + (k >> 1);
        const unsigned char cl = (unsigned char)node_class(g, bc_mode, object_id[u], i, j, k);
        cls[c] = cl; cls_nat[u] = cl;                                           // colour-compact for the row sweeps, node order for the tiled sweep
        rho_split[c] = rho[u];
    }
}
// One colour half-sweep, one block per (i,j) row: no divisions, coalesced along k, class byte instead of geometry tests.
__global__ void __launch_bounds__(128) k_sor_row(Grid g, SorParams sp, int color, double* __restrict__ phi, const double* __restrict__ rho,
                                                 const unsigned char* __restrict__ cls) {
    const int j = blockIdx.x, i = blockIdx.y;
    const size_t si = (size_t)g.nj * g.nk, sj = g.nk;
    const size_t row = ((size_t)i * g.nj + j) * g.nk;
    const int hk = (g.nk + 1) >> 1;
    const size_t crow = (size_t)color * ((size_t)g.ni * g.nj * hk) + ((size_t)i * g.nj + j) * hk;     // this row in the colour-compact arrays
    for (int k = 2 * threadIdx.x + ((i + j + color) & 1); k < g.nk; k += 2 * blockDim.x) {
        const size_t u = row + k;
        const int c = cls[crow + (k >> 1)];
        if (c == 0) continue;
        if (c < 7) { phi[u] = phi[face_neighbor(g, c, u)]; continue; }
        const double p = phi[u];
        const double ne = (sp.n0 != 0.0) ? sp.n0 * exp((p - sp.phi0) / sp.Te0) : 0.0;
        const double nw = ((rho[crow + (k >> 1)] - sp.qe * ne) * sp.inv_eps0 + (phi[u - si] + phi[u + si]) * sp.inv_d2x + (phi[u - sj] + phi[u + sj]) * sp.inv_d2y +
                           (phi[u - 1] + phi[u + 1]) * sp.inv_d2z) * sp.inv_twos;
        phi[u] = p + sp.w * (nw - p);
    }
}

// ---------------------------------------------------------------- one pass per iteration: plane-marching shared-memory tiles
// A red-black iteration as two row sweeps reads phi twice (279 MB of DRAM traffic per half-sweep at 256^3: 33 B per node and iteration).
// Here ONE kernel does both colours: a block owns a (j, k) tile and marches along i.  Step p: (A) the RED nodes of plane p are updated
// from the old BLACK values on the tile extended by one node in j and k (the ring is recomputed by the neighbour tiles: it only
// reads old values, which never change - phi is double-buffered) and kept in a three-plane ring in shared memory; (B) the BLACK nodes of
// plane p - 1 are updated from the new red values of planes p - 2, p - 1, p in shared memory, and plane p - 1 leaves for the output
// buffer as whole rows (new reds from shared memory, new blacks).  Every operand has exactly the value the two-sweep kernels see, and the
// update expression is the same: the result is bit-identical to k_sor_row.  DRAM traffic: phi read once, rho, class byte, phi
// written once = 25 B per node and iteration, plus the ring rows shared with the neighbour tiles (served by the L2).
#define ST_TJ 16
#define ST_TK 64
#define ST_ROW (ST_TK + 4)                                   // padded row of the shared-memory planes (ST_TK + 2 used)
#define ST_THREADS 256
__device__ __forceinline__ double sor_update(const SorParams& sp, double p, double rho, double im, double ip, double jm, double jp, double km, double kp) {
    const double ne = (sp.n0 != 0.0) ? sp.n0 * exp((p - sp.phi0) / sp.Te0) : 0.0;
    const double nw = ((rho - sp.qe * ne) * sp.inv_eps0 + (im + ip) * sp.inv_d2x + (jm + jp) * sp.inv_d2y + (km + kp) * sp.inv_d2z) * sp.inv_twos;
    return p + sp.w * (nw - p);
}
// Within a step a thread requests all operands of a node at once, without looking at the class byte first (neighbour offsets are
// zeroed on the mesh faces, so every address is valid): one memory round trip per node instead of two.
#define ST_HALF ((ST_TK + 2) / 2)                            // red (or black) nodes per row of the extended region
__global__ void __launch_bounds__(ST_THREADS, 4) k_sor_tiled(Grid g, SorParams sp, const double* __restrict__ in, double* __restrict__ out, const double* __restrict__ rho,
                                                            const unsigned char* __restrict__ cls, int planes_per_block) {
    __shared__ double R[3][ST_TJ + 2][ST_ROW];
    const int k0 = blockIdx.x * ST_TK, j0 = blockIdx.y * ST_TJ, ia = blockIdx.z * planes_per_block, ib = min(ia + planes_per_block, g.ni);
    const long long si = (long long)g.nj * g.nk, sj = g.nk;
    for (int p = max(ia - 1, 0); p <= ib; p++) {
        // ---- (A) new red values of plane p on the tile extended by one node in j and k -> R[p % 3]
        if (p < g.ni) {
            double (*Rp)[ST_ROW] = R[p % 3];
            const long long oim = p > 0 ? -si : 0, oip = p < g.ni - 1 ? si : 0;
            for (int t = threadIdx.x; t < (ST_TJ + 2) * ST_HALF; t += ST_THREADS) {
                const int jr = t / ST_HALF, j = j0 - 1 + jr;
                const int kr = 2 * (t - jr * ST_HALF) + ((p + j + k0 + 1) & 1), k = k0 - 1 + kr;      // (p + j + k) even: red
                if (j < 0 || j >= g.nj || k < 0 || k >= g.nk) continue;
                const size_t u = ((size_t)p * g.nj + j) * g.nk + k;
                const long long ojm = j > 0 ? -sj : 0, ojp = j < g.nj - 1 ? sj : 0, okm = k > 0 ? -1 : 0, okp = k < g.nk - 1 ? 1 : 0;
                const int c = cls[u]; const double v0 = in[u], rh = rho[u];
                const double a0 = in[u + oim], a1 = in[u + oip], a2 = in[u + ojm], a3 = in[u + ojp], a4 = in[u + okm], a5 = in[u + okp];
                double v;
                if (c == 7) v = sor_update(sp, v0, rh, a0, a1, a2, a3, a4, a5);
                else v = c == 0 ? v0 : c == 1 ? a1 : c == 2 ? a0 : c == 3 ? a3 : c == 4 ? a2 : c == 5 ? a5 : a4;       // face_neighbor: the inward neighbour
                Rp[jr][kr] = v;
            }
        }
        __syncthreads();
        // ---- (B) plane q = p - 1 of the tile: new blacks from the new reds in shared memory; the whole plane leaves for `out`
        const int q = p - 1;
        if (q >= ia && q < ib) {
            double (*Rq)[ST_ROW] = R[q % 3]; double (*Rm)[ST_ROW] = R[(q + 2) % 3]; double (*Rn)[ST_ROW] = R[(q + 1) % 3];
            for (int t = threadIdx.x; t < ST_TJ * ST_TK; t += ST_THREADS) {
                const int jr = t / ST_TK, kr = t - jr * ST_TK, j = j0 + jr, k = k0 + kr;
                if (j >= g.nj || k >= g.nk) continue;
                const int jj = jr + 1, kk = kr + 1;
                const size_t u = ((size_t)q * g.nj + j) * g.nk + k;
                double v;
                if (((q + j + k) & 1) == 0) v = Rq[jj][kk];                                 // red: computed in step q
                else {
                    const int c = cls[u]; const double v0 = in[u], rh = rho[u];
                    if (c == 7) v = sor_update(sp, v0, rh, Rm[jj][kk], Rn[jj][kk], Rq[jj - 1][kk], Rq[jj + 1][kk], Rq[jj][kk - 1], Rq[jj][kk + 1]);
                    else v = c == 0 ? v0 : c == 1 ? Rn[jj][kk] : c == 2 ? Rm[jj][kk] : c == 3 ? Rq[jj + 1][kk] : c == 4 ? Rq[jj - 1][kk] : c == 5 ? Rq[jj][kk + 1] : Rq[jj][kk - 1];
                }
                out[u] = v;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- multi-GPU: slab decomposition with peer-memory halos
// Mailbox layout (u64 words).  Written by the peers with system-scope stores, read with volatile loads.
#define MB_FLAG_LEFT 0          // last half-sweep whose boundary plane the LEFT neighbour has delivered into this rank's phi
#define MB_FLAG_RIGHT 1
#define MB_RES_VAL 8            // + r: residual partial sum of rank r (bits of a double)
#define MB_RES_TAG 72           // + r: sequence number of that sum
#define MB_GATHER 136           // + r: sequence number of the last all-gather rank r has delivered
#define MB_COUNT_LEFT 200       // local: boundary rows of the current half-sweep already delivered to the left / right neighbour
#define MB_COUNT_RIGHT 201
#define MB_SEQ 202              // local: number of half-sweeps completed before the current batch (advanced on the device: graph replays need no new arguments)
#define MB_WORDS 256
#define SLAB_MAX_RANKS 64

struct SlabArgs { int i0, i1; double* left_phi; double* right_phi; u64* mbox; u64* left_mbox; u64* right_mbox; unsigned seq_off; };

__device__ __forceinline__ void spin_until(const u64* flag, u64 want) {
    while (*(volatile const u64*)flag < want) __nanosleep(64);
    __threadfence_system();
}

// One colour half-sweep over the planes [i0, i1) of this rank.  Same update as k_sor_row.  The two boundary planes are
// scheduled first; their rows wait for the neighbour's previous half-sweep (flag in the mailbox, long set), then write every
// new value to the neighbour's copy of the plane as well (its halo) - the exchange is part of the sweep, tile by tile -
// and the last row to finish raises the neighbour's flag.
__global__ void __launch_bounds__(128) k_sor_slab(Grid g, SorParams sp, int color, double* __restrict__ phi, const double* __restrict__ rho,
                                                  const unsigned char* __restrict__ cls, SlabArgs S) {
    const int j = blockIdx.x, np = S.i1 - S.i0, y = blockIdx.y;
    const u64 seq = S.mbox[MB_SEQ] + S.seq_off;                                  // number of this half-sweep (the same on every rank)
    int i; bool to_left = false, to_right = false;
    if (y == 0) { i = S.i0; to_left = S.left_phi != nullptr; }                   // boundary planes first: their delivery (peer stores +
    else if (y == 1) { i = S.i1 - 1; to_right = S.right_phi != nullptr; }         // system fence) overlaps the interior planes
    else i = S.i0 + y - 1;
    if (to_left || to_right) {
        if (threadIdx.x == 0) spin_until(S.mbox + (to_left ? MB_FLAG_LEFT : MB_FLAG_RIGHT), seq - 1);
        __syncthreads();
    }
    double* peer = to_left ? S.left_phi : S.right_phi;
    const size_t si = (size_t)g.nj * g.nk, sj = g.nk;
    const size_t row = ((size_t)i * g.nj + j) * g.nk;
    const int hk = (g.nk + 1) >> 1;
    const size_t crow = (size_t)color * ((size_t)g.ni * g.nj * hk) + ((size_t)i * g.nj + j) * hk;     // this row in the colour-compact arrays
    for (int k = 2 * threadIdx.x + ((i + j + color) & 1); k < g.nk; k += 2 * blockDim.x) {
        const size_t u = row + k;
        const int c = cls[crow + (k >> 1)];
        if (c == 0) continue;
        double nv_;
        if (c < 7) nv_ = phi[face_neighbor(g, c, u)];
        else {
            const double p = phi[u];
            const double ne = (sp.n0 != 0.0) ? sp.n0 * exp((p - sp.phi0) / sp.Te0) : 0.0;
            // the halo plane is written by the neighbour GPU: read it from L2 (ld.cg), never from a possibly stale L1 line
            const double lo_i = to_left ? __ldcg(phi + u - si) : phi[u - si], hi_i = to_right ? __ldcg(phi + u + si) : phi[u + si];
            const double nw = ((rho[crow + (k >> 1)] - sp.qe * ne) * sp.inv_eps0 + (lo_i + hi_i) * sp.inv_d2x + (phi[u - sj] + phi[u + sj]) * sp.inv_d2y +
                               (phi[u - 1] + phi[u + 1]) * sp.inv_d2z) * sp.inv_twos;
            nv_ = p + sp.w * (nw - p);
        }
        phi[u] = nv_;
        if (to_left || to_right) peer[u] = nv_;
    }
    if (to_left || to_right) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            u64* counter = S.mbox + (to_left ? MB_COUNT_LEFT : MB_COUNT_RIGHT);
            if (atomicAdd(counter, 1ull) == (u64)g.nj - 1) {                    // every row of the plane is delivered
                *counter = 0;
                __threadfence_system();
                *(volatile u64*)((to_left ? S.left_mbox : S.right_mbox) + (to_left ? MB_FLAG_RIGHT : MB_FLAG_LEFT)) = seq;
            }
        }
    }
}
// the halos of both colours are in place once both neighbours have delivered the last half-sweep
__global__ void k_slab_wait_halos(SlabArgs S) {
    const u64 seq = S.mbox[MB_SEQ];
    if (S.left_phi) spin_until(S.mbox + MB_FLAG_LEFT, seq);
    if (S.right_phi) spin_until(S.mbox + MB_FLAG_RIGHT, seq);
}
__global__ void k_slab_advance(u64* mbox, unsigned n) { mbox[MB_SEQ] += n; }
// sum of this rank's residual partials -> every rank's mailbox; then the sum over ranks, in rank order (the same bits everywhere)
__global__ void __launch_bounds__(256) k_slab_post_residual(const double* __restrict__ partial, int n, int rank, int world, u64* const* __restrict__ peer_mbox, u64 tag) {
    __shared__ double sm[256];
    double a = 0; for (int t = threadIdx.x; t < n; t += 256) a += partial[t];
    sm[threadIdx.x] = a; __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s]; __syncthreads(); }
    if (threadIdx.x < world) {
        u64* mb = peer_mbox[threadIdx.x];
        *(volatile u64*)(mb + MB_RES_VAL + rank) = (u64)__double_as_longlong(sm[0]);
        __threadfence_system();
        *(volatile u64*)(mb + MB_RES_TAG + rank) = tag;
    }
}
__global__ void k_slab_collect_residual(const u64* __restrict__ mbox, int world, u64 tag, double* __restrict__ out) {
    double s = 0;
    for (int r = 0; r < world; r++) { spin_until(mbox + MB_RES_TAG + r, tag); s += __longlong_as_double((long long)*(volatile const u64*)(mbox + MB_RES_VAL + r)); }
    out[0] = s;
}
// all-gather of phi: this rank's planes go to every peer, then the peers are told
__global__ void __launch_bounds__(256) k_slab_push(const double* __restrict__ phi, size_t begin, size_t end, int rank, int world, double* const* __restrict__ peer_phi) {
    const size_t n2 = (end - begin) / 2;                                        // plane sizes are even or handled by the tail below
    for (int r = 0; r < world; r++) {
        if (r == rank) continue;
        double* dst = peer_phi[r];
        if (((begin & 1) == 0)) {
            const double2* s2 = reinterpret_cast<const double2*>(phi + begin); double2* d2 = reinterpret_cast<double2*>(dst + begin);
            for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n2; t += (size_t)gridDim.x * blockDim.x) d2[t] = s2[t];
            if (((end - begin) & 1) && blockIdx.x == 0 && threadIdx.x == 0) dst[end - 1] = phi[end - 1];
        } else {
            for (size_t t = begin + blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < end; t += (size_t)gridDim.x * blockDim.x) dst[t] = phi[t];
        }
    }
    __threadfence_system();
}
__global__ void k_slab_gather_signal(int rank, int world, u64* const* __restrict__ peer_mbox, u64 tag) {
    if (threadIdx.x < world && threadIdx.x != rank) { __threadfence_system(); *(volatile u64*)(peer_mbox[threadIdx.x] + MB_GATHER + rank) = tag; }
}
__global__ void k_slab_gather_wait(const u64* __restrict__ mbox, int rank, int world, u64 tag) {
    for (int r = 0; r < world; r++) if (r != rank) spin_until(mbox + MB_GATHER + r, tag);
}

__global__ void __launch_bounds__(256) k_residual(Grid g, SorParams sp, const double* __restrict__ phi, const double* __restrict__ rho,
                                                  const int* __restrict__ object_id, double* __restrict__ partial, size_t u_begin, size_t u_end) {
    __shared__ double sm[256];
    const size_t si = (size_t)g.nj * g.nk, sj = g.nk;
    double acc = 0.0;
    for (size_t u = u_begin + blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < u_end; u += (size_t)gridDim.x * blockDim.x) {
        int k = (int)(u % g.nk); size_t row = u / g.nk; int j = (int)(row % g.nj), i = (int)(row / g.nj);
        int cls = node_class(g, sp.bc_mode, object_id[u], i, j, k);
        if (cls == 0) continue;
        double R;
        if (cls < 7) R = phi[u] - phi[face_neighbor(g, cls, u)];
        else {
            double p = phi[u];
            double ne = (sp.n0 != 0.0) ? sp.n0 * exp((p - sp.phi0) / sp.Te0) : 0.0;
            R = -p * sp.twos + (rho[u] - sp.qe * ne) * sp.inv_eps0 + (phi[u - si] + phi[u + si]) * sp.inv_d2x + (phi[u - sj] + phi[u + sj]) * sp.inv_d2y +
                (phi[u - 1] + phi[u + 1]) * sp.inv_d2z;
        }
        acc += R * R;
    }
    sm[threadIdx.x] = acc; __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s]; __syncthreads(); }
    if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

// PotentialSolver::computeEF (:354-408): all nodes, central differences, 2nd-order one-sided on the faces
__global__ void __launch_bounds__(256) k_compute_ef(Grid g, double inv_2dx, double inv_2dy, double inv_2dz, const double* __restrict__ phi, double* __restrict__ ef) {
    const size_t si = (size_t)g.nj * g.nk, sj = g.nk;
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < (size_t)g.nv; u += (size_t)gridDim.x * blockDim.x) {
        int k = (int)(u % g.nk); size_t row = u / g.nk; int j = (int)(row % g.nj), i = (int)(row / g.nj);
        double p = phi[u], ex, ey, ez;
        if (i == 0) ex = __dmul_rn(__dadd_rn(__dsub_rn(__dmul_rn(3.0, p), __dmul_rn(4.0, phi[u + si])), phi[u + 2 * si]), inv_2dx);
        else if (i == g.ni - 1) ex = __dmul_rn(__dsub_rn(__dadd_rn(-phi[u - 2 * si], __dmul_rn(4.0, phi[u - si])), __dmul_rn(3.0, p)), inv_2dx);
        else ex = __dmul_rn(__dsub_rn(phi[u - si], phi[u + si]), inv_2dx);
        if (j == 0) ey = __dmul_rn(__dadd_rn(__dsub_rn(__dmul_rn(3.0, p), __dmul_rn(4.0, phi[u + sj])), phi[u + 2 * sj]), inv_2dy);
        else if (j == g.nj - 1) ey = __dmul_rn(__dsub_rn(__dadd_rn(-phi[u - 2 * sj], __dmul_rn(4.0, phi[u - sj])), __dmul_rn(3.0, p)), inv_2dy);
        else ey = __dmul_rn(__dsub_rn(phi[u - sj], phi[u + sj]), inv_2dy);
        if (k == 0) ez = __dmul_rn(__dadd_rn(__dsub_rn(__dmul_rn(3.0, p), __dmul_rn(4.0, phi[u + 1])), phi[u + 2]), inv_2dz);
        else if (k == g.nk - 1) ez = __dmul_rn(__dsub_rn(__dadd_rn(-phi[u - 2], __dmul_rn(4.0, phi[u - 1])), __dmul_rn(3.0, p)), inv_2dz);
        else ez = __dmul_rn(__dsub_rn(phi[u - 1], phi[u + 1]), inv_2dz);
        ef[3 * u] = ex; ef[3 * u + 1] = ey; ef[3 * u + 2] = ez;
    }
}

static SorParams make_params(const picg_solver_s* s) {
    const Grid& g = s->w->g;
    SorParams p;
    p.inv_d2x = 1.0 / (g.dx[0] * g.dx[0]); p.inv_d2y = 1.0 / (g.dx[1] * g.dx[1]); p.inv_d2z = 1.0 / (g.dx[2] * g.dx[2]);
    p.inv_eps0 = 1.0 / 8.85418782e-12;
    p.twos = 2.0 * (p.inv_d2x + p.inv_d2y + p.inv_d2z); p.inv_twos = 1.0 / p.twos;
    p.phi0 = s->phi0; p.n0 = s->n0; p.Te0 = s->Te0; p.qe = 1.602176565e-19; p.w = 1.4;     // SOR_weight PotentialSolver.h:38
    p.bc_mode = s->bc_mode;
    return p;
}
static const int kResidualBlocks = 1024;

static int prepare_classes(picg_solver_s* s) {
    const Grid& g = s->w->g;
    const size_t half = (size_t)g.ni * g.nj * ((g.nk + 1) >> 1);
    if (!s->cls) {
        cudaError_t e = cudaMalloc(&s->cls, 2 * half);
        if (e == cudaSuccess) e = cudaMalloc(&s->rho_split, 2 * half * 8);
        if (e == cudaSuccess) e = cudaMalloc(&s->cls_nat, (size_t)g.nv);
        if (e == cudaSuccess) e = cudaMalloc(&s->phi_alt, (size_t)g.nv * 8);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(colour-compact classes / rho)", __FILE__, __LINE__);
    }
    LAUNCH(K_MISC, k_node_classes, std::min(div_up(g.nv, 256), g_sm_count * 8), 256, 0, g, s->bc_mode, s->w->object_id, s->w->rho, s->cls, s->rho_split, s->cls_nat); CHECK_LAUNCH();
    return PICG_OK;
}
static SlabArgs slab_args(picg_solver_s* s, unsigned seq_off) {
    SlabArgs S; S.i0 = s->slab_i0; S.i1 = s->slab_i1; S.mbox = s->mbox; S.seq_off = seq_off;
    const int r = s->slab_rank, G = s->slab_world;
    S.left_phi = r > 0 ? s->peer_phi[r - 1] : nullptr; S.left_mbox = r > 0 ? s->peer_mbox[r - 1] : nullptr;
    S.right_phi = r + 1 < G ? s->peer_phi[r + 1] : nullptr; S.right_mbox = r + 1 < G ? s->peer_mbox[r + 1] : nullptr;
    return S;
}
// n iterations = 2n colour half-sweeps enqueued back to back (plain launches; `timed` adds the per-kernel bookkeeping)
static bool use_tiled(const picg_solver_s* s) { return s->slab_world <= 1 && s->sweep_mode == 1; }
static int enqueue_iterations(picg_solver_s* s, const SorParams& p, unsigned n, bool timed) {
    const Grid& g = s->w->g;
    const bool slab = s->slab_world > 1;
    if (use_tiled(s)) {                                           // one kernel per iteration, phi double-buffered (see k_sor_tiled)
        const int tiles = div_up(g.nk, ST_TK) * div_up(g.nj, ST_TJ);
        int chunks = std::max(1, std::min(g.ni, div_up((size_t)g_sm_count * 4, (size_t)tiles)));       // four blocks per SM: one wave
        const int ppb = div_up(g.ni, chunks);
        dim3 tgrid(div_up(g.nk, ST_TK), div_up(g.nj, ST_TJ), div_up(g.ni, ppb));
        double* src = s->w->phi; double* dst = s->phi_alt;
        for (unsigned h = 0; h < n; h++) {
            if (timed) LAUNCH(K_SOR_TILED, k_sor_tiled, tgrid, ST_THREADS, 0, g, p, src, dst, s->w->rho, s->cls_nat, ppb);
            else k_sor_tiled<<<tgrid, ST_THREADS, 0, g_stream>>>(g, p, src, dst, s->w->rho, s->cls_nat, ppb);
            CHECK_LAUNCH();
            std::swap(src, dst);
        }
        if (n & 1) CUDA_TRY(cudaMemcpyAsync(s->w->phi, s->phi_alt, (size_t)g.nv * 8, cudaMemcpyDeviceToDevice, g_stream));      // the result lives in the world's phi
        return PICG_OK;
    }
    dim3 grid(g.nj, slab ? s->slab_i1 - s->slab_i0 : g.ni);
    for (unsigned h = 0; h < 2 * n; h++) {
        const int color = h & 1;
        if (timed) {
            if (slab) LAUNCH(K_SOR, k_sor_slab, grid, 128, 0, g, p, color, s->w->phi, s->rho_split, s->cls, slab_args(s, h + 1));
            else LAUNCH(K_SOR, k_sor_row, grid, 128, 0, g, p, color, s->w->phi, s->rho_split, s->cls);
        } else {
            if (slab) k_sor_slab<<<grid, 128, 0, g_stream>>>(g, p, color, s->w->phi, s->rho_split, s->cls, slab_args(s, h + 1));
            else k_sor_row<<<grid, 128, 0, g_stream>>>(g, p, color, s->w->phi, s->rho_split, s->cls);
        }
        CHECK_LAUNCH();
    }
    if (slab) { k_slab_advance<<<1, 1, 0, g_stream>>>(s->mbox, 2 * n); CHECK_LAUNCH(); }
    return PICG_OK;
}
// Runs n iterations.  Batches of 4 or more replay a CUDA graph captured once per (n, parameters): a half-sweep on a slab or on a
// small mesh takes a few microseconds, less than a host-side launch.
static int run_iterations(picg_solver_s* s, const SorParams& p, unsigned n) {
    static const bool no_graph = getenv("PICG_NO_GRAPH") && atoi(getenv("PICG_NO_GRAPH")) != 0;
    if (n == 0) return PICG_OK;
    if (n < 4 || no_graph) return enqueue_iterations(s, p, n, true);
    static_assert(sizeof(SorParams) <= sizeof(picg_solver_s::SorGraph::params), "SorGraph::params too small");
    cudaGraphExec_t exec = nullptr;
    for (auto& e : s->graphs) if (e.n == n && memcmp(e.params, &p, sizeof(SorParams)) == 0) exec = (cudaGraphExec_t)e.exec;
    if (!exec) {
        cudaGraph_t graph = nullptr;
        CUDA_TRY(cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeThreadLocal));
        g_capturing = true;
        int rc = enqueue_iterations(s, p, n, false);
        g_capturing = false;
        cudaError_t e = cudaStreamEndCapture(g_stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
        e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
        picg_solver_s::SorGraph entry; entry.n = n; memset(entry.params, 0, sizeof(entry.params)); memcpy(entry.params, &p, sizeof(SorParams)); entry.exec = exec;
        s->graphs.push_back(entry);
    }
    {
        const bool tiled = use_tiled(s);
        TimerScope t(tiled ? K_SOR_TILED : K_SOR);               // the whole batch is one timed interval; every sweep kernel counts as a launch
        for (unsigned h = 0; h < (tiled ? n : 2 * n); h++) count_launch(tiled ? K_SOR_TILED : K_SOR);
        CUDA_TRY(cudaGraphLaunch(exec, g_stream));
    }
    return PICG_OK;
}
static void drop_graphs(picg_solver_s* s) {
    for (auto& e : s->graphs) cudaGraphExecDestroy((cudaGraphExec_t)e.exec);
    s->graphs.clear();
}
static int compute_residual(picg_solver_s* s, const SorParams& p, double* L2) {
    const Grid& g = s->w->g;
    const size_t plane = (size_t)g.nj * g.nk;
    if (s->slab_world > 1) {                                      // own planes only; the sum over ranks goes through the mailboxes
        const size_t ub = s->slab_i0 * plane, ue = s->slab_i1 * plane;
        int grid = std::min(div_up(ue - ub, 256), kResidualBlocks);
        LAUNCH(K_RESIDUAL, k_slab_wait_halos, 1, 1, 0, slab_args(s, 0)); CHECK_LAUNCH();
        LAUNCH(K_RESIDUAL, k_residual, grid, 256, 0, g, p, s->w->phi, s->w->rho, s->w->object_id, s->partial, ub, ue); CHECK_LAUNCH();
        s->residual_seq++;
        LAUNCH(K_RESIDUAL, k_slab_post_residual, 1, 256, 0, s->partial, grid, s->slab_rank, s->slab_world, s->peer_mbox_dev, s->residual_seq); CHECK_LAUNCH();
        LAUNCH(K_RESIDUAL, k_slab_collect_residual, 1, 1, 0, s->mbox, s->slab_world, s->residual_seq, s->partial); CHECK_LAUNCH();
        CUDA_TRY(cudaMemcpyAsync(s->w->reduce_host, s->partial, 8, cudaMemcpyDeviceToHost, g_stream));
        CUDA_TRY(cudaStreamSynchronize(g_stream));
        *L2 = std::sqrt(s->w->reduce_host[0] / g.nv);
        return PICG_OK;
    }
    int grid = std::min(div_up(g.nv, 256), kResidualBlocks);
    LAUNCH(K_RESIDUAL, k_residual, grid, 256, 0, g, p, s->w->phi, s->w->rho, s->w->object_id, s->partial, (size_t)0, (size_t)g.nv); CHECK_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(s->w->reduce_host, s->partial, grid * 8, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    double sum = 0; for (int b = 0; b < grid; b++) sum += s->w->reduce_host[b];
    *L2 = std::sqrt(sum / g.nv);                                  // normalised by all nv nodes (:154)
    return PICG_OK;
}
// slab mode: every rank ends a solve with the full, identical phi (its own planes pushed to all peers)
static int slab_allgather(picg_solver_s* s) {
    if (s->slab_world <= 1) return PICG_OK;
    const Grid& g = s->w->g;
    const size_t plane = (size_t)g.nj * g.nk;
    s->gather_seq++;
    LAUNCH(K_MISC, k_slab_wait_halos, 1, 1, 0, slab_args(s, 0)); CHECK_LAUNCH();       // both neighbours have delivered their last half-sweep
    LAUNCH(K_MISC, k_slab_push, g_sm_count * 2, 256, 0, s->w->phi, s->slab_i0 * plane, s->slab_i1 * plane, s->slab_rank, s->slab_world, s->peer_phi_dev); CHECK_LAUNCH();
    LAUNCH(K_MISC, k_slab_gather_signal, 1, SLAB_MAX_RANKS, 0, s->slab_rank, s->slab_world, s->peer_mbox_dev, s->gather_seq); CHECK_LAUNCH();
    LAUNCH(K_MISC, k_slab_gather_wait, 1, 1, 0, s->mbox, s->slab_rank, s->slab_world, s->gather_seq); CHECK_LAUNCH();
    return PICG_OK;
}

extern "C" {

int picg_solver_create(picg_world_t w, unsigned max_it, double tol, picg_solver_t* out) {
    REQUIRE_DEVICE(); REQUIRE_ARG(w && out, "picg_solver_create: null argument");
    picg_solver_s* s = new picg_solver_s();
    s->w = w; s->max_it = max_it; s->tol = tol;
    cudaError_t e = cudaMalloc(&s->partial, kResidualBlocks * 8);
    if (e != cudaSuccess) { delete s; return cuda_fail(e, "cudaMalloc(solver)", __FILE__, __LINE__); }
    *out = s; return PICG_OK;
}
int picg_solver_destroy(picg_solver_t s) {
    if (!s) return PICG_OK;
    if (g_stream) cudaStreamSynchronize(g_stream);
    drop_graphs(s);
    for (int r = 0; r < (int)s->peer_phi.size(); r++) if (r != s->slab_rank) { cudaIpcCloseMemHandle(s->peer_phi[r]); cudaIpcCloseMemHandle(s->peer_mbox[r]); }
    cudaFree(s->peer_phi_dev); cudaFree(s->peer_mbox_dev); cudaFree(s->mbox);
    cudaFree(s->partial); cudaFree(s->cls); cudaFree(s->rho_split); cudaFree(s->cls_nat); cudaFree(s->phi_alt); cudaFree(s->pcg_work); delete s; return PICG_OK;
}

// Slab decomposition over `world` ranks on one node.  (1) every rank exports 128 bytes (the CUDA IPC handles of its phi and of
// its mailbox); (2) the caller all-gathers them (torch.distributed, MPI ...) and hands the table of world x 128 bytes back.
int picg_solver_slab_export(picg_solver_t s, void* handle128) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && handle128, "picg_solver_slab_export: null argument");
    if (!s->mbox) { CUDA_TRY(cudaMalloc(&s->mbox, MB_WORDS * 8)); CUDA_TRY(cudaMemset(s->mbox, 0, MB_WORDS * 8)); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, s->w->phi)); memcpy(handle128, &h, 64);
    CUDA_TRY(cudaIpcGetMemHandle(&h, s->mbox)); memcpy((char*)handle128 + 64, &h, 64);
    return PICG_OK;
}
int picg_solver_slab_range(picg_solver_t s, size_t* node_begin, size_t* node_end) {
    REQUIRE_ARG(s && node_begin && node_end, "picg_solver_slab_range: null argument");
    const Grid& g = s->w->g;
    const size_t plane = (size_t)g.nj * g.nk;
    if (s->slab_world > 1) { *node_begin = s->slab_i0 * plane; *node_end = s->slab_i1 * plane; }
    else { *node_begin = 0; *node_end = (size_t)g.nv; }
    return PICG_OK;
}
int picg_solver_slab_enable(picg_solver_t s, int rank, int world, const void* handles) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && handles && world >= 1 && world <= SLAB_MAX_RANKS && rank >= 0 && rank < world, "picg_solver_slab_enable: bad argument");
    REQUIRE_ARG(s->mbox, "picg_solver_slab_enable: call picg_solver_slab_export first");
    const Grid& g = s->w->g;
    REQUIRE_ARG(g.ni / world >= 2, "picg_solver_slab_enable: fewer than two planes per rank");
    s->peer_phi.assign(world, nullptr); s->peer_mbox.assign(world, nullptr);
    for (int r = 0; r < world; r++) {
        if (r == rank) { s->peer_phi[r] = s->w->phi; s->peer_mbox[r] = s->mbox; continue; }
        cudaIpcMemHandle_t h; void* p = nullptr;
        memcpy(&h, (const char*)handles + (size_t)r * 128, 64);
        CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess)); s->peer_phi[r] = (double*)p;
        memcpy(&h, (const char*)handles + (size_t)r * 128 + 64, 64);
        CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess)); s->peer_mbox[r] = (u64*)p;
    }
    CUDA_TRY(cudaMalloc(&s->peer_phi_dev, world * sizeof(double*))); CUDA_TRY(cudaMalloc(&s->peer_mbox_dev, world * sizeof(u64*)));
    CUDA_TRY(cudaMemcpy(s->peer_phi_dev, s->peer_phi.data(), world * sizeof(double*), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(s->peer_mbox_dev, s->peer_mbox.data(), world * sizeof(u64*), cudaMemcpyHostToDevice));
    drop_graphs(s);
    s->slab_rank = rank; s->slab_world = world;
    s->slab_i0 = (int)(((long long)g.ni * rank) / world); s->slab_i1 = (int)(((long long)g.ni * (rank + 1)) / world);
    return PICG_OK;
}
int picg_solver_set_reference(picg_solver_t s, double phi0, double n0, double Te0) {
    REQUIRE_ARG(s, "picg_solver_set_reference: null solver"); s->phi0 = phi0; s->n0 = n0; s->Te0 = Te0; return PICG_OK;
}
// 0 (default): two row sweeps per iteration (k_sor_row); 1: one plane-marching shared-memory pass per iteration (k_sor_tiled).  Same bits.
// Measured at 256^3 (profiles/r2_sor_tiled.md): the tiled pass moves 26 B per node and iteration instead of 33, but is issue-bound
// (index arithmetic of two phases + shared-memory traffic: 2.1 x the row sweeps' time), so the row sweeps stay the default.
// The slab-decomposed multi-GPU solve always uses its own row sweeps (k_sor_slab).
int picg_solver_set_sweep(picg_solver_t s, int mode) {
    REQUIRE_ARG(s && (mode == 0 || mode == 1), "picg_solver_set_sweep: mode must be 0 (row sweeps) or 1 (tiled one-pass sweep)");
    if (mode != s->sweep_mode) drop_graphs(s);
    s->sweep_mode = mode; return PICG_OK;
}
int picg_solver_set_boundary_mode(picg_solver_t s, int mode) {
    REQUIRE_ARG(s && (mode == 0 || mode == 1), "picg_solver_set_boundary_mode: mode must be 0 or 1"); s->bc_mode = mode; return PICG_OK;
}

int picg_solver_solve_gs(picg_solver_t s, int* converged, unsigned* iterations, double* L2_out) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_solver_solve_gs: null solver");
    SorParams p = make_params(s);
    { int rc0 = prepare_classes(s); if (rc0) return rc0; }
    double L2 = 0; bool conv = false; unsigned it = 0;               // it: iterations completed
    while (it < s->max_it) {
        // the reference checks the residual right after iteration index chk, chk % 25 == 0 (:124): run up to it in one batch
        const unsigned chk = ((it + 24) / 25) * 25, last = std::min(chk, s->max_it - 1);
        int rc = run_iterations(s, p, last - it + 1); if (rc) return rc;
        it = last + 1;
        if (last == chk) {
            rc = compute_residual(s, p, &L2); if (rc) return rc;
            if (L2 < s->tol) { conv = true; break; }
        }
    }
    { int rc = slab_allgather(s); if (rc) return rc; }
    if (converged) *converged = conv; if (iterations) *iterations = it; if (L2_out) *L2_out = L2;
    return PICG_OK;
}

int picg_solver_iterate(picg_solver_t s, unsigned n) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_solver_iterate: null solver");
    SorParams p = make_params(s);
    { int rc0 = prepare_classes(s); if (rc0) return rc0; }
    { int rc = run_iterations(s, p, n); if (rc) return rc; }
    return slab_allgather(s);
}

int picg_solver_residual(picg_solver_t s, double* L2) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s && L2, "picg_solver_residual: null argument");
    return compute_residual(s, make_params(s), L2);
}

int picg_solver_compute_ef(picg_solver_t s) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_solver_compute_ef: null solver");
    const Grid& g = s->w->g;
    LAUNCH(K_COMPUTE_EF, k_compute_ef, std::min(div_up(g.nv, 256), g_sm_count * 8), 256, 0, g, 1.0 / (2 * g.dx[0]), 1.0 / (2 * g.dx[1]),
           1.0 / (2 * g.dx[2]), s->w->phi, s->w->ef);
    CHECK_LAUNCH();
    return PICG_OK;
}

}  // extern "C"
