// PotentialSolver on the device: red-black SOR with the reference's node classes, residual, E = -grad(phi).
//   k_sor_color   : one colour half-sweep of PotentialSolver::solveGS     ch4/v3/src/PotentialSolver.cpp:86-121
//   k_residual    : the convergence check (every 25 iterations)           :124-159
//   k_compute_ef  : PotentialSolver::computeEF                             :354-408
// The reference sweeps lexicographically (Gauss-Seidel); red-black ordering visits the same node
// classes with the same update formula, so the two iterations share their fixed point (parity is
// checked on converged solutions, SURVEY.md section 7 "hard parts").
// Algorithmic bytes: 25 B/node/iteration (phi R+W, rho R, mask R); computeEF 32 B/node.
#include "common.cuh"
#include <algorithm>
#include <cmath>

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

// node classes precomputed once per solve (1 byte per node) so that the sweeps do no index arithmetic or branching on geometry
__global__ void __launch_bounds__(256) k_node_classes(Grid g, int bc_mode, const int* __restrict__ object_id, unsigned char* __restrict__ cls) {
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < (size_t)g.nv; u += (size_t)gridDim.x * blockDim.x) {
        int k = (int)(u % g.nk); size_t row = u / g.nk; int j = (int)(row % g.nj), i = (int)(row / g.nj);
        cls[u] = (unsigned char)node_class(g, bc_mode, object_id[u], i, j, k);
    }
}
// One colour half-sweep, one block per (i,j) row: no divisions, coalesced along k, class byte instead of geometry tests.
__global__ void __launch_bounds__(128) k_sor_row(Grid g, SorParams sp, int color, double* __restrict__ phi, const double* __restrict__ rho,
                                                 const unsigned char* __restrict__ cls) {
    const int j = blockIdx.x, i = blockIdx.y;
    const size_t si = (size_t)g.nj * g.nk, sj = g.nk;
    const size_t row = ((size_t)i * g.nj + j) * g.nk;
    for (int k = 2 * threadIdx.x + ((i + j + color) & 1); k < g.nk; k += 2 * blockDim.x) {
        const size_t u = row + k;
        const int c = cls[u];
        if (c == 0) continue;
        if (c < 7) { phi[u] = phi[face_neighbor(g, c, u)]; continue; }
        const double p = phi[u];
        const double ne = (sp.n0 != 0.0) ? sp.n0 * exp((p - sp.phi0) / sp.Te0) : 0.0;
        const double nw = ((rho[u] - sp.qe * ne) * sp.inv_eps0 + (phi[u - si] + phi[u + si]) * sp.inv_d2x + (phi[u - sj] + phi[u + sj]) * sp.inv_d2y +
                           (phi[u - 1] + phi[u + 1]) * sp.inv_d2z) * sp.inv_twos;
        phi[u] = p + sp.w * (nw - p);
    }
}

__global__ void __launch_bounds__(256) k_sor_color(Grid g, SorParams sp, int color, double* __restrict__ phi, const double* __restrict__ rho,
                                                   const int* __restrict__ object_id) {
    const int hk = (g.nk + 1) >> 1;
    const size_t total = (size_t)g.ni * g.nj * hk;
    const size_t si = (size_t)g.nj * g.nk, sj = g.nk;
    for (size_t h = blockIdx.x * (size_t)blockDim.x + threadIdx.x; h < total; h += (size_t)gridDim.x * blockDim.x) {
        int t = (int)(h % hk); size_t row = h / hk;
        int j = (int)(row % g.nj), i = (int)(row / g.nj);
        int k = 2 * t + ((i + j + color) & 1);
        if (k >= g.nk) continue;
        size_t u = row * g.nk + k;
        int cls = node_class(g, sp.bc_mode, object_id[u], i, j, k);
        if (cls == 0) continue;
        if (cls < 7) { phi[u] = phi[face_neighbor(g, cls, u)]; continue; }
        double p = phi[u];
        double ne = (sp.n0 != 0.0) ? sp.n0 * exp((p - sp.phi0) / sp.Te0) : 0.0;
        double nw = ((rho[u] - sp.qe * ne) * sp.inv_eps0 + (phi[u - si] + phi[u + si]) * sp.inv_d2x + (phi[u - sj] + phi[u + sj]) * sp.inv_d2y +
                     (phi[u - 1] + phi[u + 1]) * sp.inv_d2z) * sp.inv_twos;
        phi[u] = p + sp.w * (nw - p);
    }
}

__global__ void __launch_bounds__(256) k_residual(Grid g, SorParams sp, const double* __restrict__ phi, const double* __restrict__ rho,
                                                  const int* __restrict__ object_id, double* __restrict__ partial) {
    __shared__ double sm[256];
    const size_t si = (size_t)g.nj * g.nk, sj = g.nk;
    double acc = 0.0;
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < (size_t)g.nv; u += (size_t)gridDim.x * blockDim.x) {
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
    if (!s->cls) { cudaError_t e = cudaMalloc(&s->cls, (size_t)g.nv); if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(node classes)", __FILE__, __LINE__); }
    LAUNCH(K_MISC, k_node_classes, std::min(div_up(g.nv, 256), g_sm_count * 8), 256, 0, g, s->bc_mode, s->w->object_id, s->cls); CHECK_LAUNCH();
    return PICG_OK;
}
static int launch_iteration(picg_solver_s* s, const SorParams& p) {
    const Grid& g = s->w->g;
    dim3 grid(g.nj, g.ni);
    LAUNCH(K_SOR, k_sor_row, grid, 128, 0, g, p, 0, s->w->phi, s->w->rho, s->cls); CHECK_LAUNCH();
    LAUNCH(K_SOR, k_sor_row, grid, 128, 0, g, p, 1, s->w->phi, s->w->rho, s->cls); CHECK_LAUNCH();
    return PICG_OK;
}
static int compute_residual(picg_solver_s* s, const SorParams& p, double* L2) {
    const Grid& g = s->w->g;
    int grid = std::min(div_up(g.nv, 256), kResidualBlocks);
    LAUNCH(K_RESIDUAL, k_residual, grid, 256, 0, g, p, s->w->phi, s->w->rho, s->w->object_id, s->partial); CHECK_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(s->w->reduce_host, s->partial, grid * 8, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    double sum = 0; for (int b = 0; b < grid; b++) sum += s->w->reduce_host[b];
    *L2 = std::sqrt(sum / g.nv);                                  // normalised by all nv nodes (:154)
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
int picg_solver_destroy(picg_solver_t s) { if (!s) return PICG_OK; if (g_stream) cudaStreamSynchronize(g_stream); cudaFree(s->partial); cudaFree(s->cls); delete s; return PICG_OK; }
int picg_solver_set_reference(picg_solver_t s, double phi0, double n0, double Te0) {
    REQUIRE_ARG(s, "picg_solver_set_reference: null solver"); s->phi0 = phi0; s->n0 = n0; s->Te0 = Te0; return PICG_OK;
}
int picg_solver_set_boundary_mode(picg_solver_t s, int mode) {
    REQUIRE_ARG(s && (mode == 0 || mode == 1), "picg_solver_set_boundary_mode: mode must be 0 or 1"); s->bc_mode = mode; return PICG_OK;
}

int picg_solver_solve_gs(picg_solver_t s, int* converged, unsigned* iterations, double* L2_out) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_solver_solve_gs: null solver");
    SorParams p = make_params(s);
    { int rc0 = prepare_classes(s); if (rc0) return rc0; }
    double L2 = 0; bool conv = false; unsigned it;
    for (it = 0; it < s->max_it; it++) {
        int rc = launch_iteration(s, p); if (rc) return rc;
        if (it % 25 == 0) {                                        // :124
            rc = compute_residual(s, p, &L2); if (rc) return rc;
            if (L2 < s->tol) { conv = true; it++; break; }
        }
    }
    if (converged) *converged = conv; if (iterations) *iterations = it; if (L2_out) *L2_out = L2;
    return PICG_OK;
}

int picg_solver_iterate(picg_solver_t s, unsigned n) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_solver_iterate: null solver");
    SorParams p = make_params(s);
    { int rc0 = prepare_classes(s); if (rc0) return rc0; }
    for (unsigned it = 0; it < n; it++) { int rc = launch_iteration(s, p); if (rc) return rc; }
    return PICG_OK;
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
