// Newton-Raphson + preconditioned conjugate gradients on the device, matrix-free.
//   picg_solver_solve_nrpcg : PotentialSolver::solveNRPCG      ch4/v3/src/PotentialSolver.cpp:178-240
//                             solvePCGlinear :242-313, solveGSlinear fallback :315-347, buildMatrix :414-470, vec::norm :36-38
// The reference assembles a sparse matrix A (Matrix.cpp) and multiplies with it; here a row of A is recomputed from the node
// type byte: DIRICHLET (object or inlet node) -> x_u; NEUMANN face (first match i0,iN,j0,jN,k0,kN) -> inv_d*(x_u - x_inward);
// REGULAR -> the 7-point Laplacian.  The Jacobian is J = A - diag(P), P = q_e n0/(eps0 Te0) exp((phi-phi0)/Te0) on REGULAR
// nodes; the preconditioner is the inverse diagonal of A (not of J), as in the reference.
// Reference quirk kept on purpose: buildMatrix puts inv_d2z on the i-neighbours and inv_d2x on the
// k-neighbours of a REGULAR row (:457-463); on meshes with dx != dz its PCG therefore solves another equation than solveGS.
// `xz_swap` (default on = the reference's matrix) selects it; off gives the operator solveGS relaxes.
// Reductions are two-stage with a fixed block count and a fixed summation order: the iteration is deterministic.  The
// reference's dot products run over the unknowns in i-fastest order; the order differs here, so iterates agree to rounding
// and converged potentials to the solver tolerance (tests compare at 1e-6 relative, the north-star bound for phi).
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>

using namespace picg;

namespace {
const int kRedBlocks = 512;

struct PcgGeom {
    int ni, nj, nk; size_t nv;
    double ci, cj, ck, diag;          // REGULAR row: coefficient of the i-, j-, k-neighbours and the diagonal (-twos)
    double inv_dx, inv_dy, inv_dz;    // NEUMANN rows
    double inv_eps0, qe, n0, phi0, Te0;
};

// node type byte: 0 DIRICHLET, 1..6 NEUMANN face (first matching rule, :432-455), 7 REGULAR
__global__ void __launch_bounds__(256) k_pcg_types(PcgGeom G, const int* __restrict__ object_id, const int* __restrict__ node_type, unsigned char* __restrict__ ty) {
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < G.nv; u += (size_t)gridDim.x * blockDim.x) {
        int k = (int)(u % G.nk); size_t r = u / G.nk; int j = (int)(r % G.nj), i = (int)(r / G.nj);
        unsigned char t;
        if (object_id[u] > 0 || node_type[u] == 2) t = 0;                          // World::DIRICHLET == 2 (World.h:17)
        else if (i == 0) t = 1; else if (i == G.ni - 1) t = 2; else if (j == 0) t = 3; else if (j == G.nj - 1) t = 4;
        else if (k == 0) t = 5; else if (k == G.nk - 1) t = 6; else t = 7;
        ty[u] = t;
    }
}
// row u of A applied to x
__device__ __forceinline__ double apply_row(const PcgGeom& G, int t, const double* __restrict__ x, size_t u) {
    const size_t si = (size_t)G.nj * G.nk, sj = G.nk;
    switch (t) {
        case 0: return x[u];
        case 1: return G.inv_dx * x[u] - G.inv_dx * x[u + si];
        case 2: return G.inv_dx * x[u] - G.inv_dx * x[u - si];
        case 3: return G.inv_dy * x[u] - G.inv_dy * x[u + sj];
        case 4: return G.inv_dy * x[u] - G.inv_dy * x[u - sj];
        case 5: return G.inv_dz * x[u] - G.inv_dz * x[u + 1];
        case 6: return G.inv_dz * x[u] - G.inv_dz * x[u - 1];
        default: return G.ck * (x[u - 1] + x[u + 1]) + G.cj * (x[u - sj] + x[u + sj]) + G.ci * (x[u - si] + x[u + si]) + G.diag * x[u];
    }
}
__device__ __forceinline__ double inv_diag(const PcgGeom& G, int t) {               // Matrix::invDiagonal of A (:468)
    switch (t) { case 0: return 1.0; case 1: case 2: return 1.0 / G.inv_dx; case 3: case 4: return 1.0 / G.inv_dy; case 5: case 6: return 1.0 / G.inv_dz; default: return 1.0 / G.diag; }
}
// block sums of up to three values -> part[q * kRedBlocks + block]
template <int NQ>
__device__ __forceinline__ void block_sums(double v0, double v1, double v2, double* __restrict__ part) {
    __shared__ double sm[NQ][256];
    sm[0][threadIdx.x] = v0; if (NQ > 1) sm[1 % NQ][threadIdx.x] = v1; if (NQ > 2) sm[2 % NQ][threadIdx.x] = v2;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) for (int q = 0; q < NQ; q++) sm[q][threadIdx.x] += sm[q][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) for (int q = 0; q < NQ; q++) part[q * kRedBlocks + blockIdx.x] = sm[q][0];
}
// final sums in block order -> out[q]
__global__ void __launch_bounds__(256) k_pcg_finish(int nq, const double* __restrict__ part, double* __restrict__ out) {
    __shared__ double sm[256];
    for (int q = 0; q < nq; q++) {
        double a = 0; for (int b = threadIdx.x; b < kRedBlocks; b += 256) a += part[q * kRedBlocks + b];
        sm[threadIdx.x] = a; __syncthreads();
        for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s]; __syncthreads(); }
        if (threadIdx.x == 0) out[q] = sm[0];
        __syncthreads();
    }
}

// Newton step set-up (:185-216): b_ (the reference's rho_), F = A phi - b_ - b(phi), P; y keeps its value (initial guess of PCG)
__global__ void __launch_bounds__(256) k_nr_residual(PcgGeom G, const unsigned char* __restrict__ ty, const double* __restrict__ phi, const double* __restrict__ rho,
                                                     double* __restrict__ F, double* __restrict__ P) {
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < G.nv; u += (size_t)gridDim.x * blockDim.x) {
        const int t = ty[u];
        const double b = t == 0 ? phi[u] : (t < 7 ? 0.0 : -rho[u] * G.inv_eps0);    // :185-189
        double f = apply_row(G, t, phi, u) - b, p = 0.0;
        if (t == 7) {
            const double e = exp((phi[u] - G.phi0) / G.Te0);
            f -= G.qe * G.n0 * e * G.inv_eps0;                                        // :199-203
            p = G.qe * G.n0 / (8.85418782e-12 * G.Te0) * e;                           // :205-212
        }
        F[u] = f; P[u] = p;
    }
}
// g = J x - b ; s = M g ; d = -s ; partial sums of g.s   (:246-252)
__global__ void __launch_bounds__(256) k_pcg_start(PcgGeom G, const unsigned char* __restrict__ ty, const double* __restrict__ P, const double* __restrict__ x,
                                                   const double* __restrict__ b, double* __restrict__ g, double* __restrict__ s, double* __restrict__ d, double* __restrict__ part) {
    double gs = 0;
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < G.nv; u += (size_t)gridDim.x * blockDim.x) {
        const int t = ty[u];
        const double gv = apply_row(G, t, x, u) - P[u] * x[u] - b[u], sv = inv_diag(G, t) * gv;
        g[u] = gv; s[u] = sv; d[u] = -1 * sv; gs += gv * sv;
    }
    block_sums<1>(gs, 0, 0, part);
}
// z = J d ; partial sums of d.z   (:268-271)
__global__ void __launch_bounds__(256) k_pcg_matvec(PcgGeom G, const unsigned char* __restrict__ ty, const double* __restrict__ P, const double* __restrict__ d,
                                                    double* __restrict__ z, double* __restrict__ part) {
    double dz = 0;
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < G.nv; u += (size_t)gridDim.x * blockDim.x) {
        const double zv = apply_row(G, ty[u], d, u) - P[u] * d[u];
        z[u] = zv; dz += d[u] * zv;
    }
    block_sums<1>(dz, 0, 0, part);
}
// x += (alpha/beta) d ; g += (alpha/beta) z ; s = M g ; partial sums of g.s and g.g   (:281-291).  sc[0] = alpha, sc[1] = beta
__global__ void __launch_bounds__(256) k_pcg_update(PcgGeom G, const unsigned char* __restrict__ ty, const double* __restrict__ sc, const double* __restrict__ d,
                                                    const double* __restrict__ z, double* __restrict__ x, double* __restrict__ g, double* __restrict__ s, double* __restrict__ part) {
    const double a = sc[0] / sc[1];
    double gs = 0, gg = 0;
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < G.nv; u += (size_t)gridDim.x * blockDim.x) {
        x[u] = x[u] + a * d[u];
        const double gv = g[u] + a * z[u], sv = inv_diag(G, ty[u]) * gv;
        g[u] = gv; s[u] = sv; gs += gv * sv; gg += gv * gv;
    }
    block_sums<2>(gs, gg, 0, part);
}
// d = (alpha_new/alpha_old) d - s   (:287-289).  sc[2] = alpha_new, sc[0] = alpha_old
__global__ void __launch_bounds__(256) k_pcg_direction(size_t nv, const double* __restrict__ sc, const double* __restrict__ s, double* __restrict__ d) {
    const double b = sc[2] / sc[0];
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < nv; u += (size_t)gridDim.x * blockDim.x) d[u] = b * d[u] - s[u];
}
__global__ void k_pcg_shift(double* sc) { sc[0] = sc[2]; }                         // alpha <- alpha_new for the next iteration
// One colour of a Gauss-Seidel sweep on J y = F (fallback, :315-347; red-black instead of lexicographic: same fixed point)
__global__ void __launch_bounds__(256) k_gs_linear(PcgGeom G, int color, const unsigned char* __restrict__ ty, const double* __restrict__ P, const double* __restrict__ b,
                                                   double* __restrict__ x) {
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < G.nv; u += (size_t)gridDim.x * blockDim.x) {
        int k = (int)(u % G.nk); size_t r = u / G.nk; int j = (int)(r % G.nj), i = (int)(r / G.nj);
        if (((i + j + k) & 1) != color) continue;
        const int t = ty[u];
        const double dg = 1.0 / inv_diag(G, t) - P[u];
        const double S = apply_row(G, t, x, u) - P[u] * x[u] - dg * x[u];
        x[u] = (b[u] - S) / dg;
    }
}
// R = J x - b, partial sums of R.R
__global__ void __launch_bounds__(256) k_lin_residual(PcgGeom G, const unsigned char* __restrict__ ty, const double* __restrict__ P, const double* __restrict__ b,
                                                      const double* __restrict__ x, double* __restrict__ part) {
    double rr = 0;
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < G.nv; u += (size_t)gridDim.x * blockDim.x) {
        const double R = apply_row(G, ty[u], x, u) - P[u] * x[u] - b[u];
        rr += R * R;
    }
    block_sums<1>(rr, 0, 0, part);
}
// Newton update (:221-229): y = 0 on DIRICHLET nodes, phi -= y, partial sums of y.y
__global__ void __launch_bounds__(256) k_nr_update(PcgGeom G, const unsigned char* __restrict__ ty, double* __restrict__ y, double* __restrict__ phi, double* __restrict__ part) {
    double yy = 0;
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < G.nv; u += (size_t)gridDim.x * blockDim.x) {
        double v = y[u];
        if (ty[u] == 0) { v = 0; y[u] = 0; }
        phi[u] = phi[u] - v; yy += v * v;
    }
    block_sums<1>(yy, 0, 0, part);
}

struct Work { double *y, *F, *P, *g, *s, *d, *z, *part, *sc; unsigned char* ty; };

int read_scalars(picg_solver_s* s, const double* dev, int n, double* host) {
    CUDA_TRY(cudaMemcpyAsync(s->w->reduce_host, dev, n * 8, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    for (int i = 0; i < n; i++) host[i] = s->w->reduce_host[i];
    return PICG_OK;
}
}  // namespace

extern "C" {

int picg_solver_solve_nrpcg(picg_solver_t s, int xz_swap, unsigned pcg_max_it, int* converged, unsigned* nr_iterations, unsigned* pcg_iterations, double* norm_out) {
    REQUIRE_DEVICE(); REQUIRE_ARG(s, "picg_solver_solve_nrpcg: null solver");
    REQUIRE_ARG(s->slab_world <= 1, "picg_solver_solve_nrpcg: not available on a slab-decomposed solver (use picg_solver_solve_gs)");
    const Grid& g = s->w->g; const size_t nv = (size_t)g.nv;
    PcgGeom G; G.ni = g.ni; G.nj = g.nj; G.nk = g.nk; G.nv = nv;
    const double inv_d2x = 1.0 / (g.dx[0] * g.dx[0]), inv_d2y = 1.0 / (g.dx[1] * g.dx[1]), inv_d2z = 1.0 / (g.dx[2] * g.dx[2]);   // precalculate :473-491
    G.ci = xz_swap ? inv_d2z : inv_d2x; G.cj = inv_d2y; G.ck = xz_swap ? inv_d2x : inv_d2z;                                           // buildMatrix :457-463
    G.diag = -(2.0 * (inv_d2x + inv_d2y + inv_d2z));
    G.inv_dx = 1.0 / g.dx[0]; G.inv_dy = 1.0 / g.dx[1]; G.inv_dz = 1.0 / g.dx[2];
    G.inv_eps0 = 1.0 / 8.85418782e-12; G.qe = 1.602176565e-19; G.n0 = s->n0; G.phi0 = s->phi0; G.Te0 = s->Te0;
    // work vectors: 7 x nv doubles + type bytes + reduction partials, in one allocation kept by the solver
    const size_t need = 7 * nv * 8 + ((nv + 255) & ~(size_t)255) + (3 * kRedBlocks + 8) * 8;
    if (s->pcg_bytes < need) {
        if (s->pcg_work) { CUDA_TRY(cudaStreamSynchronize(g_stream)); cudaFree(s->pcg_work); s->pcg_work = nullptr; s->pcg_bytes = 0; }
        cudaError_t e = cudaMalloc(&s->pcg_work, need);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(NR-PCG work vectors)", __FILE__, __LINE__);
        s->pcg_bytes = need;
    }
    Work W; double* base = (double*)s->pcg_work;
    W.y = base; W.F = base + nv; W.P = base + 2 * nv; W.g = base + 3 * nv; W.s = base + 4 * nv; W.d = base + 5 * nv; W.z = base + 6 * nv;
    W.part = base + 7 * nv; W.sc = W.part + 3 * kRedBlocks; W.ty = (unsigned char*)(W.sc + 8);
    const int grid = std::max(1, std::min(div_up(nv, 256), kRedBlocks));
    double* phi = s->w->phi;
    CUDA_TRY(cudaMemsetAsync(W.y, 0, nv * 8, g_stream));                            // tcvector y(n_unknowns), :184
    CUDA_TRY(cudaMemsetAsync(W.part, 0, 3 * kRedBlocks * 8, g_stream));             // blocks beyond `grid` contribute zero
    LAUNCH(K_PCG, k_pcg_types, grid, 256, 0, G, s->w->object_id, s->w->node_type, W.ty); CHECK_LAUNCH();
    const unsigned pcg_max = pcg_max_it ? pcg_max_it : s->max_it, gs_max = 20 * pcg_max;   // PotentialSolver.cpp:45-46
    static const bool trace = getenv("PICG_TRACE_PCG") && atoi(getenv("PICG_TRACE_PCG")) != 0;
    bool conv = false; double norm = 0; unsigned nr_it = 0, pcg_total = 0, gs_fallbacks = 0;
    for (int it = 0; it < 20; it++) {                                               // NR_MAX_IT :179
        nr_it++;
        LAUNCH(K_PCG, k_nr_residual, grid, 256, 0, G, W.ty, phi, s->w->rho, W.F, W.P); CHECK_LAUNCH();
        // ---- solvePCGlinear(J, y, F)
        bool lin_conv = false; double gg_first = 0, gg_last = 0;
        LAUNCH(K_PCG, k_pcg_start, grid, 256, 0, G, W.ty, W.P, W.y, W.F, W.g, W.s, W.d, W.part); CHECK_LAUNCH();
        LAUNCH(K_PCG, k_pcg_finish, 1, 256, 0, 1, W.part, W.sc); CHECK_LAUNCH();                       // sc[0] = alpha = g.s
        for (unsigned p = 0; p < pcg_max; p++) {
            pcg_total++;
            LAUNCH(K_PCG, k_pcg_matvec, grid, 256, 0, G, W.ty, W.P, W.d, W.z, W.part); CHECK_LAUNCH();
            LAUNCH(K_PCG, k_pcg_finish, 1, 256, 0, 1, W.part, W.sc + 1); CHECK_LAUNCH();               // sc[1] = beta = d.z
            LAUNCH(K_PCG, k_pcg_update, grid, 256, 0, G, W.ty, W.sc, W.d, W.z, W.y, W.g, W.s, W.part); CHECK_LAUNCH();
            LAUNCH(K_PCG, k_pcg_finish, 1, 256, 0, 2, W.part, W.sc + 2); CHECK_LAUNCH();               // sc[2] = alpha_new, sc[3] = g.g
            LAUNCH(K_PCG, k_pcg_direction, grid, 256, 0, nv, W.sc, W.s, W.d); CHECK_LAUNCH();
            LAUNCH(K_PCG, k_pcg_shift, 1, 1, 0, W.sc); CHECK_LAUNCH();
            double sc3[3]; int rc = read_scalars(s, W.sc + 1, 3, sc3); if (rc) return rc;                   // beta, alpha_new, g.g
            if (trace && (p < 5 || p % 500 == 0 || sc3[0] == 0)) fprintf(stderr, "[nrpcg] newton %d pcg %u: beta %.6e alpha_new %.6e g.g %.6e\n", it, p, sc3[0], sc3[1], sc3[2]);
            // beta == 0: the reference throws std::runtime_error here (:274-278), out of a noexcept call chain.  It only happens once the
            // diverging iterates have overflowed; it is counted as a PCG failure and the Gauss-Seidel fallback finishes the Newton step.
            if (sc3[0] == 0 || !std::isfinite(sc3[0])) { gg_last = HUGE_VAL; break; }
            if (std::sqrt(sc3[2] / (double)nv) < s->tol) { lin_conv = true; break; }                    // vec::norm(g) < tolerance :291-303
            if (p == 0) gg_first = sc3[2];
            gg_last = sc3[2];
            if (!std::isfinite(sc3[2])) break;                                                          // diverged: let the fallback repair it
        }
        if (!lin_conv) {                                                            // solveGSlinear(J, y, F) :215-217
            // A is not symmetric (NEUMANN rows), so CG often diverges - in the reference too, which then relaxes from the diverged
            // iterate (norm(g) ~ 1e16 in its own runs) and still reaches the fixed point within its 20 x max_it sweeps.  Same fixed
            // point, shorter way: a diverged iterate is dropped and the relaxation starts from zero.
            if (!std::isfinite(gg_last) || gg_last > gg_first) CUDA_TRY(cudaMemsetAsync(W.y, 0, nv * 8, g_stream));
            gs_fallbacks++;
            for (unsigned q = 0; q < gs_max; q++) {
                LAUNCH(K_PCG, k_gs_linear, grid, 256, 0, G, 0, W.ty, W.P, W.F, W.y); CHECK_LAUNCH();
                LAUNCH(K_PCG, k_gs_linear, grid, 256, 0, G, 1, W.ty, W.P, W.F, W.y); CHECK_LAUNCH();
                if (q % 25 == 0) {
                    LAUNCH(K_PCG, k_lin_residual, grid, 256, 0, G, W.ty, W.P, W.F, W.y, W.part); CHECK_LAUNCH();
                    LAUNCH(K_PCG, k_pcg_finish, 1, 256, 0, 1, W.part, W.sc + 4); CHECK_LAUNCH();
                    double rr; int rc = read_scalars(s, W.sc + 4, 1, &rr); if (rc) return rc;
                    if (std::sqrt(rr / (double)nv) < s->tol) break;
                }
            }
        }
        LAUNCH(K_PCG, k_nr_update, grid, 256, 0, G, W.ty, W.y, phi, W.part); CHECK_LAUNCH();
        LAUNCH(K_PCG, k_pcg_finish, 1, 256, 0, 1, W.part, W.sc + 5); CHECK_LAUNCH();
        double yy; int rc = read_scalars(s, W.sc + 5, 1, &yy); if (rc) return rc;
        norm = std::sqrt(yy / (double)nv);
        if (norm < 1e-3) { conv = true; break; }                                    // NR_TOL :180
    }
    s->pcg_gs_fallbacks = gs_fallbacks;
    if (converged) *converged = conv; if (nr_iterations) *nr_iterations = nr_it; if (pcg_iterations) *pcg_iterations = pcg_total; if (norm_out) *norm_out = norm;
    return PICG_OK;
}

}  // extern "C"
