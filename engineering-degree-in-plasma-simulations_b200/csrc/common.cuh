// Internal definitions shared by the CUDA translation units of libpicgpu.so.
// Never include reference-style headers here (Vec3.h's `double3`/`int3` aliases
// collide with CUDA's built-ins, ch4/v3/src/Vec3.h:387-389).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>
#include "../../include/picgpu.h"

#define PICG_MAX_OBJECTS 8
#define PICG_SM_COUNT_FALLBACK 148

typedef long long i64;
typedef unsigned long long u64;

// ---------------------------------------------------------------- geometry
struct ObjShape {           // Object.h: Sphere / Rectangle
    int type;               // 0 rectangle, 1 sphere
    double c[3];            // Object::pos
    double h[3];            // Rectangle::half_sides (Object.cpp:161,168) ; sphere: h[0]=r^2 (r_squared), h[1]=r
    double lo[3], hi[3];    // Rectangle::x_min / x_max (Object.cpp:163-171)
    double phi;
};

struct Grid {               // passed by value to kernels (World.cpp:63-77)
    int ni, nj, nk, nv;
    int ci, cj, ck, nc;     // cells per axis (ni-1 ..), num_cells
    unsigned div_ck_mul, div_ck_shift, div_cj_mul, div_cj_shift;   // multiply-shift division by ck / cj (exact for numerators < 2^31)
    double x0[3], xm[3], dx[3], inv_dx[3];
    int n_obj;
    ObjShape obj[PICG_MAX_OBJECTS];
};

// ---------------------------------------------------------------- handles
struct SpeciesCounters {    // lives in device memory, one per species
    u64 n;                  // live particle count (authoritative)
    u64 n_dead;             // dead-list cursor of the current push
    u64 n_hole, n_surv;     // compaction cursors
    u64 overflow;           // appends dropped because capacity was exhausted
    i64 den_max;            // max fixed-point node sum of the last finalize
    u64 den_neg;            // number of negative (overflowed) nodes seen by finalize
    u64 n_movers;           // device-side mover count (sort.cu, cellstep.cu)
    u64 n_impact;           // heavy push: particles that ended their first sub-move inside an object (handled by k_heavy_impacts)
    u64 n_listed;           // movers of slots [0, n_listed) are listed in mv_trip (written by the deposit pass that listed them)
    u64 n_movers_dep;       // n_movers as the deposit pass left it (sort.cu: list builds append the tail behind it and restore it)
};

struct picg_world_s {
    Grid g;
    double dt = 1e-4; int num_ts = 0;
    uint32_t n_species = 0, live_species = 0;      // ids handed out / species alive: when the last one goes the ids start again (a re-created plasma draws the same streams)
    double *phi = nullptr, *rho = nullptr, *node_vol = nullptr, *ef = nullptr;   // ef: 3*nv interleaved
    int *object_id = nullptr, *node_type = nullptr;
    // scratch arena shared by sort / compaction (never live at the same time)
    void* scratch = nullptr; size_t scratch_bytes = 0;
    unsigned* cbm = nullptr; size_t cbm_words = 0;   // compaction bitmap: one bit per store slot, all zero between compactions (push.cu)
    double* reduce_buf = nullptr;      // small device buffer for reductions
    double* reduce_host = nullptr;     // pinned mirror
};

struct picg_species_s {
    picg_world_s* w;
    double mass, charge, mpw0;
    uint32_t id = 0;                   // index within its world; decorrelates the species' RNG streams
    uint32_t n_load_calls = 0, n_heavy_calls = 0, n_merge_calls = 0;   // call counters that address the Philox streams (reproducible per species)
    size_t cap = 0;                   // allocated particles per array
    size_t n_host = 0;                 // last count known to the host
    bool n_host_valid = true;
    size_t n_upper = 0;                // always >= the true device count (sizes scratch buffers)
    double* a[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // x y z u v w mpw (SoA)
    double* spare = nullptr;           // 8th array: out-of-place target of the sort permutation (rotated with a[c])
    SpeciesCounters* ctr = nullptr;    // device
    SpeciesCounters* ctr_host = nullptr; // pinned
    // node fields
    i64* den_fixed = nullptr; double* den = nullptr; double* den_avg = nullptr; int avg_samples = 0;
    double *T = nullptr, *vel = nullptr, *n_sum = nullptr, *nv_sum = nullptr, *nuu = nullptr, *nvv = nullptr, *nww = nullptr;
    double* macro_count = nullptr;     // cells, Field order
    int S = 0; bool S_pinned = false; bool S_calibrated = false;
    // cell-sorted layout
    unsigned* cell_start = nullptr;    // nc+1 entries, valid when sorted_valid
    bool sorted_valid = false;         // cell_start describes the current particle order exactly
    // exact per-cell lists on top of a stale partition (sort.cu: movers)
    unsigned *home = nullptr, *in_start = nullptr, *out_start = nullptr, *mv_in = nullptr;
    size_t home_cap = 0, lists_cap = 0, mv_cap = 0, mv_stride = 0;
    unsigned* home_alt = nullptr; size_t home_alt_cap = 0;   // second home array: target of the tail merge (sort.cu), swapped with home
    bool count_valid = false;          // macro_count holds the per-cell counts of the current particle positions (a deposit produces them for free)
    bool lists_valid = false;          // cell_start + in/out mover lists describe the current cell membership exactly
    // (slot, current cell, home cell) of the particles that left their slot's home cell: three arrays of mv_trip_cap entries
    unsigned* mv_trip = nullptr; size_t mv_trip_cap = 0;
    bool wants_lists = false;          // per-cell lists have been asked for (MC collisions): deposit passes emit the movers on the fly
    bool movers_saved = false;         // ctr->n_movers_dep holds the mover count of the deposit pass that set movers_fresh
    bool movers_fresh = false;         // mv_trip[0, ctr->n_movers) lists the movers of the partition for the current particle positions
    bool part_valid = false;           // cell_start is a partition of [0, part_n) (possibly stale: particles may have drifted)
    size_t part_n = 0;                 // upper bound of the particle count at the last sort
};

struct picg_solver_s {
    picg_world_s* w;
    unsigned max_it; double tol;
    double phi0 = 0, n0 = 0, Te0 = 1;
    int bc_mode = 0;
    double* partial = nullptr;         // residual partial sums
    unsigned char* cls = nullptr;      // node class per colour-compact node (poisson.cu), rebuilt at the start of every solve
    double* rho_split = nullptr;       // rho per colour-compact node: a colour half-sweep reads its own half with unit stride
    unsigned char* cls_nat = nullptr;  // node class in node order (the tiled one-pass sweep)
    double* phi_alt = nullptr;         // second potential buffer of the tiled sweep (phi is double-buffered within a batch of iterations)
    int sweep_mode = 0;                // 0: k_sor_row (two sweeps per iteration, default: faster), 1: k_sor_tiled (one pass, less DRAM traffic)
    void* pcg_work = nullptr; size_t pcg_bytes = 0;   // NR-PCG work vectors (nrpcg.cu), allocated on first use
    unsigned pcg_gs_fallbacks = 0;     // Newton steps of the last NR-PCG solve whose linear system was finished by the Gauss-Seidel fallback
    // multi-GPU slab decomposition (poisson.cu): planes [i0, i1) of the slowest index belong to this rank; halo planes,
    // residual sums and the final all-gather go through peer memory (CUDA IPC), signalled by flags in the mailboxes
    int slab_rank = 0, slab_world = 1, slab_i0 = 0, slab_i1 = 0;
    u64* mbox = nullptr;               // this rank's mailbox (device memory, exported to the peers)
    std::vector<double*> peer_phi;     // peer_phi[r]: rank r's phi (own pointer at r == rank)
    std::vector<u64*> peer_mbox;
    double** peer_phi_dev = nullptr; u64** peer_mbox_dev = nullptr;   // the same tables in device memory
    u64 residual_seq = 0, gather_seq = 0;                             // monotone sequence numbers, identical on every rank (the half-sweep counter lives in the mailbox)
    // CUDA graphs of n back-to-back iterations (2n half-sweep kernels), keyed by n and the parameter block they were captured with
    struct SorGraph { unsigned n; unsigned char params[160]; void* exec; };
    std::vector<SorGraph> graphs;
};

struct picg_mcc_s {
    picg_species_s *neu, *ion, *ele; picg_world_s* w;
    int n_table; double *tab_E = nullptr, *tab_s = nullptr;
    double E_ion_J;
    double* wsv = nullptr;             // device: [0] current max, [1] step max
    u64* stats = nullptr;              // device: candidates, collisions, ionizations
    u64 step = 0;
    size_t last_appends[3] = {0, 0, 0};   // neutrals, electrons, ions appended by the previous apply (capacity estimate)
    int fixed_weight = 0;              // 1: the fixed-weight algorithm of ch4/v2 (Interactions.cpp:566-641) instead of v3's variable weights
};

struct picg_dsmc_s {
    picg_species_s *sp1, *sp2; picg_world_s* w;      // sp2 == sp1: collisions within one species
    double* svm = nullptr;             // device: [0] sigma_v_rel_max in force, [1] largest value sampled by the current call
    u64* stats = nullptr;              // device: candidates, collisions
    u64 step = 0;
};

struct picg_source_s {
    picg_species_s* sp; picg_world_s* w;
    double v_drift, den, T; int face;
    double L[3]; double A; double num_micro;
    u64 step = 0;
};

// ---------------------------------------------------------------- runtime
namespace picg {
extern cudaStream_t g_stream;
extern int g_device;          // -1 until picg_init succeeds
extern int g_sm_count;
extern uint64_t g_seed;
extern int g_rank, g_world_size;
extern bool g_capturing;       // a stream capture is in progress: no event records, launches are counted by the replay
extern uint64_t g_reallocs;     // device (re)allocations of particle stores / scratch since start (a timed region should see none)
void note_realloc(const char* what, size_t bytes);   // counts a device (re)allocation; PICG_TRACE_REALLOC=1 prints it
int  set_error(int code, const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void count_launch(int kernel_id);
// timers
struct TimerScope { int id; bool on; cudaEvent_t a, b; TimerScope(int id); ~TimerScope(); };
int ensure_scratch(picg_world_s* w, size_t bytes);
int species_refresh_count(picg_species_s* s);     // syncs; updates n_host
int species_ensure_capacity(picg_species_s* s, size_t cap);
}

enum KernelId {
    K_PUSH_ELECTRONS = 0, K_PUSH_DEPOSIT, K_PUSH_REFLECT, K_PUSH_HEAVY, K_COMPACT, K_DEPOSIT, K_FINALIZE_DEN,
    K_CHARGE_DENSITY, K_SOR, K_RESIDUAL, K_COMPUTE_EF, K_SORT_KEYS, K_SORT_HIST, K_SORT_SCAN, K_SORT_SCATTER,
    K_SORT_PERMUTE, K_CELL_START, K_MCC, K_SOURCE, K_ADD_PARTICLES, K_MOMENTS, K_COUNT_CELLS, K_TRANSPOSE,
    K_DIAG, K_MISC, K_PUSH_HEAVY_DEPOSIT, K_HEAVY_IMPACTS, K_DSMC, K_PUSH_NEUTRAL, K_PCG, K_DEPOSIT_TAIL, K_MCC_APPEND, K_SOR_TILED, K_NUM_KERNELS
};

#define CUDA_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return picg::cuda_fail(_e, #expr, __FILE__, __LINE__); } while (0)
#define REQUIRE_DEVICE() do { if (picg::g_device < 0) return picg::set_error(PICG_ERR_NO_DEVICE, "no CUDA device: call picg_init() on a machine with a GPU (there is no CPU fallback)"); } while (0)
#define REQUIRE_ARG(cond, msg) do { if (!(cond)) return picg::set_error(PICG_ERR_ARG, "%s", msg); } while (0)
// launch bookkeeping: counts every kernel, optional CUDA-event timing per kernel id
#define LAUNCH(id, kernel, grid, block, smem, ...) do { picg::TimerScope _t(id); picg::count_launch(id); \
    kernel<<<grid, block, smem, picg::g_stream>>>(__VA_ARGS__); } while (0)
#define CHECK_LAUNCH() CUDA_TRY(cudaGetLastError())

static inline int div_up(size_t a, size_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
// World::XtoL (World.cpp:123-127): (x - x0) * inv_dx, no FMA so that the truncation and the
// fractional weights are the reference's bits.
__device__ __forceinline__ double x_to_l(double x, double x0, double inv_dx) { return __dmul_rn(__dsub_rn(x, x0), inv_dx); }

// World::inBounds (World.cpp:201-205): x0 <= p < xm on every axis
__device__ __forceinline__ bool in_bounds(const Grid& g, double x, double y, double z) {
    return !(x < g.x0[0] || x >= g.xm[0] || y < g.x0[1] || y >= g.xm[1] || z < g.x0[2] || z >= g.xm[2]);
}
// World::inObject (World.cpp:293-301) -> 1-based index of the first object containing p, 0 if none.
// Rectangle::inObject (Object.cpp:231-238): |x-c| > half  => outside (closed test)
// Sphere::inObject    (Object.cpp:111-115): r.r <= R^2
__device__ __forceinline__ int in_object(const Grid& g, double x, double y, double z) {
    for (int o = 0; o < g.n_obj; o++) {
        const ObjShape& s = g.obj[o];
        double rx = __dsub_rn(x, s.c[0]), ry = __dsub_rn(y, s.c[1]), rz = __dsub_rn(z, s.c[2]);
        if (s.type == 0) {
            if (!(fabs(rx) > s.h[0]) && !(fabs(ry) > s.h[1]) && !(fabs(rz) > s.h[2])) return o + 1;
        } else {
            double r2 = __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz));
            if (r2 <= s.h[0]) return o + 1;
        }
    }
    return 0;
}
// cell index used for the device-side sorted layout: k fastest, matching the Field layout so
// that consecutive cells touch consecutive nodes.  (The reference's XtoC is i-fastest,
// World.cpp:134-137; the order of cells is not observable through the API.)
__device__ __forceinline__ int cell_of(const Grid& g, int i, int j, int k) { return (i * g.cj + j) * g.ck + k; }

// Field<Vec3>::gather (Field.h:201-232): eight terms (F*w2)*w1 summed left to right.
__device__ __forceinline__ void gather_ef(const Grid& g, const double* __restrict__ ef, double lx, double ly, double lz,
                                          double& ex, double& ey, double& ez) {
    // the min() only acts when (x-x0)*inv_dx rounds up to exactly n-1 (reference: out-of-range read)
    int i = min((int)lx, g.ni - 2), j = min((int)ly, g.nj - 2), k = min((int)lz, g.nk - 2);
    double di = __dsub_rn(lx, (double)i), dj = __dsub_rn(ly, (double)j), dk = __dsub_rn(lz, (double)k);
    double odi = __dsub_rn(1.0, di), odj = __dsub_rn(1.0, dj), odk = __dsub_rn(1.0, dk);
    double wa = __dmul_rn(odi, odj), wb = __dmul_rn(odi, dj), wc = __dmul_rn(di, odj), wd = __dmul_rn(di, dj);
    size_t r00 = ((size_t)(i * g.nj + j) * g.nk + k) * 3;
    size_t r01 = r00 + (size_t)g.nk * 3;
    size_t r10 = r00 + (size_t)g.nj * g.nk * 3;
    size_t r11 = r10 + (size_t)g.nk * 3;
    double acc[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double v;
        v = __dmul_rn(__dmul_rn(__ldg(ef + r00 + c), wa), odk);
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(__ldg(ef + r00 + 3 + c), wa), dk));
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(__ldg(ef + r01 + c), wb), odk));
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(__ldg(ef + r01 + 3 + c), wb), dk));
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(__ldg(ef + r10 + c), wc), odk));
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(__ldg(ef + r10 + 3 + c), wc), dk));
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(__ldg(ef + r11 + c), wd), odk));
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(__ldg(ef + r11 + 3 + c), wd), dk));
        acc[c] = v;
    }
    ex = acc[0]; ey = acc[1]; ez = acc[2];
}

// Field<double>::scatter (Field.h:157-199): the eight contributions val*w, each formed with the
// reference's association ((val*wi)*wj)*wk, then quantised to fixed point: llrint(c * 2^S).
// q[] order: (i,j,k) (i,j,k+1) (i,j+1,k) (i,j+1,k+1) (i+1,j,k) (i+1,j,k+1) (i+1,j+1,k) (i+1,j+1,k+1)
__device__ __forceinline__ void scatter_weights_fixed(const Grid& g, double lx, double ly, double lz, double val, double scale,
                                                      int& i, int& j, int& k, i64 q[8]) {
    // the min() only acts when (x-x0)*inv_dx rounds up to exactly n-1 (reference: out-of-range write)
    i = min((int)lx, g.ni - 2); j = min((int)ly, g.nj - 2); k = min((int)lz, g.nk - 2);
    double di = __dsub_rn(lx, (double)i), dj = __dsub_rn(ly, (double)j), dk = __dsub_rn(lz, (double)k);
    double odi = __dsub_rn(1.0, di), odj = __dsub_rn(1.0, dj), odk = __dsub_rn(1.0, dk);
    // scale = 2^S is applied to val first: multiplying by a power of two commutes exactly with every rounding below, so
    // ((val*2^S)*wi*wj)*wk == ((val*wi*wj)*wk)*2^S bit for bit (no overflow/underflow in range) and 8 multiplications are saved
    const double vs = __dmul_rn(val, scale);
    double w00 = __dmul_rn(__dmul_rn(vs, odi), odj);
    double w01 = __dmul_rn(__dmul_rn(vs, odi), dj);
    double w10 = __dmul_rn(__dmul_rn(vs, di), odj);
    double w11 = __dmul_rn(__dmul_rn(vs, di), dj);
    q[0] = __double2ll_rn(__dmul_rn(w00, odk));
    q[1] = __double2ll_rn(__dmul_rn(w00, dk));
    q[2] = __double2ll_rn(__dmul_rn(w01, odk));
    q[3] = __double2ll_rn(__dmul_rn(w01, dk));
    q[4] = __double2ll_rn(__dmul_rn(w10, odk));
    q[5] = __double2ll_rn(__dmul_rn(w10, dk));
    q[6] = __double2ll_rn(__dmul_rn(w11, odk));
    q[7] = __double2ll_rn(__dmul_rn(w11, dk));
}
#endif
