// World: mesh geometry, node fields, object table, charge density.
// Replaces the device-relevant parts of ch4/v3/src/World.cpp (see include/picgpu.h for the map).
#include "common.cuh"
#include <cmath>
#include <cstring>

using namespace picg;

// ---------------------------------------------------------------- kernels
// World::computeChargeDensity (World.cpp:193-200): rho = 0; rho += q_s * den_s, species in call order.
struct ChargeArgs { int n; const double* den[8]; double q[8]; };
__global__ void __launch_bounds__(256) k_charge_density(int u_begin, int u_end, ChargeArgs a, double* __restrict__ rho) {
    for (int u = u_begin + blockIdx.x * blockDim.x + threadIdx.x; u < u_end; u += gridDim.x * blockDim.x) {
        double r = 0.0;
        for (int s = 0; s < a.n; s++) r = __dadd_rn(r, __dmul_rn(a.q[s], a.den[s][u]));
        rho[u] = r;
    }
}

// World::getPE (World.cpp:108-118): sum ef.ef*node_vol  (the 0.5*eps0 factor is applied on the host)
__global__ void __launch_bounds__(256) k_pe(int nv, const double* __restrict__ ef, const double* __restrict__ vol, double* __restrict__ out) {
    __shared__ double sm[256];
    double acc = 0.0;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < nv; u += gridDim.x * blockDim.x) {
        double ex = ef[3 * (size_t)u], ey = ef[3 * (size_t)u + 1], ez = ef[3 * (size_t)u + 2];
        acc += (ex * ex + ey * ey + ez * ez) * vol[u];
    }
    sm[threadIdx.x] = acc; __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s]; __syncthreads(); }
    if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

__global__ void k_int_to_double(int n, const int* __restrict__ in, double* __restrict__ out) {
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n; u += gridDim.x * blockDim.x) out[u] = (double)in[u];
}
__global__ void k_double_to_int(int n, const double* __restrict__ in, int* __restrict__ out) {
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n; u += gridDim.x * blockDim.x) out[u] = (int)in[u];
}

// ---------------------------------------------------------------- host
static int rect_in(const ObjShape& s, const double p[3]) {      // Rectangle::inObject Object.cpp:231-238
    for (int i = 0; i < 3; i++) if (std::fabs(p[i] - s.c[i]) > s.h[i]) return 0;
    return 1;
}
static int sphere_in(const ObjShape& s, const double p[3]) {    // Sphere::inObject Object.cpp:111-115
    double r[3] = {p[0] - s.c[0], p[1] - s.c[1], p[2] - s.c[2]};
    return (r[0] * r[0] + r[1] * r[1] + r[2] * r[2] <= s.h[0]) ? 1 : 0;
}

extern "C" {

int picg_world_create(int ni, int nj, int nk, const double x0[3], const double xm[3], picg_world_t* out) {
    REQUIRE_DEVICE();
    REQUIRE_ARG(out && x0 && xm, "picg_world_create: null argument");
    REQUIRE_ARG(ni >= 3 && nj >= 3 && nk >= 3, "picg_world_create: need at least 3 nodes per axis");
    REQUIRE_ARG((double)ni * nj * nk < 2147483647.0, "picg_world_create: node count exceeds int range (reference uses int nv, World.h:48)");
    picg_world_s* w = new picg_world_s();
    Grid& g = w->g;
    memset(&g, 0, sizeof(g));
    g.ni = ni; g.nj = nj; g.nk = nk; g.nv = ni * nj * nk;
    g.ci = ni - 1; g.cj = nj - 1; g.ck = nk - 1; g.nc = g.ci * g.cj * g.ck;
    auto magic = [](unsigned d, unsigned& mul, unsigned& shift) {      // n / d == (n * mul) >> shift for 0 <= n < 2^31
        unsigned l = 0; while ((1u << l) < d) l++;
        shift = 31 + l; mul = (unsigned)(((1ull << shift) / d) + 1);
    };
    magic((unsigned)g.ck, g.div_ck_mul, g.div_ck_shift); magic((unsigned)g.cj, g.div_cj_mul, g.div_cj_shift);
    int nn[3] = {ni, nj, nk};
    for (int a = 0; a < 3; a++) {                                // World::setExtents World.cpp:63-77
        g.x0[a] = x0[a]; g.xm[a] = xm[a];
        g.dx[a] = (xm[a] - x0[a]) / (nn[a] - 1);
        g.inv_dx[a] = 1 / g.dx[a];
    }
    size_t nv = g.nv;
    int rc = PICG_OK;
    cudaError_t e;
    if ((e = cudaMalloc(&w->phi, nv * 8)) != cudaSuccess || (e = cudaMalloc(&w->rho, nv * 8)) != cudaSuccess ||
        (e = cudaMalloc(&w->node_vol, nv * 8)) != cudaSuccess || (e = cudaMalloc(&w->ef, nv * 24)) != cudaSuccess ||
        (e = cudaMalloc(&w->object_id, nv * 4)) != cudaSuccess || (e = cudaMalloc(&w->node_type, nv * 4)) != cudaSuccess ||
        (e = cudaMalloc(&w->reduce_buf, 4096 * 8)) != cudaSuccess || (e = cudaMallocHost(&w->reduce_host, 4096 * 8)) != cudaSuccess) {
        rc = cuda_fail(e, "cudaMalloc(world fields)", __FILE__, __LINE__);
        picg_world_destroy(w); return rc;
    }
    cudaMemsetAsync(w->phi, 0, nv * 8, g_stream); cudaMemsetAsync(w->rho, 0, nv * 8, g_stream);
    cudaMemsetAsync(w->ef, 0, nv * 24, g_stream); cudaMemsetAsync(w->object_id, 0, nv * 4, g_stream);
    cudaMemsetAsync(w->node_type, 0, nv * 4, g_stream);
    // World::computeNodeVolumes World.cpp:353-367 (host, then upload)
    std::vector<double> vol(nv);
    double base = g.dx[0] * g.dx[1] * g.dx[2];
    size_t u = 0;
    for (int i = 0; i < ni; i++) for (int j = 0; j < nj; j++) for (int k = 0; k < nk; k++) {
        double v = base;
        if (i == 0 || i == ni - 1) v *= 0.5;
        if (j == 0 || j == nj - 1) v *= 0.5;
        if (k == 0 || k == nk - 1) v *= 0.5;
        vol[u++] = v;
    }
    CUDA_TRY(cudaMemcpyAsync(w->node_vol, vol.data(), nv * 8, cudaMemcpyHostToDevice, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    *out = w;
    return PICG_OK;
}

int picg_world_destroy(picg_world_t w) {
    if (!w) return PICG_OK;
    if (g_stream) cudaStreamSynchronize(g_stream);
    cudaFree(w->phi); cudaFree(w->rho); cudaFree(w->node_vol); cudaFree(w->ef); cudaFree(w->object_id); cudaFree(w->node_type);
    cudaFree(w->scratch); cudaFree(w->cbm); cudaFree(w->reduce_buf); if (w->reduce_host) cudaFreeHost(w->reduce_host);
    delete w;
    return PICG_OK;
}

int picg_world_set_time(picg_world_t w, double dt, int num_ts) {
    REQUIRE_ARG(w, "picg_world_set_time: null world");
    w->dt = dt; w->num_ts = num_ts; return PICG_OK;
}

int picg_world_add_rectangle(picg_world_t w, const double c[3], double phi, const double sides[3]) {
    REQUIRE_ARG(w && c && sides, "picg_world_add_rectangle: null argument");
    REQUIRE_ARG(w->g.n_obj < PICG_MAX_OBJECTS, "picg_world_add_rectangle: too many objects");
    ObjShape& s = w->g.obj[w->g.n_obj++];
    memset(&s, 0, sizeof(s));
    s.type = 0; s.phi = phi;
    for (int a = 0; a < 3; a++) {                                // Rectangle ctor Object.cpp:166-172
        s.c[a] = c[a];
        s.h[a] = sides[a] * 0.5;                                 // half_sides from the ctor argument (not |sides|)
        s.lo[a] = c[a] - s.h[a]; s.hi[a] = c[a] + s.h[a];
    }
    return PICG_OK;
}

int picg_world_add_sphere(picg_world_t w, const double c[3], double phi, double radius) {
    REQUIRE_ARG(w && c, "picg_world_add_sphere: null argument");
    REQUIRE_ARG(w->g.n_obj < PICG_MAX_OBJECTS, "picg_world_add_sphere: too many objects");
    ObjShape& s = w->g.obj[w->g.n_obj++];
    memset(&s, 0, sizeof(s));
    s.type = 1; s.phi = phi;
    for (int a = 0; a < 3; a++) s.c[a] = c[a];
    s.h[0] = radius * radius; s.h[1] = radius;
    return PICG_OK;
}

int picg_world_compute_object_id(picg_world_t w) {              // World::computeObjectID World.cpp:276-292
    REQUIRE_DEVICE();
    REQUIRE_ARG(w, "picg_world_compute_object_id: null world");
    const Grid& g = w->g;
    size_t nv = g.nv;
    std::vector<int> oid(nv), ntype(nv);
    std::vector<double> phi(nv);
    CUDA_TRY(cudaMemcpyAsync(oid.data(), w->object_id, nv * 4, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaMemcpyAsync(ntype.data(), w->node_type, nv * 4, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaMemcpyAsync(phi.data(), w->phi, nv * 8, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    for (int o = 0; o < g.n_obj; o++) {
        const ObjShape& s = g.obj[o];
        size_t u = 0;
        for (int i = 0; i < g.ni; i++) for (int j = 0; j < g.nj; j++) for (int k = 0; k < g.nk; k++, u++) {
            double p[3] = {g.x0[0] + (double)i * g.dx[0], g.x0[1] + (double)j * g.dx[1], g.x0[2] + (double)k * g.dx[2]};  // World::LtoX :153-160
            int in = s.type == 0 ? rect_in(s, p) : sphere_in(s, p);
            if (in) { oid[u] = 1; ntype[u] = 2 /*DIRICHLET*/; phi[u] = s.phi; }   // id is the bool
        }
    }
    CUDA_TRY(cudaMemcpyAsync(w->object_id, oid.data(), nv * 4, cudaMemcpyHostToDevice, g_stream));
    CUDA_TRY(cudaMemcpyAsync(w->node_type, ntype.data(), nv * 4, cudaMemcpyHostToDevice, g_stream));
    CUDA_TRY(cudaMemcpyAsync(w->phi, phi.data(), nv * 8, cudaMemcpyHostToDevice, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return PICG_OK;
}

static int world_field(picg_world_t w, int field, void** p, size_t* bytes, bool* is_int) {
    size_t nv = w->g.nv; *is_int = false;
    switch (field) {
        case PICG_F_PHI: *p = w->phi; *bytes = nv * 8; break;
        case PICG_F_RHO: *p = w->rho; *bytes = nv * 8; break;
        case PICG_F_NODE_VOL: *p = w->node_vol; *bytes = nv * 8; break;
        case PICG_F_EF: *p = w->ef; *bytes = nv * 24; break;
        case PICG_F_OBJECT_ID: *p = w->object_id; *bytes = nv * 4; *is_int = true; break;
        case PICG_F_NODE_TYPE: *p = w->node_type; *bytes = nv * 4; *is_int = true; break;
        default: return set_error(PICG_ERR_ARG, "unknown world field %d", field);
    }
    return PICG_OK;
}

int picg_world_device_ptr(picg_world_t w, int field, void** dptr, size_t* bytes) {
    REQUIRE_ARG(w && dptr, "picg_world_device_ptr: null argument");
    bool is_int; size_t b; int rc = world_field(w, field, dptr, &b, &is_int);
    if (bytes) *bytes = b;
    return rc;
}

int picg_world_download(picg_world_t w, int field, double* host) {
    REQUIRE_DEVICE();
    REQUIRE_ARG(w && host, "picg_world_download: null argument");
    void* p; size_t bytes; bool is_int;
    int rc = world_field(w, field, &p, &bytes, &is_int); if (rc) return rc;
    if (!is_int) {
        CUDA_TRY(cudaMemcpyAsync(host, p, bytes, cudaMemcpyDeviceToHost, g_stream));
    } else {
        int nv = w->g.nv;
        rc = ensure_scratch(w, (size_t)nv * 8); if (rc) return rc;
        LAUNCH(K_MISC, k_int_to_double, div_up(nv, 256), 256, 0, nv, (const int*)p, (double*)w->scratch); CHECK_LAUNCH();
        CUDA_TRY(cudaMemcpyAsync(host, w->scratch, (size_t)nv * 8, cudaMemcpyDeviceToHost, g_stream));
    }
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return PICG_OK;
}

// Asynchronous form of the download for callers that have other device work to queue meanwhile (the per-step read-back of rho next
// to the diagnostics pass): _begin orders the copy after everything queued so far and runs it on a copy stream, _end waits for it.
// Work that overwrites the field must not be queued before _end.  `host` should be page-locked; one download in flight.
static cudaStream_t g_copy_stream = nullptr; static cudaEvent_t g_copy_ready = nullptr, g_copy_done = nullptr;
int picg_world_download_begin(picg_world_t w, int field, double* host) {
    REQUIRE_DEVICE();
    REQUIRE_ARG(w && host, "picg_world_download_begin: null argument");
    void* p; size_t bytes; bool is_int;
    int rc = world_field(w, field, &p, &bytes, &is_int); if (rc) return rc;
    REQUIRE_ARG(!is_int, "picg_world_download_begin: integer fields go through picg_world_download");
    if (!g_copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&g_copy_ready, cudaEventDisableTiming)); CUDA_TRY(cudaEventCreateWithFlags(&g_copy_done, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventRecord(g_copy_ready, g_stream));
    CUDA_TRY(cudaStreamWaitEvent(g_copy_stream, g_copy_ready, 0));
    CUDA_TRY(cudaMemcpyAsync(host, p, bytes, cudaMemcpyDeviceToHost, g_copy_stream));
    CUDA_TRY(cudaEventRecord(g_copy_done, g_copy_stream));
    return PICG_OK;
}
int picg_world_download_end(picg_world_t w) {
    REQUIRE_DEVICE(); REQUIRE_ARG(w, "picg_world_download_end: null world");
    if (g_copy_done) CUDA_TRY(cudaEventSynchronize(g_copy_done));
    return PICG_OK;
}

int picg_world_upload(picg_world_t w, int field, const double* host) {
    REQUIRE_DEVICE();
    REQUIRE_ARG(w && host, "picg_world_upload: null argument");
    void* p; size_t bytes; bool is_int;
    int rc = world_field(w, field, &p, &bytes, &is_int); if (rc) return rc;
    if (!is_int) {
        CUDA_TRY(cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, g_stream));
    } else {
        int nv = w->g.nv;
        rc = ensure_scratch(w, (size_t)nv * 8); if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(w->scratch, host, (size_t)nv * 8, cudaMemcpyHostToDevice, g_stream));
        LAUNCH(K_MISC, k_double_to_int, div_up(nv, 256), 256, 0, nv, (const double*)w->scratch, (int*)p); CHECK_LAUNCH();
    }
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return PICG_OK;
}

static int charge_density_range(picg_world_t w, const picg_species_t* species, int n, size_t u_begin, size_t u_end) {
    ChargeArgs a; a.n = 0;
    for (int s = 0; s < n; s++) {
        REQUIRE_ARG(species[s] && species[s]->w == w, "picg_world_charge_density: species does not belong to this world");
        if (species[s]->charge == 0) continue;                   // World.cpp:197
        REQUIRE_ARG(a.n < 8, "picg_world_charge_density: more than 8 charged species");
        a.den[a.n] = species[s]->den; a.q[a.n] = species[s]->charge; a.n++;
    }
    u_end = std::min(u_end, (size_t)w->g.nv);
    if (u_begin >= u_end) return PICG_OK;
    int grid = std::min(div_up(u_end - u_begin, 256), g_sm_count * 8);
    LAUNCH(K_CHARGE_DENSITY, k_charge_density, grid, 256, 0, (int)u_begin, (int)u_end, a, w->rho); CHECK_LAUNCH();
    return PICG_OK;
}
int picg_world_charge_density(picg_world_t w, const picg_species_t* species, int n) {
    REQUIRE_DEVICE();
    REQUIRE_ARG(w && (species || n == 0), "picg_world_charge_density: null argument");
    return charge_density_range(w, species, n, 0, (size_t)-1);
}
int picg_world_charge_density_range(picg_world_t w, const picg_species_t* species, int n, size_t node_begin, size_t node_end) {
    REQUIRE_DEVICE();
    REQUIRE_ARG(w && (species || n == 0) && node_begin <= node_end, "picg_world_charge_density_range: bad argument");
    return charge_density_range(w, species, n, node_begin, node_end);
}

int picg_world_potential_energy(picg_world_t w, double* pe) {
    REQUIRE_DEVICE();
    REQUIRE_ARG(w && pe, "picg_world_potential_energy: null argument");
    int nv = w->g.nv;
    int grid = std::min(div_up(nv, 256), 1024);
    LAUNCH(K_DIAG, k_pe, grid, 256, 0, nv, w->ef, w->node_vol, w->reduce_buf); CHECK_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(w->reduce_host, w->reduce_buf, grid * 8, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    double s = 0; for (int i = 0; i < grid; i++) s += w->reduce_host[i];
    *pe = 0.5 * 8.85418782e-12 * s;                              // Const::eps_0 all.h:15
    return PICG_OK;
}

}  // extern "C"
