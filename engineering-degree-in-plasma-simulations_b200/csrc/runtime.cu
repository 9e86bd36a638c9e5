// Runtime plumbing of libpicgpu.so: device selection, the single stream, error reporting,
// launch counting and per-kernel CUDA-event timers.
#include "common.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace picg {
cudaStream_t g_stream = nullptr;
int g_device = -1;
int g_sm_count = PICG_SM_COUNT_FALLBACK;
uint64_t g_seed = 0x5EED0000ull;
int g_rank = 0, g_world_size = 1;
uint64_t g_reallocs = 0;
void note_realloc(const char* what, size_t bytes) {
    static const bool trace = getenv("PICG_TRACE_REALLOC") && atoi(getenv("PICG_TRACE_REALLOC")) != 0;
    g_reallocs++;
    if (trace) fprintf(stderr, "[picgpu] device allocation #%llu: %s, %.1f MB\n", (unsigned long long)g_reallocs, what, bytes / 1e6);
}
bool g_capturing = false;
static thread_local char g_err[1024] = "";
static uint64_t g_launches = 0;
static bool g_timers_on = false;
struct Pending { int id; cudaEvent_t a, b; };
static std::vector<Pending> g_pending;
static std::vector<cudaEvent_t> g_event_pool;
static double g_timer_ms[K_NUM_KERNELS];
static uint64_t g_timer_n[K_NUM_KERNELS];

int set_error(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    int code = (e == cudaErrorMemoryAllocation) ? PICG_ERR_OOM : PICG_ERR_CUDA;
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    cudaGetLastError();   // clear the sticky-less error state
    return code;
}
void count_launch(int id) { if (g_capturing) return; g_launches++; if (id >= 0 && id < K_NUM_KERNELS) g_timer_n[id]++; }

static cudaEvent_t get_event() {
    if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
TimerScope::TimerScope(int id_) : id(id_), on(g_timers_on && !g_capturing) {
    if (on) { a = get_event(); b = get_event(); cudaEventRecord(a, g_stream); }
}
TimerScope::~TimerScope() {
    if (on) { cudaEventRecord(b, g_stream); g_pending.push_back({id, a, b}); }
}
static void drain_timers() {
    if (g_pending.empty()) return;
    cudaStreamSynchronize(g_stream);
    for (Pending& p : g_pending) {
        float ms = 0; cudaEventElapsedTime(&ms, p.a, p.b);
        g_timer_ms[p.id] += ms;
        g_event_pool.push_back(p.a); g_event_pool.push_back(p.b);
    }
    g_pending.clear();
}

int ensure_scratch(picg_world_s* w, size_t bytes) {
    if (w->scratch_bytes >= bytes) return PICG_OK;
    if (w->scratch) { cudaStreamSynchronize(g_stream); cudaFree(w->scratch); w->scratch = nullptr; w->scratch_bytes = 0; }
    size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaMalloc(&w->scratch, want);
    if (e != cudaSuccess) { e = cudaMalloc(&w->scratch, bytes); want = bytes; }
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(scratch)", __FILE__, __LINE__);
    w->scratch_bytes = want; note_realloc("scratch arena", want);
    return PICG_OK;
}
}  // namespace picg

using namespace picg;

static const char* kKernelNames[K_NUM_KERNELS] = {
    "push_electrons", "push_electrons_deposit", "push_reflect", "push_heavy", "compact", "deposit_density",
    "finalize_density", "charge_density", "sor_redblack", "residual_l2", "compute_ef", "sort_keys", "sort_hist",
    "sort_scan", "sort_scatter", "sort_permute", "cell_start", "mc_ionize", "source_inject", "add_particles",
    "sample_moments", "count_per_cell", "transpose", "diagnostics", "misc", "push_heavy_deposit", "heavy_impacts", "dsmc_collide", "push_neutral", "nr_pcg", "deposit_tail", "mc_append", "sor_tiled"};

extern "C" {

int picg_device_count(int* n) {
    int c = 0; cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { cudaGetLastError(); c = 0; }
    if (n) *n = c;
    return PICG_OK;
}

int picg_init(int device) {
    int c = 0; picg_device_count(&c);
    if (c <= 0) return set_error(PICG_ERR_NO_DEVICE, "picg_init: no CUDA device visible (there is no CPU fallback)");
    if (device < 0 || device >= c) return set_error(PICG_ERR_ARG, "picg_init: device %d out of range (0..%d)", device, c - 1);
    if (g_device == device && g_stream) return PICG_OK;
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop; CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return set_error(PICG_ERR_NO_DEVICE, "picg_init: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    g_sm_count = prop.multiProcessorCount;
    if (g_stream) { cudaStreamDestroy(g_stream); g_stream = nullptr; }
    CUDA_TRY(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    g_device = device;
    memset(g_timer_ms, 0, sizeof(g_timer_ms)); memset(g_timer_n, 0, sizeof(g_timer_n));
    return PICG_OK;
}

int picg_shutdown(void) {
    if (g_device < 0) return PICG_OK;
    cudaStreamSynchronize(g_stream);
    for (cudaEvent_t e : g_event_pool) cudaEventDestroy(e);
    g_event_pool.clear();
    cudaStreamDestroy(g_stream); g_stream = nullptr; g_device = -1;
    return PICG_OK;
}

const char* picg_last_error(void) { return g_err; }
const char* picg_version(void) { return "picgpu 0.1 (sm_100a)"; }
void* picg_stream(void) { return (void*)g_stream; }
int picg_synchronize(void) { REQUIRE_DEVICE(); CUDA_TRY(cudaStreamSynchronize(g_stream)); return PICG_OK; }
int picg_seed(uint64_t seed) { g_seed = seed; return PICG_OK; }
int picg_set_rank(int rank, int world_size) {
    REQUIRE_ARG(world_size >= 1 && rank >= 0 && rank < world_size, "picg_set_rank: need 0 <= rank < world_size");
    g_rank = rank; g_world_size = world_size; return PICG_OK;
}
uint64_t picg_realloc_count(void) { return g_reallocs; }
uint64_t picg_launch_count(void) { return g_launches; }
void picg_launch_count_reset(void) { g_launches = 0; }
int picg_timers_enable(int on) {
    drain_timers();
    if (on && g_event_pool.size() < 8192) {            // create the events up front: cudaEventCreate inside a timed region would distort it
        g_event_pool.reserve(8192);
        while (g_event_pool.size() < 8192) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) break; g_event_pool.push_back(e); }
    }
    g_timers_on = on != 0; return PICG_OK;
}
int picg_timers_reset(void) { drain_timers(); memset(g_timer_ms, 0, sizeof(g_timer_ms)); memset(g_timer_n, 0, sizeof(g_timer_n)); return PICG_OK; }
int picg_timer_read(int id, double* total_ms, uint64_t* launches) {
    if (id < 0 || id >= K_NUM_KERNELS) return set_error(PICG_ERR_ARG, "picg_timer_read: bad kernel id %d", id);
    drain_timers();
    if (total_ms) *total_ms = g_timer_ms[id];
    if (launches) *launches = g_timer_n[id];
    return PICG_OK;
}
const char* picg_timer_name(int id) { return (id >= 0 && id < K_NUM_KERNELS) ? kKernelNames[id] : nullptr; }

}  // extern "C"
