// Binary outputs and restart files: the device-side replacement of the reference's ASCII writers.
//   picg_write_fields_vti : Output::fieldsOutput    ch4/v3/src/Outputs.cpp:9-123  (same VTK ImageData arrays, names and order;
//                           appended raw binary instead of ASCII: a 256^3 mesh is 3 GB of text per snapshot in the reference)
//   picg_checkpoint_save / picg_checkpoint_load : the reference has no restart format; this one holds what the time loop
//                           carries from step to step (fields, particle stores, averages and moment sums, RNG stream positions).
// Fields leave the device through a pinned staging buffer in chunks; the VTK point order (i fastest, Field.h:676-690
// operator<<) is produced on the device by a transpose kernel, so the host only streams bytes to the file.
#include "common.cuh"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>

using namespace picg;

namespace {
const size_t kStageBytes = (size_t)32 << 20;

struct Stage {                                   // pinned host buffer + FILE, RAII
    void* host = nullptr; FILE* f = nullptr;
    ~Stage() { if (host) cudaFreeHost(host); if (f) fclose(f); }
};

// Field order (k fastest) -> VTK order (i fastest), `comps` interleaved components per node
__global__ void __launch_bounds__(256) k_to_vtk_order(int ni, int nj, int nk, int comps, const double* __restrict__ in, double* __restrict__ out) {
    const size_t n = (size_t)ni * nj * nk;
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < n; v += (size_t)gridDim.x * blockDim.x) {
        int i = (int)(v % ni); size_t r = v / ni; int j = (int)(r % nj), k = (int)(r / nj);
        const size_t u = ((size_t)i * nj + j) * nk + k;
        for (int c = 0; c < comps; c++) out[v * comps + c] = in[u * comps + c];
    }
}
__global__ void __launch_bounds__(256) k_int_to_vtk_order(int ni, int nj, int nk, const int* __restrict__ in, double* __restrict__ out) {
    const size_t n = (size_t)ni * nj * nk;
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < n; v += (size_t)gridDim.x * blockDim.x) {
        int i = (int)(v % ni); size_t r = v / ni; int j = (int)(r % nj), k = (int)(r / nj);
        out[v] = (double)in[((size_t)i * nj + j) * nk + k];
    }
}

// device bytes -> file, through the pinned stage
int dev_to_file(Stage& st, const void* dev, size_t bytes) {
    for (size_t off = 0; off < bytes; off += kStageBytes) {
        const size_t m = std::min(kStageBytes, bytes - off);
        CUDA_TRY(cudaMemcpyAsync(st.host, (const char*)dev + off, m, cudaMemcpyDeviceToHost, g_stream));
        CUDA_TRY(cudaStreamSynchronize(g_stream));
        if (fwrite(st.host, 1, m, st.f) != m) return set_error(PICG_ERR_IO, "write failed (disk full?)");
    }
    return PICG_OK;
}
int file_to_dev(Stage& st, void* dev, size_t bytes) {
    for (size_t off = 0; off < bytes; off += kStageBytes) {
        const size_t m = std::min(kStageBytes, bytes - off);
        if (fread(st.host, 1, m, st.f) != m) return set_error(PICG_ERR_IO, "checkpoint file is truncated");
        CUDA_TRY(cudaMemcpyAsync((char*)dev + off, st.host, m, cudaMemcpyHostToDevice, g_stream));
        CUDA_TRY(cudaStreamSynchronize(g_stream));
    }
    return PICG_OK;
}
int open_stage(Stage& st, const char* path, const char* mode) {
    st.f = fopen(path, mode);
    if (!st.f) return set_error(PICG_ERR_IO, "could not open %s", path);
    cudaError_t e = cudaMallocHost(&st.host, kStageBytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMallocHost(io stage)", __FILE__, __LINE__);
    return PICG_OK;
}

struct VtkArray { std::string name; int comps; const double* dptr; const int* iptr; bool cell; };

// ---- checkpoint records (little-endian, fixed layout) ----
struct CkpHeader {
    char magic[8]; uint64_t version, user_ts, seed;
    int32_t ni, nj, nk, pad; double x0[3], xm[3], dt;
    uint32_t n_species, n_mcc, n_dsmc, n_sources;
};
struct CkpSpecies {
    uint64_t n; double mass, charge, mpw0;
    int32_t S, avg_samples; uint8_t S_pinned, S_calibrated, pad[6];
    uint32_t n_load_calls, n_heavy_calls, n_merge_calls, pad2;
};
struct CkpScalar { double value; uint64_t step; };
const char kMagic[8] = {'P', 'I', 'C', 'G', 'C', 'K', 'P', '1'};
}  // namespace

extern "C" {

int picg_write_fields_vti(const char* path, picg_world_t w, const picg_species_t* species, const char* const* names, int n) {
    REQUIRE_DEVICE(); REQUIRE_ARG(path && w && (n == 0 || (species && names)) && n >= 0, "picg_write_fields_vti: bad argument");
    const Grid& g = w->g;
    const size_t nv = (size_t)g.nv, nc = (size_t)g.nc;
    std::vector<VtkArray> arr;                                                  // Outputs.cpp:38-112, in the reference's order
    arr.push_back({"NodeVol", 1, w->node_vol, nullptr, false});
    arr.push_back({"ObjectID", 1, nullptr, w->object_id, false});
    arr.push_back({"NodeType", 1, nullptr, w->node_type, false});
    arr.push_back({"phi", 1, w->phi, nullptr, false});
    arr.push_back({"rho", 1, w->rho, nullptr, false});
    for (int s = 0; s < n; s++) arr.push_back({std::string("nd.") + names[s], 1, species[s]->den, nullptr, false});
    for (int s = 0; s < n; s++) arr.push_back({std::string("avg_nd.") + names[s], 1, species[s]->den_avg, nullptr, false});
    for (int s = 0; s < n; s++) arr.push_back({std::string("vel.") + names[s], 3, species[s]->vel, nullptr, false});
    for (int s = 0; s < n; s++) arr.push_back({std::string("T.") + names[s], 1, species[s]->T, nullptr, false});
    arr.push_back({"ef", 3, w->ef, nullptr, false});
    for (int s = 0; s < n; s++) arr.push_back({std::string("mpc.") + names[s], 1, species[s]->macro_count, nullptr, true});
    int rc = ensure_scratch(w, nv * 3 * 8); if (rc) return rc;
    Stage st; rc = open_stage(st, path, "wb"); if (rc) return rc;
    // XML head: every DataArray points into the appended block (UInt64 byte count, then the raw doubles)
    std::string head = "<VTKFile type=\"ImageData\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n";
    char buf[512];
    snprintf(buf, sizeof buf, "<ImageData Origin=\"%.17g %.17g %.17g\" Spacing=\"%.17g %.17g %.17g\" WholeExtent=\"0 %d 0 %d 0 %d\">\n",
             g.x0[0], g.x0[1], g.x0[2], g.dx[0], g.dx[1], g.dx[2], g.ni - 1, g.nj - 1, g.nk - 1);
    head += buf;
    snprintf(buf, sizeof buf, "<Piece Extent=\"0 %d 0 %d 0 %d\">\n", g.ni - 1, g.nj - 1, g.nk - 1); head += buf;
    size_t offset = 0; bool in_cells = false;
    head += "<PointData>\n";
    for (const VtkArray& a : arr) {
        if (a.cell && !in_cells) { head += "</PointData>\n<CellData>\n"; in_cells = true; }
        snprintf(buf, sizeof buf, "<DataArray Name=\"%s\" NumberOfComponents=\"%d\" format=\"appended\" type=\"Float64\" offset=\"%zu\"/>\n", a.name.c_str(), a.comps, offset);
        head += buf;
        offset += 8 + (a.cell ? nc : nv) * a.comps * 8;
    }
    head += in_cells ? "</CellData>\n" : "</PointData>\n<CellData>\n</CellData>\n";
    head += "</Piece>\n</ImageData>\n<AppendedData encoding=\"raw\">\n_";
    if (fwrite(head.data(), 1, head.size(), st.f) != head.size()) return set_error(PICG_ERR_IO, "write failed: %s", path);
    double* tmp = (double*)w->scratch;
    for (const VtkArray& a : arr) {
        const int ni = a.cell ? g.ci : g.ni, nj = a.cell ? g.cj : g.nj, nk = a.cell ? g.ck : g.nk;
        const size_t cnt = (size_t)ni * nj * nk;
        const int grid = std::min(div_up(cnt, 256), g_sm_count * 8);
        if (a.iptr) LAUNCH(K_MISC, k_int_to_vtk_order, grid, 256, 0, ni, nj, nk, a.iptr, tmp);
        else LAUNCH(K_MISC, k_to_vtk_order, grid, 256, 0, ni, nj, nk, a.comps, a.dptr, tmp);
        CHECK_LAUNCH();
        const uint64_t bytes = (uint64_t)cnt * a.comps * 8;
        if (fwrite(&bytes, 8, 1, st.f) != 1) return set_error(PICG_ERR_IO, "write failed: %s", path);
        rc = dev_to_file(st, tmp, bytes); if (rc) return rc;
    }
    const char tail[] = "\n</AppendedData>\n</VTKFile>\n";
    if (fwrite(tail, 1, sizeof tail - 1, st.f) != sizeof tail - 1) return set_error(PICG_ERR_IO, "write failed: %s", path);
    return PICG_OK;
}

int picg_checkpoint_save(const char* path, const picg_checkpoint_set* set, uint64_t user_ts) {
    REQUIRE_DEVICE(); REQUIRE_ARG(path && set && set->world && set->n_species >= 0 && (set->n_species == 0 || set->species), "picg_checkpoint_save: bad argument");
    picg_world_s* w = set->world; const Grid& g = w->g; const size_t nv = (size_t)g.nv;
    Stage st; int rc = open_stage(st, path, "wb"); if (rc) return rc;
    CkpHeader h; memset(&h, 0, sizeof h);
    memcpy(h.magic, kMagic, 8); h.version = 1; h.user_ts = user_ts; h.seed = g_seed;
    h.ni = g.ni; h.nj = g.nj; h.nk = g.nk; for (int c = 0; c < 3; c++) { h.x0[c] = g.x0[c]; h.xm[c] = g.xm[c]; } h.dt = w->dt;
    h.n_species = (uint32_t)set->n_species; h.n_mcc = (uint32_t)set->n_mcc; h.n_dsmc = (uint32_t)set->n_dsmc; h.n_sources = (uint32_t)set->n_sources;
    if (fwrite(&h, sizeof h, 1, st.f) != 1) return set_error(PICG_ERR_IO, "write failed: %s", path);
    rc = dev_to_file(st, w->phi, nv * 8); if (rc) return rc;
    rc = dev_to_file(st, w->rho, nv * 8); if (rc) return rc;
    rc = dev_to_file(st, w->ef, nv * 24); if (rc) return rc;
    for (int k = 0; k < set->n_species; k++) {
        picg_species_s* s = set->species[k];
        REQUIRE_ARG(s && s->w == w, "picg_checkpoint_save: species of another world");
        s->n_host_valid = false; rc = species_refresh_count(s); if (rc) return rc;
        CkpSpecies r; memset(&r, 0, sizeof r);
        r.n = s->n_host; r.mass = s->mass; r.charge = s->charge; r.mpw0 = s->mpw0; r.S = s->S; r.avg_samples = s->avg_samples;
        r.S_pinned = s->S_pinned; r.S_calibrated = s->S_calibrated;
        r.n_load_calls = s->n_load_calls; r.n_heavy_calls = s->n_heavy_calls; r.n_merge_calls = s->n_merge_calls;
        if (fwrite(&r, sizeof r, 1, st.f) != 1) return set_error(PICG_ERR_IO, "write failed: %s", path);
        for (int c = 0; c < 7; c++) { rc = dev_to_file(st, s->a[c], (size_t)r.n * 8); if (rc) return rc; }
        const double* f1[] = {s->den, s->den_avg, s->n_sum, s->nuu, s->nvv, s->nww};
        for (const double* f : f1) { rc = dev_to_file(st, f, nv * 8); if (rc) return rc; }
        rc = dev_to_file(st, s->nv_sum, nv * 24); if (rc) return rc;
    }
    for (int k = 0; k < set->n_mcc; k++) {
        CkpScalar r; r.step = set->mcc[k]->step;
        CUDA_TRY(cudaMemcpyAsync(&r.value, set->mcc[k]->wsv, 8, cudaMemcpyDeviceToHost, g_stream)); CUDA_TRY(cudaStreamSynchronize(g_stream));
        if (fwrite(&r, sizeof r, 1, st.f) != 1) return set_error(PICG_ERR_IO, "write failed: %s", path);
    }
    for (int k = 0; k < set->n_dsmc; k++) {
        CkpScalar r; r.step = set->dsmc[k]->step;
        CUDA_TRY(cudaMemcpyAsync(&r.value, set->dsmc[k]->svm, 8, cudaMemcpyDeviceToHost, g_stream)); CUDA_TRY(cudaStreamSynchronize(g_stream));
        if (fwrite(&r, sizeof r, 1, st.f) != 1) return set_error(PICG_ERR_IO, "write failed: %s", path);
    }
    for (int k = 0; k < set->n_sources; k++) {
        CkpScalar r; r.value = 0; r.step = set->sources[k]->step;
        if (fwrite(&r, sizeof r, 1, st.f) != 1) return set_error(PICG_ERR_IO, "write failed: %s", path);
    }
    return PICG_OK;
}

int picg_checkpoint_load(const char* path, const picg_checkpoint_set* set, uint64_t* user_ts) {
    REQUIRE_DEVICE(); REQUIRE_ARG(path && set && set->world && set->n_species >= 0 && (set->n_species == 0 || set->species), "picg_checkpoint_load: bad argument");
    picg_world_s* w = set->world; const Grid& g = w->g; const size_t nv = (size_t)g.nv;
    Stage st; int rc = open_stage(st, path, "rb"); if (rc) return rc;
    CkpHeader h;
    if (fread(&h, sizeof h, 1, st.f) != 1 || memcmp(h.magic, kMagic, 8) != 0 || h.version != 1) return set_error(PICG_ERR_ARG, "%s is not a picgpu checkpoint (version 1)", path);
    bool same = h.ni == g.ni && h.nj == g.nj && h.nk == g.nk;
    for (int c = 0; c < 3; c++) same = same && h.x0[c] == g.x0[c] && h.xm[c] == g.xm[c];
    if (!same) return set_error(PICG_ERR_ARG, "checkpoint mesh %dx%dx%d does not match the world (%dx%dx%d, or other extents)", h.ni, h.nj, h.nk, g.ni, g.nj, g.nk);
    if (h.n_species != (uint32_t)set->n_species || h.n_mcc != (uint32_t)set->n_mcc || h.n_dsmc != (uint32_t)set->n_dsmc || h.n_sources != (uint32_t)set->n_sources)
        return set_error(PICG_ERR_ARG, "checkpoint holds %u species / %u + %u interactions / %u sources, the caller passed %d / %d + %d / %d", h.n_species, h.n_mcc, h.n_dsmc,
                         h.n_sources, set->n_species, set->n_mcc, set->n_dsmc, set->n_sources);
    g_seed = h.seed; w->dt = h.dt;
    rc = file_to_dev(st, w->phi, nv * 8); if (rc) return rc;
    rc = file_to_dev(st, w->rho, nv * 8); if (rc) return rc;
    rc = file_to_dev(st, w->ef, nv * 24); if (rc) return rc;
    for (int k = 0; k < set->n_species; k++) {
        picg_species_s* s = set->species[k];
        REQUIRE_ARG(s && s->w == w, "picg_checkpoint_load: species of another world");
        CkpSpecies r;
        if (fread(&r, sizeof r, 1, st.f) != 1) return set_error(PICG_ERR_IO, "checkpoint file is truncated");
        if (r.mass != s->mass || r.charge != s->charge || r.mpw0 != s->mpw0) return set_error(PICG_ERR_ARG, "checkpoint species %d has other mass / charge / mpw0 than the species passed", k);
        rc = species_ensure_capacity(s, (size_t)r.n); if (rc) return rc;
        for (int c = 0; c < 7; c++) { rc = file_to_dev(st, s->a[c], (size_t)r.n * 8); if (rc) return rc; }
        double* f1[] = {s->den, s->den_avg, s->n_sum, s->nuu, s->nvv, s->nww};
        for (double* f : f1) { rc = file_to_dev(st, f, nv * 8); if (rc) return rc; }
        rc = file_to_dev(st, s->nv_sum, nv * 24); if (rc) return rc;
        u64 n64 = r.n;
        CUDA_TRY(cudaMemcpyAsync(&s->ctr->n, &n64, 8, cudaMemcpyHostToDevice, g_stream)); CUDA_TRY(cudaStreamSynchronize(g_stream));
        s->n_host = (size_t)r.n; s->n_host_valid = true; s->n_upper = (size_t)r.n;
        s->S = r.S; s->S_pinned = r.S_pinned != 0; s->S_calibrated = r.S_calibrated != 0; s->avg_samples = r.avg_samples;
        s->n_load_calls = r.n_load_calls; s->n_heavy_calls = r.n_heavy_calls; s->n_merge_calls = r.n_merge_calls;
        // the cell partition is not part of the file: it is rebuilt by the next sort
        s->sorted_valid = false; s->part_valid = false; s->lists_valid = false; s->count_valid = false; s->movers_fresh = false; s->part_n = 0;
    }
    for (int k = 0; k < set->n_mcc; k++) {
        CkpScalar r; if (fread(&r, sizeof r, 1, st.f) != 1) return set_error(PICG_ERR_IO, "checkpoint file is truncated");
        set->mcc[k]->step = r.step;
        CUDA_TRY(cudaMemcpyAsync(set->mcc[k]->wsv, &r.value, 8, cudaMemcpyHostToDevice, g_stream)); CUDA_TRY(cudaStreamSynchronize(g_stream));
    }
    for (int k = 0; k < set->n_dsmc; k++) {
        CkpScalar r; if (fread(&r, sizeof r, 1, st.f) != 1) return set_error(PICG_ERR_IO, "checkpoint file is truncated");
        set->dsmc[k]->step = r.step;
        CUDA_TRY(cudaMemcpyAsync(set->dsmc[k]->svm, &r.value, 8, cudaMemcpyHostToDevice, g_stream)); CUDA_TRY(cudaStreamSynchronize(g_stream));
    }
    for (int k = 0; k < set->n_sources; k++) {
        CkpScalar r; if (fread(&r, sizeof r, 1, st.f) != 1) return set_error(PICG_ERR_IO, "checkpoint file is truncated");
        set->sources[k]->step = r.step;
    }
    if (user_ts) *user_ts = h.user_ts;
    return PICG_OK;
}

}  // extern "C"
