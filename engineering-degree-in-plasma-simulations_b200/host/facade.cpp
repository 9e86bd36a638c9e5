// Implementation of the host-side C++ facade: every member forwards to the C ABI of libpicgpu.so
// (include/picgpu.h).  No particle or field arithmetic happens here; without a GPU the first call throws.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <map>
#include <sstream>
#include <thread>
#include "Config.h"
#include "Interactions.h"
#include "Outputs.h"
#include "PotentialSolver.h"
#include "Rnd.h"
#include "Source.h"
#include "Species.h"
#include "World.h"
#include "funkc.h"

namespace {
// C-ABI status -> the reference's error conventions: argument errors become std::invalid_argument,
// everything else std::runtime_error.
void check(int rc) {
    if (rc == PICG_OK) return;
    std::string msg = picg_last_error();
    if (rc == PICG_ERR_ARG) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}
void ensure_runtime() {
    static bool ready = false;
    if (ready) return;
    const char* dev = std::getenv("PICG_DEVICE");
    check(picg_init(dev ? std::atoi(dev) : 0));
    ready = true;
}
}  // namespace

// ------------------------------------------------------------------ Rnd / Config / funkc
Rnd::Rnd() : mt_gen{std::random_device()()}, rnd_dist{0.0, 1.0} {}
Rnd::Rnd(unsigned seed) : mt_gen{seed}, rnd_dist{0.0, 1.0} { picg_seed(seed); }
Rnd rnd;

Config::Config() {
    unsigned hc = std::thread::hardware_concurrency();
    if (hc <= 1) { m_NUM_THREADS = 1; m_MULTITHREADING = false; } else m_NUM_THREADS = hc - 1;     // Config.cpp:68-75
}
Config& Config::getInstance() { static Config c; return c; }
std::vector<size_t> splitIntoChunks(size_t size) {                                                  // Config.cpp:79-99
    unsigned T = Config::getInstance().getNUM_THREADS();
    if (size < T) { std::vector<size_t> r(size + 1); for (size_t i = 0; i <= size; i++) r[i] = i; return r; }
    std::vector<size_t> idx(T + 1, 0);
    for (unsigned i = 1; i <= T; i++) idx[i] = idx[i - 1] + size / T + (i <= size % T ? 1 : 0);
    return idx;
}
void print_help() {
    std::cout << "commands:\n\t--help\n\t--subcycling <bool>\n\t--multithreading <bool>\n\t--num_threads <n>\n\t--merging <bool>\n\t--output <mode>\n\t"
                 "--s_type GS|PCG\n\t--s_max_it <n>\n\t--s_tol <x>\n\t--phi <V>\n\t--num_ts <n>\n\t--dt <s>\n";
}
std::string lower(std::string& s) { std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return std::tolower(c); }); return s; }
bool parseArgument(const std::vector<std::string>& args, const std::string& option) { return std::find(args.begin(), args.end(), option) != args.end(); }

int inletName2Index(std::string f) {                                                                // World.cpp:369-389
    static const std::map<std::string, int> faces{{"x-", 0}, {"x+", 1}, {"y-", 2}, {"y+", 3}, {"z-", 4}, {"z+", 5},
                                                  {"-x", 0}, {"+x", 1}, {"-y", 2}, {"+y", 3}, {"-z", 4}, {"+z", 5}};
    lower(f);
    auto it = faces.find(f);
    if (it == faces.end()) throw std::invalid_argument("Passed wrong face name.");
    return it->second;
}

// ------------------------------------------------------------------ World
World::World(int ni_, int nj_, int nk_, type_calc x1, type_calc y1, type_calc z1, type_calc x2, type_calc y2, type_calc z2)
    : time_start{std::chrono::high_resolution_clock::now()}, nn{ni_, nj_, nk_}, ni{ni_}, nj{nj_}, nk{nk_}, ni_1{ni_ - 1}, nj_1{nj_ - 1}, nk_1{nk_ - 1},
      nv{ni_ * nj_ * nk_}, num_cells{(ni_ - 1) * (nj_ - 1) * (nk_ - 1)}, phi(nn), rho(nn), node_vol(nn), ef(nn), object_id(nn), object_phi(nn), node_type(nn) {
    ensure_runtime();
    x0 = {x1, y1, z1}; xm = {x2, y2, z2};
    for (int i = 0; i < 3; i++) { dx[i] = (xm[i] - x0[i]) / (nn[i] - 1); inv_dx[i] = 1 / dx[i]; xc[i] = 0.5 * (xm[i] + x0[i]); }   // World.cpp:63-77
    double a[3] = {x1, y1, z1}, b[3] = {x2, y2, z2};
    picg_world_t h = nullptr;
    check(picg_world_create(ni, nj, nk, a, b, &h));
    handle = std::shared_ptr<picg_world_s>(h, [](picg_world_s* p) { picg_world_destroy(p); });
    bindFields();
}
World::World(int ni_, int nj_, int nk_, type_calc3 v1, type_calc3 v2) : World(ni_, nj_, nk_, v1[0], v1[1], v1[2], v2[0], v2[1], v2[2]) {}

void World::bindFields() {
    picg_world_s* h = handle.get();
    auto fetchd = [h](int id) { return [h, id](type_calc* p) { check(picg_world_download(h, id, p)); }; };
    phi.bind(fetchd(PICG_F_PHI)); rho.bind(fetchd(PICG_F_RHO)); node_vol.bind(fetchd(PICG_F_NODE_VOL));
    ef.bind([h](type_calc3* p) { check(picg_world_download(h, PICG_F_EF, reinterpret_cast<double*>(p))); });
    int n = nv;
    auto fetchi = [h, n](int id) { return [h, id, n](int* p) { std::vector<double> t(n); check(picg_world_download(h, id, t.data())); for (int u = 0; u < n; u++) p[u] = (int)t[u]; }; };
    object_id.bind(fetchi(PICG_F_OBJECT_ID)); node_type.bind(fetchi(PICG_F_NODE_TYPE));
    for (int f : {PICG_F_PHI, PICG_F_RHO, PICG_F_NODE_VOL, PICG_F_EF, PICG_F_OBJECT_ID, PICG_F_NODE_TYPE}) deviceChanged(f);
}
void World::deviceChanged(int f) {
    switch (f) {
        case PICG_F_PHI: phi.invalidate(); break; case PICG_F_RHO: rho.invalidate(); break; case PICG_F_NODE_VOL: node_vol.invalidate(); break;
        case PICG_F_EF: ef.invalidate(); break; case PICG_F_OBJECT_ID: object_id.invalidate(); break; case PICG_F_NODE_TYPE: node_type.invalidate(); break;
    }
}
void World::syncToDevice() {
    if (phi.takeHostModified()) check(picg_world_upload(dev(), PICG_F_PHI, phi.raw()));
    if (rho.takeHostModified()) check(picg_world_upload(dev(), PICG_F_RHO, rho.raw()));
    if (ef.takeHostModified()) check(picg_world_upload(dev(), PICG_F_EF, reinterpret_cast<const double*>(ef.raw())));
}
void World::registerObject(Object* o) {
    type_calc3 c = o->getPos(); double cc[3] = {c[0], c[1], c[2]};
    if (auto* r = dynamic_cast<Rectangle*>(o)) { type_calc3 s = r->getSides(); double ss[3] = {s[0], s[1], s[2]}; check(picg_world_add_rectangle(dev(), cc, o->getPhi(), ss)); }
    else if (auto* s = dynamic_cast<Sphere*>(o)) check(picg_world_add_sphere(dev(), cc, o->getPhi(), s->getRadius()));
    else throw std::invalid_argument("World::addObject: only Rectangle and Sphere have a device representation");
}
void World::computeObjectID() { syncToDevice(); check(picg_world_compute_object_id(dev())); deviceChanged(PICG_F_PHI); deviceChanged(PICG_F_OBJECT_ID); deviceChanged(PICG_F_NODE_TYPE); }
int World::inObject(const type_calc3& pos) const { int i = 1; for (const auto& o : objects) { if (o->inObject(pos)) return i; i++; } return 0; }   // World.cpp:293-301
bool World::inBounds(const type_calc3& p) const { for (int i = 0; i < 3; i++) if (p[i] < x0[i] || p[i] >= xm[i]) return false; return true; }
std::string World::printObjects() const { std::stringstream ss; for (const auto& o : objects) ss << *o << "\n"; return ss.str(); }
void World::setTime(type_calc dt_, int n) { dt = dt_; num_ts = n; check(picg_world_set_time(dev(), dt_, n)); }
type_calc World::getWallTime() { return (std::chrono::high_resolution_clock::now() - time_start).count() * 1e-9; }
type_calc World::getPE() { syncToDevice(); double pe = 0; check(picg_world_potential_energy(dev(), &pe)); return pe; }
void World::addInlet(std::string face, type_calc phi_set, int type) {                              // World.cpp:206-261
    int f = inletName2Index(face), axis = f >> 1; bool plus = f & 1;
    int lim[3] = {ni, nj, nk};
    for (int i = 0; i < ni; i++) for (int j = 0; j < nj; j++) for (int k = 0; k < nk; k++) {
        int c[3] = {i, j, k};
        if (c[axis] != (plus ? lim[axis] - 1 : 0)) continue;
        node_type[i][j][k] = type; phi[i][j][k] = phi_set;
    }
    syncToDevice();
    std::vector<double> t(nv); for (int u = 0; u < nv; u++) t[u] = node_type.raw()[u];
    node_type.takeHostModified();
    check(picg_world_upload(dev(), PICG_F_NODE_TYPE, t.data()));
}
void World::computeChargeDensity(std::vector<Species>& species) {                                   // World.cpp:193-200
    std::vector<picg_species_t> h; for (Species& s : species) h.push_back(s.dev());
    check(picg_world_charge_density(dev(), h.data(), (int)h.size()));
    deviceChanged(PICG_F_RHO);
}

// ------------------------------------------------------------------ Species
Species::Species(std::string name_, type_calc mass_, type_calc charge_, World& world_, type_calc mpw0_)
    : Species(name_, mass_, charge_, world_, mpw0_, -666) {}
Species::Species(std::string name_, type_calc mass_, type_calc charge_, World& world_, type_calc mpw0_, type_calc E_ion_)
    : world{world_}, name{name_}, mass{mass_}, charge{charge_}, mpw0{mpw0_}, E_ion{E_ion_}, den{world_.nn}, den_avg{world_.nn}, T{world_.nn}, vel{world_.nn},
      macro_part_count{world_.ni - 1, world_.nj - 1, world_.nk - 1} {
    picg_species_t h = nullptr;
    check(picg_species_create(world.dev(), mass, charge, mpw0, &h));
    handle = std::shared_ptr<picg_species_s>(h, [](picg_species_s* p) { picg_species_destroy(p); });
    bindFields();
}
Species::Species(const Species& o)
    : world{o.world}, handle{o.handle}, sorted{o.sorted}, name{o.name}, mass{o.mass}, charge{o.charge}, mpw0{o.mpw0}, E_ion{o.E_ion}, den{o.den}, den_avg{o.den_avg},
      T{o.T}, vel{o.vel}, macro_part_count{o.macro_part_count} { bindFields(); }
void Species::bindFields() {
    picg_species_s* h = handle.get();
    auto f = [h](int id) { return [h, id](type_calc* p) { check(picg_species_download_field(h, id, p)); }; };
    den.bind(f(PICG_SF_DEN)); den_avg.bind(f(PICG_SF_DEN_AVG)); T.bind(f(PICG_SF_T)); macro_part_count.bind(f(PICG_SF_MACRO_COUNT));
    vel.bind([h](type_calc3* p) { check(picg_species_download_field(h, PICG_SF_VEL, p)); });
}
size_t Species::getNumParticles() const { size_t n = 0; check(picg_species_count(dev(), &n)); return n; }
void Species::advanceElectrons(type_calc dt) { world.syncToDevice(); check(picg_species_push_electrons(dev(), dt)); sorted = false; }
void Species::advanceNonElectron(Species& neutrals, Species& spherium, type_calc dt) {
    world.syncToDevice();
    check(picg_species_push_heavy(dev(), neutrals.dev(), spherium.dev(), dt, Config::getInstance().getSPUTTERING() ? 1 : 0));
    sorted = false;
}
void Species::computeNumberDensity() { check(picg_species_deposit_density(dev())); den.invalidate(); }
void Species::addParticle(type_calc x, type_calc y, type_calc z, type_calc u, type_calc v, type_calc w, type_calc m) { addParticle({x, y, z}, {u, v, w}, m); }
void Species::addParticle(type_calc3 p, type_calc3 v, type_calc m) {
    world.syncToDevice();
    double a[7] = {p[0], p[1], p[2], v[0], v[1], v[2], m};
    check(picg_species_add_particles(dev(), 1, a, nullptr));
}
void Species::loadParticleBoxThermal(type_calc3 c, type_calc3 sides, type_calc num_den, type_calc T_) {
    world.syncToDevice();
    double cc[3] = {c[0], c[1], c[2]}, ss[3] = {sides[0], sides[1], sides[2]};
    size_t loaded = 0;
    check(picg_species_load_box_thermal(dev(), cc, ss, num_den, T_, &loaded));
    std::cout << "Loaded number of macroparticles: " << loaded << " (" << name << ")\n";
}
type_calc Species::getMicroCount() { double m; check(picg_species_diagnostics(dev(), &m, nullptr, nullptr)); return m; }
type_calc3 Species::getMomentum() { double p[3]; check(picg_species_diagnostics(dev(), nullptr, p, nullptr)); return {p[0], p[1], p[2]}; }
type_calc Species::getKE() { double ke; check(picg_species_diagnostics(dev(), nullptr, nullptr, &ke)); return ke; }
void Species::updateAverages() { check(picg_species_update_averages(dev())); den_avg.invalidate(); }
void Species::sampleMoments() { check(picg_species_sample_moments(dev())); }
void Species::computeGasProperties() { check(picg_species_compute_gas_properties(dev())); vel.invalidate(); T.invalidate(); }
void Species::clearSamples() { check(picg_species_clear_samples(dev())); }
void Species::computeMacroParticlesCount() { check(picg_species_count_per_cell(dev())); macro_part_count.invalidate(); }
void Species::sortByCell() { check(picg_species_sort(dev())); sorted = true; }
void Species::merge() {                                  // Species.cpp:1037-1145; same console report as the reference (:1038-1040, :1142-1143)
    std::stringstream out;
    out << "Merging " << name << "\nBefore merge: particles size: " << getNumParticles() << ", energy: " << getKE() << ", momentum: " << getMomentum() << "\n";
    uint64_t n0 = 0, n1 = 0, st[4] = {0, 0, 0, 0};
    check(picg_species_merge(dev(), &n0, &n1, st));
    if (n1 != n0) sorted = false;
    out << name << " after merge: particles size: " << n1 << ", energy: " << getKE() << ", momentum: " << getMomentum() << "\n";
    std::cout << out.str();
}
const std::vector<Particle>& Species::getConstPartRef() {
    size_t n = getNumParticles(), got = 0;
    std::vector<double> a(n * 7);
    check(picg_species_download(dev(), n, a.data(), &got));
    particles_mirror.clear(); particles_mirror.reserve(got);
    for (size_t i = 0; i < got; i++) particles_mirror.emplace_back(a[7 * i], a[7 * i + 1], a[7 * i + 2], a[7 * i + 3], a[7 * i + 4], a[7 * i + 5], a[7 * i + 6]);
    return particles_mirror;
}
const Particle& Species::getConstPartRef(int i) { if ((size_t)i >= particles_mirror.size()) getConstPartRef(); return particles_mirror.at(i); }
void Species::setParticles(const std::vector<Particle>& p) {
    std::vector<double> a(p.size() * 7);
    for (size_t i = 0; i < p.size(); i++) { for (int c = 0; c < 3; c++) { a[7 * i + c] = p[i].pos[c]; a[7 * i + 3 + c] = p[i].vel[c]; } a[7 * i + 6] = p[i].macro_weight; }
    check(picg_species_upload(dev(), p.size(), a.data()));
    sorted = false;
}

// ------------------------------------------------------------------ PotentialSolver
PotentialSolver::PotentialSolver(World& w, unsigned max_it, type_calc tol, SolverType type) : world(w), solver_type(type), PCG_max_solver_it(max_it), tolerance(tol) {
    GS_max_solver_it = (type == GS) ? max_it : 20 * max_it;                                         // PotentialSolver.cpp:46-49
    picg_solver_t h = nullptr;
    check(picg_solver_create(world.dev(), GS_max_solver_it, tol, &h));
    handle = std::shared_ptr<picg_solver_s>(h, [](picg_solver_s* p) { picg_solver_destroy(p); });
}
void PotentialSolver::setReferenceValues(type_calc phi0, type_calc n0, type_calc Te0) { check(picg_solver_set_reference(handle.get(), phi0, n0, Te0)); }
bool PotentialSolver::solve() {
    if (solver_type == QN) { std::cerr << "quasi neutral not implemented yet\n"; return false; }      // PotentialSolver.cpp:349-352
    return solveGS();
}
bool PotentialSolver::solveGS() {
    world.syncToDevice();
    int conv = 0;
    check(picg_solver_solve_gs(handle.get(), &conv, &last_iterations, &last_L2));
    world.deviceChanged(PICG_F_PHI);
    if (!conv) std::cerr << "GS SOR failed to converge, L2 = " << last_L2 << " tolerance = " << tolerance << " time step = " << world.getTs() << std::endl;   // :162-165
    return conv != 0;
}
bool PotentialSolver::solveNRPCG() {                                                                // PotentialSolver.cpp:178-240
    world.syncToDevice();
    int conv = 0; unsigned nr = 0, pcg = 0; double norm = 0;
    check(picg_solver_solve_nrpcg(handle.get(), 1, PCG_max_solver_it, &conv, &nr, &pcg, &norm));
    world.deviceChanged(PICG_F_PHI);
    last_iterations = pcg; last_L2 = norm;
    if (!conv) std::cerr << "NR+PCG failed to converge, norm = " << norm << " tolerance = " << 1e-3 << " time step = " << world.getTs() << std::endl;   // :236-238
    return conv != 0;
}
void PotentialSolver::computeEF() { world.syncToDevice(); check(picg_solver_compute_ef(handle.get())); world.deviceChanged(PICG_F_EF); }
std::ostream& operator<<(std::ostream& out, SolverType& t) { return out << (t == GS ? "GS" : t == PCG ? "PCG" : "QN"); }
std::istream& operator>>(std::istream& in, SolverType& t) {
    std::string s; in >> s;
    if (s == "GS") t = GS; else if (s == "PCG") t = PCG; else if (s == "QN") t = QN; else throw std::invalid_argument("Wrong Solver Type");
    return in;
}

// ------------------------------------------------------------------ Source
Source::Source(Species& species, World& w, type_calc v_drift, type_calc den, type_calc T, std::string face) noexcept : sp(species), world(w) {
    int f = 0;
    try { f = inletName2Index(face); } catch (const std::invalid_argument& e) { std::cerr << e.what() << std::endl; f = 0; }   // Source.cpp:5-14
    picg_source_t h = nullptr;
    if (picg_source_create(sp.dev(), world.dev(), v_drift, den, T, f, &h) != PICG_OK) { std::cerr << "Source: " << picg_last_error() << std::endl; return; }
    handle = std::shared_ptr<picg_source_s>(h, [](picg_source_s* p) { picg_source_destroy(p); });
    world.deviceChanged(PICG_F_PHI); world.deviceChanged(PICG_F_NODE_TYPE);                          // World::addInlet side effect (:28)
}
void Source::sample() const noexcept {
    if (!handle) return;
    if (picg_source_sample(handle.get(), nullptr) != PICG_OK) std::cerr << "Source::sample: " << picg_last_error() << std::endl;
    sp.setSorted(false);
}

// ------------------------------------------------------------------ MC_MEX_Ionization
MC_MEX_Ionization::MC_MEX_Ionization(Species& n, Species& i, Species& e, World& w, std::string path, int)
    : neutrals(n), ions(i), electrons(e), world(w) {
    std::ifstream in(path);
    if (!in.is_open()) throw std::invalid_argument("Couldn't open file " + path + " for crosssection data of collisions");   // Interactions.cpp:519-522
    std::vector<double> E, S; double a, b;
    while (in >> a >> b) { E.push_back(a); S.push_back(b); }
    picg_mcc_t h = nullptr;
    check(picg_mcc_create(n.dev(), i.dev(), e.dev(), w.dev(), E.data(), S.data(), (int)E.size(), n.E_ion, &h));
    handle = std::shared_ptr<picg_mcc_s>(h, [](picg_mcc_s* p) { picg_mcc_destroy(p); });
}
void MC_MEX_Ionization::apply(type_calc dt) noexcept {
    if (picg_mcc_apply(handle.get(), dt, &last) != PICG_OK) { std::cerr << "MC_MEX_Ionization::apply: " << picg_last_error() << std::endl; return; }
    if (last.collisions) { ions.setSorted(false); electrons.setSorted(false); neutrals.setSorted(false); }       // Interactions.cpp:751-754
}

// ------------------------------------------------------------------ DSMC_MEX
DSMC_MEX::DSMC_MEX(Species& sp, World& w) noexcept : species1(sp), species2(sp), world(w) {
    picg_dsmc_t h = nullptr;
    if (picg_dsmc_create(sp.dev(), nullptr, w.dev(), &h) != PICG_OK) { std::cerr << "DSMC_MEX: " << picg_last_error() << std::endl; return; }
    handle = std::shared_ptr<picg_dsmc_s>(h, [](picg_dsmc_s* p) { picg_dsmc_destroy(p); });
}
DSMC_MEX::DSMC_MEX(Species& s1, Species& s2, World& w) : species1(s1), species2(s2), world(w) {
    picg_dsmc_t h = nullptr;
    check(picg_dsmc_create(s1.dev(), s2.dev(), w.dev(), &h));                                                  // Interactions.cpp:160-162
    handle = std::shared_ptr<picg_dsmc_s>(h, [](picg_dsmc_s* p) { picg_dsmc_destroy(p); });
}
void DSMC_MEX::apply(type_calc dt) noexcept {
    if (!handle) return;
    if (picg_dsmc_apply(handle.get(), dt, &last) != PICG_OK) std::cerr << "DSMC_MEX::apply: " << picg_last_error() << std::endl;
}

// ------------------------------------------------------------------ Output (minimal)
namespace Output {
static std::ofstream f_diag;
void screenOutput(World& world, std::vector<Species>& species) {
    std::cout << "ts: " << world.getTs();
    for (Species& sp : species) std::cout << "\t " << sp.name << ":" << sp.getNumParticles();
    std::cout << std::endl;
}
void diagOutput(World& world, std::vector<Species>& species) {                                      // Outputs.cpp:143-179
    if (!f_diag.is_open()) {
        f_diag.open("results/runtime_diags.csv");
        f_diag << "ts,time,wall_time";
        for (Species& sp : species) f_diag << ",mp_count." << sp.name << ",real_count." << sp.name << ",px." << sp.name << ",py." << sp.name << ",pz." << sp.name << ",KE." << sp.name;
        f_diag << ",PE,E_total" << std::endl;
    }
    f_diag << world.getTs() << "," << world.getTime() << "," << world.getWallTime();
    double tot = 0;
    for (Species& sp : species) {
        double ke = sp.getKE(); tot += ke; double3 m = sp.getMomentum();
        f_diag << "," << sp.getNumParticles() << "," << sp.getMicroCount() << "," << m[0] << "," << m[1] << "," << m[2] << "," << ke;
    }
    double pe = world.getPE();
    f_diag << "," << pe << "," << (tot + pe) << "\n";
    if (world.getTs() % 25 == 0) f_diag.flush();
}
void fieldsOutput(World& world, std::vector<Species>& species, std::string name1) {                 // Outputs.cpp:9-123, binary (appended raw) .vti
    for (Species& sp : species) sp.computeGasProperties();                                          // :11-14
    std::stringstream name; name << "results/" << name1 << "fields_" << std::setfill('0') << std::setw(5) << world.getTs() << ".vti";   // :16-17
    std::vector<picg_species_t> hs; std::vector<const char*> names;
    for (Species& sp : species) { hs.push_back(sp.dev()); names.push_back(sp.name.c_str()); }
    if (picg_write_fields_vti(name.str().c_str(), world.dev(), hs.data(), names.data(), (int)hs.size()) != PICG_OK)
        std::cerr << "Could not write " << name.str() << ": " << picg_last_error() << std::endl;      // :22-25
    for (Species& sp : species) sp.clearSamples();                                                  // :119-121
}
void particlesOutput(World& world, std::vector<Species>& species, int base, std::string name1) {
    for (Species& sp : species) {
        std::stringstream name; name << "results/" << name1 << sp.name << "_" << std::setfill('0') << std::setw(5) << world.getTs() << ".csv";
        std::ofstream out(name.str()); if (!out.is_open()) continue;
        const std::vector<Particle>& p = sp.getConstPartRef();
        size_t stride = std::max<size_t>(1, p.size() / std::max(1, base));
        out << "x,y,z,u,v,w,mpw\n";
        for (size_t i = 0; i < p.size(); i += stride) out << p[i].pos[0] << "," << p[i].pos[1] << "," << p[i].pos[2] << "," << p[i].vel[0] << "," << p[i].vel[1] << "," << p[i].vel[2] << "," << p[i].macro_weight << "\n";
    }
}
static bool checkpoint_io(bool save, const std::string& path, World& world, std::vector<Species>& species) {
    std::vector<picg_species_t> hs; for (Species& sp : species) hs.push_back(sp.dev());
    picg_checkpoint_set set{}; set.world = world.dev(); set.species = hs.data(); set.n_species = (int)hs.size();
    uint64_t ts = (uint64_t)world.getTs();
    world.syncToDevice();                                      // pending host-side edits of phi / rho / ef go to the device first (save: they belong to the state; load: they are overwritten)
    int rc = save ? picg_checkpoint_save(path.c_str(), &set, ts) : picg_checkpoint_load(path.c_str(), &set, &ts);
    if (rc != PICG_OK) { std::cerr << (save ? "saveCheckpoint: " : "loadCheckpoint: ") << picg_last_error() << std::endl; return false; }
    if (!save) {
        world.deviceChanged(PICG_F_PHI); world.deviceChanged(PICG_F_RHO); world.deviceChanged(PICG_F_EF);
        for (Species& sp : species) sp.setSorted(false);
    }
    return true;
}
bool saveCheckpoint(const std::string& path, World& world, std::vector<Species>& species) { return checkpoint_io(true, path, world, species); }
bool loadCheckpoint(const std::string& path, World& world, std::vector<Species>& species) { return checkpoint_io(false, path, world, species); }
std::ostream& operator<<(std::ostream& out, Output::modes& t) {
    static const char* n[] = {"none", "all", "screen", "fields", "particles", "diagnostics", "convergence"};
    return out << n[(int)t];
}
std::istream& operator>>(std::istream& in, Output::modes& t) {                                      // Outputs.cpp:323-354
    std::string s; in >> s; lower(s);
    static const std::map<std::string, modes> m{{"none", none}, {"0", none}, {"all", all}, {"1", all}, {"screen", screen}, {"2", screen}, {"fields", fields}, {"3", fields},
                                                {"particles", particles}, {"4", particles}, {"diagnostics", diagnostics}, {"5", diagnostics}, {"convergence", convergence}, {"6", convergence}};
    auto it = m.find(s);
    if (it == m.end()) { std::cerr << "Wrong Output::modes value, setting Output::modes::fields \n"; t = fields; } else t = it->second;
    return in;
}
}  // namespace Output
