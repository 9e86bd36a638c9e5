/* picgpu.h -- C ABI of the B200-native PIC-DSMC particle loop (libpicgpu.so).
 *
 * This is the drop-in boundary for the per-timestep particle loop of the
 * reference's ch4/v3 main (ch4/v3/src/main.cpp:177-288).  The reference has no
 * FFI of its own; its "API" is the public C++ surface of World / Species /
 * PotentialSolver / Source / Interaction.  Each entry point below names the
 * reference member (file:line under /root/reference) whose device-side
 * replacement it is.  The host-side C++ facade in
 * engineering-degree-in-plasma-simulations_b200/host/ forwards those members
 * to these functions; see INTEGRATION.md for the binding a maintainer adds.
 *
 * Conventions
 *  - plain C types only; every function returns 0 on success or a negative
 *    picg_status; picg_last_error() gives a human-readable message
 *    (thread-local).  No exception crosses this boundary.
 *  - one host thread drives one GPU; all kernels are issued on one CUDA stream
 *    per process (picg_stream()), so call order == execution order, matching
 *    the reference's sequential semantics.  Not re-entrant per handle.
 *  - node fields are double[nv] in Field<T> order u=(i*nj+j)*nk+k
 *    (ch4/v3/src/Field.h:16,88); vector fields are double[3*nv] xyz-interleaved;
 *    cell fields are double[(ni-1)(nj-1)(nk-1)] in Field order
 *    (i*(nj-1)+j)*(nk-1)+k; particles cross the boundary as the reference's
 *    AoS record x y z u v w mpw (7 doubles, ch4/v3/src/Species.h:12-29).
 *  - there is NO CPU fallback: without a CUDA device every compute call fails
 *    with PICG_ERR_NO_DEVICE.
 */
#ifndef PICGPU_H
#define PICGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PICG_API __attribute__((visibility("default")))

typedef enum {
    PICG_OK = 0,
    PICG_ERR_NO_DEVICE = -1,   /* no CUDA device / picg_init not called          */
    PICG_ERR_CUDA = -2,        /* a CUDA runtime call or kernel failed           */
    PICG_ERR_ARG = -3,         /* invalid argument (maps to std::invalid_argument)*/
    PICG_ERR_OOM = -4,         /* device allocation failed                       */
    PICG_ERR_OVERFLOW = -5,    /* fixed-point density accumulator overflowed     */
    PICG_ERR_STATE = -6,       /* call not valid in the current state            */
    PICG_ERR_IO = -7           /* a file could not be opened, written or is truncated */
} picg_status;

typedef struct picg_world_s*   picg_world_t;
typedef struct picg_species_s* picg_species_t;
typedef struct picg_solver_s*  picg_solver_t;
typedef struct picg_mcc_s*     picg_mcc_t;
typedef struct picg_dsmc_s*    picg_dsmc_t;
typedef struct picg_source_s*  picg_source_t;
typedef struct picg_species32_s* picg_species32_t;   /* fp32 secondary store, see the end of this file */

/* ------------------------------------------------------------------ runtime */
PICG_API int         picg_init(int device);                 /* select device, create the stream */
PICG_API int         picg_shutdown(void);
PICG_API const char* picg_last_error(void);
PICG_API const char* picg_version(void);
PICG_API int         picg_device_count(int* n);
PICG_API void*       picg_stream(void);                     /* cudaStream_t all kernels run on  */
PICG_API int         picg_synchronize(void);
PICG_API int         picg_seed(uint64_t seed);              /* Philox key for every stochastic kernel (replaces `rnd = Rnd(seed)`, Rnd.cpp:7) */
PICG_API int         picg_set_rank(int rank, int world_size); /* multi-GPU: decorrelates RNG streams, rescales MC candidate counts */
/* number of kernels this library launched since the last reset (bench.py's gpu_launches) */
PICG_API uint64_t    picg_launch_count(void);
PICG_API void        picg_launch_count_reset(void);
/* number of device (re)allocations of particle stores / scratch so far: a timed region should not see it change */
PICG_API uint64_t    picg_realloc_count(void);
/* per-kernel CUDA-event timers: enable, then read accumulated ms and launch counts by kernel id */
PICG_API int         picg_timers_enable(int on);
PICG_API int         picg_timers_reset(void);
PICG_API int         picg_timer_read(int kernel_id, double* total_ms, uint64_t* launches);
PICG_API const char* picg_timer_name(int kernel_id);        /* NULL past the last id */

/* -------------------------------------------------------------------- World */
/* World::World(int,int,int,type_calc3,type_calc3)  ch4/v3/src/World.cpp:23-27, setExtents :63-77, computeNodeVolumes :353-367 */
PICG_API int picg_world_create(int ni, int nj, int nk, const double x0[3], const double xm[3], picg_world_t* out);
PICG_API int picg_world_destroy(picg_world_t w);
/* World::setTime  World.cpp:323-326 (dt is what addParticle's half-step rewind uses, Species.cpp:431) */
PICG_API int picg_world_set_time(picg_world_t w, double dt, int num_ts);
/* World::addObject<Rectangle>/<Sphere>  World.h:119-131, Object.cpp:163-171,111-115 */
PICG_API int picg_world_add_rectangle(picg_world_t w, const double centre[3], double phi, const double sides[3]);
PICG_API int picg_world_add_sphere(picg_world_t w, const double centre[3], double phi, double radius);
/* World::computeObjectID  World.cpp:276-292 -- rasterised on the host with the reference's closed fp64 test
 * (SURVEY B18), then uploaded: sets object_id, node_type=DIRICHLET and phi on object nodes. */
PICG_API int picg_world_compute_object_id(picg_world_t w);

typedef enum {
    PICG_F_PHI = 0, PICG_F_RHO = 1, PICG_F_NODE_VOL = 2, PICG_F_EF = 3 /*3*nv*/,
    PICG_F_OBJECT_ID = 4 /*as double*/, PICG_F_NODE_TYPE = 5 /*as double*/
} picg_world_field;
/* public Field members World::{phi,rho,node_vol,ef,object_id,node_type}  World.h:51-57 */
PICG_API int picg_world_download(picg_world_t w, int field, double* host);
/* the same read-back split in two (replaces nothing in the reference, whose fields live in host memory): _begin queues the copy behind the
 * work submitted so far on a copy stream, _end waits for it; device work queued in between runs alongside the transfer and must not
 * overwrite the field.  `host` should be page-locked.  Real-valued fields only; one download in flight. */
PICG_API int picg_world_download_begin(picg_world_t w, int field, double* host);
PICG_API int picg_world_download_end(picg_world_t w);
PICG_API int picg_world_upload(picg_world_t w, int field, const double* host);
/* World::computeChargeDensity  World.cpp:193-200 : rho = sum_s charge_s * den_s over charged species */
PICG_API int picg_world_charge_density(picg_world_t w, const picg_species_t* species, int n);
/* the same on the nodes [node_begin, node_end) only (multi-GPU: the planes this rank solves on) */
PICG_API int picg_world_charge_density_range(picg_world_t w, const picg_species_t* species, int n, size_t node_begin, size_t node_end);
/* World::getPE  World.cpp:108-118 */
PICG_API int picg_world_potential_energy(picg_world_t w, double* pe);
/* device pointers for zero-copy interop (torch.distributed all-reduce of the grids) */
PICG_API int picg_world_device_ptr(picg_world_t w, int field, void** dptr, size_t* bytes);

/* ------------------------------------------------------------------ Species */
/* Species::Species(name,mass,charge,World&,mpw0[,E_ion])  Species.cpp:29-40 */
PICG_API int picg_species_create(picg_world_t w, double mass, double charge, double mpw0, picg_species_t* out);
PICG_API int picg_species_destroy(picg_species_t s);
PICG_API int picg_species_reserve(picg_species_t s, size_t capacity);
/* number of leading store slots the cell partition of the last sort covers (0: none); particles appended since lie beyond it and
 * take the generic deposit kernel ("deposit_tail" in the timers).  Diagnostics for bench.py's byte accounting. */
PICG_API int picg_species_partition_size(picg_species_t s, size_t* n);
/* Species::getNumParticles  Species.cpp:44-46 */
PICG_API int picg_species_count(picg_species_t s, size_t* n);
/* raw store access == Species::getPartRef() / getConstPartRef()  Species.cpp:820-828 */
PICG_API int picg_species_upload(picg_species_t s, size_t n, const double* aos7);
PICG_API int picg_species_download(picg_species_t s, size_t capacity, double* aos7, size_t* n);
/* Species::addParticle(pos,vel,mpw)  Species.cpp:420-434: reject NaN / out of bounds / in object, then
 * vel -= charge/mass*E(pos)*(0.5*world.dt).  n particles at once; *accepted returns how many were kept. */
PICG_API int picg_species_add_particles(picg_species_t s, size_t n, const double* aos7, size_t* accepted);
/* Species::loadParticleBoxThermal(x0, sides, num_den, T)  Species.cpp:560-598, generated on the device:
 * (size_t)(num_den*volume/mpw0) particles uniform in the box centred at x0, velocities by sampleV3th(T) (:855-869),
 * each passed through addParticle.  With picg_set_rank(r, G) every rank loads its 1/G share. */
PICG_API int picg_species_load_box_thermal(picg_species_t s, const double centre[3], const double sides[3], double num_den, double T, size_t* loaded);
/* Species::advanceElectrons(dt)  Species.cpp:258-399 (gather, kick, drift, absorb on walls/objects, remove) */
PICG_API int picg_species_push_electrons(picg_species_t s, double dt);
/* Species::advanceNonElectron(neutrals, spherium, dt)  Species.cpp:47-256 */
PICG_API int picg_species_push_heavy(picg_species_t s, picg_species_t neutrals, picg_species_t spherium, double dt, int sputtering);
/* ch2 Species::advance(): kick, drift, specular reflection on the six faces  ch2/v2/Species.cpp:18-55 */
PICG_API int picg_species_push_reflect(picg_species_t s, double dt);
/* ch3 Species::advance(): kick, drift, delete in object / out of bounds  ch3/v1/Species.cpp:28-45 (== electron push arithmetic) */
/* Species::computeNumberDensity  Species.cpp:401-416 + Field::scatter Field.h:157-199 + operator/= :563-583.
 * Deterministic: contributions are quantised llrint(c*2^S) and summed in int64. */
PICG_API int picg_species_deposit_density(picg_species_t s);
/* fused Species::advanceElectrons + computeNumberDensity (+ computeMacroParticlesCount): one pass over the particles */
PICG_API int picg_species_push_electrons_deposit(picg_species_t s, double dt, int count_cells);
/* fused Species::advanceNonElectron + computeNumberDensity (+ computeMacroParticlesCount) */
PICG_API int picg_species_push_heavy_deposit(picg_species_t s, picg_species_t neutrals, picg_species_t spherium, double dt, int sputtering, int count_cells);
/* multi-GPU: fused push + deposit into the raw int64 accumulator only (scale pinned on every rank); all-reduce
 * PICG_SF_DEN_FIXED across ranks, then picg_species_finalize_density.  heavy != 0 selects advanceNonElectron. */
PICG_API int picg_species_push_deposit_partial(picg_species_t s, picg_species_t neutrals, picg_species_t spherium, double dt, int heavy, int sputtering, int count_cells);
PICG_API int picg_species_density_scale(picg_species_t s, int* S);           /* the S of the last deposit */
PICG_API int picg_species_set_density_scale(picg_species_t s, int S);        /* pin S (tests); <-1000 = automatic */
/* Species::sampleMoments :767-776, computeGasProperties :777-804, clearSamples :805-812 */
PICG_API int picg_species_sample_moments(picg_species_t s);
PICG_API int picg_species_compute_gas_properties(picg_species_t s);
PICG_API int picg_species_clear_samples(picg_species_t s);
/* Species::updateAverages -> Field::updateMovingAverage  Field.h:246-261 */
PICG_API int picg_species_update_averages(picg_species_t s);
/* Species::computeMacroParticlesCount  Species.cpp:813-819 */
PICG_API int picg_species_count_per_cell(picg_species_t s);
/* Species::sortIndexes  Species.cpp:905-929 -> device: cell-sorted SoA layout + cell_start[].
 * Between sorts the exact per-cell lists that MC collisions need are kept up to date by listing the particles that
 * changed cell ("movers"); when more than `f` of a species moved, it is re-sorted (periodic radix sort).  f = 0 forces a
 * full sort whenever the order is stale (the reference re-sorts every step). */
PICG_API int picg_set_mover_fraction(double f);
/* Particles appended since the last sort (MC products, sources, re-emitted neutrals) form a tail behind the cell partition.  When the
 * tail exceeds the fraction `f` of the store (default 0.12) it is merged into the partition: only the tail is sorted, the partition is
 * shifted to open the gaps (a streaming pass; replaces the full re-sort the reference does every step, Species.cpp:905-929).  0: never. */
PICG_API int picg_set_merge_fraction(double f);
PICG_API int picg_tail_merge_count(uint64_t* merges);
/* How the mover lists were obtained since start: passes that re-used the list a deposit pass produced on the fly (only the
 * appended tail is scanned), full scans of the store, and fall-backs to a full radix sort. */
PICG_API int picg_mover_stats(uint64_t* from_deposit, uint64_t* full_scans, uint64_t* resorts);
PICG_API int picg_species_sort(picg_species_t s);
/* Species::merge()  Species.cpp:1037-1145 (+ sortVelocitiesInCell :981-1035): in every cell with >= 10 particles, the particles of a
 * 15^3 velocity bin that holds more than two are replaced by two (weight, momentum and per-axis energy of the bin kept).
 * stats (optional, 4 words): merged bins, particles removed, new particles dropped by addParticle's filter, cells left unmerged
 * because they hold more than 1024 particles. */
PICG_API int picg_species_merge(picg_species_t s, uint64_t* n_before, uint64_t* n_after, uint64_t stats[4]);
/* diagnostics: getMicroCount, getMomentum, getKE  Species.cpp:726-755 */
PICG_API int picg_species_diagnostics(picg_species_t s, double* micro_count, double momentum[3], double* ke);

typedef enum {
    PICG_SF_DEN = 0, PICG_SF_DEN_AVG = 1, PICG_SF_T = 2, PICG_SF_VEL = 3 /*3*nv*/,
    PICG_SF_MACRO_COUNT = 4 /*cells*/, PICG_SF_N_SUM = 5, PICG_SF_NV_SUM = 6 /*3*nv*/,
    PICG_SF_NUU_SUM = 7, PICG_SF_NVV_SUM = 8, PICG_SF_NWW_SUM = 9,
    PICG_SF_DEN_FIXED = 10 /* int64[nv]: the raw fixed-point accumulator */
} picg_species_field;
PICG_API int picg_species_download_field(picg_species_t s, int field, void* host);
PICG_API int picg_species_device_ptr(picg_species_t s, int field, void** dptr, size_t* bytes);
/* Device-resident hand-over (no reference counterpart: Species::getPartRef, Species.h:104, hands out the host vector): the seven SoA
 * arrays x y z u v w mpw of the store, sized for at least `capacity` particles; the caller fills n of them on the library's stream
 * (picg_stream) and calls picg_species_adopt(s, n), which replaces the contents like picg_species_upload does. */
PICG_API int picg_species_particle_arrays(picg_species_t s, size_t capacity, void* arrays7[7], size_t* capacity_out /*may be NULL*/);
PICG_API int picg_species_adopt(picg_species_t s, size_t n);
/* multi-GPU: after all-reducing DEN_FIXED across ranks, turn it into den (divide by 2^S and node_vol) */
PICG_API int picg_species_finalize_density(picg_species_t s);
PICG_API int picg_species_finalize_density_range(picg_species_t s, size_t node_begin, size_t node_end);
/* multi-GPU: deposit into the fixed-point accumulator only (no finalize) */
PICG_API int picg_species_deposit_density_partial(picg_species_t s);

/* ---------------------------------------------------------- PotentialSolver */
/* PotentialSolver(World&, max_it, tol, GS)  PotentialSolver.cpp:44-52, precalculate :473-491 */
PICG_API int picg_solver_create(picg_world_t w, unsigned max_it, double tol, picg_solver_t* out);
PICG_API int picg_solver_destroy(picg_solver_t s);
/* setReferenceValues(phi0,n0,Te0)  PotentialSolver.cpp:409-413 */
PICG_API int picg_solver_set_reference(picg_solver_t s, double phi0, double n0, double Te0);
/* boundary mode: 0 = v3/ch3 zero-gradient faces updated in-sweep (PotentialSolver.cpp:96-107),
 *                1 = ch2 interior-only sweep, faces keep their (Dirichlet) values (ch2/v2/PotentialSolver.cpp:39-52) */
/* How an iteration sweeps the mesh: 0 (default) = two colour sweeps, one block per (i, j) row; 1 = ONE plane-marching pass over
 * shared-memory tiles that updates both colours (26 B of DRAM traffic per node and iteration instead of 33, but more instructions:
 * slower on B200, kept as an option).  Bit-identical results. */
PICG_API int picg_solver_set_sweep(picg_solver_t s, int mode);
PICG_API int picg_solver_set_boundary_mode(picg_solver_t s, int mode);
/* solveGS  PotentialSolver.cpp:69-166 as red-black SOR (w=1.4), residual every 25 iterations normalised by nv */
PICG_API int picg_solver_solve_gs(picg_solver_t s, int* converged, unsigned* iterations, double* L2);
/* run exactly n iterations without convergence checks (benchmarks) */
PICG_API int picg_solver_iterate(picg_solver_t s, unsigned n);
PICG_API int picg_solver_residual(picg_solver_t s, double* L2);
/* computeEF  PotentialSolver.cpp:354-408 */
PICG_API int picg_solver_compute_ef(picg_solver_t s);
/* PotentialSolver::solveNRPCG  PotentialSolver.cpp:178-240 (Newton-Raphson, <= 20 steps, NR_TOL 1e-3) with solvePCGlinear :242-313
 * (Jacobi-preconditioned CG, <= pcg_max_it iterations, vec::norm(g) < tol) and the solveGSlinear fallback :315-347, matrix-free on the
 * device.  xz_swap != 0 reproduces buildMatrix's REGULAR rows as written (inv_d2z on the i-neighbours, inv_d2x on the k-neighbours,
 * :457-463: the reference's PCG results on meshes with dx != dz); 0 uses the operator solveGS relaxes.  Single GPU only. */
PICG_API int picg_solver_solve_nrpcg(picg_solver_t s, int xz_swap, unsigned pcg_max_it /* 0: the solver's max_it; the GS fallback gets 20x */, int* converged, unsigned* nr_iterations, unsigned* pcg_iterations, double* norm);
/* multi-GPU (one process per GPU, one node): split the planes of the slowest index over `world` ranks.  Halo planes, the
 * residual sum and the final all-gather of phi go through peer memory (CUDA IPC over NVLink), inside the sweep kernels.
 * export: 128 bytes per rank (IPC handles of phi and of the mailbox); the caller all-gathers them and passes the table of
 * world x 128 bytes to enable (a collective: every rank must call both).  Every rank must then make the same solver calls. */
PICG_API int picg_solver_slab_export(picg_solver_t s, void* handle128);
PICG_API int picg_solver_slab_enable(picg_solver_t s, int rank, int world, const void* handles);
/* the node range [begin, end) of this rank's planes (the whole grid when slabs are off): the solve reads rho only there, so a
 * multi-GPU loop may reduce-scatter the density accumulators onto the slabs and finalize / sum charges on the owned range only */
PICG_API int picg_solver_slab_range(picg_solver_t s, size_t* node_begin, size_t* node_end);

/* -------------------------------------------------------- MC_MEX_Ionization */
/* MC_MEX_Ionization(neutrals, ions, electrons, world, table)  Interactions.cpp:476-539; the cross-section table
 * (2 columns eV, m^2) is passed in memory instead of by path; the ctor preconditions are checked (PICG_ERR_ARG). */
PICG_API int picg_mcc_create(picg_species_t neutrals, picg_species_t ions, picg_species_t electrons, picg_world_t w,
                             const double* table_E, const double* table_sigma, int n_table, double E_ion_J, picg_mcc_t* out);
PICG_API int picg_mcc_destroy(picg_mcc_t m);
typedef struct { uint64_t candidates, collisions, ionizations; double w_sigma_v_max;
                 uint64_t dropped; /* collisions skipped (untouched) because a product store was full; 0 in a healthy run */
                 uint64_t extras_capped; /* split-off neutrals beyond 16 per cell and call: created, but not selectable by later candidates of the same call (Interactions.cpp:699-701 has no bound) */
                 uint64_t nan_products;  /* fixed-weight variant: ejected electrons created with a NaN velocity (ionisation below the threshold; ch4/v2 appends them, they die at the next push between the electrodes) */
               } picg_mcc_stats;
/* Interaction::apply(dt) -> MC_MEX_Ionization::apply_vector_indexes  Interactions.cpp:567-762 */
PICG_API int picg_mcc_apply(picg_mcc_t m, double dt, picg_mcc_stats* stats /*may be NULL*/);
PICG_API int picg_mcc_set_wsv_max(picg_mcc_t m, double v);
/* variant 0 (default): the variable-weight algorithm of ch4/v3.  variant 1: MC_MEX_Ionization::apply of ch4/v2 with fixed weights
 * (ch4/v2/Interactions.cpp:566-641, collide :678-735, evaluateSigmaIon :560-563; BASELINE config 3): candidate count
 * 0.5*np_n*np_e*neutrals.mpw0*(sigma v)_max*dt/dV, unweighted acceptance, no ionisation threshold guard, products through
 * Species::addParticle (ch4/v2/Species.cpp:226-237: bounds / object filter, half-step rewind), neutrals never depleted. */
PICG_API int picg_mcc_set_variant(picg_mcc_t m, int variant);
/* debug / tests: lengths of the exact per-cell particle lists the collision kernel uses (0 neutrals, 1 electrons) */
PICG_API int picg_mcc_list_counts(picg_mcc_t m, int which, double* cells);
PICG_API int picg_mcc_sigma(picg_mcc_t m, int n, const double* E_eV, double* sigma_coll, double* sigma_ion); /* evaluateSigmaColl/Ion :541-566 */

/* ----------------------------------------------------------------- DSMC_MEX */
/* DSMC_MEX(species, world) / DSMC_MEX(species1, species2, world)  Interactions.cpp:143-176: Bird NTC collisions with the
 * variable-hard-sphere cross-section between macro-particles of equal weight.  species2 == NULL (or == species1): collisions
 * within one species.  Different mpw0 -> PICG_ERR_ARG (the reference's std::invalid_argument, :160-162). */
PICG_API int picg_dsmc_create(picg_species_t species1, picg_species_t species2 /*may be NULL*/, picg_world_t w, picg_dsmc_t* out);
PICG_API int picg_dsmc_destroy(picg_dsmc_t m);
typedef struct { uint64_t candidates, collisions; double sigma_v_max; } picg_dsmc_stats;
/* Interaction::apply(dt) -> DSMC_MEX::applyOneSpecies / applyTwoSpecies  Interactions.cpp:183-265 (collide :267-285).
 * stats == NULL: nothing is read back and the call does not synchronise. */
PICG_API int picg_dsmc_apply(picg_dsmc_t m, double dt, picg_dsmc_stats* stats /*may be NULL*/);
PICG_API int picg_dsmc_set_sigma_v_max(picg_dsmc_t m, double v);                             /* sigma_v_rel_max  Interactions.h:58 */
PICG_API int picg_dsmc_sigma(picg_dsmc_t m, int n, const double* v_rel, double* sigma);      /* evaluateSigma :178-181 */

/* ------------------------------------------------------------------- Source */
/* ColdBeamSource / WarmBeamSource  Source.cpp:3-191; face: 0 x- 1 x+ 2 y- 3 y+ 4 z- 5 z+ ; T<=0 => cold */
PICG_API int picg_source_create(picg_species_t s, picg_world_t w, double v_drift, double den, double T, int face, picg_source_t* out);
PICG_API int picg_source_destroy(picg_source_t src);
/* Source::sample()  Source.cpp:99-103,187-191 */
PICG_API int picg_source_sample(picg_source_t src, size_t* injected /*may be NULL*/);

/* ------------------------------------------------------- outputs and restart */
/* Output::fieldsOutput  Outputs.cpp:9-123: the same VTK ImageData file (arrays NodeVol, ObjectID, NodeType, phi, rho, nd.<sp>,
 * avg_nd.<sp>, vel.<sp>, T.<sp>, ef as PointData and mpc.<sp> as CellData, in that order, all Float64, points in VTK order = i
 * fastest) with the data as one appended raw binary block instead of ASCII.  names[k] is Species::name of species[k].  The
 * caller runs computeGasProperties / clearSamples around it like the reference (:11-14, :119-121). */
PICG_API int picg_write_fields_vti(const char* path, picg_world_t w, const picg_species_t* species, const char* const* names, int n);
/* Restart files (the reference has none).  A checkpoint holds what the time loop carries from step to step: phi, rho, ef; per
 * species the particle store, den, den_avg (+ sample count), the moment sums, the fixed-point scale and the positions of its
 * RNG streams; W_sigma_v_rel_max / sigma_v_rel_max and the call counters of the interactions; the call counters of the sources;
 * the seed.  The caller rebuilds World (mesh, objects), Species, interactions and sources as at start-up and passes them in the
 * same order to load; mesh, species constants and counts are checked (PICG_ERR_ARG).  Deterministic kernels (push, deposit,
 * fields) continue bit for bit; stochastic kernels continue on the same streams (the cell partition is rebuilt, so the order of
 * a cell's particle list - hence the pairs drawn - may differ from the uninterrupted run). */
typedef struct {
    picg_world_t world;
    const picg_species_t* species; int n_species;
    const picg_mcc_t* mcc; int n_mcc;
    const picg_dsmc_t* dsmc; int n_dsmc;
    const picg_source_t* sources; int n_sources;
} picg_checkpoint_set;
PICG_API int picg_checkpoint_save(const char* path, const picg_checkpoint_set* set, uint64_t user_ts /* e.g. World::getTs() */);
PICG_API int picg_checkpoint_load(const char* path, const picg_checkpoint_set* set, uint64_t* user_ts /*may be NULL*/);

/* ------------------------------------------------------- fp32 secondary path (north star: "1e-4 relative for fp32", float4 loads)
 * The reference is written against one scalar type (all.h:11 `using type_calc = double`) and rebuilds with float.  Here a species can
 * live in a single-precision store instead: 32 bytes per particle - the cell index (u32) and the cell-relative coordinates fx, fy, fz
 * in [0,1), velocities and weight as floats - because an absolute fp32 position cannot carry the motion of the slow species (a neutral
 * moves 4e-10 m per step, the fp32 spacing at 2.5 cm is 1.9e-9 m).  Node fields stay fp64.  Covered: the electron-type push
 * (Species::advanceElectrons, Species.cpp:258-399: gather, kick, drift, absorption outside the box / inside an object), number density +
 * per-cell count (:401-416, :813-819; same int64 fixed-point grid as the fp64 path, deterministic), cell sort, diagnostics (:731-752),
 * charge density (World.cpp:193-200).  The heavy species' wall interaction and the collision kernels are fp64 only.
 * Host buffers are the same AoS of doubles (x y z u v w mpw) as for the fp64 store. */
PICG_API int picg_species32_create(picg_world_t w, double mass, double charge, double mpw0, picg_species32_t* out);
PICG_API int picg_species32_destroy(picg_species32_t s);
PICG_API int picg_species32_reserve(picg_species32_t s, size_t capacity);
PICG_API int picg_species32_count(picg_species32_t s, size_t* n);
PICG_API int picg_species32_upload(picg_species32_t s, size_t n, const double* aos7);
PICG_API int picg_species32_from_species(picg_species32_t s, picg_species_t src);      /* device-side conversion of an fp64 store of the same world */
PICG_API int picg_species32_download(picg_species32_t s, size_t capacity, double* aos7, size_t* n_out);
PICG_API int picg_species32_push_electrons(picg_species32_t s, double dt);
PICG_API int picg_species32_deposit_density(picg_species32_t s);                        /* + computeMacroParticlesCount as a by-product */
PICG_API int picg_species32_set_density_scale(picg_species32_t s, int S);
PICG_API int picg_species32_density_scale(picg_species32_t s, int* S);
PICG_API int picg_species32_sort(picg_species32_t s);
PICG_API int picg_species32_diagnostics(picg_species32_t s, double* micro_count, double momentum[3], double* ke);
PICG_API int picg_species32_download_field(picg_species32_t s, int field /* PICG_SF_DEN, _DEN_FIXED, _MACRO_COUNT */, void* host);
PICG_API int picg_world_charge_density32(picg_world_t w, picg_species32_t* species, int n);

#ifdef __cplusplus
}
#endif
#endif /* PICGPU_H */
