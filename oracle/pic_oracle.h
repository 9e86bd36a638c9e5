/* TEST INFRASTRUCTURE -- CPU restatement of the reference's per-timestep particle loop.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this library;
 * the product path (libpicgpu.so) never links, loads or calls it.
 *
 * Parity status: PINNED against the compiled reference (oracle/_ref/libref_v3.so, built from the
 * unmodified sources under /root/reference/ch4/v3/src) by tests/test_oracle_vs_reference.py and
 * against the golden vectors in tests/golden/ generated from that library (tests/golden/make_golden.py).
 * The reference itself ships no tests or golden vectors.
 *
 * Layouts: node fields double[nv] with u=(i*nj+j)*nk+k (Field.h:16,88); ef double[3*nv] interleaved;
 * particles AoS 7 doubles x y z u v w mpw (Species.h:12-29); cells (i*(nj-1)+j)*(nk-1)+k.
 */
#ifndef PIC_ORACLE_H
#define PIC_ORACLE_H
#include <stddef.h>
#include <stdint.h>

typedef struct {
    int type;            /* 0 rectangle, 1 sphere */
    double c[3];         /* centre */
    double h[3];         /* rectangle: half sides ; sphere: h[0] = r^2, h[1] = r */
    double lo[3], hi[3]; /* rectangle x_min / x_max */
    double phi;
} orc_object;

typedef struct {
    int ni, nj, nk;
    double x0[3], xm[3], dx[3], inv_dx[3];
    int n_obj;
    orc_object obj[8];
} orc_grid;

void orc_grid_init(orc_grid* g, int ni, int nj, int nk, const double x0[3], const double xm[3]);           /* World.cpp:63-77 */
void orc_add_rectangle(orc_grid* g, const double c[3], double phi, const double sides[3]);                /* Object.cpp:166-172 */
void orc_add_sphere(orc_grid* g, const double c[3], double phi, double r);
void orc_node_volumes(const orc_grid* g, double* vol);                                                     /* World.cpp:353-367 */
void orc_compute_object_id(const orc_grid* g, int* object_id, double* phi);                               /* World.cpp:276-292 */
int  orc_in_bounds(const orc_grid* g, const double p[3]);                                                  /* World.cpp:201-205 */
int  orc_in_object(const orc_grid* g, const double p[3]);                                                  /* World.cpp:293-301 */
void orc_gather_ef(const orc_grid* g, const double* ef, const double p[3], double e[3]);                  /* Field.h:201-232 */

/* Species::advanceElectronsSerial (Species.cpp:356-399) WITHOUT the removal: alive[p]=0 marks particles the
 * reference would delete; positions/velocities of dead particles are left as pushed. */
void orc_push_electrons(const orc_grid* g, const double* ef, double charge, double mass, double dt, size_t n, double* aos7, unsigned char* alive);
/* ch2 Species::advance (ch2/v2/Species.cpp:18-55) with v3 XtoL/gather arithmetic */
void orc_push_reflect(const orc_grid* g, const double* ef, double charge, double mass, double dt, size_t n, double* aos7);
/* Species::addParticle filter + half-step rewind (Species.cpp:420-434); returns number kept, compacted in place */
size_t orc_add_particles(const orc_grid* g, const double* ef, double charge, double mass, double world_dt, size_t n, double* aos7);

/* fixed-point restatement of Species::computeNumberDensity: sum over particles of llrint(c*2^S), c formed as Field.h:172-197 */
void orc_deposit_fixed(const orc_grid* g, size_t n, const double* aos7, int S, int64_t* fixed);
/* den = (fixed*2^-S)/node_vol with zero-divisor guard (Field.h:563-583) */
void orc_finalize_density(const orc_grid* g, const int64_t* fixed, int S, const double* vol, double* den);
/* the reference's own fp64, particle-order summation (for normwise comparison) */
void orc_deposit_fp64(const orc_grid* g, size_t n, const double* aos7, const double* vol, double* den);
void orc_count_per_cell(const orc_grid* g, size_t n, const double* aos7, double* count);                  /* Species.cpp:813-819 */
void orc_sample_moments(const orc_grid* g, size_t n, const double* aos7, double* n_sum, double* nv_sum, double* nuu, double* nvv, double* nww);
void orc_charge_density(const orc_grid* g, int ns, const double* const* den, const double* charge, double* rho); /* World.cpp:193-200 */

/* PotentialSolver::solveGS (PotentialSolver.cpp:69-166), lexicographic; bc_mode 1 = ch2 interior-only sweep.
 * Returns converged flag; *iters = sweeps done; *L2 = last residual. */
int  orc_solve_gs(const orc_grid* g, const int* object_id, const double* rho, double* phi, unsigned max_it, double tol,
                  double phi0, double n0, double Te0, int bc_mode, unsigned* iters, double* L2);
/* the same node classes and update in red-black order (what the device runs) */
int  orc_solve_rb(const orc_grid* g, const int* object_id, const double* rho, double* phi, unsigned max_it, double tol,
                  double phi0, double n0, double Te0, int bc_mode, unsigned* iters, double* L2);
double orc_residual(const orc_grid* g, const int* object_id, const double* rho, const double* phi, double phi0, double n0, double Te0, int bc_mode);
void orc_compute_ef(const orc_grid* g, const double* phi, double* ef);                                     /* PotentialSolver.cpp:354-408 */

/* DSMC_MEX (Interactions.cpp:143-285): VHS cross-section of evaluateSigma (:178-181) for the reduced mass of m1, m2, and the
 * isotropic centre-of-mass scattering of collide (:267-285) with the two uniform draws passed in (r1 -> cos_ksi, r2 -> eps). */
double orc_dsmc_sigma(double m1, double m2, double v_rel);
void orc_dsmc_collide(double m1, double m2, double r1, double r2, double v1[3], double v2[3]);

/* Philox4x32-10 (counter-based RNG used by every stochastic device kernel) */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
#endif
