/* TEST INFRASTRUCTURE -- see pic_oracle.h.  Plain C restatement of the reference's hot path.
 * Compile with -ffp-contract=off and without -march=native so no FMA is formed. */
#include "pic_oracle.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>

#define EPS0 8.85418782e-12    /* all.h:15 */
#define QE   1.602176565e-19   /* all.h:16 */

static size_t U(const orc_grid* g, int i, int j, int k) { return ((size_t)i * g->nj + j) * g->nk + k; }

void orc_grid_init(orc_grid* g, int ni, int nj, int nk, const double x0[3], const double xm[3]) {
    memset(g, 0, sizeof(*g));
    g->ni = ni; g->nj = nj; g->nk = nk;
    int nn[3] = {ni, nj, nk};
    for (int a = 0; a < 3; a++) {                      /* World::setExtents World.cpp:63-77 */
        g->x0[a] = x0[a]; g->xm[a] = xm[a];
        g->dx[a] = (xm[a] - x0[a]) / (nn[a] - 1);
        g->inv_dx[a] = 1 / g->dx[a];
    }
}
void orc_add_rectangle(orc_grid* g, const double c[3], double phi, const double sides[3]) {
    orc_object* o = &g->obj[g->n_obj++];
    memset(o, 0, sizeof(*o)); o->type = 0; o->phi = phi;
    for (int a = 0; a < 3; a++) { o->c[a] = c[a]; o->h[a] = sides[a] * 0.5; o->lo[a] = c[a] - o->h[a]; o->hi[a] = c[a] + o->h[a]; }
}
void orc_add_sphere(orc_grid* g, const double c[3], double phi, double r) {
    orc_object* o = &g->obj[g->n_obj++];
    memset(o, 0, sizeof(*o)); o->type = 1; o->phi = phi;
    for (int a = 0; a < 3; a++) o->c[a] = c[a];
    o->h[0] = r * r; o->h[1] = r;
}
void orc_node_volumes(const orc_grid* g, double* vol) {       /* World.cpp:353-367 */
    double base = g->dx[0] * g->dx[1] * g->dx[2];
    for (int i = 0; i < g->ni; i++) for (int j = 0; j < g->nj; j++) for (int k = 0; k < g->nk; k++) {
        double v = base;
        if (i == 0 || i == g->ni - 1) v *= 0.5;
        if (j == 0 || j == g->nj - 1) v *= 0.5;
        if (k == 0 || k == g->nk - 1) v *= 0.5;
        vol[U(g, i, j, k)] = v;
    }
}
static int obj_contains(const orc_object* o, const double p[3]) {
    double r[3] = {p[0] - o->c[0], p[1] - o->c[1], p[2] - o->c[2]};
    if (o->type == 0) {                                          /* Rectangle::inObject Object.cpp:231-238 */
        for (int a = 0; a < 3; a++) if (fabs(r[a]) > o->h[a]) return 0;
        return 1;
    }
    return (r[0] * r[0] + r[1] * r[1] + r[2] * r[2] <= o->h[0]); /* Sphere::inObject Object.cpp:111-115 */
}
int orc_in_object(const orc_grid* g, const double p[3]) {       /* World.cpp:293-301 */
    for (int o = 0; o < g->n_obj; o++) if (obj_contains(&g->obj[o], p)) return o + 1;
    return 0;
}
int orc_in_bounds(const orc_grid* g, const double p[3]) {       /* World.cpp:201-205 */
    for (int a = 0; a < 3; a++) if (p[a] < g->x0[a] || p[a] >= g->xm[a]) return 0;
    return 1;
}
void orc_compute_object_id(const orc_grid* g, int* object_id, double* phi) {   /* World.cpp:276-292 */
    for (int o = 0; o < g->n_obj; o++)
        for (int i = 0; i < g->ni; i++) for (int j = 0; j < g->nj; j++) for (int k = 0; k < g->nk; k++) {
            double p[3] = {g->x0[0] + (double)i * g->dx[0], g->x0[1] + (double)j * g->dx[1], g->x0[2] + (double)k * g->dx[2]};
            if (obj_contains(&g->obj[o], p)) { object_id[U(g, i, j, k)] = 1; if (phi) phi[U(g, i, j, k)] = g->obj[o].phi; }
        }
}

static void x_to_l(const orc_grid* g, const double p[3], double lc[3]) {       /* World::XtoL World.cpp:123-127 */
    for (int a = 0; a < 3; a++) lc[a] = (p[a] - g->x0[a]) * g->inv_dx[a];
}
static int imin(int a, int b) { return a < b ? a : b; }

void orc_gather_ef(const orc_grid* g, const double* ef, const double p[3], double e[3]) {   /* Field.h:201-232 */
    double lc[3]; x_to_l(g, p, lc);
    int i = imin((int)lc[0], g->ni - 2), j = imin((int)lc[1], g->nj - 2), k = imin((int)lc[2], g->nk - 2);
    double di = lc[0] - i, dj = lc[1] - j, dk = lc[2] - k;
    double odi = 1 - di, odj = 1 - dj, odk = 1 - dk;
    double wa = odi * odj, wb = odi * dj, wc = di * odj, wd = di * dj;
    const double* r00 = ef + U(g, i, j, k) * 3; const double* r01 = ef + U(g, i, j + 1, k) * 3;
    const double* r10 = ef + U(g, i + 1, j, k) * 3; const double* r11 = ef + U(g, i + 1, j + 1, k) * 3;
    for (int c = 0; c < 3; c++) {
        double v = r00[c] * wa * odk;
        v = v + r00[3 + c] * wa * dk;
        v = v + r01[c] * wb * odk;
        v = v + r01[3 + c] * wb * dk;
        v = v + r10[c] * wc * odk;
        v = v + r10[3 + c] * wc * dk;
        v = v + r11[c] * wd * odk;
        v = v + r11[3 + c] * wd * dk;
        e[c] = v;
    }
}

static void kick_drift(const orc_grid* g, const double* ef, double qm_dt, double dt, double* a) {  /* Species.cpp:368-373 */
    double e[3]; orc_gather_ef(g, ef, a, e);
    for (int c = 0; c < 3; c++) a[3 + c] = a[3 + c] + e[c] * qm_dt;
    for (int c = 0; c < 3; c++) a[c] = a[c] + a[3 + c] * dt;
}
void orc_push_electrons(const orc_grid* g, const double* ef, double charge, double mass, double dt, size_t n, double* aos7, unsigned char* alive) {
    double qm_dt = dt * charge / mass;
    for (size_t p = 0; p < n; p++) {
        double* a = aos7 + 7 * p;
        kick_drift(g, ef, qm_dt, dt, a);
        alive[p] = (orc_in_bounds(g, a) && !orc_in_object(g, a)) ? 1 : 0;      /* Species.cpp:375-388 */
    }
}
void orc_push_reflect(const orc_grid* g, const double* ef, double charge, double mass, double dt, size_t n, double* aos7) {
    double qm_dt = dt * charge / mass;
    for (size_t p = 0; p < n; p++) {
        double* a = aos7 + 7 * p;
        kick_drift(g, ef, qm_dt, dt, a);
        for (int c = 0; c < 3; c++) {                                              /* ch2/v2/Species.cpp:44-53 */
            if (a[c] < g->x0[c]) { a[c] = 2.0 * g->x0[c] - a[c]; a[3 + c] *= -1.0; }
            else if (a[c] >= g->xm[c]) { a[c] = 2.0 * g->xm[c] - a[c]; a[3 + c] *= -1.0; }
        }
    }
}
size_t orc_add_particles(const orc_grid* g, const double* ef, double charge, double mass, double world_dt, size_t n, double* aos7) {
    size_t m = 0;
    double q_over_m = charge / mass, half_dt = 0.5 * world_dt;
    for (size_t p = 0; p < n; p++) {
        double a[7]; memcpy(a, aos7 + 7 * p, sizeof(a));
        int bad = 0; for (int c = 0; c < 6; c++) if (isnan(a[c])) bad = 1;
        if (bad || !orc_in_bounds(g, a) || orc_in_object(g, a)) continue;
        double e[3]; orc_gather_ef(g, ef, a, e);
        for (int c = 0; c < 3; c++) a[3 + c] = a[3 + c] - (e[c] * q_over_m) * half_dt;   /* Species.cpp:431 */
        memcpy(aos7 + 7 * m, a, sizeof(a)); m++;
    }
    return m;
}

/* the eight contributions of Field::scatter (Field.h:172-197), reference association */
static void scatter_contrib(const orc_grid* g, const double* a, double val, int* i, int* j, int* k, double c[8]) {
    double lc[3]; x_to_l(g, a, lc);
    *i = imin((int)lc[0], g->ni - 2); *j = imin((int)lc[1], g->nj - 2); *k = imin((int)lc[2], g->nk - 2);
    double di = lc[0] - *i, dj = lc[1] - *j, dk = lc[2] - *k;
    double odi = 1 - di, odj = 1 - dj, odk = 1 - dk;
    double w00 = val * odi * odj, w01 = val * odi * dj, w10 = val * di * odj, w11 = val * di * dj;
    c[0] = w00 * odk; c[1] = w00 * dk; c[2] = w01 * odk; c[3] = w01 * dk;
    c[4] = w10 * odk; c[5] = w10 * dk; c[6] = w11 * odk; c[7] = w11 * dk;
}
static size_t corner(const orc_grid* g, int i, int j, int k, int c) { return U(g, i + (c >> 2), j + ((c >> 1) & 1), k + (c & 1)); }

void orc_deposit_fixed(const orc_grid* g, size_t n, const double* aos7, int S, int64_t* fixed) {
    size_t nv = (size_t)g->ni * g->nj * g->nk;
    memset(fixed, 0, nv * sizeof(int64_t));
    double scale = ldexp(1.0, S);
    for (size_t p = 0; p < n; p++) {
        const double* a = aos7 + 7 * p;
        int i, j, k; double c[8];
        scatter_contrib(g, a, a[6], &i, &j, &k, c);
        for (int q = 0; q < 8; q++) fixed[corner(g, i, j, k, q)] += llrint(c[q] * scale);
    }
}
void orc_finalize_density(const orc_grid* g, const int64_t* fixed, int S, const double* vol, double* den) {
    size_t nv = (size_t)g->ni * g->nj * g->nk;
    double inv = ldexp(1.0, -S);
    for (size_t u = 0; u < nv; u++) { double d = (double)fixed[u] * inv; den[u] = vol[u] != 0 ? d / vol[u] : 0; }
}
void orc_deposit_fp64(const orc_grid* g, size_t n, const double* aos7, const double* vol, double* den) {
    size_t nv = (size_t)g->ni * g->nj * g->nk;
    memset(den, 0, nv * sizeof(double));
    for (size_t p = 0; p < n; p++) {
        const double* a = aos7 + 7 * p;
        int i, j, k; double c[8];
        scatter_contrib(g, a, a[6], &i, &j, &k, c);
        for (int q = 0; q < 8; q++) den[corner(g, i, j, k, q)] += c[q];
    }
    for (size_t u = 0; u < nv; u++) { if (vol[u] != 0) den[u] /= vol[u]; else den[u] = 0; }
}
void orc_count_per_cell(const orc_grid* g, size_t n, const double* aos7, double* count) {
    size_t nc = (size_t)(g->ni - 1) * (g->nj - 1) * (g->nk - 1);
    memset(count, 0, nc * sizeof(double));
    for (size_t p = 0; p < n; p++) {
        double lc[3]; x_to_l(g, aos7 + 7 * p, lc);
        int i = imin((int)lc[0], g->ni - 2), j = imin((int)lc[1], g->nj - 2), k = imin((int)lc[2], g->nk - 2);
        count[((size_t)i * (g->nj - 1) + j) * (g->nk - 1) + k] += 1;
    }
}
void orc_sample_moments(const orc_grid* g, size_t n, const double* aos7, double* n_sum, double* nv_sum, double* nuu, double* nvv, double* nww) {
    for (size_t p = 0; p < n; p++) {                               /* Species.cpp:767-776, accumulates (no clear) */
        const double* a = aos7 + 7 * p;
        double m = a[6], vals[7] = {m, m * a[3], m * a[4], m * a[5], m * a[3] * a[3], m * a[4] * a[4], m * a[5] * a[5]};
        for (int f = 0; f < 7; f++) {
            int i, j, k; double c[8];
            scatter_contrib(g, a, vals[f], &i, &j, &k, c);
            for (int q = 0; q < 8; q++) {
                size_t u = corner(g, i, j, k, q);
                if (f == 0) n_sum[u] += c[q]; else if (f <= 3) nv_sum[3 * u + f - 1] += c[q];
                else if (f == 4) nuu[u] += c[q]; else if (f == 5) nvv[u] += c[q]; else nww[u] += c[q];
            }
        }
    }
}
void orc_charge_density(const orc_grid* g, int ns, const double* const* den, const double* charge, double* rho) {
    size_t nv = (size_t)g->ni * g->nj * g->nk;
    for (size_t u = 0; u < nv; u++) {
        double r = 0;
        for (int s = 0; s < ns; s++) { if (charge[s] == 0) continue; r = r + charge[s] * den[s][u]; }
        rho[u] = r;
    }
}

/* ------------------------------------------------------------------ Poisson */
typedef struct { double inv_d2x, inv_d2y, inv_d2z, inv_eps0, twos, inv_twos; } sor_consts;
static sor_consts precalc(const orc_grid* g) {                   /* PotentialSolver::precalculate :473-491 */
    sor_consts c;
    c.inv_d2x = 1.0 / (g->dx[0] * g->dx[0]); c.inv_d2y = 1.0 / (g->dx[1] * g->dx[1]); c.inv_d2z = 1.0 / (g->dx[2] * g->dx[2]);
    c.inv_eps0 = 1.0 / EPS0; c.twos = 2.0 * (c.inv_d2x + c.inv_d2y + c.inv_d2z); c.inv_twos = 1.0 / c.twos;
    return c;
}
/* 0 skip, 1..6 zero-gradient face (first match i0,iN,j0,jN,k0,kN), 7 interior */
static int node_class(const orc_grid* g, int bc_mode, int oid, int i, int j, int k) {
    if (oid > 0) return 0;
    int face = (i == 0 || i == g->ni - 1 || j == 0 || j == g->nj - 1 || k == 0 || k == g->nk - 1);
    if (!face) return 7;
    if (bc_mode == 1) return 0;
    if (i == 0) return 1; if (i == g->ni - 1) return 2; if (j == 0) return 3; if (j == g->nj - 1) return 4; if (k == 0) return 5; return 6;
}
static size_t face_nb(const orc_grid* g, int cls, size_t u) {
    size_t si = (size_t)g->nj * g->nk, sj = g->nk;
    switch (cls) { case 1: return u + si; case 2: return u - si; case 3: return u + sj; case 4: return u - sj; case 5: return u + 1; default: return u - 1; }
}
static void update_node(const orc_grid* g, const sor_consts* c, int cls, size_t u, const double* rho, double* phi, double phi0, double n0, double Te0) {
    size_t si = (size_t)g->nj * g->nk, sj = g->nk;
    if (cls < 7) { phi[u] = phi[face_nb(g, cls, u)]; return; }
    double ne = n0 * exp((phi[u] - phi0) / Te0);
    double nw = ((rho[u] - QE * ne) * c->inv_eps0 + (phi[u - si] + phi[u + si]) * c->inv_d2x + (phi[u - sj] + phi[u + sj]) * c->inv_d2y +
                 (phi[u - 1] + phi[u + 1]) * c->inv_d2z) * c->inv_twos;
    phi[u] = phi[u] + 1.4 * (nw - phi[u]);
}
double orc_residual(const orc_grid* g, const int* object_id, const double* rho, const double* phi, double phi0, double n0, double Te0, int bc_mode) {
    sor_consts c = precalc(g);
    size_t si = (size_t)g->nj * g->nk, sj = g->nk;
    double sum = 0;
    for (int i = 0; i < g->ni; i++) for (int j = 0; j < g->nj; j++) for (int k = 0; k < g->nk; k++) {
        size_t u = U(g, i, j, k);
        int cls = node_class(g, bc_mode, object_id[u], i, j, k);
        if (cls == 0) continue;
        double R;
        if (cls < 7) R = phi[u] - phi[face_nb(g, cls, u)];
        else {
            double ne = n0 * exp((phi[u] - phi0) / Te0);
            R = -phi[u] * c.twos + (rho[u] - QE * ne) * c.inv_eps0 + (phi[u - si] + phi[u + si]) * c.inv_d2x + (phi[u - sj] + phi[u + sj]) * c.inv_d2y +
                (phi[u - 1] + phi[u + 1]) * c.inv_d2z;
        }
        sum += R * R;
    }
    return sqrt(sum / ((double)g->ni * g->nj * g->nk));
}
static int solve_generic(const orc_grid* g, const int* object_id, const double* rho, double* phi, unsigned max_it, double tol,
                         double phi0, double n0, double Te0, int bc_mode, unsigned* iters, double* L2out, int redblack) {
    sor_consts c = precalc(g);
    double L2 = 0; int conv = 0; unsigned it;
    for (it = 0; it < max_it; it++) {
        if (!redblack) {
            for (int i = 0; i < g->ni; i++) for (int j = 0; j < g->nj; j++) for (int k = 0; k < g->nk; k++) {
                size_t u = U(g, i, j, k);
                int cls = node_class(g, bc_mode, object_id[u], i, j, k);
                if (cls) update_node(g, &c, cls, u, rho, phi, phi0, n0, Te0);
            }
        } else {
            for (int color = 0; color < 2; color++)
                for (int i = 0; i < g->ni; i++) for (int j = 0; j < g->nj; j++) for (int k = (i + j + color) & 1; k < g->nk; k += 2) {
                    size_t u = U(g, i, j, k);
                    int cls = node_class(g, bc_mode, object_id[u], i, j, k);
                    if (cls) update_node(g, &c, cls, u, rho, phi, phi0, n0, Te0);
                }
        }
        if (it % 25 == 0) {
            L2 = orc_residual(g, object_id, rho, phi, phi0, n0, Te0, bc_mode);
            if (L2 < tol) { conv = 1; it++; break; }
        }
    }
    if (iters) *iters = it; if (L2out) *L2out = L2;
    return conv;
}
int orc_solve_gs(const orc_grid* g, const int* object_id, const double* rho, double* phi, unsigned max_it, double tol,
                 double phi0, double n0, double Te0, int bc_mode, unsigned* iters, double* L2) {
    return solve_generic(g, object_id, rho, phi, max_it, tol, phi0, n0, Te0, bc_mode, iters, L2, 0);
}
int orc_solve_rb(const orc_grid* g, const int* object_id, const double* rho, double* phi, unsigned max_it, double tol,
                 double phi0, double n0, double Te0, int bc_mode, unsigned* iters, double* L2) {
    return solve_generic(g, object_id, rho, phi, max_it, tol, phi0, n0, Te0, bc_mode, iters, L2, 1);
}
void orc_compute_ef(const orc_grid* g, const double* phi, double* ef) {        /* PotentialSolver.cpp:354-408 */
    size_t si = (size_t)g->nj * g->nk, sj = g->nk;
    double i2x = 1.0 / (2 * g->dx[0]), i2y = 1.0 / (2 * g->dx[1]), i2z = 1.0 / (2 * g->dx[2]);
    for (int i = 0; i < g->ni; i++) for (int j = 0; j < g->nj; j++) for (int k = 0; k < g->nk; k++) {
        size_t u = U(g, i, j, k);
        double* e = ef + 3 * u;
        if (i == 0) e[0] = (3 * phi[u] - 4 * phi[u + si] + phi[u + 2 * si]) * i2x;
        else if (i == g->ni - 1) e[0] = (-phi[u - 2 * si] + 4 * phi[u - si] - 3 * phi[u]) * i2x;
        else e[0] = (phi[u - si] - phi[u + si]) * i2x;
        if (j == 0) e[1] = (3 * phi[u] - 4 * phi[u + sj] + phi[u + 2 * sj]) * i2y;
        else if (j == g->nj - 1) e[1] = (-phi[u - 2 * sj] + 4 * phi[u - sj] - 3 * phi[u]) * i2y;
        else e[1] = (phi[u - sj] - phi[u + sj]) * i2y;
        if (k == 0) e[2] = (3 * phi[u] - 4 * phi[u + 1] + phi[u + 2]) * i2z;
        else if (k == g->nk - 1) e[2] = (-phi[u - 2] + 4 * phi[u - 1] - 3 * phi[u]) * i2z;
        else e[2] = (phi[u - 1] - phi[u + 1]) * i2z;
    }
}

/* ------------------------------------------------------------------ Philox4x32-10 */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ---------------------------------------------------------------- DSMC_MEX (Interactions.cpp:143-285) */
double orc_dsmc_sigma(double m1, double m2, double v_rel) {
    const double m_reduced = m1 * m2 / (m1 + m2);                 /* :146 */
    const double c0 = 4.07e-10, c1 = 0.77;                        /* :151-152 */
    const double c2 = 2 * 1.380648e-23 * 273.15 / m_reduced;      /* :153, Const::k all.h:19 */
    const double c3 = tgamma(2.5 - c1);                           /* :154 */
    return 3.141592653 * c0 * c0 * pow(c2 / (v_rel * v_rel), c1 - 0.5) / c3;   /* :179 */
}
void orc_dsmc_collide(double m1, double m2, double r1, double r2, double v1[3], double v2[3]) {
    const double sum_mass = m1 + m2;
    double cm[3], g[3];
    const double inv_sum = 1.0 / sum_mass;                        /* Vec3::operator/(scalar) multiplies by the reciprocal, Vec3.h:180-192 */
    for (int c = 0; c < 3; c++) { cm[c] = (m1 * v1[c] + m2 * v2[c]) * inv_sum; g[c] = v1[c] - v2[c]; }   /* :268-270 */
    const double g_mag = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    const double cos_ksi = 2 * r1 - 1;                            /* :273-275 */
    const double sin_ksi = sqrt(1 - cos_ksi * cos_ksi);
    const double eps = 2 * 3.141592653 * r2;
    g[0] = g_mag * cos_ksi; g[1] = g_mag * sin_ksi * cos(eps); g[2] = g_mag * sin_ksi * sin(eps);          /* :278-280 */
    for (int c = 0; c < 3; c++) { v1[c] = cm[c] + m2 / sum_mass * g[c]; v2[c] = cm[c] - m1 / sum_mass * g[c]; }   /* :282-283 */
}
