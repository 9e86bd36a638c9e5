// Velocity samplers and segment/object intersections of the reference, as host+device inline functions
// over a PhiloxStream.  Shared by source.cu (loaders, inlet sources) and heavy.cu (wall interaction).
#pragma once
#include "common.cuh"
#include "philox.cuh"
#include <math.h>

#define CONST_K  1.380648e-23      // all.h:19
#define CONST_PI 3.141592653       // all.h:20 (10 digits, as in the reference)

// Species::sampleVth (Species.cpp:855-861): |v| of three components vth*(r+r+r-1.5)
PHILOX_HD double sample_vth(PhiloxStream& r, double T, double mass) {
    double v_th = sqrt(2 * CONST_K * T / mass);
    double a = r.next(), b = r.next(), c = r.next();
    double v1 = v_th * (a + b + c - 1.5);
    a = r.next(); b = r.next(); c = r.next();
    double v2 = v_th * (a + b + c - 1.5);
    a = r.next(); b = r.next(); c = r.next();
    double v3 = v_th * (a + b + c - 1.5);
    return sqrt(v1 * v1 + v2 * v2 + v3 * v3);
}
// Species::sampleV3th (Species.cpp:862-869): that speed in an isotropic direction; draw order: 9 for the speed, theta, r
PHILOX_HD void sample_v3th(PhiloxStream& r, double T, double mass, double v[3]) {
    double v_th = sample_vth(r, T, mass);
    double theta = 2 * CONST_PI * r.next();
    double rr = -1.0 + 2 * r.next();
    double a = sqrt(1 - rr * rr);
    v[0] = v_th * rr; v[1] = v_th * (cos(theta) * a); v[2] = v_th * (sin(theta) * a);
}
// Species::sampleReflectedVelocity (Species.cpp:835-853): cosine-law re-emission at T=300 K, a_th=1.
// Reproduces the reference literally: the t2 term lacks sin(theta) and t1 is not normalised.
PHILOX_HD void sample_reflected(PhiloxStream& r, double v_mag1, const double n[3], double mass, double out[3]) {
    double v_th = sample_vth(r, 300, mass);
    double v_mag2 = v_mag1 + 1.0 * (v_th - v_mag1);
    double sin_t = r.next();
    double cos_t = sqrt(1 - sin_t * sin_t);
    double psi = 2 * CONST_PI * r.next();
    double t1[3], t2[3];
    if (n[0] * 1.0 + n[1] * 0.0 + n[2] * 0.0 != 0) { t1[0] = n[1] * 0.0 - n[2] * 0.0; t1[1] = n[2] * 1.0 - n[0] * 0.0; t1[2] = n[0] * 0.0 - n[1] * 1.0; }   // n x (1,0,0)
    else { t1[0] = n[1] * 0.0 - n[2] * 1.0; t1[1] = n[2] * 0.0 - n[0] * 0.0; t1[2] = n[0] * 1.0 - n[1] * 0.0; }                                            // n x (0,1,0)
    t2[0] = n[1] * t1[2] - n[2] * t1[1]; t2[1] = n[2] * t1[0] - n[0] * t1[2]; t2[2] = n[0] * t1[1] - n[1] * t1[0];
    double a = sin_t * cos(psi), b = sin(psi);
    for (int c = 0; c < 3; c++) out[c] = v_mag2 * (a * t1[c] + b * t2[c] + cos_t * n[c]);
}

// Rectangle::lineIntersect (Object.cpp:239-290) with find_n (:292-317); precondition x2 inside the box.
PHILOX_HD void rect_line_intersect(const ObjShape& o, const double x1[3], const double x2[3], double* t_entry, double pos[3], double n[3]) {
    double tmin[3], tmax[3]; int side[3] = {0, 1, 2};
    for (int a = 0; a < 3; a++) {
        double A = x2[a] - x1[a];
        tmin[a] = (o.lo[a] - x1[a]) / A; tmax[a] = (o.hi[a] - x1[a]) / A;
        if (tmin[a] > tmax[a]) { double t = tmin[a]; tmin[a] = tmax[a]; tmax[a] = t; side[a] += 3; }
    }
    double te = tmin[0]; int j = 0, k = 0, t = 0;
    for (int i = 0; i < 2; i++) {
        if (te < tmin[i + 1]) { te = tmin[i + 1]; j = i + 1; k = 0; }
        else if (fabs(te - tmin[i + 1]) < 1e-6) { k++; t = i + 1; }
    }
    n[0] = n[1] = n[2] = 0;
#define FIND_N(s) do { int s_ = (s); n[s_ % 3] += (s_ < 3) ? -1.0 : 1.0; } while (0)
    FIND_N(side[j]);
    if (k != 0) {
        if (k < 2) FIND_N(side[t]);
        else { FIND_N(side[2]); FIND_N(side[1]); }
        double inv = 1.0 / sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);     // Vec3::normalise -> operator/=(scalar)
        n[0] *= inv; n[1] *= inv; n[2] *= inv;
    }
#undef FIND_N
    *t_entry = te;
    for (int a = 0; a < 3; a++) pos[a] = x1[a] + te * (x2[a] - x1[a]);
}
// Sphere::lineIntersect (Object.cpp:116-140)
PHILOX_HD void sphere_line_intersect(const ObjShape& o, const double x1[3], const double x2[3], double* t_entry, double pos[3], double n[3]) {
    double B[3], A[3];
    for (int a = 0; a < 3; a++) { B[a] = x2[a] - x1[a]; A[a] = x1[a] - o.c[a]; }
    double aa = B[0] * B[0] + B[1] * B[1] + B[2] * B[2];
    double bb = 2 * (A[0] * B[0] + A[1] * B[1] + A[2] * B[2]);
    double cc = A[0] * A[0] + A[1] * A[1] + A[2] * A[2] - o.h[0];
    double det = bb * bb - 4 * aa * cc, te;
    if (det < 0) te = 0.5;
    else {
        te = (-bb + sqrt(det)) / (2 * aa);
        if (te < 0 || te > 1.0) { te = (-bb - sqrt(det)) / (2 * aa); if (te < 0 || te > 1.0) te = 0.5; }
    }
    double r[3];
    for (int a = 0; a < 3; a++) { pos[a] = x1[a] + te * B[a]; r[a] = pos[a] - o.c[a]; }
    double inv = 1.0 / sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);           // Vec3::unit -> operator/(scalar)
    for (int a = 0; a < 3; a++) n[a] = r[a] * inv;
    *t_entry = te;
}
