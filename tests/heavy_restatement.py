"""TEST INFRASTRUCTURE: plain-Python restatement of the neutral branch of Species::advanceNoSputteringSerial
(ch4/v3/src/Species.cpp:170-256) with Rectangle::lineIntersect (Object.cpp:239-317) and the uniform draws passed in.
Pinned bit for bit against the compiled reference on CPU (tests/test_oracle_vs_reference.py); used to replay the device's
Philox streams in tests/test_gpu_stochastic.py."""
import math

import sampler_restatement as S


def rect_line_intersect(lo, hi, x1, x2):
    """Rectangle::lineIntersect (Object.cpp:239-290) + find_n (:292-317); x2 inside the box.  Returns (t_entry, pos, n)."""
    tmin, tmax, side = [0.0] * 3, [0.0] * 3, [0, 1, 2]
    for a in range(3):
        A = x2[a] - x1[a]
        tmin[a] = (lo[a] - x1[a]) / A if A != 0 else math.copysign(math.inf, lo[a] - x1[a]) if lo[a] != x1[a] else math.nan
        tmax[a] = (hi[a] - x1[a]) / A if A != 0 else math.copysign(math.inf, hi[a] - x1[a]) if hi[a] != x1[a] else math.nan
        if tmin[a] > tmax[a]:
            tmin[a], tmax[a] = tmax[a], tmin[a]
            side[a] += 3
    te, j, k, t = tmin[0], 0, 0, 0
    for i in range(2):
        if te < tmin[i + 1]:
            te, j, k = tmin[i + 1], i + 1, 0
        elif abs(te - tmin[i + 1]) < 1e-6:
            k += 1
            t = i + 1
    n = [0.0, 0.0, 0.0]

    def find_n(s):
        n[s % 3] += -1.0 if s < 3 else 1.0
    find_n(side[j])
    if k != 0:
        if k < 2:
            find_n(side[t])
        else:
            find_n(side[2]); find_n(side[1])
        inv = 1.0 / math.sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2])      # Vec3::normalise -> operator/=(scalar) multiplies by the reciprocal
        n = [n[0] * inv, n[1] * inv, n[2] * inv]
    return te, [x1[a] + te * (x2[a] - x1[a]) for a in range(3)], n


def advance_neutral(r, grid, rects, mass, pos, vel, dt):
    """One neutral through Species.cpp:179-249 (charge == 0: the kick adds E*0).  grid: oracle Grid (inBounds / inObject);
    rects: [(lo, hi)] in object order.  Returns (pos, vel) or None when the particle is removed."""
    x, v = [float(c) for c in pos], [float(c) for c in vel]
    t_rem, n_bounces = 1.0, 0
    while t_rem > 0:
        n_bounces += 1
        if n_bounces > 20:
            return None
        old = list(x)
        x = [x[a] + v[a] * t_rem * dt for a in range(3)]
        obj = grid.in_object(x)
        if not grid.in_bounds(x):
            return None
        if obj:
            lo, hi = rects[obj - 1]
            tp, hit, n = rect_line_intersect(lo, hi, old, x)
            x = hit
            v_mag = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
            v = S.sample_reflected(r, v_mag, n, mass)
            t_rem *= (1 - tp)
            continue
        t_rem = 0
    return x, v


def advance_ion(r, grid, rects, mass, mpw, neutral_mpw0, pos, vel, dt):
    """One ion through Species.cpp:179-249 in a zero field (the kick adds 0): it flies free, leaves the box, or is neutralised on
    a surface - int(mpw / neutrals.mpw0 + rnd()) neutrals are re-emitted from the hit point through neutrals.addParticle
    (:225-232; rejected when the rounded hit point tests as inside the object or out of bounds).
    Returns (particle or None, [emitted neutral rows])."""
    x, v = [float(c) for c in pos], [float(c) for c in vel]
    old = list(x)
    x = [x[a] + v[a] * 1.0 * dt for a in range(3)]
    obj = grid.in_object(x)
    if not grid.in_bounds(x):
        return None, []
    if not obj:
        return (x, v), []
    lo, hi = rects[obj - 1]
    tp, hit, n = rect_line_intersect(lo, hi, old, x)
    v_mag = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    emitted = []
    for _ in range(int(mpw / neutral_mpw0 + next(r))):
        nv = S.sample_reflected(r, v_mag, n, mass)
        if grid.in_bounds(hit) and not grid.in_object(hit):
            emitted.append(list(hit) + nv + [neutral_mpw0])
    return None, emitted
