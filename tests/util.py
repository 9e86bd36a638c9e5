"""Shared builders for the parity tests: the reference's geometries and seeded particle sets."""
import numpy as np

# all.h:13-23
EPS0 = 8.85418782e-12
QE = 1.602176565e-19
AMU = 1.660538921e-27
ME = 9.10938215e-31
KB = 1.380648e-23
PI = 3.141592653
NA = 6.02214076e23


def discharge_geometry(ni=41, nj=41, nk=61, phi=-4000.0):
    """ch4/v3/src/main.cpp:89-98: box + two electrode Rectangles, each 0.1*Lz thick, centred on the z faces."""
    x0 = np.array([-0.004, -0.004, 0.0])
    xm = np.array([0.004, 0.004, 0.005])
    dx = (xm - x0) / (np.array([ni, nj, nk]) - 1)
    L = dx * (np.array([ni, nj, nk]) - 1)            # World::getL World.cpp:94-100
    xc = 0.5 * (xm + x0)
    rects = [((xc[0], xc[1], x0[2]), phi, (L[0], L[1], L[2] * 0.1)),
             ((xc[0], xc[1], xm[2]), -phi, (L[0], L[1], L[2] * 0.1))]
    return x0, xm, rects


def build_world(cls, ni, nj, nk, x0, xm, rects=(), spheres=(), dt=1e-12, num_ts=100):
    """cls is any of the World-like classes (product binding, reference wrapper)."""
    w = cls(ni, nj, nk, x0, xm)
    w.setTime(dt, num_ts)
    for c, phi, sides in rects:
        w.addRectangle(c, phi, sides)
    for c, phi, r in spheres:
        w.addSphere(c, phi, r)
    if rects or spheres:
        w.computeObjectID()
    return w


def build_grid(orc, ni, nj, nk, x0, xm, rects=(), spheres=()):
    g = orc.Grid(ni, nj, nk, x0, xm)
    for c, phi, sides in rects:
        g.add_rectangle(c, phi, sides)
    for c, phi, r in spheres:
        g.add_sphere(c, phi, r)
    return g


def random_particles(n, x0, xm, seed, vth=1e5, mpw=(1.0, 100.0), lo_frac=0.0, hi_frac=1.0):
    """n particles uniform in the box [x0 + lo_frac*L, x0 + hi_frac*L) (per axis), gaussian velocities, log-uniform weights."""
    rng = np.random.default_rng(seed)
    L = np.asarray(xm) - np.asarray(x0)
    lo = np.asarray(x0) + np.asarray(lo_frac) * L
    hi = np.asarray(x0) + np.asarray(hi_frac) * L
    a = np.empty((n, 7))
    a[:, 0:3] = lo + rng.random((n, 3)) * (hi - lo)
    a[:, 0:3] = np.minimum(a[:, 0:3], np.nextafter(np.asarray(xm), -np.inf))
    a[:, 3:6] = rng.normal(0.0, vth, (n, 3))
    a[:, 6] = np.exp(rng.uniform(np.log(mpw[0]), np.log(mpw[1]), n))
    return a


def smooth_ef(shape, x0, xm, seed=0, amp=1e5):
    """A smooth analytic E field on the nodes (so gathers exercise every weight)."""
    ni, nj, nk = shape
    x = np.linspace(0, 1, ni)[:, None, None]
    y = np.linspace(0, 1, nj)[None, :, None]
    z = np.linspace(0, 1, nk)[None, None, :]
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 6.28, 9)
    ef = np.empty((ni, nj, nk, 3))
    ef[..., 0] = amp * (np.sin(3 * x + ph[0]) * np.cos(2 * y + ph[1]) + 0.3 * np.sin(5 * z + ph[2])) * np.ones((ni, nj, nk))
    ef[..., 1] = amp * (np.cos(2 * x + ph[3]) * np.sin(4 * z + ph[4]) + 0.2 * np.cos(3 * y + ph[5])) * np.ones((ni, nj, nk))
    ef[..., 2] = amp * (np.sin(2 * y + ph[6]) * np.sin(3 * z + ph[7]) + 0.5 * np.cos(x + ph[8])) * np.ones((ni, nj, nk))
    return ef


def sort_rows(a):
    """Canonical order for multiset comparison of particle arrays."""
    a = np.asarray(a)
    if a.shape[0] == 0:
        return a
    idx = np.lexsort(tuple(a[:, c] for c in range(a.shape[1] - 1, -1, -1)))
    return a[idx]


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    return np.max(np.abs(a - b) / den) if a.size else 0.0


def norm_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    s = np.max(np.abs(b))
    return np.max(np.abs(a - b)) / s if s > 0 else np.max(np.abs(a - b))


def momentum_transfer_table():
    """Harness-supplied O elastic momentum-transfer cross-section (eV, m^2).  The reference's own
    data/Oxygen_momentum_transfer.txt is absent from its tree (.gitignore:17-18); this table follows the
    shape of ch4/v3/data/sigmas.png.  The same table feeds both sides."""
    E = np.array([1e-3, 1e-2, 0.1, 0.3, 1.0, 2.0, 4.0, 8.0, 15.0, 30.0, 60.0, 100.0, 300.0, 1e3, 3e3, 1e4, 1e5, 1e6])
    s = np.array([1.6e-20, 1.8e-20, 2.4e-20, 3.2e-20, 4.6e-20, 5.8e-20, 7.0e-20, 8.0e-20, 7.2e-20, 5.2e-20, 3.4e-20, 2.5e-20,
                  1.1e-20, 4.0e-21, 1.6e-21, 7.0e-22, 1.2e-22, 2.7e-23])
    return E, s


def write_table(path):
    E, s = momentum_transfer_table()
    with open(path, "w") as f:
        for e, v in zip(E, s):
            f.write(f"{e:.17g} {v:.17g}\n")
    return path
