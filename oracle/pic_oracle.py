"""TEST INFRASTRUCTURE: ctypes wrapper of oracle/libpic_oracle.so (the plain-C restatement, pic_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpic_oracle.so")


class OrcObject(C.Structure):
    _fields_ = [("type", C.c_int), ("c", C.c_double * 3), ("h", C.c_double * 3), ("lo", C.c_double * 3), ("hi", C.c_double * 3), ("phi", C.c_double)]


class OrcGrid(C.Structure):
    _fields_ = [("ni", C.c_int), ("nj", C.c_int), ("nk", C.c_int), ("x0", C.c_double * 3), ("xm", C.c_double * 3), ("dx", C.c_double * 3),
                ("inv_dx", C.c_double * 3), ("n_obj", C.c_int), ("obj", OrcObject * 8)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libpic_oracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_residual.restype = C.c_double
        _lib.orc_add_particles.restype = C.c_size_t
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


class Grid:
    """The mesh + object table (World.cpp:63-77, Object.cpp)."""

    def __init__(self, ni, nj, nk, x0, xm):
        self.g = OrcGrid()
        lib().orc_grid_init(C.byref(self.g), ni, nj, nk, _d3(x0), _d3(xm))
        self.ni, self.nj, self.nk = ni, nj, nk
        self.nv = ni * nj * nk
        self.shape = (ni, nj, nk)

    def add_rectangle(self, c, phi, sides):
        lib().orc_add_rectangle(C.byref(self.g), _d3(c), C.c_double(phi), _d3(sides))

    def add_sphere(self, c, phi, r):
        lib().orc_add_sphere(C.byref(self.g), _d3(c), C.c_double(phi), C.c_double(r))

    def node_volumes(self):
        v = np.empty(self.shape)
        lib().orc_node_volumes(C.byref(self.g), _dp(v))
        return v

    def compute_object_id(self, phi=None):
        oid = np.zeros(self.shape, dtype=np.int32)
        if phi is None:
            phi = np.zeros(self.shape)
        phi = np.ascontiguousarray(phi, dtype=np.float64)
        lib().orc_compute_object_id(C.byref(self.g), oid.ctypes.data_as(C.POINTER(C.c_int)), _dp(phi))
        return oid, phi

    def in_object(self, p):
        return lib().orc_in_object(C.byref(self.g), _d3(p))

    def in_bounds(self, p):
        return lib().orc_in_bounds(C.byref(self.g), _d3(p))

    def push_electrons(self, ef, charge, mass, dt, aos7):
        a = np.array(aos7, dtype=np.float64, order="C").reshape(-1, 7)
        alive = np.empty(a.shape[0], dtype=np.uint8)
        ef = np.ascontiguousarray(ef, dtype=np.float64)
        lib().orc_push_electrons(C.byref(self.g), _dp(ef), C.c_double(charge), C.c_double(mass), C.c_double(dt), C.c_size_t(a.shape[0]), _dp(a),
                                 alive.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return a, alive.astype(bool)

    def push_reflect(self, ef, charge, mass, dt, aos7):
        a = np.array(aos7, dtype=np.float64, order="C").reshape(-1, 7)
        ef = np.ascontiguousarray(ef, dtype=np.float64)
        lib().orc_push_reflect(C.byref(self.g), _dp(ef), C.c_double(charge), C.c_double(mass), C.c_double(dt), C.c_size_t(a.shape[0]), _dp(a))
        return a

    def add_particles(self, ef, charge, mass, world_dt, aos7):
        a = np.array(aos7, dtype=np.float64, order="C").reshape(-1, 7)
        ef = np.ascontiguousarray(ef, dtype=np.float64)
        m = lib().orc_add_particles(C.byref(self.g), _dp(ef), C.c_double(charge), C.c_double(mass), C.c_double(world_dt), C.c_size_t(a.shape[0]), _dp(a))
        return a[:m]

    def deposit_fixed(self, aos7, S):
        a = np.ascontiguousarray(aos7, dtype=np.float64).reshape(-1, 7)
        f = np.empty(self.shape, dtype=np.int64)
        lib().orc_deposit_fixed(C.byref(self.g), C.c_size_t(a.shape[0]), _dp(a), int(S), f.ctypes.data_as(C.POINTER(C.c_int64)))
        return f

    def finalize_density(self, fixed, S, vol):
        den = np.empty(self.shape)
        fixed = np.ascontiguousarray(fixed, dtype=np.int64)
        lib().orc_finalize_density(C.byref(self.g), fixed.ctypes.data_as(C.POINTER(C.c_int64)), int(S), _dp(np.ascontiguousarray(vol)), _dp(den))
        return den

    def deposit_fp64(self, aos7, vol):
        a = np.ascontiguousarray(aos7, dtype=np.float64).reshape(-1, 7)
        den = np.empty(self.shape)
        lib().orc_deposit_fp64(C.byref(self.g), C.c_size_t(a.shape[0]), _dp(a), _dp(np.ascontiguousarray(vol)), _dp(den))
        return den

    def count_per_cell(self, aos7):
        a = np.ascontiguousarray(aos7, dtype=np.float64).reshape(-1, 7)
        cnt = np.empty((self.ni - 1, self.nj - 1, self.nk - 1))
        lib().orc_count_per_cell(C.byref(self.g), C.c_size_t(a.shape[0]), _dp(a), _dp(cnt))
        return cnt

    def sample_moments(self, aos7, sums=None):
        a = np.ascontiguousarray(aos7, dtype=np.float64).reshape(-1, 7)
        if sums is None:
            sums = [np.zeros(self.shape), np.zeros(self.shape + (3,)), np.zeros(self.shape), np.zeros(self.shape), np.zeros(self.shape)]
        lib().orc_sample_moments(C.byref(self.g), C.c_size_t(a.shape[0]), _dp(a), *[_dp(s) for s in sums])
        return sums

    def charge_density(self, dens, charges):
        rho = np.empty(self.shape)
        dens = [np.ascontiguousarray(d, dtype=np.float64) for d in dens]
        ptrs = (C.POINTER(C.c_double) * len(dens))(*[_dp(d) for d in dens])
        q = np.ascontiguousarray(charges, dtype=np.float64)
        lib().orc_charge_density(C.byref(self.g), len(dens), ptrs, _dp(q), _dp(rho))
        return rho

    def _solve(self, fn, object_id, rho, phi, max_it, tol, phi0, n0, Te0, bc_mode):
        phi = np.array(phi, dtype=np.float64, order="C")
        oid = np.ascontiguousarray(object_id, dtype=np.int32)
        rho = np.ascontiguousarray(rho, dtype=np.float64)
        it = C.c_uint(0)
        l2 = C.c_double(0)
        conv = fn(C.byref(self.g), oid.ctypes.data_as(C.POINTER(C.c_int)), _dp(rho), _dp(phi), C.c_uint(max_it), C.c_double(tol),
                  C.c_double(phi0), C.c_double(n0), C.c_double(Te0), int(bc_mode), C.byref(it), C.byref(l2))
        return phi, bool(conv), it.value, l2.value

    def solve_gs(self, object_id, rho, phi, max_it, tol, phi0=0.0, n0=0.0, Te0=1.0, bc_mode=0):
        return self._solve(lib().orc_solve_gs, object_id, rho, phi, max_it, tol, phi0, n0, Te0, bc_mode)

    def solve_rb(self, object_id, rho, phi, max_it, tol, phi0=0.0, n0=0.0, Te0=1.0, bc_mode=0):
        return self._solve(lib().orc_solve_rb, object_id, rho, phi, max_it, tol, phi0, n0, Te0, bc_mode)

    def residual(self, object_id, rho, phi, phi0=0.0, n0=0.0, Te0=1.0, bc_mode=0):
        oid = np.ascontiguousarray(object_id, dtype=np.int32)
        return lib().orc_residual(C.byref(self.g), oid.ctypes.data_as(C.POINTER(C.c_int)), _dp(np.ascontiguousarray(rho)), _dp(np.ascontiguousarray(phi)),
                                  C.c_double(phi0), C.c_double(n0), C.c_double(Te0), int(bc_mode))

    def compute_ef(self, phi):
        ef = np.empty(self.shape + (3,))
        lib().orc_compute_ef(C.byref(self.g), _dp(np.ascontiguousarray(phi, dtype=np.float64)), _dp(ef))
        return ef


def philox4x32(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32(c, k, o)
    return list(o)


def dsmc_sigma(m1, m2, v_rel):
    """DSMC_MEX::evaluateSigma (v3/Interactions.cpp:178-181)."""
    lib().orc_dsmc_sigma.restype = C.c_double
    return lib().orc_dsmc_sigma(C.c_double(m1), C.c_double(m2), C.c_double(v_rel))


def dsmc_collide(m1, m2, r1, r2, v1, v2):
    """DSMC_MEX::collide (v3/Interactions.cpp:267-285) with the two uniform draws given."""
    a = (C.c_double * 3)(*v1)
    b = (C.c_double * 3)(*v2)
    lib().orc_dsmc_collide(C.c_double(m1), C.c_double(m2), C.c_double(r1), C.c_double(r2), a, b)
    return np.array(list(a)), np.array(list(b))
