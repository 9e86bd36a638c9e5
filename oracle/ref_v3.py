"""TEST INFRASTRUCTURE: ctypes wrapper of oracle/_ref/libref_v3.so -- the UNMODIFIED reference ch4/v3
sources compiled by oracle/Makefile together with oracle/ref_harness_v3.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` arm may import this.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_v3.so")

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
        for f in ("refv3_world_create", "refv3_species_create", "refv3_solver_create", "refv3_mcc_create", "refv3_dsmc_create", "refv3_source_cold", "refv3_source_warm"):
            getattr(l, f).restype = C.c_void_p
        for f in ("refv3_rnd", "refv3_world_get_pe", "refv3_species_ke", "refv3_species_micro_count", "refv3_mcc_sigma_coll", "refv3_mcc_sigma_ion",
                  "refv3_mcc_get_wsv_max", "refv3_dsmc_sigma", "refv3_dsmc_get_sigma_v_max"):
            getattr(l, f).restype = C.c_double
        for f in ("refv3_species_count", "refv3_species_sort_counts"):
            getattr(l, f).restype = C.c_size_t
        _lib = l
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def _h(x):
    return C.c_void_p(x)


def seed(s):
    lib().refv3_seed(C.c_uint(int(s)))


def rnd():
    return lib().refv3_rnd()


def config(subcycling=False, multithreading=False, num_threads=0, merging=False, sputtering=False):
    lib().refv3_config(int(subcycling), int(multithreading), int(num_threads), int(merging), int(sputtering))


def num_threads():
    return lib().refv3_num_threads()


class World:
    def __init__(self, ni, nj, nk, x0, xm):
        self.ni, self.nj, self.nk = ni, nj, nk
        self.shape = (ni, nj, nk)
        self.h = lib().refv3_world_create(ni, nj, nk, _d3(x0), _d3(xm))

    def close(self):
        if self.h:
            lib().refv3_world_destroy(_h(self.h))
            self.h = None

    def setTime(self, dt, num_ts):
        lib().refv3_world_set_time(_h(self.h), C.c_double(dt), int(num_ts))

    def advanceTime(self):
        return bool(lib().refv3_world_advance_time(_h(self.h)))

    def addRectangle(self, c, phi, sides):
        lib().refv3_world_add_rectangle(_h(self.h), _d3(c), C.c_double(phi), _d3(sides))

    def addSphere(self, c, phi, r):
        lib().refv3_world_add_sphere(_h(self.h), _d3(c), C.c_double(phi), C.c_double(r))

    def computeObjectID(self):
        lib().refv3_world_compute_object_id(_h(self.h))

    def inObject(self, p):
        return lib().refv3_world_in_object(_h(self.h), _d3(p))

    def inBounds(self, p):
        return lib().refv3_world_in_bounds(_h(self.h), _d3(p))

    def lineIntersect(self, x1, x2, in_object):
        tp = C.c_double(0)
        pos = (C.c_double * 3)()
        n = (C.c_double * 3)()
        lib().refv3_world_line_intersect(_h(self.h), _d3(x1), _d3(x2), int(in_object), C.byref(tp), pos, n)
        return tp.value, np.array(list(pos)), np.array(list(n))

    def get(self, field):
        out = np.empty(self.shape + ((3,) if field == 3 else ()))
        lib().refv3_world_get_field(_h(self.h), int(field), _dp(out))
        return out

    def set(self, field, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        lib().refv3_world_set_field(_h(self.h), int(field), _dp(a))

    def getPE(self):
        return lib().refv3_world_get_pe(_h(self.h))

    def computeChargeDensity(self, species):
        arr = (C.c_void_p * len(species))(*[s.h for s in species])
        lib().refv3_world_compute_charge_density(_h(self.h), arr, len(species))


class Species:
    def __init__(self, name, mass, charge, world, mpw0, E_ion=-666.0):
        self.world, self.mass, self.charge, self.mpw0, self.E_ion = world, mass, charge, mpw0, E_ion
        self.h = lib().refv3_species_create(_h(world.h), name.encode(), C.c_double(mass), C.c_double(charge), C.c_double(mpw0), C.c_double(E_ion))

    def close(self):
        if self.h:
            lib().refv3_species_destroy(_h(self.h))
            self.h = None

    def getNumParticles(self):
        return lib().refv3_species_count(_h(self.h))

    def setParticles(self, aos7):
        a = np.ascontiguousarray(aos7, dtype=np.float64).reshape(-1, 7)
        lib().refv3_species_set_particles(_h(self.h), C.c_size_t(a.shape[0]), _dp(a))

    def getParticles(self):
        out = np.empty((self.getNumParticles(), 7))
        lib().refv3_species_get_particles(_h(self.h), _dp(out))
        return out

    def addParticle(self, a7):
        a = np.ascontiguousarray(a7, dtype=np.float64)
        lib().refv3_species_add_particle(_h(self.h), _dp(a))

    def loadParticleBoxThermal(self, x0, sides, den, T):
        lib().refv3_species_load_box_thermal(_h(self.h), _d3(x0), _d3(sides), C.c_double(den), C.c_double(T))

    def advanceElectrons(self, dt):
        lib().refv3_species_advance_electrons(_h(self.h), C.c_double(dt))

    def advanceNonElectron(self, neutrals, spherium, dt):
        lib().refv3_species_advance_non_electron(_h(self.h), _h(neutrals.h), _h(spherium.h), C.c_double(dt))

    def computeNumberDensity(self):
        lib().refv3_species_compute_number_density(_h(self.h))

    def sampleMoments(self):
        lib().refv3_species_sample_moments(_h(self.h))

    def computeGasProperties(self):
        lib().refv3_species_compute_gas_properties(_h(self.h))

    def clearSamples(self):
        lib().refv3_species_clear_samples(_h(self.h))

    def updateAverages(self):
        lib().refv3_species_update_averages(_h(self.h))

    def computeMacroParticlesCount(self):
        lib().refv3_species_compute_macro_count(_h(self.h))

    def merge(self):
        lib().refv3_species_merge(_h(self.h))

    def getKE(self):
        return lib().refv3_species_ke(_h(self.h))

    def getMicroCount(self):
        return lib().refv3_species_micro_count(_h(self.h))

    def getMomentum(self):
        m = (C.c_double * 3)()
        lib().refv3_species_momentum(_h(self.h), m)
        return np.array(list(m))

    def sampleV3th(self, T):
        v = (C.c_double * 3)()
        lib().refv3_species_sample_v3th(_h(self.h), C.c_double(T), v)
        return np.array(list(v))

    def sampleReflectedVelocity(self, pos, vmag, n):
        v = (C.c_double * 3)()
        lib().refv3_species_sample_reflected(_h(self.h), _d3(pos), C.c_double(vmag), _d3(n), v)
        return np.array(list(v))

    def sortCounts(self):
        w = self.world
        cnt = np.empty((w.ni - 1) * (w.nj - 1) * (w.nk - 1), dtype=np.int32)
        lib().refv3_species_sort_counts(_h(self.h), _h(w.h), cnt.ctypes.data_as(C.POINTER(C.c_int)))
        return cnt

    def get(self, field):
        w = self.world
        if field in (3, 6):
            out = np.empty(w.shape + (3,))
        elif field == 4:
            out = np.empty((w.ni - 1, w.nj - 1, w.nk - 1))
        else:
            out = np.empty(w.shape)
        lib().refv3_species_get_field(_h(self.h), int(field), _dp(out))
        return out


class PotentialSolver:
    GS, PCG, QN = 0, 1, 2

    def __init__(self, world, max_it, tol, solver_type=0):
        self.h = lib().refv3_solver_create(_h(world.h), C.c_uint(int(max_it)), C.c_double(tol), int(solver_type))

    def close(self):
        if self.h:
            lib().refv3_solver_destroy(_h(self.h))
            self.h = None

    def setReferenceValues(self, phi0, n0, Te0):
        lib().refv3_solver_set_reference(_h(self.h), C.c_double(phi0), C.c_double(n0), C.c_double(Te0))

    def solveGS(self):
        return bool(lib().refv3_solver_solve_gs(_h(self.h)))

    def solve(self):
        """PotentialSolver::solve(): dispatches on the solver type (PCG -> solveNRPCG, v3/PotentialSolver.cpp:55-67)."""
        return bool(lib().refv3_solver_solve(_h(self.h)))

    def computeEF(self):
        lib().refv3_solver_compute_ef(_h(self.h))


class MC_MEX_Ionization:
    def __init__(self, neutrals, ions, electrons, world, table_path):
        self.h = lib().refv3_mcc_create(_h(neutrals.h), _h(ions.h), _h(electrons.h), _h(world.h), table_path.encode())
        if not self.h:
            raise ValueError("reference MC_MEX_Ionization constructor threw")

    def close(self):
        if self.h:
            lib().refv3_mcc_destroy(_h(self.h))
            self.h = None

    def apply(self, dt):
        lib().refv3_mcc_apply(_h(self.h), C.c_double(dt))

    def sigmaColl(self, E):
        return lib().refv3_mcc_sigma_coll(_h(self.h), C.c_double(E))

    def sigmaIon(self, E):
        return lib().refv3_mcc_sigma_ion(_h(self.h), C.c_double(E))

    def getWsvMax(self):
        return lib().refv3_mcc_get_wsv_max(_h(self.h))

    def setWsvMax(self, v):
        lib().refv3_mcc_set_wsv_max(_h(self.h), C.c_double(v))

    def collide(self, vn, ve, sigma_coll):
        a = (C.c_double * 3)(*vn)
        b = (C.c_double * 3)(*ve)
        c = (C.c_double * 3)()
        ion = lib().refv3_mcc_collide(_h(self.h), a, b, c, C.c_double(sigma_coll))
        return bool(ion), np.array(list(a)), np.array(list(b)), np.array(list(c))


class DSMC_MEX:
    """DSMC_MEX(species, world) / DSMC_MEX(species1, species2, world) of the compiled reference (v3/Interactions.cpp:143-285)."""

    def __init__(self, species1, species2_or_world, world=None):
        if world is None:
            species2, world = None, species2_or_world
        else:
            species2 = species2_or_world
        self.h = lib().refv3_dsmc_create(_h(species1.h), _h(species2.h) if species2 is not None else None, _h(world.h))
        if not self.h:
            raise ValueError("reference DSMC_MEX constructor threw")

    def close(self):
        if self.h:
            lib().refv3_dsmc_destroy(_h(self.h))
            self.h = None

    def apply(self, dt):
        lib().refv3_dsmc_apply(_h(self.h), C.c_double(dt))

    def sigma(self, v_rel):
        return np.array([lib().refv3_dsmc_sigma(_h(self.h), C.c_double(float(v))) for v in np.atleast_1d(v_rel)])

    def getSigmaVMax(self):
        return lib().refv3_dsmc_get_sigma_v_max(_h(self.h))

    def setSigmaVMax(self, v):
        lib().refv3_dsmc_set_sigma_v_max(_h(self.h), C.c_double(v))

    def collide(self, v1, v2):
        a = (C.c_double * 3)(*v1)
        b = (C.c_double * 3)(*v2)
        lib().refv3_dsmc_collide(_h(self.h), a, b)
        return np.array(list(a)), np.array(list(b))


class Source:
    def __init__(self, species, world, v_drift, den, face="-z", T=None):
        if T is None:
            self.h = lib().refv3_source_cold(_h(species.h), _h(world.h), C.c_double(v_drift), C.c_double(den), face.encode())
        else:
            self.h = lib().refv3_source_warm(_h(species.h), _h(world.h), C.c_double(v_drift), C.c_double(den), C.c_double(T), face.encode())

    def close(self):
        if self.h:
            lib().refv3_source_destroy(_h(self.h))
            self.h = None

    def sample(self):
        lib().refv3_source_sample(_h(self.h))
