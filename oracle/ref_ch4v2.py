"""TEST INFRASTRUCTURE: ctypes wrapper of oracle/_ref/libref_ch4v2.so -- the UNMODIFIED reference ch4/v2 sources
(BASELINE config 3: the fixed-weight MC_MEX_Ionization, ch4/v2/Interactions.cpp:476-735) compiled by oracle/Makefile
together with oracle/ref_harness_ch4v2.cpp.

Only tests/ may import this.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_ch4v2.so")

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
        for f in ("refv2_world_create", "refv2_species_create", "refv2_mcc_create"):
            getattr(l, f).restype = C.c_void_p
        for f in ("refv2_rnd", "refv2_mcc_sigma_coll", "refv2_mcc_sigma_ion", "refv2_mcc_get_sv_max"):
            getattr(l, f).restype = C.c_double
        l.refv2_species_count.restype = C.c_size_t
        _lib = l
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def _h(x):
    return C.c_void_p(x)


def seed(s):
    lib().refv2_seed(C.c_uint(int(s)))


def rnd():
    return lib().refv2_rnd()


class World:
    def __init__(self, ni, nj, nk, x0, xm):
        self.ni, self.nj, self.nk = ni, nj, nk
        self.shape = (ni, nj, nk)
        self.h = lib().refv2_world_create(ni, nj, nk, _d3(x0), _d3(xm))

    def close(self):
        if self.h:
            lib().refv2_world_destroy(_h(self.h))
            self.h = None

    def setTime(self, dt, num_ts):
        lib().refv2_world_set_time(_h(self.h), C.c_double(dt), int(num_ts))

    def addRectangle(self, c, phi, sides):
        lib().refv2_world_add_rectangle(_h(self.h), _d3(c), C.c_double(phi), _d3(sides))

    def computeObjectID(self):
        lib().refv2_world_compute_object_id(_h(self.h))

    def setEF(self, ef):
        a = np.ascontiguousarray(ef, dtype=np.float64)
        assert a.size == 3 * self.ni * self.nj * self.nk
        lib().refv2_world_set_ef(_h(self.h), _dp(a))


class Species:
    def __init__(self, name, mass, charge, world, mpw0, E_ion=-666.0):
        self.world, self.mass, self.charge, self.mpw0, self.E_ion = world, mass, charge, mpw0, E_ion
        self.h = lib().refv2_species_create(_h(world.h), name.encode(), C.c_double(mass), C.c_double(charge), C.c_double(mpw0), C.c_double(E_ion))

    def close(self):
        if self.h:
            lib().refv2_species_destroy(_h(self.h))
            self.h = None

    def getNumParticles(self):
        return lib().refv2_species_count(_h(self.h))

    def setParticles(self, aos7):
        a = np.ascontiguousarray(aos7, dtype=np.float64).reshape(-1, 7)
        lib().refv2_species_set_particles(_h(self.h), C.c_size_t(a.shape[0]), _dp(a))

    def getParticles(self):
        out = np.empty((self.getNumParticles(), 7))
        lib().refv2_species_get_particles(_h(self.h), _dp(out))
        return out

    def addParticle(self, a7):
        a = np.ascontiguousarray(a7, dtype=np.float64)
        lib().refv2_species_add_particle(_h(self.h), _dp(a))

    def computeMacroParticlesCount(self):
        lib().refv2_species_compute_macro_count(_h(self.h))


class MC_MEX_Ionization:
    """MC_MEX_Ionization of ch4/v2 (fixed weights).  apply() reads Species::macro_part_count through map_indexes
    (ch4/v2/Species.cpp:740-776), so the per-cell counts of both collision partners are refreshed first, as the v2 main loop does."""

    def __init__(self, neutrals, ions, electrons, world, table_path):
        self.neutrals, self.electrons = neutrals, electrons
        self.h = lib().refv2_mcc_create(_h(neutrals.h), _h(ions.h), _h(electrons.h), _h(world.h), table_path.encode())
        if not self.h:
            raise ValueError("reference MC_MEX_Ionization (ch4/v2) constructor threw")

    def close(self):
        if self.h:
            lib().refv2_mcc_destroy(_h(self.h))
            self.h = None

    def apply(self, dt):
        self.neutrals.computeMacroParticlesCount(); self.electrons.computeMacroParticlesCount()
        lib().refv2_mcc_apply(_h(self.h), C.c_double(dt))

    def sigmaColl(self, E):
        return lib().refv2_mcc_sigma_coll(_h(self.h), C.c_double(E))

    def sigmaIon(self, E):
        return lib().refv2_mcc_sigma_ion(_h(self.h), C.c_double(E))

    def getWsvMax(self):
        return lib().refv2_mcc_get_sv_max(_h(self.h))

    def setWsvMax(self, v):
        lib().refv2_mcc_set_sv_max(_h(self.h), C.c_double(v))

    def collide(self, vn, ve, sigma_coll, n_atoms=0):
        a = (C.c_double * 3)(*vn)
        b = (C.c_double * 3)(*ve)
        c = (C.c_double * 3)()
        ion = lib().refv2_mcc_collide(_h(self.h), a, b, c, int(n_atoms), C.c_double(sigma_coll))
        return bool(ion), np.array(list(a)), np.array(list(b)), np.array(list(c))
