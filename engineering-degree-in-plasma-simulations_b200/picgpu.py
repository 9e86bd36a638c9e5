"""ctypes binding of libpicgpu.so (include/picgpu.h) with the reference's class names.

The classes mirror the public surface of the reference's ch4/v3 classes that the main loop
uses (ch4/v3/src/main.cpp:177-288): World, Species, PotentialSolver, MC_MEX_Ionization,
ColdBeamSource / WarmBeamSource.  Everything forwards to the C ABI; nothing is computed in
Python and there is no CPU fallback: without the CUDA library or a GPU, calls raise.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpicgpu.so")
if os.environ.get("PICGPU_SO"):                 # A/B builds of the same library (profiles/scripts): never a different implementation
    LIB_PATH = os.path.abspath(os.environ["PICGPU_SO"])


class PicgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"picgpu error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Loads libpicgpu.so (fails loudly if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback for the particle loop)")
        l = C.CDLL(LIB_PATH)
        l.picg_last_error.restype = C.c_char_p
        l.picg_version.restype = C.c_char_p
        l.picg_timer_name.restype = C.c_char_p
        l.picg_stream.restype = C.c_void_p
        l.picg_launch_count.restype = C.c_uint64
        l.picg_realloc_count.restype = C.c_uint64
        l.picg_launch_count_reset.restype = None
        _lib = l
    return _lib


def _chk(rc):
    if rc != 0:
        raise PicgError(rc, lib().picg_last_error().decode())


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def init(device=0):
    _chk(lib().picg_init(int(device)))


def device_count():
    n = C.c_int(0)
    lib().picg_device_count(C.byref(n))
    return n.value


def seed(s):
    _chk(lib().picg_seed(C.c_uint64(int(s))))


def set_rank(rank, world_size):
    _chk(lib().picg_set_rank(int(rank), int(world_size)))


def set_mover_fraction(f):
    _chk(lib().picg_set_mover_fraction(C.c_double(f)))


def set_merge_fraction(f):
    _chk(lib().picg_set_merge_fraction(C.c_double(f)))


def tail_merge_count():
    n = C.c_uint64(0)
    _chk(lib().picg_tail_merge_count(C.byref(n)))
    return n.value


def mover_stats():
    """(lists re-used from a deposit pass, full mover scans, fall-backs to a full sort) since start."""
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    _chk(lib().picg_mover_stats(C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def synchronize():
    _chk(lib().picg_synchronize())


def stream_ptr():
    return lib().picg_stream()


def launch_count():
    return int(lib().picg_launch_count())


def realloc_count():
    return int(lib().picg_realloc_count())


def launch_count_reset():
    lib().picg_launch_count_reset()


def timers_enable(on=True):
    _chk(lib().picg_timers_enable(int(bool(on))))


def timers_reset():
    _chk(lib().picg_timers_reset())


def timers_read():
    """{kernel name: (total ms, launches)} for kernels launched since the last reset."""
    out = {}
    i = 0
    while True:
        name = lib().picg_timer_name(i)
        if name is None:
            break
        ms = C.c_double(0)
        n = C.c_uint64(0)
        _chk(lib().picg_timer_read(i, C.byref(ms), C.byref(n)))
        if n.value:
            out[name.decode()] = (ms.value, n.value)
        i += 1
    return out


# field ids (include/picgpu.h)
F_PHI, F_RHO, F_NODE_VOL, F_EF, F_OBJECT_ID, F_NODE_TYPE = range(6)
(SF_DEN, SF_DEN_AVG, SF_T, SF_VEL, SF_MACRO_COUNT, SF_N_SUM, SF_NV_SUM, SF_NUU_SUM, SF_NVV_SUM, SF_NWW_SUM,
 SF_DEN_FIXED) = range(11)


class World:
    """World (ch4/v3/src/World.h:14-116)."""

    def __init__(self, ni, nj, nk, x0, xm):
        self.ni, self.nj, self.nk = int(ni), int(nj), int(nk)
        self.nv = self.ni * self.nj * self.nk
        self.num_cells = (self.ni - 1) * (self.nj - 1) * (self.nk - 1)
        self.x0 = np.asarray(x0, dtype=np.float64)
        self.xm = np.asarray(xm, dtype=np.float64)
        self.dx = (self.xm - self.x0) / (np.array([ni, nj, nk]) - 1)
        self.h = C.c_void_p()
        _chk(lib().picg_world_create(self.ni, self.nj, self.nk, _d3(x0), _d3(xm), C.byref(self.h)))
        self.dt = 1e-4

    def close(self):
        if self.h:
            lib().picg_world_destroy(self.h)
            self.h = C.c_void_p()

    def setTime(self, dt, num_ts):
        self.dt = float(dt)
        _chk(lib().picg_world_set_time(self.h, C.c_double(dt), int(num_ts)))

    def addRectangle(self, centre, phi, sides):
        _chk(lib().picg_world_add_rectangle(self.h, _d3(centre), C.c_double(phi), _d3(sides)))

    def addSphere(self, centre, phi, radius):
        _chk(lib().picg_world_add_sphere(self.h, _d3(centre), C.c_double(phi), C.c_double(radius)))

    def computeObjectID(self):
        _chk(lib().picg_world_compute_object_id(self.h))

    def _shape(self, field):
        return (self.ni, self.nj, self.nk, 3) if field == F_EF else (self.ni, self.nj, self.nk)

    def download(self, field):
        out = np.empty(self._shape(field), dtype=np.float64)
        _chk(lib().picg_world_download(self.h, int(field), _dp(out)))
        return out

    def upload(self, field, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        assert a.size == int(np.prod(self._shape(field)))
        _chk(lib().picg_world_upload(self.h, int(field), _dp(a)))

    phi = property(lambda self: self.download(F_PHI))
    rho = property(lambda self: self.download(F_RHO))
    ef = property(lambda self: self.download(F_EF))
    node_vol = property(lambda self: self.download(F_NODE_VOL))
    object_id = property(lambda self: self.download(F_OBJECT_ID))

    def computeChargeDensity(self, species, node_range=None):
        arr = (C.c_void_p * len(species))(*[s.h for s in species])
        if node_range is None:
            _chk(lib().picg_world_charge_density(self.h, arr, len(species)))
        else:                                               # multi-GPU: only the planes this rank solves on
            _chk(lib().picg_world_charge_density_range(self.h, arr, len(species), C.c_size_t(node_range[0]), C.c_size_t(node_range[1])))

    def getPE(self):
        pe = C.c_double(0)
        _chk(lib().picg_world_potential_energy(self.h, C.byref(pe)))
        return pe.value

    def device_ptr(self, field):
        p = C.c_void_p()
        n = C.c_size_t(0)
        _chk(lib().picg_world_device_ptr(self.h, int(field), C.byref(p), C.byref(n)))
        return p.value, n.value


class Species:
    """Species (ch4/v3/src/Species.h:31-131) on a device SoA store."""

    def __init__(self, name, mass, charge, world, mpw0, E_ion=-666.0):
        self.name, self.mass, self.charge, self.world, self.mpw0, self.E_ion = name, float(mass), float(charge), world, float(mpw0), float(E_ion)
        self.h = C.c_void_p()
        _chk(lib().picg_species_create(world.h, C.c_double(mass), C.c_double(charge), C.c_double(mpw0), C.byref(self.h)))

    def close(self):
        if self.h:
            lib().picg_species_destroy(self.h)
            self.h = C.c_void_p()

    def reserve(self, n):
        _chk(lib().picg_species_reserve(self.h, C.c_size_t(int(n))))

    def partitionSize(self):
        n = C.c_size_t(0)
        _chk(lib().picg_species_partition_size(self.h, C.byref(n)))
        return n.value

    def getNumParticles(self):
        n = C.c_size_t(0)
        _chk(lib().picg_species_count(self.h, C.byref(n)))
        return n.value

    def setParticles(self, aos7):
        a = np.ascontiguousarray(aos7, dtype=np.float64).reshape(-1, 7)
        _chk(lib().picg_species_upload(self.h, C.c_size_t(a.shape[0]), _dp(a)))

    def getParticles(self):
        n = self.getNumParticles()
        out = np.empty((n, 7), dtype=np.float64)
        m = C.c_size_t(0)
        _chk(lib().picg_species_download(self.h, C.c_size_t(n), _dp(out), C.byref(m)))
        return out[:m.value]

    def particleArrays(self, capacity):
        """Device pointers of the seven SoA arrays (x y z u v w mpw) sized for `capacity` particles, and the capacity."""
        arr = (C.c_void_p * 7)()
        cap = C.c_size_t(0)
        _chk(lib().picg_species_particle_arrays(self.h, C.c_size_t(int(capacity)), arr, C.byref(cap)))
        return [arr[c] for c in range(7)], cap.value

    def adopt(self, n):
        """The first n particles written into particleArrays() become the contents of the store."""
        _chk(lib().picg_species_adopt(self.h, C.c_size_t(int(n))))

    def addParticles(self, aos7):
        a = np.ascontiguousarray(aos7, dtype=np.float64).reshape(-1, 7)
        acc = C.c_size_t(0)
        _chk(lib().picg_species_add_particles(self.h, C.c_size_t(a.shape[0]), _dp(a), C.byref(acc)))
        return acc.value

    def loadParticleBoxThermal(self, x0, sides, num_den, T):
        n = C.c_size_t(0)
        _chk(lib().picg_species_load_box_thermal(self.h, _d3(x0), _d3(sides), C.c_double(num_den), C.c_double(T), C.byref(n)))
        return n.value

    def advanceElectrons(self, dt):
        _chk(lib().picg_species_push_electrons(self.h, C.c_double(dt)))

    def advanceNonElectron(self, neutrals, spherium, dt, sputtering=False):
        _chk(lib().picg_species_push_heavy(self.h, neutrals.h, spherium.h, C.c_double(dt), int(sputtering)))

    def advanceReflect(self, dt):
        _chk(lib().picg_species_push_reflect(self.h, C.c_double(dt)))

    def advanceElectronsDeposit(self, dt, count_cells=False):
        _chk(lib().picg_species_push_electrons_deposit(self.h, C.c_double(dt), int(count_cells)))

    def advanceNonElectronDeposit(self, neutrals, spherium, dt, sputtering=False, count_cells=False):
        _chk(lib().picg_species_push_heavy_deposit(self.h, neutrals.h, spherium.h, C.c_double(dt), int(sputtering), int(count_cells)))

    def advanceDepositPartial(self, dt, neutrals=None, spherium=None, heavy=False, sputtering=False, count_cells=False):
        _chk(lib().picg_species_push_deposit_partial(self.h, neutrals.h if neutrals else None, spherium.h if spherium else None, C.c_double(dt),
                                                     int(heavy), int(sputtering), int(count_cells)))

    def computeNumberDensity(self):
        _chk(lib().picg_species_deposit_density(self.h))

    def merge(self):
        """Species::merge (Species.cpp:1037-1145).  Returns (n_before, n_after, (merged bins, removed, dropped, cells too large))."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        st = (C.c_uint64 * 4)()
        _chk(lib().picg_species_merge(self.h, C.byref(a), C.byref(b), st))
        return a.value, b.value, tuple(int(v) for v in st)

    def depositPartial(self):
        _chk(lib().picg_species_deposit_density_partial(self.h))

    def finalizeDensity(self, node_range=None):
        if node_range is None:
            _chk(lib().picg_species_finalize_density(self.h))
        else:
            _chk(lib().picg_species_finalize_density_range(self.h, C.c_size_t(node_range[0]), C.c_size_t(node_range[1])))

    def densityScale(self):
        s = C.c_int(0)
        _chk(lib().picg_species_density_scale(self.h, C.byref(s)))
        return s.value

    def setDensityScale(self, S):
        _chk(lib().picg_species_set_density_scale(self.h, int(S)))

    def sampleMoments(self):
        _chk(lib().picg_species_sample_moments(self.h))

    def computeGasProperties(self):
        _chk(lib().picg_species_compute_gas_properties(self.h))

    def clearSamples(self):
        _chk(lib().picg_species_clear_samples(self.h))

    def updateAverages(self):
        _chk(lib().picg_species_update_averages(self.h))

    def computeMacroParticlesCount(self):
        _chk(lib().picg_species_count_per_cell(self.h))

    def sort(self):
        _chk(lib().picg_species_sort(self.h))

    def diagnostics(self):
        mc = C.c_double(0)
        ke = C.c_double(0)
        mom = (C.c_double * 3)()
        _chk(lib().picg_species_diagnostics(self.h, C.byref(mc), mom, C.byref(ke)))
        return mc.value, np.array(list(mom)), ke.value

    def download(self, field):
        w = self.world
        if field in (SF_VEL, SF_NV_SUM):
            out = np.empty((w.ni, w.nj, w.nk, 3), dtype=np.float64)
        elif field == SF_MACRO_COUNT:
            out = np.empty((w.ni - 1, w.nj - 1, w.nk - 1), dtype=np.float64)
        elif field == SF_DEN_FIXED:
            out = np.empty((w.ni, w.nj, w.nk), dtype=np.int64)
        else:
            out = np.empty((w.ni, w.nj, w.nk), dtype=np.float64)
        _chk(lib().picg_species_download_field(self.h, int(field), out.ctypes.data_as(C.c_void_p)))
        return out

    den = property(lambda self: self.download(SF_DEN))
    den_fixed = property(lambda self: self.download(SF_DEN_FIXED))
    macro_part_count = property(lambda self: self.download(SF_MACRO_COUNT))

    def device_ptr(self, field):
        p = C.c_void_p()
        n = C.c_size_t(0)
        _chk(lib().picg_species_device_ptr(self.h, int(field), C.byref(p), C.byref(n)))
        return p.value, n.value


class Species32:
    """A species in the fp32 secondary store (csrc/f32.cu): cell index + cell-relative coordinates, 32 bytes per particle.  Same method
    names as Species for the part of the loop it covers; host arrays are AoS of doubles like Species."""

    def __init__(self, name, mass, charge, world, mpw0):
        self.name, self.mass, self.charge, self.mpw0, self.world = name, float(mass), float(charge), float(mpw0), world
        self.h = C.c_void_p()
        _chk(lib().picg_species32_create(world.h, C.c_double(mass), C.c_double(charge), C.c_double(mpw0), C.byref(self.h)))

    def close(self):
        if self.h:
            lib().picg_species32_destroy(self.h)
            self.h = C.c_void_p()

    def reserve(self, n):
        _chk(lib().picg_species32_reserve(self.h, C.c_size_t(int(n))))

    def getNumParticles(self):
        n = C.c_size_t(0)
        _chk(lib().picg_species32_count(self.h, C.byref(n)))
        return n.value

    def setParticles(self, aos7):
        a = np.ascontiguousarray(aos7, dtype=np.float64).reshape(-1, 7)
        _chk(lib().picg_species32_upload(self.h, C.c_size_t(a.shape[0]), _dp(a)))

    def fromSpecies(self, species):
        _chk(lib().picg_species32_from_species(self.h, species.h))

    def getParticles(self):
        n = self.getNumParticles()
        out = np.empty((n, 7), dtype=np.float64)
        got = C.c_size_t(0)
        _chk(lib().picg_species32_download(self.h, C.c_size_t(n), _dp(out), C.byref(got)))
        return out[:got.value]

    def advanceElectrons(self, dt):
        _chk(lib().picg_species32_push_electrons(self.h, C.c_double(dt)))

    def computeNumberDensity(self):
        _chk(lib().picg_species32_deposit_density(self.h))

    def computeMacroParticlesCount(self):
        pass                                              # a by-product of the deposit pass

    def setDensityScale(self, S):
        _chk(lib().picg_species32_set_density_scale(self.h, int(S)))

    def densityScale(self):
        S = C.c_int(0)
        _chk(lib().picg_species32_density_scale(self.h, C.byref(S)))
        return S.value

    def sort(self):
        _chk(lib().picg_species32_sort(self.h))

    def diagnostics(self):
        mc = C.c_double(0)
        ke = C.c_double(0)
        mom = (C.c_double * 3)()
        _chk(lib().picg_species32_diagnostics(self.h, C.byref(mc), mom, C.byref(ke)))
        return mc.value, np.array(list(mom)), ke.value

    def download(self, field):
        w = self.world
        if field == SF_MACRO_COUNT:
            out = np.empty((w.ni - 1, w.nj - 1, w.nk - 1), dtype=np.float64)
        elif field == SF_DEN_FIXED:
            out = np.empty((w.ni, w.nj, w.nk), dtype=np.int64)
        else:
            out = np.empty((w.ni, w.nj, w.nk), dtype=np.float64)
        _chk(lib().picg_species32_download_field(self.h, int(field), out.ctypes.data_as(C.c_void_p)))
        return out

    den = property(lambda self: self.download(SF_DEN))
    den_fixed = property(lambda self: self.download(SF_DEN_FIXED))
    macro_part_count = property(lambda self: self.download(SF_MACRO_COUNT))


def charge_density32(world, species32):
    """World::computeChargeDensity over fp32 species."""
    arr = (C.c_void_p * len(species32))(*[s.h for s in species32])
    _chk(lib().picg_world_charge_density32(world.h, arr, len(species32)))


class PotentialSolver:
    """PotentialSolver with SolverType GS (ch4/v3/src/PotentialSolver.h:26-92)."""

    def __init__(self, world, max_solver_it, tolerance):
        self.world = world
        self.h = C.c_void_p()
        _chk(lib().picg_solver_create(world.h, C.c_uint(int(max_solver_it)), C.c_double(tolerance), C.byref(self.h)))
        self.iterations = 0
        self.L2 = 0.0

    def close(self):
        if self.h:
            lib().picg_solver_destroy(self.h)
            self.h = C.c_void_p()

    def setReferenceValues(self, phi0, n0, Te0):
        _chk(lib().picg_solver_set_reference(self.h, C.c_double(phi0), C.c_double(n0), C.c_double(Te0)))

    def setSweep(self, mode):
        """0: row sweeps (two per iteration, default); 1: tiled one-pass sweep.  Same bits."""
        _chk(lib().picg_solver_set_sweep(self.h, int(mode)))

    def setBoundaryMode(self, mode):
        _chk(lib().picg_solver_set_boundary_mode(self.h, int(mode)))

    def solveGS(self):
        conv = C.c_int(0)
        it = C.c_uint(0)
        l2 = C.c_double(0)
        _chk(lib().picg_solver_solve_gs(self.h, C.byref(conv), C.byref(it), C.byref(l2)))
        self.iterations, self.L2 = it.value, l2.value
        return bool(conv.value)

    solve = solveGS

    def iterate(self, n):
        _chk(lib().picg_solver_iterate(self.h, C.c_uint(int(n))))

    def solveNRPCG(self, xz_swap=True):
        """PotentialSolver::solveNRPCG (ch4/v3/src/PotentialSolver.cpp:178-240); xz_swap=True is the reference's matrix."""
        conv, nr, pcg, norm = C.c_int(0), C.c_uint(0), C.c_uint(0), C.c_double(0)
        _chk(lib().picg_solver_solve_nrpcg(self.h, int(bool(xz_swap)), C.c_uint(0), C.byref(conv), C.byref(nr), C.byref(pcg), C.byref(norm)))
        self.nr_iterations, self.pcg_iterations, self.norm = nr.value, pcg.value, norm.value
        return bool(conv.value)

    def residual(self):
        l2 = C.c_double(0)
        _chk(lib().picg_solver_residual(self.h, C.byref(l2)))
        return l2.value

    def computeEF(self):
        _chk(lib().picg_solver_compute_ef(self.h))

    def slabRange(self):
        """Node range [begin, end) of the planes this rank solves on (the whole grid when slabs are off)."""
        a, b = C.c_size_t(0), C.c_size_t(0)
        _chk(lib().picg_solver_slab_range(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def enableSlabs(self, rank, world, all_gather_bytes):
        """Multi-GPU slab decomposition of the solve.  all_gather_bytes(b: bytes) -> list of `world` bytes objects (rank order)."""
        mine = C.create_string_buffer(128)
        _chk(lib().picg_solver_slab_export(self.h, mine))
        table = b"".join(all_gather_bytes(bytes(mine.raw)))
        assert len(table) == 128 * world
        _chk(lib().picg_solver_slab_enable(self.h, int(rank), int(world), table))


class MccStats(C.Structure):
    _fields_ = [("candidates", C.c_uint64), ("collisions", C.c_uint64), ("ionizations", C.c_uint64), ("w_sigma_v_max", C.c_double), ("dropped", C.c_uint64),
                ("extras_capped", C.c_uint64), ("nan_products", C.c_uint64)]


class MC_MEX_Ionization:
    """MC_MEX_Ionization (ch4/v3/src/Interactions.h:99-144); the cross-section table is passed as arrays."""

    def __init__(self, neutrals, ions, electrons, world, table_E, table_sigma):
        e = np.ascontiguousarray(table_E, dtype=np.float64)
        s = np.ascontiguousarray(table_sigma, dtype=np.float64)
        self.h = C.c_void_p()
        _chk(lib().picg_mcc_create(neutrals.h, ions.h, electrons.h, world.h, _dp(e), _dp(s), int(e.size), C.c_double(neutrals.E_ion), C.byref(self.h)))
        self.stats = MccStats()

    def close(self):
        if self.h:
            lib().picg_mcc_destroy(self.h)
            self.h = C.c_void_p()

    def apply(self, dt):
        _chk(lib().picg_mcc_apply(self.h, C.c_double(dt), C.byref(self.stats)))
        return self.stats

    def setWsvMax(self, v):
        _chk(lib().picg_mcc_set_wsv_max(self.h, C.c_double(v)))

    def setVariant(self, variant):
        """0: variable weights (ch4/v3, default); 1: the fixed-weight algorithm of ch4/v2 (ch4/v2/Interactions.cpp:566-641)."""
        _chk(lib().picg_mcc_set_variant(self.h, int(variant)))

    def listCounts(self, which, world):
        out = np.empty((world.ni - 1, world.nj - 1, world.nk - 1), dtype=np.float64)
        _chk(lib().picg_mcc_list_counts(self.h, int(which), _dp(out)))
        return out

    def sigma(self, E_eV):
        e = np.ascontiguousarray(E_eV, dtype=np.float64)
        sc = np.empty_like(e)
        si = np.empty_like(e)
        _chk(lib().picg_mcc_sigma(self.h, int(e.size), _dp(e), _dp(sc), _dp(si)))
        return sc, si


class _CheckpointSet(C.Structure):
    _fields_ = [("world", C.c_void_p), ("species", C.POINTER(C.c_void_p)), ("n_species", C.c_int), ("mcc", C.POINTER(C.c_void_p)), ("n_mcc", C.c_int),
                ("dsmc", C.POINTER(C.c_void_p)), ("n_dsmc", C.c_int), ("sources", C.POINTER(C.c_void_p)), ("n_sources", C.c_int)]


def _handle_array(objs):
    arr = (C.c_void_p * max(len(objs), 1))(*[o.h.value if isinstance(o.h, C.c_void_p) else o.h for o in objs])
    return arr


def _checkpoint_set(world, species, mcc, dsmc, sources):
    keep = [_handle_array(list(x)) for x in (species, mcc, dsmc, sources)]
    cs = _CheckpointSet(world.h, C.cast(keep[0], C.POINTER(C.c_void_p)), len(species), C.cast(keep[1], C.POINTER(C.c_void_p)), len(mcc),
                        C.cast(keep[2], C.POINTER(C.c_void_p)), len(dsmc), C.cast(keep[3], C.POINTER(C.c_void_p)), len(sources))
    return cs, keep


def checkpoint_save(path, world, species, mcc=(), dsmc=(), sources=(), ts=0):
    """picg_checkpoint_save: binary restart file of the loop state (fields, particle stores, averages, RNG stream positions)."""
    cs, keep = _checkpoint_set(world, species, mcc, dsmc, sources)
    _chk(lib().picg_checkpoint_save(os.fsencode(path), C.byref(cs), C.c_uint64(int(ts))))


def checkpoint_load(path, world, species, mcc=(), dsmc=(), sources=()):
    """picg_checkpoint_load into objects rebuilt as at start-up (same order); returns the stored time-step number."""
    cs, keep = _checkpoint_set(world, species, mcc, dsmc, sources)
    ts = C.c_uint64(0)
    _chk(lib().picg_checkpoint_load(os.fsencode(path), C.byref(cs), C.byref(ts)))
    return ts.value


def write_fields_vti(path, world, species):
    """Output::fieldsOutput (ch4/v3/src/Outputs.cpp:9-123) as binary VTK ImageData (appended raw)."""
    hs = _handle_array(list(species))
    names = (C.c_char_p * max(len(species), 1))(*[sp.name.encode() for sp in species])
    _chk(lib().picg_write_fields_vti(os.fsencode(path), world.h, C.cast(hs, C.POINTER(C.c_void_p)), names, len(species)))


class DsmcStats(C.Structure):
    _fields_ = [("candidates", C.c_uint64), ("collisions", C.c_uint64), ("sigma_v_max", C.c_double)]


class DSMC_MEX:
    """DSMC_MEX (ch4/v3/src/Interactions.h:42-66): DSMC_MEX(species, world) or DSMC_MEX(species1, species2, world)."""

    def __init__(self, species1, species2_or_world, world=None):
        if world is None:
            species2, world = None, species2_or_world
        else:
            species2 = species2_or_world
        self.h = C.c_void_p()
        _chk(lib().picg_dsmc_create(species1.h, species2.h if species2 is not None else None, world.h, C.byref(self.h)))
        self.stats = DsmcStats()

    def close(self):
        if self.h:
            lib().picg_dsmc_destroy(self.h)
            self.h = C.c_void_p()

    def apply(self, dt):
        _chk(lib().picg_dsmc_apply(self.h, C.c_double(dt), C.byref(self.stats)))
        return self.stats

    def setSigmaVMax(self, v):
        _chk(lib().picg_dsmc_set_sigma_v_max(self.h, C.c_double(v)))

    def sigma(self, v_rel):
        v = np.ascontiguousarray(v_rel, dtype=np.float64)
        out = np.empty_like(v)
        _chk(lib().picg_dsmc_sigma(self.h, int(v.size), _dp(v), _dp(out)))
        return out


_FACES = {"x-": 0, "-x": 0, "x+": 1, "+x": 1, "y-": 2, "-y": 2, "y+": 3, "+y": 3, "z-": 4, "-z": 4, "z+": 5, "+z": 5}


class _Source:
    def __init__(self, species, world, v_drift, den, T, inlet_face):
        self.h = C.c_void_p()
        _chk(lib().picg_source_create(species.h, world.h, C.c_double(v_drift), C.c_double(den), C.c_double(T), _FACES[inlet_face.lower()], C.byref(self.h)))

    def close(self):
        if self.h:
            lib().picg_source_destroy(self.h)
            self.h = C.c_void_p()

    def sample(self):
        n = C.c_size_t(0)
        _chk(lib().picg_source_sample(self.h, C.byref(n)))
        return n.value


class ColdBeamSource(_Source):
    """ColdBeamSource (ch4/v3/src/Source.h:41-52)."""

    def __init__(self, species, world, v_drift, den, inlet_face="-z"):
        super().__init__(species, world, v_drift, den, 0.0, inlet_face)


class WarmBeamSource(_Source):
    """WarmBeamSource (ch4/v3/src/Source.h:54-66)."""

    def __init__(self, species, world, v_drift, den, T, inlet_face="-z"):
        super().__init__(species, world, v_drift, den, T, inlet_face)
