#!/usr/bin/env python
"""Benchmark of the PIC-DSMC per-timestep particle loop (BASELINE.json: particle-steps/s, Poisson ms/step, % HBM roofline).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path through the C ABI (libpicgpu.so)
  python bench.py --impl reference --steps K --warmup W     # the reference's own CPU code (oracle/_ref, or the C port)

Workload (BASELINE.json configs[4], SURVEY.md 8d "C5"): World(256,256,256) with dx=1e-4 m, the two electrode
Rectangles of ch4/v3/src/main.cpp:94-98 at -/+4000 V, dt=1e-12 s, 1e9 macro-particles uniform in the gap
(0.1..0.9 Lz): 5e8 O (300 K), 2.5e8 O+ (300 K), 2.5e8 e- (3000 K), velocities by the reference's sampleV3th law,
generated on the device with Philox.  One step = the body of the v3 main loop (main.cpp:177-288) with sub-cycling
off and the Poisson solve live (as in ch2/ch3/ch4-v1): MC ionisation (cell sort of the two collision partners
included), push of every species, number-density deposit + per-cell macro-particle count, charge density,
red-black SOR solve (warm start, reference tolerance) and E = -grad(phi).
Particles are split evenly across the N GPUs (strong scaling); every GPU deposits onto a full-grid fixed-point
accumulator; the accumulators are reduce-scattered (NCCL) onto the slabs of the Poisson solve, whose halo planes, residual
sum and final all-gather go through NVLink peer memory inside the sweep kernels (--poisson replicated / --allreduce_density
select the simpler variants).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
PKG = "engineering-degree-in-plasma-simulations_b200"

AMU, QE, ME, NA = 1.660538921e-27, 1.602176565e-19, 9.10938215e-31, 6.02214076e23
# algorithmic bytes (SURVEY.md 8d / DESIGN.md): per particle or per node, per launch of the kernel
ALG_BYTES_PER_PARTICLE = {"push_electrons": 96, "push_heavy": 96, "push_neutral": 72, "push_electrons_deposit": 104, "push_heavy_deposit": 104, "deposit_density": 32}
ALG_BYTES_PER_NODE = {"sor_redblack": 12.5, "sor_tiled": 25.0, "compute_ef": 32, "charge_density": 24, "finalize_density": 24}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mesh", type=int, default=256, help="nodes per axis")
    ap.add_argument("--particles", type=float, default=1e9, help="total macro-particles over all GPUs")
    ap.add_argument("--s_max_it", type=int, default=50, help="max SOR iterations per step (warm start)")
    ap.add_argument("--s_tol", type=float, default=1.0, help="L2 tolerance of the solve (v3 default, main.cpp:83)")
    ap.add_argument("--sort_every", type=int, default=1, help="cell-sort period of species that MC ionisation does not sort itself")
    ap.add_argument("--no_mcc", action="store_true")
    ap.add_argument("--moments", action="store_true", help="also run Species::sampleMoments every step (SURVEY 8f row 1)")
    ap.add_argument("--cpu_sample_nodes", type=int, default=49, help="nodes per axis of the CPU-baseline sub-volume (same dx, same particles per cell)")
    ap.add_argument("--skip_cpu_baseline", action="store_true")
    ap.add_argument("--cpu_leg", default=None, help="internal: 'steps,warmup,no_mcc' - run one CPU reference leg and print its JSON (child process of the default run)")
    ap.add_argument("--allreduce_density", action="store_true", help="multi-GPU: all-reduce the density accumulators (full grids everywhere) instead of a reduce-scatter onto the slabs")
    ap.add_argument("--poisson", choices=["auto", "replicated", "slab"], default="auto",
                    help="multi-GPU solve: every rank solves the whole grid, or planes split over the ranks with peer-memory halos (auto: slab when N > 1)")
    ap.add_argument("--poisson_full_max_it", type=int, default=8000, help="iteration budget of the once-per-run solve to the reference's tolerance (main.cpp:82); 0: skip")
    ap.add_argument("--init_max_it", type=int, default=20000, help="iteration cap of the initial vacuum solve (profiling runs use a small value)")
    ap.add_argument("--subcycled_steps", type=int, default=None, help="steps of the reference's subcycled loop timed after the headline (0: skip; default 10 on one GPU, 0 on several)")
    ap.add_argument("--fp32_steps", type=int, default=5, help="steps of the fp32 secondary path timed after everything else on one GPU (0: skip)")
    ap.add_argument("--inject", type=int, default=1 << 20, help="e2e: electrons injected from pinned host memory per step")
    return ap.parse_args()


# --------------------------------------------------------------------------------------- workload definition
def workload(mesh, n_total):
    """Geometry and species of the synthetic discharge; particle counts follow from densities and weights."""
    dx = 1e-4
    L = dx * (mesh - 1)
    x0 = np.zeros(3)
    xm = np.full(3, L)
    xc = 0.5 * (x0 + xm)
    rects = [((xc[0], xc[1], x0[2]), -4000.0, (L, L, L * 0.1)), ((xc[0], xc[1], xm[2]), 4000.0, (L, L, L * 0.1))]
    box_c = xc.copy()
    box_s = np.array([L, L, 0.8 * L])
    vol = box_s.prod()
    frac = {"O": 0.5, "O+": 0.25, "e-": 0.25}
    den = {"O": 1e24, "O+": 1e16, "e-": 1e16}
    T = {"O": 300.0, "O+": 300.0, "e-": 3000.0}
    mass = {"O": 16 * AMU, "O+": 16 * AMU, "e-": ME}
    charge = {"O": 0.0, "O+": QE, "e-": -QE}
    ppc_total = n_total / ((mesh - 1) ** 2 * (mesh - 1) * 0.8)
    species = []
    for name in ("O", "O+", "e-"):
        count = n_total * frac[name]
        mpw0 = den[name] * vol / count
        species.append(dict(name=name, mass=mass[name], charge=charge[name], mpw0=mpw0, den=den[name], T=T[name], count=int(count)))
    return dict(mesh=mesh, dx=dx, x0=x0, xm=xm, rects=rects, box_c=box_c, box_s=box_s, species=species, dt=1e-12, ppc=ppc_total,
                E_ion=1313.9 * 1000 / NA)


def sub_volume(wl, nodes):
    """The CPU-baseline sample: a sub-box of the same plasma (same dx, same particles per cell, electrodes kept)."""
    frac_cells = ((nodes - 1) ** 3) / float((wl["mesh"] - 1) ** 3)
    n_total = sum(s["count"] for s in wl["species"]) * frac_cells
    return workload(nodes, n_total), n_total


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self, first=0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw = [], [], []
        reasons = set()
        rows = self.rows[first:] if len(self.rows) - first >= 3 else self.rows      # prefer samples taken inside the timed region
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "samples_in_timed_region": max(0, len(self.rows) - first), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pg = importlib.import_module(PKG + ".picgpu")
    pg.init(local)
    pg.set_rank(rank, world)
    pg.seed(0x5EED0000)
    if os.environ.get("PICG_MERGE_FRACTION"):                  # tuning switches (profiles/r2_list_policy.md)
        pg.set_merge_fraction(float(os.environ["PICG_MERGE_FRACTION"]))
    if os.environ.get("PICG_MOVER_FRACTION"):
        pg.set_mover_fraction(float(os.environ["PICG_MOVER_FRACTION"]))
    stream = torch.cuda.ExternalStream(pg.stream_ptr(), device=local)

    if args.subcycled_steps is None:                       # a secondary, single-GPU measurement unless asked for
        args.subcycled_steps = 10 if world == 1 else 0
    wl = workload(args.mesh, args.particles)
    m = wl["mesh"]
    nv_total = m ** 3
    w = pg.World(m, m, m, wl["x0"], wl["xm"])
    w.setTime(wl["dt"], 1 << 30)
    for c, phi, sides in wl["rects"]:
        w.addRectangle(c, phi, sides)
    w.computeObjectID()
    sol = pg.PotentialSolver(w, args.s_max_it, args.s_tol)
    sol.setReferenceValues(0.0, 0.0, 1e20)                 # main.cpp:138
    cold = pg.PotentialSolver(w, args.init_max_it, args.s_tol)   # initial vacuum solve (main.cpp:172-173)
    cold.setReferenceValues(0.0, 0.0, 1e20)
    poisson_mode = "replicated"
    poisson_probe = None
    if world > 1 and args.poisson in ("auto", "slab"):
        def all_gather_bytes(b):
            t = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
            out = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            return [o.cpu().numpy().tobytes() for o in out]

        def probe(solver, iters=20):
            """ms per SOR iteration of `solver`, device-timed, max over ranks (every rank must take the same decision)."""
            solver.iterate(4)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            pg.synchronize(); dist.barrier(); torch.cuda.synchronize()
            e0.record(stream); solver.iterate(iters); e1.record(stream); pg.synchronize(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        try:
            use_slabs = True
            if args.poisson == "auto":
                # north star: "solved either redundantly or slab-decomposed, whichever is faster at the given mesh size": measure both
                # (the probe iterations are iterations of the initial vacuum solve, nothing is thrown away)
                rep = pg.PotentialSolver(w, args.init_max_it, args.s_tol); rep.setReferenceValues(0.0, 0.0, 1e20)
                slb = pg.PotentialSolver(w, args.init_max_it, args.s_tol); slb.setReferenceValues(0.0, 0.0, 1e20)
                slb.enableSlabs(rank, world, all_gather_bytes)
                poisson_probe = {"replicated_ms_per_iteration": round(probe(rep), 4), "slab_ms_per_iteration": round(probe(slb), 4)}
                use_slabs = poisson_probe["slab_ms_per_iteration"] < poisson_probe["replicated_ms_per_iteration"]
                rep.close(); slb.close()
            if use_slabs:
                cold.enableSlabs(rank, world, all_gather_bytes)
                sol.enableSlabs(rank, world, all_gather_bytes)
                poisson_mode = "slab (planes of i split over %d ranks; halos, residual sum and all-gather through NVLink peer memory)" % world
            else:
                poisson_mode = "replicated (measured faster than slabs at this mesh size)"
        except pg.PicgError as e:                              # no peer access on this node: every rank solves the whole grid
            if args.poisson == "slab":
                raise
            poisson_mode = "replicated (slab mode unavailable: %s)" % e
    t0 = time.time()
    cold.solveGS(); cold.computeEF()
    init_iters = cold.iterations

    mg = importlib.import_module(PKG + ".multigpu")

    def reduce_min(v):
        t = torch.tensor([v], device="cuda", dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t.item())

    def reduce_max(v):
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    phi_initial = w.download(pg.F_PHI).copy()                # the potential every (re-)created plasma starts from
    species = {}
    neu = ion = ele = mcc = None
    order = []
    fixed_views = {}
    slab_range, scatter_chunks, density_mode = None, None, "all-reduce (full grids on every rank)"

    def create_plasma():
        """Species, particles, collision handler and fixed-point scales of the run.  Called again before the end-to-end leg: the runs
        are reproducible, so the second plasma is the first one over again, particle for particle."""
        nonlocal neu, ion, ele, mcc, order, fixed_views, slab_range, scatter_chunks, density_mode
        for s in wl["species"]:
            sp = pg.Species(s["name"], s["mass"], s["charge"], w, s["mpw0"], wl["E_ion"] if s["name"] == "O" else -666.0)
            per_rank = s["count"] // world + 1
            # no store may be re-allocated inside a timed region: the neutral store grows by the split-off neutrals of the MC collisions
            # (about 0.5 % per step of this workload), also over the steps of the subcycled loop
            head = 1.25 + (0.85 if s["name"] == "O" else 0.0)     # the neutral store doubles within the run (split-off neutrals of the MC collisions)
            sp.reserve(int(per_rank * head) + args.inject * (args.steps + args.warmup + 8))
            sp.loadParticleBoxThermal(wl["box_c"], wl["box_s"], s["den"], s["T"])
            sp.sort()
            species[s["name"]] = sp
        neu, ion, ele = species["O"], species["O+"], species["e-"]
        order = [neu, ion, ele]
        mcc = None
        if not args.no_mcc:
            import util
            E, sg = util.momentum_transfer_table()
            mcc = pg.MC_MEX_Ionization(neu, ion, ele, w, E, sg)
        # a common fixed-point scale on every rank (the all-reduce sums raw int64 accumulators)
        for sp in order:
            sp.computeNumberDensity()
            sp.setDensityScale(mg.common_scale(sp.densityScale(), world, reduce_min))
        fixed_views = {sp.name: mg.fixed_view(torch, sp, pg.SF_DEN_FIXED) for sp in order} if world > 1 else {}
        # Slab Poisson reads rho only on this rank's planes: reduce-scatter the accumulators onto the slabs (in place: rank r's output is
        # the r-th chunk of its own input) and finalize / sum charges there only.  Needs equal chunks = planes divisible by the ranks.
        slab_range, scatter_chunks, density_mode = None, None, "all-reduce (full grids on every rank)"
        if world > 1 and poisson_mode.startswith("slab") and m % world == 0 and not args.allreduce_density:
            slab_range = sol.slabRange()
            scatter_chunks = {name: list(v.chunk(world)) for name, v in fixed_views.items()}
            assert slab_range == (rank * (nv_total // world), (rank + 1) * (nv_total // world))
            density_mode = "reduce-scatter onto the Poisson slabs (density and rho finalised on the owned planes only)"

    def destroy_plasma():
        nonlocal mcc, fixed_views, scatter_chunks
        fixed_views = {}; scatter_chunks = None
        if mcc is not None:
            mcc.close(); mcc = None
        for sp in order:
            sp.close()
        species.clear()

    create_plasma()
    comm_stream = torch.cuda.Stream(device=local) if world > 1 else None
    setup_s = time.time() - t0

    counts = {}

    prof = {} if os.environ.get("PICG_E2E_PROFILE") else None   # host wall time per phase (debugging aid; synchronises)

    def stamp(name, t0):
        if prof is not None:
            pg.synchronize(); prof[name] = prof.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return time.perf_counter()

    def step(ts, inject=None, subcycling=False, advanced=None):
        """One pass of the v3 main-loop body (main.cpp:197-261) through the C ABI.  subcycling: Config::SUBCYCLING of the reference
        (main.cpp:211-236): electrons every step, ions every 10th step with 10 dt, neutrals every 100th step with 50 dt; a species that is
        not advanced keeps its density.  advanced: list that receives the species advanced in this step."""
        t0 = time.perf_counter()
        if inject is not None:                              # Source::sample on the host side: H2D of this step's new particles
            ele.addParticles(inject)
            t0 = stamp("inject", t0)
        if mcc is not None:
            st = mcc.apply(wl["dt"])
            counts["mcc"] = (st.candidates, st.collisions, st.ionizations, st.dropped)
            if counts.get("sampling"):
                # populations the kernels of THIS step work on (the apply call has just refreshed the host-side counts: no extra
                # synchronisation) and the part of each store the cell partition covers: the per-launch algorithmic bytes follow from these
                counts["samples"].append({sp.name: (sp.getNumParticles(), sp.partitionSize()) for sp in order})
                counts["mcc_samples"].append(counts["mcc"])
            if world > 1:                                   # the acceptance ceiling must be the same on every rank
                mcc.setWsvMax(mg.common_ceiling(st.w_sigma_v_max, reduce_max))
            t0 = stamp("mcc", t0)
        if mcc is None and counts.get("sampling"):
            counts["samples"].append({sp.name: (sp.getNumParticles(), sp.partitionSize()) for sp in order})
        pending = []
        for sp in order:
            dt_sp = wl["dt"]
            if subcycling and sp is not ele:
                if sp.charge != 0:
                    if ts % 10 != 0:
                        continue
                    dt_sp = wl["dt"] * 10
                else:
                    if ts % 100 != 0:
                        continue
                    dt_sp = wl["dt"] * 50
            if advanced is not None:
                advanced.append(sp)
            if sp is ele:
                sp.advanceElectrons(dt_sp)
            else:
                sp.advanceNonElectron(neu, neu, dt_sp)
            if world == 1:
                sp.computeNumberDensity()
                sp.computeMacroParticlesCount()
            else:
                sp.depositPartial()
                sp.computeMacroParticlesCount()
                # int64 sum over NVLink on a side stream: it overlaps the next species' push; the finalize waits for it below
                ev = torch.cuda.Event(); ev.record(stream)
                comm_stream.wait_event(ev)
                with torch.cuda.stream(comm_stream):
                    if scatter_chunks is not None:                              # each rank only needs the sums on the planes it solves on
                        dist.reduce_scatter_tensor(scatter_chunks[sp.name][rank], fixed_views[sp.name])
                    else:
                        dist.all_reduce(fixed_views[sp.name])
                    done = torch.cuda.Event(); done.record(comm_stream)
                pending.append((sp, done))
            if args.moments:
                sp.sampleMoments()
            if mcc is None and args.sort_every and ts % args.sort_every == 0:
                sp.sort()
            t0 = stamp("species " + sp.name, t0)
        for sp, done in pending:
            stream.wait_event(done)
            sp.finalizeDensity(slab_range)
        if ts > 5:
            for sp in order:
                sp.updateAverages()
        w.computeChargeDensity(order, slab_range)
        sol.solveGS()
        sol.computeEF()
        stamp("fields", t0)
        return sol.iterations

    def barrier():
        pg.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    trace_steps = bool(os.environ.get("PICG_STEP_TRACE"))
    trace_prev = {}

    def timed(n_steps, ts0, e2e=False, inject_bufs=None, rho_host=None):
        """Times n_steps on the device (CUDA events on the library's stream); returns ms, particle-steps, iterations."""
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        barrier()
        psteps = 0
        its = 0
        step_events = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps)]
        ev0.record(stream)
        for k in range(n_steps):
            n_now = sum(sp.getNumParticles() for sp in order) if e2e else None
            it = step(ts0 + k, inject_bufs[k % len(inject_bufs)] if e2e else None)
            its += it
            if trace_steps and rank == 0:                     # diagnosis only (synchronises every step): per-step kernel times and populations
                cur = pg.timers_read()
                delta = {kk: round(v[0] - trace_prev.get(kk, (0.0, 0))[0], 3) for kk, v in cur.items() if v[0] - trace_prev.get(kk, (0.0, 0))[0] > 0.02}
                trace_prev.clear(); trace_prev.update(cur)
                print("step %d: n=%s part=%s movers(stats)=%s mcc(cand,coll,ion,drop)=%s ms=%s" % (ts0 + k, {sp.name: sp.getNumParticles() for sp in order}, {sp.name: sp.partitionSize() for sp in order},
                      pg.mover_stats(), counts.get("mcc"), json.dumps(delta)), file=sys.stderr)
            if e2e:                                           # what the reference loop reads every step: counts + diagnostics + rho
                t0 = time.perf_counter()
                pg._chk(pg.lib().picg_world_download_begin(w.h, pg.F_RHO, rho_host.ctypes.data_as(pg.C.POINTER(pg.C.c_double))))   # rho travels while the diagnostics run
                for sp in order:
                    sp.diagnostics()
                t0 = stamp("diagnostics", t0)
                pg._chk(pg.lib().picg_world_download_end(w.h))
                stamp("rho download (rest)", t0)
                psteps += n_now
            step_events[k].record(stream)
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        counts["step_ms"] = [round(([ev0] + step_events)[k].elapsed_time(step_events[k]), 3) for k in range(n_steps)]
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, psteps, its

    def global_count():
        n = sum(sp.getNumParticles() for sp in order)
        if world > 1:
            t = torch.tensor([n], device="cuda", dtype=torch.int64)
            dist.all_reduce(t)
            n = int(t.item())
        return n

    def multi_gpu_self_check(ts):
        """N > 1 only, outside the timed region: one more step of the multi-GPU path, then rank 0 REDOES its grid work on one GPU from the
        same inputs and compares bit for bit:
          * every rank's particles (x, y, z, mpw as they are after the step's push) are gathered onto rank 0 over NCCL and deposited by
            the single-GPU path into a second World (same scale S): den_fixed and den against the reduce-scattered / all-reduced slabs
            of the ranks, rho (World::computeChargeDensity of the single path) against the ranks' rho slabs;
          * the potential: a replicated solve on rank 0 from the potential before the step and the single-path rho, against the slab
            solve's phi (same iteration count), E against E; and all ranks must hold the same phi and E (checksums)."""
        def view(ptr_bytes, typestr="<f8"):
            ptr, nbytes = ptr_bytes
            return torch.as_tensor(mg.CudaArray(ptr, nbytes // 8, typestr), device="cuda")

        phi_live = view(w.device_ptr(pg.F_PHI)); ef_live = view(w.device_ptr(pg.F_EF)); rho_live = view(w.device_ptr(pg.F_RHO))
        barrier()
        phi_before = phi_live.clone()
        step(ts)
        barrier()
        detail = {}

        def assemble(local_full, dtype):
            """The global array on rank 0: the owned slabs of every rank (slab mode) or rank 0's own full copy."""
            if slab_range is None:
                return local_full
            a, b = slab_range
            chunk = local_full[a:b].contiguous()
            if rank == 0:
                out = torch.empty(nv_total, dtype=dtype, device="cuda")
                out[a:b] = chunk
                for r in range(1, world):
                    dist.recv(out[r * (nv_total // world):(r + 1) * (nv_total // world)], src=r)
                return out
            dist.send(chunk, dst=0)
            return None

        def mismatches(a, b):
            return int((a.view(torch.int64) != b.view(torch.int64)).sum().item())

        w2 = chk = None
        if rank == 0:
            w2 = pg.World(m, m, m, wl["x0"], wl["xm"]); w2.setTime(wl["dt"], 1 << 30)
            for c, phi, sides in wl["rects"]:
                w2.addRectangle(c, phi, sides)
            w2.computeObjectID()
            chk = []
        for sp, sdef in zip(order, wl["species"]):
            n_local = sp.getNumParticles()
            cnt = torch.tensor([n_local], device="cuda", dtype=torch.int64)
            allc = [torch.zeros_like(cnt) for _ in range(world)]
            dist.all_gather(allc, cnt)
            allc = [int(c.item()) for c in allc]
            src_ptrs, src_cap = sp.particleArrays(0)
            src = {k: torch.as_tensor(mg.CudaArray(src_ptrs[k], src_cap, "<f8"), device="cuda") for k in (0, 1, 2, 6)}
            if rank == 0:
                c = pg.Species(sdef["name"], sdef["mass"], sdef["charge"], w2, sdef["mpw0"])
                ptrs, cap = c.particleArrays(sum(allc))
                dst = {k: torch.as_tensor(mg.CudaArray(ptrs[k], cap, "<f8"), device="cuda") for k in range(7)}
                for k in (3, 4, 5):
                    dst[k][:sum(allc)].zero_()                       # velocities are not part of the deposit
                for k in (0, 1, 2, 6):
                    dst[k][:allc[0]].copy_(src[k][:allc[0]])
                    off = allc[0]
                    for r in range(1, world):
                        dist.recv(dst[k][off:off + allc[r]], src=r); off += allc[r]
                torch.cuda.synchronize()
                c.adopt(sum(allc)); c.setDensityScale(sp.densityScale())
                c.computeNumberDensity()                                # the single-GPU path: deposit of ALL particles + finalize on the full grid
                pg.synchronize()
                chk.append(c)
            else:
                for k in (0, 1, 2, 6):
                    dist.send(src[k][:n_local].contiguous(), dst=0)
            fixed_all = assemble(view(sp.device_ptr(pg.SF_DEN_FIXED), "<i8"), torch.int64)
            den_all = assemble(view(sp.device_ptr(pg.SF_DEN)), torch.float64)
            if rank == 0:
                detail["den_fixed " + sp.name] = mismatches(fixed_all, view(c.device_ptr(pg.SF_DEN_FIXED), "<i8"))
                detail["den " + sp.name] = mismatches(den_all, view(c.device_ptr(pg.SF_DEN)))
                detail["particles " + sp.name] = sum(allc)
            del fixed_all, den_all
        rho_all = assemble(rho_live, torch.float64)
        sums = torch.stack([phi_live.view(torch.int64).sum(), ef_live.view(torch.int64).sum()])       # wrap-around checksums of the bit patterns
        all_sums = [torch.zeros_like(sums) for _ in range(world)]
        dist.all_gather(all_sums, sums)
        if rank == 0:
            w2.computeChargeDensity(chk); pg.synchronize()        # (the comparison runs on torch's stream: the library's must have finished)
            detail["rho"] = mismatches(rho_all, view(w2.device_ptr(pg.F_RHO)))
            if detail["rho"] and os.environ.get("PICG_SELF_CHECK_DEBUG"):
                r2 = view(w2.device_ptr(pg.F_RHO))
                bad = (rho_all.view(torch.int64) != r2.view(torch.int64)).nonzero().flatten()
                for u in bad[:8].tolist() + bad[-3:].tolist():
                    print("rho mismatch at node %d = (i %d, j %d, k %d): multi-GPU %r, one GPU %r; den(one GPU): %s" % (
                        u, u // (m * m), (u // m) % m, u % m, float(rho_all[u]), float(r2[u]), [float(view(c.device_ptr(pg.SF_DEN))[u]) for c in chk]), file=sys.stderr)
            view(w2.device_ptr(pg.F_PHI)).copy_(phi_before); torch.cuda.synchronize()
            sol2 = pg.PotentialSolver(w2, args.s_max_it, args.s_tol); sol2.setReferenceValues(0.0, 0.0, 1e20)
            sol2.solveGS(); sol2.computeEF(); pg.synchronize()
            detail["phi"] = mismatches(phi_live, view(w2.device_ptr(pg.F_PHI)))
            detail["ef"] = mismatches(ef_live, view(w2.device_ptr(pg.F_EF)))
            detail["iterations"] = [int(sol.iterations), int(sol2.iterations)]
            detail["ranks_hold_the_same_phi_and_ef"] = all(bool((t == all_sums[0]).all().item()) for t in all_sums)
            for o in [sol2] + chk + [w2]:
                o.close()
        del phi_before, rho_all
        torch.cuda.empty_cache()
        barrier()
        if rank != 0:
            return None, None
        ok = all(detail[k] == 0 for k in detail if k.split(" ")[0] in ("den_fixed", "den", "rho", "phi", "ef")) and \
            detail["iterations"][0] == detail["iterations"][1] and detail["ranks_hold_the_same_phi_and_ef"]
        return ("bit_identical" if ok else "MISMATCH"), detail

    # ---- warm-up (the clock sampler starts here: nvidia-smi needs ~0.5 s before its first sample)
    clocks = ClockSampler(local); clocks.start()
    ts = 1
    for _ in range(args.warmup):
        step(ts); ts += 1
    mg_parity, mg_detail = (None, None)
    if world > 1 and not os.environ.get("PICG_SKIP_SELF_CHECK"):
        mg_parity, mg_detail = multi_gpu_self_check(ts); ts += 1
    n_start = global_count()
    # ---- timed region: device-resident inputs, per-kernel CUDA-event timers on
    pg.timers_reset(); pg.timers_enable(True); pg.launch_count_reset()
    n_before = len(clocks.rows)
    reallocs0 = pg.realloc_count()
    counts["sampling"] = True; counts["samples"] = []; counts["mcc_samples"] = []
    ms, _, its = timed(args.steps, ts)
    counts["sampling"] = False
    step_ms = counts["step_ms"]
    samples = counts["samples"]
    reallocs_timed = pg.realloc_count() - reallocs0
    clk = clocks.stop(first=n_before)
    launches = pg.launch_count()
    pg.timers_enable(False)
    kt = pg.timers_read()
    ts += args.steps
    n_end = global_count()
    # particle-steps = sum over the timed steps of the particles each step advanced (sampled per step), all ranks
    psteps_timed = float(sum(n for smp in counts["samples"] for (n, _part) in smp.values()))
    if world > 1:
        t = torch.tensor([psteps_timed], device="cuda", dtype=torch.float64); dist.all_reduce(t); psteps_timed = float(t.item())
    n_avg = psteps_timed / args.steps
    value = psteps_timed / (ms * 1e-3)

    # ---- roofline of the dominant kernel (largest accumulated device time)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    per_rank_counts = {sp.name: sp.getNumParticles() for sp in order}
    nv = m ** 3
    # Algorithmic bytes per launch, from the populations AT each launch (sampled every step, see step()): sum over the steps of
    # bytes-per-particle x particles of that step, divided by the number of launches.  Node kernels: the nodes THIS rank's launch
    # touches (the owned slab when Poisson / the density finalisation run on slabs).
    n_sum = {name: float(sum(smp[name][0] for smp in samples)) for name in per_rank_counts}                         # particle-launches per species
    cov_sum = float(sum(min(n, part) if part else n for smp in samples for (n, part) in smp.values()))             # covered by a cell partition
    tail_sum = float(sum(max(n - part, 0) for smp in samples for (n, part) in smp.values() if part))               # appended beyond it
    slab_nodes = (slab_range[1] - slab_range[0]) if slab_range is not None else nv
    sor_nodes = slab_nodes if poisson_mode.startswith("slab") else nv
    kernels = {}
    for name, (tot_ms, n_l) in kt.items():
        avg = tot_ms / n_l
        entry = {"ms_total": round(tot_ms, 4), "launches": int(n_l), "ms_avg": round(avg, 5)}
        alg_total = None                                                            # algorithmic bytes of ALL launches in the timed region
        if name in ("push_electrons_deposit", "push_electrons"):
            alg_total = ALG_BYTES_PER_PARTICLE[name] * n_sum["e-"]
        elif name == "push_heavy":
            alg_total = 96 * (n_sum["O+"] if "push_neutral" in kt else n_sum["O"] + n_sum["O+"])
        elif name == "push_neutral":                                               # drift only: pos + vel read, pos written (no kick for charge 0)
            alg_total = 72 * n_sum["O"]
        elif name == "deposit_density":
            # the cell-group kernel, one launch per species and step, over the part of the store the cell partition covers; the
            # particles appended since the last sort are deposited by the thread-run kernel ("deposit_tail", timed on its own)
            alg_total = 32 * cov_sum
        elif name == "deposit_tail":
            alg_total = 32 * tail_sum
        elif name == "sor_redblack":
            alg_total = ALG_BYTES_PER_NODE[name] * sor_nodes * n_l
        elif name in ("finalize_density", "charge_density"):
            alg_total = ALG_BYTES_PER_NODE[name] * slab_nodes * n_l
        elif name in ALG_BYTES_PER_NODE:
            alg_total = ALG_BYTES_PER_NODE[name] * nv * n_l
        if alg_total:
            entry["alg_GB_per_launch"] = round(alg_total / n_l / 1e9, 4)
            entry["GBps"] = round(alg_total / (tot_ms * 1e-3) / 1e9, 1)
            entry["frac_of_peak"] = round(entry["GBps"] / peak, 4)
        kernels[name] = entry
    # the deposit stage as a whole: cell-group kernel + tail kernel, all species (north star: ">= 60 % on the push and deposit kernels")
    if "deposit_density" in kernels:
        t_dep = kernels["deposit_density"]["ms_total"] + kernels.get("deposit_tail", {"ms_total": 0.0})["ms_total"]
        deposit_stage = {"alg_GB_per_step": round(32 * (cov_sum + tail_sum) / args.steps / 1e9, 3), "ms_per_step": round(t_dep / args.steps, 3),
                         "GBps": round(32 * (cov_sum + tail_sum) / (t_dep * 1e-3) / 1e9, 1)}
        deposit_stage["frac_of_peak"] = round(deposit_stage["GBps"] / peak, 4)
    else:
        deposit_stage = None
    dom = max((k for k in kernels if "GBps" in kernels[k]), key=lambda k: kernels[k]["ms_total"])
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture of this workload (profiles/capture.sh ->
    # profiles/ncu_traffic.py); only quoted when the run IS that workload (same mesh, particle count, one GPU)
    traffic, traffic_src = None, None
    ncu_names = {"deposit_density": "k_cell_deposit", "push_heavy": "k_run<1, 1, 0, 0, 0>", "push_neutral": "k_run<1, 1, 0, 0, 1>", "push_electrons": "k_run<1, 0, 0, 0, 0>", "sor_redblack": "k_sor_row"}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r2b.json")))
        if world == 1 and m == 256 and abs(args.particles - 1e9) < 1 and dom in ncu_names:
            ent = [e for k, v in tj["kernels"].items() if k.startswith(ncu_names[dom]) for e in v]    # deposit: all lane-group variants
            traffic = float(np.mean([e["dram_bytes"] for e in ent]))
            traffic_src = "profiles/ncu_traffic_r2b.json (%s, mean of %d captured launches)" % (ncu_names[dom], len(ent))
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["GBps"], "peak": peak, "unit": "GB/s", "frac": kernels[dom]["frac_of_peak"],
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "alg_bytes_per_launch": kernels[dom]["alg_GB_per_launch"] * 1e9,
                "note": "peak is the measured COPY bandwidth (1 read : 1 write); a stream that reads more than it writes can exceed it (push_neutral reads 48 B and writes 24 B per particle)"}
    poisson_ms = sum(kernels[k]["ms_total"] for k in ("sor_redblack", "sor_tiled", "residual_l2", "compute_ef") if k in kernels) / args.steps

    # ---- end to end through the C ABI with host buffers (pinned): inject + diagnostics + rho download every step
    inj = []
    rng = np.random.default_rng(7 + rank)
    n_inj = args.inject // world
    for _ in range(2):
        buf = torch.empty((n_inj, 7), dtype=torch.float64).pin_memory()
        a = buf.numpy()
        a[:, 0:3] = wl["box_c"] + (rng.random((n_inj, 3)) - 0.5) * wl["box_s"]
        a[:, 3:6] = rng.normal(0, 2e5, (n_inj, 3)); a[:, 6] = ele.mpw0
        inj.append(a)
    rho_pin = torch.empty(nv, dtype=torch.float64).pin_memory()
    rho_host = rho_pin.numpy()
    # The same steps as the device-resident leg, on the same plasma: the run is re-created (runs are reproducible, so this is the first
    # plasma over again), warmed up like the first (the multi-GPU self-check took one step on every rank), and the K steps are then taken
    # through the host-buffer path.  PICG_E2E_AFTER=1: the round-1 behaviour (five steps following the timed ones).
    did_self_check = world > 1 and not os.environ.get("PICG_SKIP_SELF_CHECK")          # the same on every rank (mg_parity is only set on rank 0)
    if not os.environ.get("PICG_E2E_AFTER"):
        e2e_steps = args.steps
        barrier()
        destroy_plasma()
        barrier()
        w.upload(pg.F_PHI, phi_initial); sol.computeEF()
        barrier()                                              # no rank starts a slab sweep (peer stores into the neighbours' phi) before every rank holds the initial potential
        create_plasma()
        ts = 1
        for _ in range(args.warmup + (1 if did_self_check else 0)):
            step(ts); ts += 1
    else:
        e2e_steps = max(2, min(args.steps, 5))
    e2e_first_step_population = global_count()
    reallocs1 = pg.realloc_count()
    pg.timers_reset(); pg.timers_enable(True)
    if prof is not None:
        prof.clear()
    ms_e2e, ps_local, _ = timed(e2e_steps, ts, e2e=True, inject_bufs=inj, rho_host=rho_host)
    ts += e2e_steps
    if prof is not None and rank == 0:
        print("e2e host profile (ms/step): " + json.dumps({k: round(v / e2e_steps, 2) for k, v in prof.items()}), file=sys.stderr)
    pg.timers_enable(False)
    kt_e2e = {k: round(v[0] / e2e_steps, 3) for k, v in pg.timers_read().items()}
    if world > 1:
        t = torch.tensor([ps_local], device="cuda", dtype=torch.int64); dist.all_reduce(t); ps_local = int(t.item())
    e2e = {"value": ps_local / (ms_e2e * 1e-3), "unit": "particle-steps/s", "h2d_bytes_per_step": int(n_inj * 56 * world),
           "d2h_bytes_per_step": int((nv * 8 + 3 * 5 * 8 + 3 * 64) * world), "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
           "kernel_ms_per_step": kt_e2e, "device_reallocs": int(pg.realloc_count() - reallocs1),
           "what": "C-ABI step with host buffers: H2D of injected electrons (pinned), D2H of per-species counters + diagnostics and of rho every step; "
                   + ("the plasma of the device-resident leg re-created (same seed: the same particles) and taken through the SAME step numbers" if not os.environ.get("PICG_E2E_AFTER") else
                      "the steps that follow the device-resident leg"), "same_start_as_value": (bool(e2e_first_step_population == n_start) if not os.environ.get("PICG_E2E_AFTER") else None)}

    # ---- Poisson to the reference's tolerance.  The step above runs the solve warm-started with a cap of --s_max_it iterations (both arms);
    # here the SAME solve is run once with the reference's own budget (main.cpp:82 --s_max_it 8000, tolerance main.cpp:83) from the
    # current potential and charge density, timed on the device: iterations to tolerance (or the residual left at the cap) and its cost.
    poisson_full = None
    if args.poisson_full_max_it > 0:
        full = pg.PotentialSolver(w, args.poisson_full_max_it, args.s_tol)
        full.setReferenceValues(0.0, 0.0, 1e20)
        if poisson_mode.startswith("slab"):
            full.enableSlabs(rank, world, all_gather_bytes)
        l2_start = sol.L2
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream); conv = full.solveGS(); e1.record(stream)
        barrier()
        ms_full = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_full], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_full = float(t.item())
        poisson_full = {"max_it": args.poisson_full_max_it, "tol": args.s_tol, "converged": bool(conv), "iterations": int(full.iterations), "ms": round(ms_full, 3),
                        "ms_per_iteration": round(ms_full / max(1, full.iterations), 5), "l2_before": l2_start, "l2_after": full.L2,
                        "what": "one warm-started red-black SOR solve with the reference's iteration budget (main.cpp:82) and tolerance on the state the timed steps left; "
                                "the L2 residual is the reference's (PotentialSolver.cpp:124-159: in V/m^2, i.e. 1/dx^2 = 1e8 times a potential error)"}
        full.close()

    # ---- the reference's subcycled loop (Config::SUBCYCLING + MERGING), reported separately: steps ts = 300, 301, ... so that
    # the window starts with a Species::merge round (main.cpp:179-193: ts > 250 and ts % 50 == 0), an ion push and a neutral push; the default
    # 10 steps hold 10 electron pushes, 1 ion push, 1 neutral push and 1 merge.  (The window cannot be long: in this synthetic discharge the
    # electrons gain energy in the 8 kV gap, the MC collision rate rises to 3e7 per step within 30 steps and every non-ionising collision
    # splits a neutral - the neutral store doubles in 30 steps.)  particle-steps = particles actually advanced
    subcycled = None
    if args.subcycled_steps > 0:
        try:
            ts_sub = 300
            merge_stats = []
            ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
            barrier()
            adv_total = 0
            ev0.record(stream)
            for k in range(args.subcycled_steps):
                adv = []
                if (ts_sub + k) % 50 == 0:
                    for sp in order:
                        merge_stats.append((sp.name,) + sp.merge()[:2])
                step(ts_sub + k, subcycling=True, advanced=adv)
                adv_total += sum(sp.getNumParticles() for sp in adv)
                if os.environ.get("PICG_TRACE_SUBCYCLED") and rank == 0 and k % 5 == 0:
                    print("subcycled ts %d: %s mcc %s merges %s" % (ts_sub + k, {sp.name: sp.getNumParticles() for sp in order},
                          (mcc.stats.candidates, mcc.stats.collisions, mcc.stats.ionizations, mcc.stats.w_sigma_v_max) if mcc else None, merge_stats[-3:]), file=sys.stderr)
            ev1.record(stream)
            barrier()
            ms_sub = ev0.elapsed_time(ev1)
            if world > 1:
                t = torch.tensor([ms_sub], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_sub = float(t.item())
                t = torch.tensor([adv_total], device="cuda", dtype=torch.int64); dist.all_reduce(t); adv_total = int(t.item())
            subcycled = {"steps": args.subcycled_steps, "ms_per_step": ms_sub / args.subcycled_steps, "value": adv_total / (ms_sub * 1e-3), "unit": "particle-steps/s",
                         "merges": [{"species": a, "before": int(b), "after": int(c)} for a, b, c in merge_stats],
                         "what": "Config::SUBCYCLING + MERGING loop of the reference (main.cpp:179-236): electrons every step, ions every 10th (10 dt), neutrals every 100th (50 dt), "
                                 "Species::merge every 50th; MC ionisation, charge density, Poisson and E every step; particle-steps count the particles actually advanced"}
        except Exception as e:                         # the secondary measurement must never take the headline line down with it
            if world > 1:
                raise                                   # (a rank that stops here would leave the others in a collective)
            subcycled = {"error": str(e)[:300]}

    # ---- fp32 secondary path (SURVEY 8d "fp64 headline, fp32 secondary"): the same plasma converted on the device to the cell-relative
    # single-precision store (csrc/f32.cu) and advanced with the kernels that path has: electron-type push of every species (kick, drift,
    # absorption), fixed-point deposit + per-cell count, charge density, the same capped Poisson solve and E.  No collisions, no wall
    # re-emission (fp64 only).  One GPU; last leg of the run (the fp64 stores are released one by one as they are converted).
    fp32 = None
    if args.fp32_steps > 0 and world == 1:
        try:
            if mcc is not None:
                mcc.close()
            sp32 = []
            for sp in (ele, ion, neu):                      # the small stores first: their fp64 memory is free before the large one is converted
                c = pg.Species32(sp.name, sp.mass, sp.charge, w, sp.mpw0)
                c.fromSpecies(sp); sp.close()
                c.sort()
                sp32.append(c)
            for c in sp32:
                c.computeNumberDensity()                    # fixes the scale S, sizes the scratch arena
            counts32 = [c.getNumParticles() for c in sp32]

            def step32():
                for c in sp32:
                    c.advanceElectrons(wl["dt"]); c.computeNumberDensity()
                pg.charge_density32(w, sp32)
                sol.solveGS(); sol.computeEF()
            step32(); step32()
            pg.timers_reset(); pg.timers_enable(True)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            barrier(); e0.record(stream)
            for _ in range(args.fp32_steps):
                step32()
            e1.record(stream); barrier()
            pg.timers_enable(False)
            ms32 = e0.elapsed_time(e1)
            k32 = pg.timers_read()
            n32 = float(sum(c.getNumParticles() for c in sp32))
            n32_avg = 0.5 * (n32 + float(sum(counts32)))
            kern32 = {}
            for name, bpp in (("push_electrons", 56), ("deposit_density", 20)):      # 28 R + 28 W; cell + 3 fractions + weight
                if name in k32:
                    tot_ms, n_l = k32[name]
                    gbps = bpp * n32_avg * args.fp32_steps / (tot_ms * 1e-3) / 1e9
                    kern32[name] = {"ms_per_step": round(tot_ms / args.fp32_steps, 3), "launches": int(n_l), "alg_bytes_per_particle": bpp, "GBps": round(gbps, 1), "frac_of_peak": round(gbps / peak, 4)}
            fp32 = {"dtype": "f32", "steps": args.fp32_steps, "ms_per_step": ms32 / args.fp32_steps, "value": n32_avg * args.fp32_steps / (ms32 * 1e-3), "unit": "particle-steps/s",
                    "particles": int(n32_avg), "kernels": kern32,
                    "what": "cell-relative fp32 store (32 B per particle): push (electron-type: gather, kick, drift, absorption) + fixed-point deposit + per-cell count of every species, "
                            "charge density, capped SOR solve and E each step; no collisions, no wall re-emission (fp64 path only)"}
            for c in sp32:
                c.close()
        except Exception as e:                                  # a secondary leg never takes the headline line down
            fp32 = {"error": str(e)[:300]}

    out = None
    if rank == 0:
        out = {"metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": "synthetic %d^3 mesh, %.3g macro-particles (O/O+/e- = 2:1:1), full PIC-DSMC step; Poisson: warm-started SOR (omega 1.4) capped at %d iterations per step "
                                      "(tolerance %g is not reached within the cap, see poisson_to_tolerance)" % (m, args.particles, args.s_max_it, args.s_tol),
                          "mesh": [m, m, m], "particles_global": int(n_avg), "species": {k: int(v) for k, v in per_rank_counts.items()},
                          "steps_per_sort": 1 if mcc else args.sort_every, "mcc": mcc is not None, "moments": bool(args.moments),
                          "poisson": {"max_it": args.s_max_it, "tol": args.s_tol, "mode": poisson_mode, "probe": poisson_probe, "iterations_per_step": its / args.steps,
                                      "initial_solve_iterations": init_iters},
                          "parallelism": "particles split by index over %d GPU(s); int64 density %s; Poisson %s" % (world, density_mode if world > 1 else "on one GPU", poisson_mode.split(" ")[0]),
                          "l2_policy": "inputs larger than L2 (%.1f GB of particle arrays per GPU vs 126 MB L2)" % (sum(per_rank_counts.values()) * 56 / 1e9)},
               "poisson_ms_per_step": poisson_ms, "poisson_to_tolerance": poisson_full, "step_ms": step_ms, "deposit_stage": deposit_stage,
               "mcc_per_step": {"candidates": [c[0] for c in counts["mcc_samples"]], "collisions": [c[1] for c in counts["mcc_samples"]]} if mcc else None,
               "populations_per_step": [{k: v[0] for k, v in smp.items()} for smp in samples],
               "gpu_launches": int(launches), "device_reallocs_in_timed_region": int(reallocs_timed), "clocks": clk, "roofline": roofline, "kernels": kernels, "e2e": e2e,
               "subcycled": subcycled, "fp32": fp32, "setup_s": round(setup_s, 1)}
        if world > 1:
            # rank 0 redid the step's grid work on ONE GPU from all ranks' particles (gathered over NCCL) and compared bit patterns
            out["multi_gpu_parity"] = mg_parity
            out["multi_gpu_parity_detail"] = mg_detail
        if not args.skip_cpu_baseline:
            # Each CPU leg runs in a child process: the reference's thread pool is racy and a crash of the reference must not
            # take the device line down with it.
            out["cpu_baseline"] = cpu_leg_in_child(args, steps=2, warmup=1, no_mcc=args.no_mcc)
            if not args.no_mcc:
                # the reference's multithreaded configuration (Config.cpp:68-75: hardware_concurrency() - 1 workers; its thread pool serves the
                # electron push only) is only safe without MC ionisation in the loop: the same step with the interaction off
                mt = cpu_leg_in_child(args, steps=2, warmup=1, no_mcc=True, tries=1)
                if "error" in mt:
                    mt["finding"] = ("the reference's thread-pool electron push cannot run this workload: ThreadPool::AddTask binds its arguments by value (ThreadPool.h:63), so the "
                                     "per-thread delete buffers of Species.cpp:276-282 stay empty, electrons that leave the domain or hit an electrode are never removed and the next "
                                     "Field::scatter writes out of range (SIGSEGV in computeNumberDensity); the serial leg above is the reference's only working configuration here")
                out["cpu_baseline_multithreaded"] = mt
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


# --------------------------------------------------------------------------------------- reference arm (CPU)

def cpu_leg_in_child(args, steps, warmup, no_mcc, tries=2):
    """Runs cpu_reference_run in a child python (same flags) and returns its dict; a crash or a time-out becomes {"error": ...}."""
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu_leg", "%d,%d,%d" % (steps, warmup, 1 if no_mcc else 0), "--mesh", str(args.mesh), "--particles", repr(args.particles),
           "--s_max_it", str(args.s_max_it), "--s_tol", repr(args.s_tol), "--cpu_sample_nodes", str(args.cpu_sample_nodes)] + (["--moments"] if args.moments else [])
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    err = "not run"
    for _ in range(tries):
        try:
            p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=600, env=env)
        except subprocess.TimeoutExpired:
            err = "CPU leg timed out after 600 s"; continue
        lines = [l for l in p.stdout.decode(errors="replace").splitlines() if l.startswith("{")]
        if p.returncode == 0 and lines:
            return json.loads(lines[-1])
        err = "CPU leg exited with code %d (the reference crashed)" % p.returncode
    return {"error": err}


def cpu_reference_run(args, wl_full, steps, warmup, no_mcc=None):
    """Times the reference's own CPU implementation of the same step on a bounded sample of the workload: a sub-volume of
    the same plasma (same dx, dt, densities and particles per cell, electrodes kept).  Uses oracle/_ref (the unmodified
    reference compiled here) when present, else the C port of the same loops."""
    from oracle import ref_v3
    import util
    no_mcc = args.no_mcc if no_mcc is None else no_mcc
    wl, n_total = sub_volume(wl_full, args.cpu_sample_nodes)
    m = wl["mesh"]
    cores = os.cpu_count() or 1
    if ref_v3.available():
        ref_v3.lib()
        # Config.cpp:68-75 default is hardware_concurrency()-1 threads, used by the electron push only.  With MC ionisation
        # in the loop the thread-pool push leaves stale per-cell index lists behind (the reference then reads
        # particles past the end of its store and crashes), so the reference is run serial in that case.
        threads = 1 if not no_mcc else max(1, cores - 1)
        ref_v3.config(subcycling=False, multithreading=threads > 1, num_threads=threads, merging=False, sputtering=False)
        saved_stdout = os.dup(1)                             # the reference prints progress to stdout: keep the JSON line clean
        sys.stdout.flush()
        devnull = os.open(os.devnull, os.O_WRONLY); os.dup2(devnull, 1)
        ref_v3.seed(12345)
        w = ref_v3.World(m, m, m, wl["x0"], wl["xm"])
        w.setTime(wl["dt"], 1 << 30)
        for c, phi, sides in wl["rects"]:
            w.addRectangle(c, phi, sides)
        w.computeObjectID()
        sol = ref_v3.PotentialSolver(w, args.s_max_it, args.s_tol, ref_v3.PotentialSolver.GS)
        sol.setReferenceValues(0.0, 0.0, 1e20)
        cold = ref_v3.PotentialSolver(w, 20000, args.s_tol, ref_v3.PotentialSolver.GS)
        cold.setReferenceValues(0.0, 0.0, 1e20)
        cold.solveGS(); cold.computeEF()
        sp = {}
        for s in wl["species"]:
            o = ref_v3.Species(s["name"], s["mass"], s["charge"], w, s["mpw0"], wl["E_ion"] if s["name"] == "O" else -666.0)
            o.loadParticleBoxThermal(wl["box_c"], wl["box_s"], s["den"], s["T"])
            sp[s["name"]] = o
        order = [sp["O"], sp["O+"], sp["e-"]]
        tmp = tempfile.mkdtemp()
        mcc = None if no_mcc else ref_v3.MC_MEX_Ionization(sp["O"], sp["O+"], sp["e-"], w, util.write_table(os.path.join(tmp, "Oxygen_momentum_transfer.txt")))

        def step(ts):
            if mcc:
                mcc.apply(wl["dt"])
            for o in order:
                if o is sp["e-"]:
                    o.advanceElectrons(wl["dt"])
                else:
                    o.advanceNonElectron(sp["O"], sp["O"], wl["dt"])
                o.computeNumberDensity()
                if args.moments:
                    o.sampleMoments()
                o.computeMacroParticlesCount()
            if ts > 5:
                for o in order:
                    o.updateAverages()
            w.computeChargeDensity(order)
            sol.solveGS(); sol.computeEF()

        ts = 1
        for _ in range(warmup):
            step(ts); ts += 1
        n0 = sum(o.getNumParticles() for o in order)
        t0 = time.perf_counter()
        for _ in range(steps):
            step(ts); ts += 1
        dt = time.perf_counter() - t0
        n1 = sum(o.getNumParticles() for o in order)
        os.dup2(saved_stdout, 1); os.close(saved_stdout); os.close(devnull)
        kind = "reference"
    else:
        raise SystemExit("oracle/_ref is absent and the C-port whole-step baseline is not wired: build oracle/_ref where /root/reference exists")
    val = 0.5 * (n0 + n1) * steps / dt
    return {"value": val, "unit": "particle-steps/s", "cores": threads, "host_cores": cores, "kind": kind, "ms_per_step": dt / steps * 1e3,
            "sample": "sub-volume of the same plasma: %d^3 nodes (same dx, dt, densities, %.1f particles/cell), %d macro-particles, %d steps; "
                      "reference on %d thread(s)%s (only its electron push can use the thread pool, Species.cpp:258-355; serial when MC ionisation is on, SURVEY B20)"
                      % (m, wl["ppc"], int(n0), steps, threads, ", MC ionisation off" if no_mcc and not args.no_mcc else "")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args.mesh, args.particles)
    cb = cpu_reference_run(args, wl, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 3)))
    out = {"impl": "reference", "metric": "particle-steps/s", "value": cb["value"], "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": "synthetic %d^3 mesh, %.3g macro-particles (O/O+/e- = 2:1:1), full PIC-DSMC step; Poisson: warm-started SOR (omega 1.4) capped at %d iterations per step "
                                  "(tolerance %g is not reached within the cap, see poisson_to_tolerance)" % (args.mesh, args.particles, args.s_max_it, args.s_tol),
                      "note": "CPU arm timed on a bounded sample of this workload, see cpu_baseline.sample"},
           "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse()
    if a.cpu_leg:
        st, wu, nm = (int(v) for v in a.cpu_leg.split(","))
        print(json.dumps(cpu_reference_run(a, workload(a.mesh, a.particles), steps=st, warmup=wu, no_mcc=bool(nm))))
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
