#!/usr/bin/env python
"""Generates the committed golden fixtures from the REFERENCE ITSELF (run where /root/reference exists).

  tests/golden/v3_small.npz        inputs + outputs of the compiled ch4/v3 reference (oracle/_ref/libref_v3.so) on small
                                   seeded cases: electron push, addParticle, number density, per-cell count, moments,
                                   charge density, Gauss-Seidel potential (tight and loose), E field, heavy push (ions).
  tests/golden/ch2_trajectory.csv  the first rows of runtime_diags.csv written by the STOCK ch2/v2 binary (BASELINE
                                   config 1: RNG-free quiet start, so the trajectory is reproducible bit for bit).

The reference ships no golden vectors; these are outputs of its unmodified sources.
    python tests/golden/make_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from oracle import ref_v3 as ref  # noqa: E402

REF = "/root/reference"


def v3_small():
    ref.lib(); ref.config(subcycling=False, multithreading=False, merging=False, sputtering=False)
    ni, nj, nk = 11, 9, 13
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    sph = [((0.0, 0.0, 0.0025), -100.0, 0.0007)]
    out = dict(ni=ni, nj=nj, nk=nk, x0=x0, xm=xm, rect_c=np.array([r[0] for r in rects]), rect_phi=np.array([r[1] for r in rects]),
               rect_s=np.array([r[2] for r in rects]), sph_c=np.array(sph[0][0]), sph_phi=sph[0][1], sph_r=sph[0][2])
    w = util.build_world(ref.World, ni, nj, nk, x0, xm, rects, sph, dt=1e-12)
    out["node_vol"] = w.get(2); out["object_id"] = w.get(4); out["phi0"] = w.get(0)
    ef = util.smooth_ef((ni, nj, nk), x0, xm, seed=3, amp=3e6)
    w.set(3, ef); out["ef"] = ef
    # electron push
    parts = util.random_particles(6000, x0, xm, seed=4, vth=2e6)
    e = ref.Species("e-", util.ME, -util.QE, w, 100.0); e.setParticles(parts)
    e.advanceElectrons(2e-10)
    out["push_in"] = parts; out["push_dt"] = 2e-10; out["push_out_sorted"] = util.sort_rows(e.getParticles())
    # deposit / count / moments on the survivors
    surv = e.getParticles()
    e.computeNumberDensity(); e.computeMacroParticlesCount(); e.sampleMoments()
    out["dep_in"] = surv; out["den"] = e.get(0); out["macro_count"] = e.get(4)
    out["n_sum"] = e.get(5); out["nv_sum"] = e.get(6); out["nuu"] = e.get(7)
    # addParticle (filter + rewind)
    cand = util.random_particles(1500, x0 - 0.1 * (xm - x0), xm + 0.1 * (xm - x0), seed=6)
    ion = ref.Species("O+", 16 * util.AMU, util.QE, w, 100.0)
    for p in cand:
        ion.addParticle(p)
    out["add_in"] = cand; out["add_out"] = ion.getParticles()
    # heavy push of ions (deterministic part: no RNG is consumed, the injected-neutral count is 0)
    neu = ref.Species("O", 16 * util.AMU, 0.0, w, 5e11)
    ions_in = util.random_particles(8000, x0, xm, seed=5, vth=4e4, mpw=(100.0, 100.0), lo_frac=(0, 0, 0.1), hi_frac=(1, 1, 0.9))
    ion2 = ref.Species("O+", 16 * util.AMU, util.QE, w, 100.0); ion2.setParticles(ions_in)
    ref.seed(7)
    for _ in range(2):
        ion2.advanceNonElectron(neu, neu, 2e-8)
    out["heavy_in"] = ions_in; out["heavy_dt"] = 2e-8; out["heavy_out_sorted"] = util.sort_rows(ion2.getParticles())
    # charge density from two charged species
    ion.computeNumberDensity()
    w.computeChargeDensity([e, ion])
    out["den_ion"] = ion.get(0); out["rho"] = w.get(1)
    # potential: the reference's GS on a second world without the sphere, random rho; loose (as shipped) and tight
    w2 = util.build_world(ref.World, 13, 11, 17, *util.discharge_geometry(13, 11, 17)[:2], util.discharge_geometry(13, 11, 17)[2])
    rng = np.random.default_rng(9)
    rho = rng.normal(0, 1e-7, (13, 11, 17)); w2.set(1, rho)
    out["gs_rho"] = rho; out["gs_phi_start"] = w2.get(0); out["gs_object_id"] = w2.get(4)
    sol = ref.PotentialSolver(w2, 20000, 1e-4, ref.PotentialSolver.GS); sol.setReferenceValues(0.0, 0.0, 1e20)
    out["gs_converged"] = sol.solveGS(); out["gs_phi"] = w2.get(0)
    sol.computeEF(); out["gs_ef"] = w2.get(3)
    np.savez_compressed(os.path.join(HERE, "v3_small.npz"), **out)
    print("wrote v3_small.npz with", len(out), "arrays")


def ch2_trajectory(rows=151):
    src = os.path.join(REF, "ch2", "v2")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "pic_ch2")
        cpps = [os.path.join(src, f) for f in os.listdir(src) if f.endswith(".cpp")]
        subprocess.check_call(["/usr/bin/g++", "-std=c++20", "-O2", "-ffp-contract=off", "-w", "-o", exe] + cpps)
        os.makedirs(os.path.join(d, "results"), exist_ok=True)
        try:
            subprocess.run([exe], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=25, input=b"")
        except subprocess.TimeoutExpired:
            pass
        lines = open(os.path.join(d, "runtime_diags.csv")).read().strip().splitlines()
    keep = lines[:1 + rows]
    # drop the wall-clock column (not reproducible)
    hdr = keep[0].split(","); iw = hdr.index("wall_time")
    outp = [",".join(c for i, c in enumerate(l.split(",")) if i != iw) for l in keep]
    open(os.path.join(HERE, "ch2_trajectory.csv"), "w").write("\n".join(outp) + "\n")
    print("wrote ch2_trajectory.csv with", len(outp) - 1, "steps")


if __name__ == "__main__":
    v3_small()
    ch2_trajectory()
