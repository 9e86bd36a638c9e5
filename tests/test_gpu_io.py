"""Binary outputs and restart files: picg_write_fields_vti, picg_checkpoint_save / _load.

The .vti is parsed back with a few lines of Python (VTK ImageData, appended raw block, UInt64 headers) and every array is
compared bit for bit with the device fields; the array names and their order are the reference's (Outputs.cpp:38-112).
A run resumed from a checkpoint must continue bit for bit on the deterministic kernels (push, deposit, fields).
"""
import os
import re
import tempfile

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

NI, NJ, NK = 11, 9, 13


def _setup(pg, seed=5):
    x0, xm, rects = util.discharge_geometry(NI, NJ, NK)
    w = util.build_world(pg.World, NI, NJ, NK, x0, xm, rects, dt=2e-11)
    sol = pg.PotentialSolver(w, 400, 1e-3)
    sol.setReferenceValues(0.0, 0.0, 1e20)
    sol.solveGS(); sol.computeEF()
    ele = pg.Species("e-", util.ME, -util.QE, w, 100.0)
    ion = pg.Species("O+", 16 * util.AMU, util.QE, w, 100.0)
    ele.setParticles(util.random_particles(30000, x0, xm, seed=seed, vth=1e6, mpw=(100.0, 100.0), lo_frac=(0, 0, 0.15), hi_frac=(1, 1, 0.85)))
    ion.setParticles(util.random_particles(20000, x0, xm, seed=seed + 1, vth=3e3, mpw=(100.0, 100.0), lo_frac=(0, 0, 0.15), hi_frac=(1, 1, 0.85)))
    return w, sol, ele, ion


def _step(w, sol, ele, ion, dt=2e-11):
    ele.advanceElectrons(dt); ion.advanceNonElectron(ion, ion, dt)
    for sp in (ele, ion):
        sp.computeNumberDensity(); sp.sampleMoments(); sp.computeMacroParticlesCount(); sp.updateAverages()
    w.computeChargeDensity([ele, ion])
    sol.solveGS(); sol.computeEF()


def _state(pg, w, species):
    out = {"phi": w.phi, "rho": w.rho, "ef": w.ef}
    for sp in species:
        out["p." + sp.name] = util.sort_rows(sp.getParticles())
        for f, name in ((pg.SF_DEN, "den"), (pg.SF_DEN_AVG, "den_avg"), (pg.SF_N_SUM, "n_sum"), (pg.SF_NV_SUM, "nv_sum"), (pg.SF_NUU_SUM, "nuu")):
            out[name + "." + sp.name] = sp.download(f)
    return out


def _parse_vti(path):
    raw = open(path, "rb").read()
    marker = b'<AppendedData encoding="raw">\n_'
    head, blob = raw[:raw.index(marker)].decode(), raw[raw.index(marker) + len(marker):]
    ext = [int(v) for v in re.search(r'WholeExtent="([^"]+)"', head).group(1).split()]
    ni, nj, nk = ext[1] + 1, ext[3] + 1, ext[5] + 1
    cells_at = head.index("<CellData>")
    arrays = []
    for m in re.finditer(r'<DataArray Name="([^"]+)" NumberOfComponents="(\d)" format="appended" type="Float64" offset="(\d+)"/>', head):
        name, comps, off = m.group(1), int(m.group(2)), int(m.group(3))
        cell = m.start() > cells_at
        dims = (ni - 1, nj - 1, nk - 1) if cell else (ni, nj, nk)
        nbytes = int(np.frombuffer(blob[off:off + 8], dtype="<u8")[0])
        assert nbytes == dims[0] * dims[1] * dims[2] * comps * 8
        a = np.frombuffer(blob[off + 8:off + 8 + nbytes], dtype="<f8").reshape(dims[2], dims[1], dims[0], comps)     # VTK order: i fastest
        a = a.transpose(2, 1, 0, 3)
        arrays.append((name, a[..., 0] if comps == 1 else a))
    return head, arrays


def test_fields_vti_binary_holds_the_device_fields(picgpu):
    pg = picgpu
    w, sol, ele, ion = _setup(pg)
    for _ in range(2):
        _step(w, sol, ele, ion)
    for sp in (ele, ion):
        sp.computeGasProperties()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "fields_00002.vti")
        pg.write_fields_vti(path, w, [ele, ion])
        head, arrays = _parse_vti(path)
    # the reference's arrays, names and order (Outputs.cpp:38-112)
    want_names = ["NodeVol", "ObjectID", "NodeType", "phi", "rho", "nd.e-", "nd.O+", "avg_nd.e-", "avg_nd.O+", "vel.e-", "vel.O+", "T.e-", "T.O+", "ef", "mpc.e-", "mpc.O+"]
    assert [n for n, _ in arrays] == want_names
    got = dict(arrays)
    want = {"NodeVol": w.node_vol, "ObjectID": w.object_id, "NodeType": w.download(pg.F_NODE_TYPE), "phi": w.phi, "rho": w.rho, "ef": w.ef}
    for sp in (ele, ion):
        want.update({"nd." + sp.name: sp.den, "avg_nd." + sp.name: sp.download(pg.SF_DEN_AVG), "vel." + sp.name: sp.download(pg.SF_VEL),
                     "T." + sp.name: sp.download(pg.SF_T), "mpc." + sp.name: sp.macro_part_count})
    for name in want_names:
        assert np.array_equal(got[name], want[name], equal_nan=True), name
    assert got["nd.e-"].sum() > 0 and got["mpc.O+"].sum() == ion.getNumParticles()
    assert 'WholeExtent="0 %d 0 %d 0 %d"' % (NI - 1, NJ - 1, NK - 1) in head
    for o in (ele, ion, sol, w):
        o.close()


def test_checkpoint_resume_is_bit_exact(picgpu):
    pg = picgpu
    E, sg = util.momentum_transfer_table()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "run.ckp")
        w, sol, ele, ion = _setup(pg)
        neu = pg.Species("O", 16 * util.AMU, 0.0, w, 5e11, 1313.9 * 1000 / util.NA)
        mcc = pg.MC_MEX_Ionization(neu, ion, ele, w, E, sg); mcc.setWsvMax(0.25)
        for _ in range(3):
            _step(w, sol, ele, ion)
        pg.checkpoint_save(path, w, [ele, ion, neu], mcc=[mcc], ts=3)
        saved = _state(pg, w, (ele, ion))
        for _ in range(3):
            _step(w, sol, ele, ion)
        want = _state(pg, w, (ele, ion))
        for o in (mcc, neu, ele, ion, sol, w):
            o.close()

        # a fresh process would rebuild the objects as at start-up (other particles on purpose) and load
        w, sol, ele, ion = _setup(pg, seed=99)
        neu = pg.Species("O", 16 * util.AMU, 0.0, w, 5e11, 1313.9 * 1000 / util.NA)
        mcc = pg.MC_MEX_Ionization(neu, ion, ele, w, E, sg)
        assert pg.checkpoint_load(path, w, [ele, ion, neu], mcc=[mcc]) == 3
        loaded = _state(pg, w, (ele, ion))
        for k in saved:
            assert np.array_equal(loaded[k], saved[k]), "state after load: " + k
        assert mcc.apply(2e-11).w_sigma_v_max == 0.25          # no electrons-neutral pairs (no neutrals): the ceiling is the stored one
        for _ in range(3):
            _step(w, sol, ele, ion)
        got = _state(pg, w, (ele, ion))
        for k in want:
            if k.split(".")[0] in ("n_sum", "nv_sum", "nuu"):        # sampleMoments accumulates with fp64 atomics: order-dependent last bits
                assert util.norm_err(got[k], want[k]) < 1e-12, "resumed run: " + k          # norm-wise: the velocity sums cancel
            else:
                assert np.array_equal(got[k], want[k]), "resumed run: " + k

        # mismatches are refused
        with pytest.raises(pg.PicgError):
            pg.checkpoint_load(path, w, [ele, ion], mcc=[mcc])                  # species count
        with pytest.raises(pg.PicgError):
            pg.checkpoint_load(path, w, [ion, ele, neu], mcc=[mcc])             # species constants
        x0, xm, rects = util.discharge_geometry(NI + 2, NJ, NK)
        w2 = util.build_world(pg.World, NI + 2, NJ, NK, x0, xm, rects)
        e2 = pg.Species("e-", util.ME, -util.QE, w2, 100.0)
        with pytest.raises(pg.PicgError):
            pg.checkpoint_load(path, w2, [e2])                                  # mesh
        with open(path, "r+b") as f:
            f.truncate(os.path.getsize(path) // 2)
        with pytest.raises(pg.PicgError):
            pg.checkpoint_load(path, w, [ele, ion, neu], mcc=[mcc])             # truncated file
        for o in (e2, w2, mcc, neu, ele, ion, sol, w):
            o.close()
