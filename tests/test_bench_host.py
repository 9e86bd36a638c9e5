"""Host-side logic of bench.py that needs no GPU: the workload definition, the CPU sample of the reference arm and the isolation of the
reference's CPU legs in child processes (a crash of the reference must never take the device line down)."""
import argparse
import json
import subprocess
import sys

import numpy as np
import pytest

import bench
import util


def test_workload_is_baseline_config_5():
    wl = bench.workload(256, 1e9)
    assert wl["mesh"] == 256 and wl["dx"] == 1e-4 and wl["dt"] == 1e-12
    counts = {s["name"]: s["count"] for s in wl["species"]}
    assert counts == {"O": 500_000_000, "O+": 250_000_000, "e-": 250_000_000}
    for s in wl["species"]:                                   # the weights follow from the densities: count * mpw0 = density * loader volume
        assert s["count"] * s["mpw0"] == pytest.approx(s["den"] * wl["box_s"].prod(), rel=1e-12)
    (c0, phi0, s0), (c1, phi1, s1) = wl["rects"]              # the v3 electrodes: 0.1 Lz thick, centred on the z faces, -/+ 4000 V (main.cpp:92-98)
    L = wl["xm"][2] - wl["x0"][2]
    assert (phi0, phi1) == (-4000.0, 4000.0) and c0[2] == wl["x0"][2] and c1[2] == wl["xm"][2] and s0[2] == pytest.approx(0.1 * L)


def test_cpu_sample_keeps_the_plasma():
    wl = bench.workload(256, 1e9)
    sub, n = bench.sub_volume(wl, 49)
    assert sub["dx"] == wl["dx"] and sub["dt"] == wl["dt"] and sub["ppc"] == pytest.approx(wl["ppc"], rel=1e-12)      # same cell size, same particles per cell
    assert n == pytest.approx(1e9 * (48 / 255) ** 3, rel=1e-12)
    for a, b in zip(sub["species"], wl["species"]):
        assert a["den"] == b["den"] and a["T"] == b["T"] and a["mpw0"] == pytest.approx(b["mpw0"], rel=1e-9)           # same densities, temperatures and weights


def _args(**kw):
    base = dict(mesh=32, particles=2e5, s_max_it=50, s_tol=1.0, cpu_sample_nodes=13, moments=False, no_mcc=False)
    base.update(kw)
    return argparse.Namespace(**base)


@pytest.mark.reference
def test_cpu_leg_runs_in_a_child_and_returns_its_line(ref):
    out = bench.cpu_leg_in_child(_args(), steps=1, warmup=1, no_mcc=False)
    assert "error" not in out, out
    assert out["kind"] == "reference" and out["cores"] == 1 and out["value"] > 0 and "13^3 nodes" in out["sample"]


def test_a_crashing_cpu_leg_becomes_an_error_entry():
    # a one-node sample can not be built: the child dies, the parent gets a dictionary with the reason and goes on
    out = bench.cpu_leg_in_child(_args(cpu_sample_nodes=1), steps=1, warmup=1, no_mcc=False, tries=1)
    assert set(out) == {"error"} and "exited with code" in out["error"]


@pytest.mark.reference
def test_reference_arm_prints_the_contract_line(ref):
    p = subprocess.run([sys.executable, bench.__file__, "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1", "--mesh", "32", "--particles", "2e5",
                        "--cpu_sample_nodes", "13"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "particle-steps/s" and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["e2e"] == {"value": line["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["value"] == line["value"]
