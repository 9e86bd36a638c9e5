#!/usr/bin/env python
"""SASS excerpts of the hot kernels from the built objects (no GPU needed): opcode counts and the instructions that prove the data path
(bulk async copies with mbarrier completion, wide streaming loads / stores, 64-bit integer reductions, no FMA contraction).
usage: python profiles/sass_excerpts.py > profiles/r2_sass_excerpts.md   (after `make` in the package directory)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "engineering-degree-in-plasma-simulations_b200", "build")
KERNELS = [("cellstep.o", r"k_cell_deposit<\(bool\)1, \(bool\)1, \(int\)3>|k_cell_depositILb1ELb1ELi3E"),
           ("cellstep.o", r"k_cell_depositILb1ELb1ELi2E"),
           ("step.o", r"k_runILb1ELb0ELb0ELb0ELb0E"), ("step.o", r"k_runILb1ELb1ELb0ELb0ELb0E"), ("step.o", r"k_runILb1ELb1ELb0ELb0ELb1E"),
           ("step.o", r"k_runILb0ELb0ELb1ELb1ELb0E"), ("poisson.o", r"k_sor_row"), ("mcc.o", r"k_mccILi0E"), ("mcc.o", r"k_stage_commit"), ("push.o", r"k_bm_emit")]
PROOF = re.compile(r"UBLKCP|SYNCS|LDG\.E\.(EF\.)?(ENL2\.)?(128|256)|STG\.E\.(EF\.)?(ENL2\.)?(128|256)|RED\.E\.ADD\.64|ATOMG\.E\.ADD\.64|ATOMS|DFMA|LDGSTS|MATCH|REDUX")
print("# SASS excerpts of the hot kernels (sm_100a; `cuobjdump -sass build/<file>.o`, nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a -O3 -fmad=false`)\n")
print("What to look for: bulk asynchronous copies global -> shared with mbarrier completion (`UBLKCP.S.G`, `SYNCS.*`), wide streaming loads / stores (`LDG.E.EF.*.256` / `.128`, `STG.E.EF.*`), "
      "64-bit integer reductions (`RED.E.ADD.64`, `ATOMG.E.ADD.64`, `ATOMS`), no FMA contraction in the parity-critical fp64 (`DMUL` / `DADD`; a `DFMA` count of 0 in the push / deposit bodies).\n"
      "The `DFMA`s of `k_sor_row` and `k_mcc` belong to the division and libm sequences (1/x, sqrt, log, exp, pow, sin, cos), which the compiler emits as library code whatever `-fmad` says.\n"
      "Regenerate with `python profiles/sass_excerpts.py` after a build.\n")
cache = {}
for obj, pat in KERNELS:
    path = os.path.join(BUILD, obj)
    if obj not in cache:
        cache[obj] = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    txt = cache[obj]
    blocks = re.split(r"\n\s*Function : ", txt)
    hit = [b for b in blocks[1:] if re.search(pat, b.split("\n", 1)[0])]
    if not hit:
        print("## %s : `%s` not found\n" % (obj, pat)); continue
    b = hit[0]
    name = b.split("\n", 1)[0].strip()
    try:
        name = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        pass
    ins = re.findall(r"/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)([^;]*);", b)
    ops = collections.Counter(i[2].split(".")[0] for i in ins)
    full = collections.Counter(i[2] for i in ins if PROOF.search(i[2]))
    print("## `%s` (%s)" % (name[:160], obj))
    print("%d SASS instructions; opcode counts: %s" % (len(ins), ", ".join("%s %d" % kv for kv in ops.most_common(14))))
    print("DFMA: %d, DMUL: %d, DADD: %d" % (ops.get("DFMA", 0), ops.get("DMUL", 0), ops.get("DADD", 0)))
    print("matching instructions: " + (", ".join("%s x%d" % kv for kv in full.most_common(12)) or "-"))
    shown = 0
    print("```")
    for addr, pred, op, rest in ins:
        if PROOF.search(op) and not op.startswith("DFMA") and shown < 8:
            print("/*%s*/ %s%s%s ;" % (addr, pred or "", op, rest.rstrip())); shown += 1
    print("```\n")
