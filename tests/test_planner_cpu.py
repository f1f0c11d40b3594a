"""The launch plan of the benchmark's workload, checked on CPU through the planner context (TNE_PLAN_ONLY=1: programs are
built, costed and dumped, nothing is ever computed).  Pins the host logic that decides what runs on the GPU: segments,
slices, fused site kernels, fused contributions, flop count and workspace of the 6x6 D=4 <H>+grad programme."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _plan(*args):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "plan_dump.py")] + list(args), capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    return res.stdout


def test_benchmark_programme_plan():
    out = _plan("6", "4")
    m = re.search(r"total model time ([0-9.]+) s, flops ([0-9.e+]+)", out)
    assert m, out[-2000:]
    flops = float(m.group(2))
    # energy programme: term-DP forward + per-column recompute + reverse pass (8.8e13 with the norm programme)
    assert abs(flops - 7.88e13) / 7.88e13 < 0.02, flops
    rows = {l.split()[0]: l.split() for l in out.splitlines() if l and l.split()[0] in ("site2", "multi", "reduce", "generic")}
    assert int(rows["site2"][1]) == 3072          # fused bra+ket absorption of the bulk sites, every slice of every sliced column
    assert int(rows["multi"][1]) > 0 and int(rows["reduce"][1]) > 0
    phases = [l for l in out.splitlines() if l.startswith("phase")]
    assert len(phases) == 11                      # 5 forward columns (the last one with its reverse ops) + 5 reverse columns ... 6 columns
    assert sum("bwd" in p for p in phases) == 6 and sum("fwd" in p for p in phases) == 5


def test_planner_reports_workspace_and_refuses_to_compute():
    code = r'''
import os, sys
os.environ["TNE_PLAN_ONLY"] = "1"
sys.path.insert(0, %r)
import numpy as np
import tne_b200 as T
from tne_b200 import peps as P, einsum as es
ctx = T.context(); ctx.set_option("debug", 2)
a = T.B200Array.from_numpy(np.ones((2, 2)))
try:
    a.to_numpy(); print("DOWNLOAD-OK")
except T.TneError as e:
    print("DOWNLOAD-REFUSED", e)
L, D = 8, 4
g = T.grid([L, L]); order, segs = T.sweep_plan(g, L)
p = T.rand_simplepeps(g, D, np.random.default_rng(1), optimizer=order, segments=segs)
leaves = [P._dev(t) for t in p.alltensors()] * 2
try:
    es.contract_tree_device(p.code_inner_product, leaves, [1] * 64 + [0] * 64); print("CONTRACT-OK")
except T.TneError as e:
    print("CONTRACT-REFUSED", e)
''' % ROOT
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "DOWNLOAD-REFUSED" in res.stdout and "CONTRACT-REFUSED" in res.stdout, res.stdout + res.stderr[-1500:]
    m = re.search(r"\[tne-plan\] (\d+) phases, (\d+) launches, ([0-9.e+]+) flops, workspace ([0-9.]+) GiB", res.stderr)
    assert m, res.stderr[-1500:]
    # the memory wall of the exact 8x8 D=4 norm: a 64 GiB boundary per checkpoint
    assert abs(float(m.group(3)) - 1.354e15) / 1.354e15 < 0.02 and float(m.group(4)) > 180.0
