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
    rows = {l.split()[0]: l.split() for l in out.splitlines() if l and l.split()[0] in ("site2", "site2b", "multi", "reduce", "generic")}
    assert int(rows["site2"][1]) == 3072          # fused bra+ket absorption of the bulk sites, every slice of every sliced column
    # fused reverse pass (dT, dA, dB1 in one launch): 4 bulk sites x 6 single-contribution variants x 16 slices x 4 sliced columns
    assert int(rows["site2b"][1]) == 1536
    off = _plan("6", "4", "no_site2b=1")
    assert "site2b" not in off and abs(float(re.search(r"total model time [0-9.]+ s, flops ([0-9.e+]+)", off).group(1)) - flops) < 1e-6 * flops
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


def test_sweep_plan_layout_on_cpu():
    """The gate list of a TEBD sweep as ``tne_tebd_sweep`` receives it (SweepPlan, filled through a numpy view of the ctypes
    array): 0-based sites, shared-leg positions, bond indices, gate matrices M[o_i + 2 o_j, i_i + 2 i_j] and, for Vidal, the
    bond vector of every virtual leg -- checked field by field against the label bookkeeping of src/peps.jl:353-372."""
    code = r'''
import os, sys
os.environ["TNE_PLAN_ONLY"] = "1"
sys.path.insert(0, %r)
import numpy as np
import tne_b200 as T
from tne_b200 import peps as P, tebd, _lib
g = T.grid([4, 4])
for kind in ("simple", "vidal"):
    p = (T.zero_simplepeps if kind == "simple" else T.zero_vidalpeps)(g, 3)
    gates = [(s, d, np.reshape(T.yao.time_evolve(tebd.bond_hamiltonian(10.0, 0.2, 0.3, g.degree(s), g.degree(d)), 0.05), (2, 2, 2, 2), order="F"))
             for s, d in g.edges()]
    plan = P.SweepPlan(p, gates)
    assert len(plan.arr) == g.ne() == 24
    bi = {l: k for k, l in enumerate(p.virtual_labels)}
    for q, (i, j, mat) in enumerate(gates):
        a = plan.arr[q]
        li, lj = P.getvlabel(p, i), P.getvlabel(p, j)
        sh = [l for l in li if l in lj]
        assert len(sh) == 1
        assert (a.site_i, a.site_j, a.axis_i, a.axis_j, a.bond) == (i - 1, j - 1, li.index(sh[0]), lj.index(sh[0]), bi[sh[0]])
        m = np.asarray(list(a.gate)).view(np.complex128).reshape(4, 4, order="F")
        assert np.array_equal(m, mat.reshape(4, 4, order="F"))
        want_i = [-1] + ([bi[l] for l in li[1:]] if kind == "vidal" else [-1] * (len(li) - 1))
        assert list(a.lambda_i)[:len(li)] == want_i and all(x == -1 for x in list(a.lambda_i)[len(li):])
        want_j = [-1] + ([bi[l] for l in lj[1:]] if kind == "vidal" else [-1] * (len(lj) - 1))
        assert list(a.lambda_j)[:len(lj)] == want_j
    # the levelling the library applies (stated on the host by tebd.wavefronts): site-disjoint, order-preserving
    lv = tebd.wavefronts(g.edges())
    assert sum(len(b) for b in lv) == 24 and all(len({s for e in b for s in e}) == 2 * len(b) for b in lv)
print("SWEEP-PLAN-OK")
''' % ROOT
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "SWEEP-PLAN-OK" in res.stdout, res.stdout + res.stderr[-2000:]


def test_checkpoint_exchange_plan_for_8_ranks():
    """Option rs_comm (planner: its value is the world size to plan for): at 6x6 D=4 every forward checkpoint of a sliced column
    qualifies for a reduce-scatter (its consumers loop over its slowest legs) and every cotangent checkpoint for an all-gather."""
    code = r'''
import os, sys
os.environ["TNE_PLAN_ONLY"] = "1"
sys.path.insert(0, %r)
import numpy as np
import tne_b200 as T
from tne_b200 import peps as P, einsum as es
ctx = T.context(); ctx.set_option("debug", 2); ctx.set_option("rs_comm", 8)
L, D = 6, 4
g = T.grid([L, L]); order, segs = T.sweep_plan(g, L)
p = T.rand_simplepeps(g, D, np.random.default_rng(1), optimizer=order, segments=segs)
h = T.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
nt = L * L
cache = {}
terms = [P._term_replacements(p, t, cache) for t in h.blocks]
tl = [(c, [(nt + s - 1, t) for s, t in sorted(r.items())]) for c, r in terms]
leaves = [P._dev(t) for t in p.alltensors()] * 2
try:
    es.expect_terms_device(p.code_inner_product, leaves, [1] * nt + [0] * nt, tl, list(range(nt)), 1 + 0j)
except T.TneError:
    pass
''' % ROOT
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    rows = re.findall(r"checkpoints combined by all-reduce (\d+) .*?reduce-scatter (\d+) .*?all-gather (\d+)", res.stderr)
    assert len(rows) == 9, res.stderr[-2000:]
    ar, rs, ag = (sum(int(r[k]) for r in rows) for k in range(3))
    assert ar == 0 and rs == 32 and ag == 40


def test_two_bond_sweep_fits_the_8x8_d4_norm_on_cpu():
    """sweep_plan(slice_first_row=True) through the planner context: the exact 8x8 D=4 norm needs two 64 GiB boundaries +
    12 GiB (140 GiB: one B200) instead of 320 GiB, at +1.5 % flops; planned for 8 ranks every boundary is combined by a
    reduce-scatter (each rank only needs the slices it will read), never an all-reduce."""
    code = r'''
import os, sys
os.environ["TNE_PLAN_ONLY"] = "1"
sys.path.insert(0, %r)
import numpy as np
import tne_b200 as T
from tne_b200 import peps as P, einsum as es
ctx = T.context(); ctx.set_option("debug", 2); ctx.set_option("rs_comm", 8)
L, D = 8, 4
g = T.grid([L, L]); order, segs = T.sweep_plan(g, L, slice_first_row=True)
assert segs[0] == (1, [g.nv() + 1 + list(g.edges()).index((1, 9))]) and len(segs) == L and all(len(ls) == 2 for _, ls in segs[1:-1]) and len(segs[-1][1]) == 1
p = T.zero_simplepeps(g, D, optimizer=order, segments=segs)
p.code_inner_product.seg_shard = True
leaves = [P._dev(t) for t in p.alltensors()] * 2
try:
    es.contract_tree_device(p.code_inner_product, leaves, [1] * 64 + [0] * 64)
except T.TneError as e:
    print("REFUSED", e)
''' % ROOT
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "REFUSED" in res.stdout, res.stdout + res.stderr[-1500:]
    m = re.search(r"\[tne-plan\] (\d+) phases, (\d+) launches, ([0-9.e+]+) flops, workspace ([0-9.]+) GiB \(checkpoints ([0-9.]+) \+ segment ([0-9.]+)\)", res.stderr)
    assert m, res.stderr[-1500:]
    assert int(m.group(1)) == 8 and abs(float(m.group(3)) - 1.374e15) / 1.374e15 < 0.01
    assert float(m.group(4)) == 140.0 and float(m.group(5)) == 128.0
    comm = re.findall(r"checkpoints combined by all-reduce (\d+) \(([0-9.]+) GiB\), reduce-scatter (\d+) \(([0-9.]+) GiB\)", res.stderr)
    assert len(comm) == 7 and all(c[0] == "0" and c[2] == "1" and float(c[3]) == 64.0 for c in comm), comm


def test_tape_retention_plan_on_cpu():
    """plan_tape through the planner context (a 151 GiB budget stands in for a B200): one rank keeps the forward tapes of 12 of
    the 64 (column, slice) units of the 6x6 D=4 <H>+grad programme and skips their recompute launches; planned for 8 ranks
    every one of a rank's 8 units fits; tape_gib = 0 restores the all-recompute plan."""
    base = _plan("6", "4", "tape_gib=0", "--raw")
    auto = _plan("6", "4", "--raw")
    m0 = re.search(r"\[tne-plan\] (\d+) phases, (\d+) launches, ([0-9.e+]+) flops", base)
    m1 = re.search(r"\[tne-plan\] (\d+) phases, (\d+) launches, ([0-9.e+]+) flops", auto)
    assert m0 and m1, auto[-1500:]
    assert "tape retention" not in base
    t = re.search(r"tape retention: (\d+) of this rank's (\d+) \(segment, slice\) units keep their forward tape \(([0-9.]+) GiB\)", auto)
    assert t and int(t.group(1)) == 12 and int(t.group(2)) == 65 and 115.0 < float(t.group(3)) <= 121.0, auto[-1500:]
    assert float(m0.group(3)) == 7.8777e13 and 7.4e13 < float(m1.group(3)) < 7.5e13 and int(m1.group(2)) < int(m0.group(2))
    w8 = _plan("6", "4", "rs_comm=8", "--raw")
    t8 = re.search(r"tape retention: (\d+) of this rank's (\d+)", w8)
    assert t8 and t8.group(1) == t8.group(2) == "9", w8[-1500:]


def test_automatic_plans_are_accepted_and_their_slices_are_free():
    """optimizer="sweep" + segments="auto" on rectangles, a ladder, a triangular strip and a ring, through the planner context:
    the library accepts every plan, one segment per lattice column where the boundary has minima (none on the strip / ring:
    plain tree), and the programme executes exactly the flops of the unsliced tree -- the entering-bond slices repeat nothing."""
    code = r'''
import os, sys, re
os.environ["TNE_PLAN_ONLY"] = "1"
sys.path.insert(0, %r)
import tne_b200 as T
from tne_b200 import peps as P, einsum as es
from tne_b200._lib import TneError
ctx = T.context(); ctx.set_option("debug", 2)
def cost(p):
    sizes = {}
    for vl, t in zip(p.vertex_labels, p.vertex_tensors):
        for l, s_ in zip(vl, t.shape): sizes[l] = s_
    for k, l in enumerate(p.virtual_labels): sizes[p.max_index + k + 1] = sizes[l]
    return es.tree_cost(p.code_inner_product, sizes)[0]
n = 10
cases = [("3x4", T.grid([3, 4]), 4), ("4x6", T.grid([4, 6]), 6), ("2x7", T.grid([2, 7]), 7),
         ("strip", T.graph_from_edges(n, [(i, i + 1) for i in range(1, n)] + [(i, i + 2) for i in range(1, n - 1)]), 0),
         ("ring", T.graph_from_edges(8, [(i, i %% 8 + 1) for i in range(1, 9)]), 0)]
for name, g, nseg in cases:
    p = T.zero_simplepeps(g, 3, optimizer="sweep", segments="auto")
    assert len(p.code_inner_product.segments or []) == nseg, (name, p.code_inner_product.segments)
    leaves = [P._dev(t) for t in p.alltensors()] * 2
    try:
        es.contract_tree_device(p.code_inner_product, leaves, [1] * g.nv() + [0] * g.nv())
        print("COMPUTED?!", name)
    except TneError as e:
        print("CASE", name, "%%.6e" %% cost(p), "planner context" in str(e))
''' % ROOT
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    cases = [l.split() for l in res.stdout.splitlines() if l.startswith("CASE")]
    assert len(cases) == 5 and all(c[3] == "True" for c in cases), res.stdout + res.stderr[-1500:]
    planned = [float(m) for m in re.findall(r"\[tne-plan\] \d+ phases, \d+ launches, ([0-9.e+]+) flops", res.stderr)]
    assert len(planned) == 5
    for c, fl in zip(cases, planned):
        assert abs(float(c[2]) - fl) <= 2e-3 * fl, (c, fl)       # (the pair formation of the last sites is counted by the planner only)
