"""TEBD driver on the device: ``rydberg_tebd!`` of src/tebd.jl with the same keyword arguments.
Each edge's 4x4 gate ``exp(-i dt h_ij)`` is built on the host (it depends only on C, delta, omega
and the two vertex degrees) and applied by the fused device bond update."""
import numpy as np

from . import yao
from . import peps as P
from .graphs import SimpleGraph


def cterm(C):  # src/tebd.jl:3-5
    return C * yao.kron_mat(yao.P1, yao.P1)


def zterm(delta):  # src/tebd.jl:7-9 (Z, not n -- as in the reference)
    return delta * yao.Z


def xterm(omega):  # src/tebd.jl:11-13
    return omega / 2 * yao.X


def bond_hamiltonian(C, delta, omega, di, dj):
    """``hi`` of src/tebd.jl:19 as a 4x4 matrix (first site = least-significant bit)."""
    I = yao.I2
    return (cterm(C) + yao.kron_mat(xterm(omega / di), I) + yao.kron_mat(I, xterm(omega / dj))
            + yao.kron_mat(zterm(delta / di), I) + yao.kron_mat(I, zterm(delta / dj)))


def wavefronts(edges):
    """Group the gates of one sweep into batches of site-disjoint bonds without changing the order of any two
    gates that share a site: gates on disjoint sites commute exactly (truncation included), so the batched
    sweep gives the same tensors as the reference's strictly sequential loop (src/tebd.jl:16-21)."""
    level_of_site, levels = {}, []
    for e in edges:
        lv = 1 + max(level_of_site.get(e[0], -1), level_of_site.get(e[1], -1))
        level_of_site[e[0]] = level_of_site[e[1]] = lv
        while len(levels) <= lv:
            levels.append([])
        levels[lv].append(e)
    return levels


def rydberg_tebd_sweep_(reg, g, C, delta, omega, dt):  # src/tebd.jl:15-23
    n = reg.n if isinstance(reg, yao.ArrayReg) else P.nsite(reg)
    cache = {}   # the gate depends on the two vertex degrees only: a handful of distinct 4x4 matrices per sweep

    def gate(src, dst):
        key = (g.degree(src), g.degree(dst))
        if key not in cache:
            cache[key] = yao.time_evolve(bond_hamiltonian(C, delta, omega, key[0], key[1]), dt)
        return cache[key]
    if isinstance(reg, yao.ArrayReg):
        for (src, dst) in g.edges():
            reg.apply_(yao.put(n, (src, dst), gate(src, dst)))
        return reg
    # the whole sweep in one library call (tne_tebd_sweep): wavefronts of site-disjoint gates, one launch each, no host
    # synchronisation in between; `wavefronts` above states the same levelling for the host-side tests.  The label
    # bookkeeping of the gate list (a SweepPlan) depends on the PEPS structure and the parameters only: it is kept on the
    # PEPS object and reused by the following sweeps of rydberg_tebd!
    key = (float(C), float(delta), float(omega), float(dt), tuple(g.edges()))
    cached = getattr(reg, "_sweep_plan", None)
    if cached is None or cached[0] != key:
        plan = P.SweepPlan(reg, [(src, dst, np.reshape(gate(src, dst), (2, 2, 2, 2), order="F")) for src, dst in g.edges()])
        cached = (key, plan)
        reg._sweep_plan = cached
    P.apply_sweep_(reg, cached[1])
    return reg


def _sweep_plan_for(reg, g, C, delta, omega, dt):
    cache = {}

    def gate(src, dst):
        key = (g.degree(src), g.degree(dst))
        if key not in cache:
            cache[key] = yao.time_evolve(bond_hamiltonian(C, delta, omega, key[0], key[1]), dt)
        return cache[key]
    key = (float(C), float(delta), float(omega), float(dt), tuple(g.edges()))
    cached = getattr(reg, "_sweep_plan", None)
    if cached is None or cached[0] != key:
        plan = P.SweepPlan(reg, [(src, dst, np.reshape(gate(src, dst), (2, 2, 2, 2), order="F")) for src, dst in g.edges()])
        cached = (key, plan)
        reg._sweep_plan = cached
    return cached[1]


def rydberg_tebd_sweep_batch_(regs, g, C, delta, omega, dt):
    """``rydberg_tebd_sweep!`` (src/tebd.jl:15-23) of R replica PEPS with one structure in ONE library call: the R copies of
    every gate run in the same wavefront launch (``peps.apply_sweep_batch_``).  Replica r ends up with exactly the tensors
    ``rydberg_tebd_sweep_(regs[r], ...)`` gives."""
    regs = list(regs)
    P.apply_sweep_batch_(regs, _sweep_plan_for(regs[0], g, C, delta, omega, dt))
    return regs


def rydberg_tebd_batch_(regs, g, *, t, C, delta, omega, nstep):
    """``rydberg_tebd!`` (src/tebd.jl:25-38, nstep-1 sweeps) for R replicas."""
    dt = t / nstep
    regs = list(regs)
    P.apply_sweep_batch_(regs, _sweep_plan_for(regs[0], g, C, delta, omega, dt), n_sweeps=nstep - 1)
    return regs


def rydberg_tebd_(reg, g=None, *, t, C, delta, omega, nstep):  # src/tebd.jl:25-38
    if g is None:
        g = SimpleGraph(P.nsite(reg))
        for i, j in P.virtualbonds(reg):
            g.add_edge(i, j)
    dt = t / nstep
    if isinstance(reg, yao.ArrayReg):
        for _ in range(nstep - 1):  # sic: nstep-1 sweeps (src/tebd.jl:27)
            rydberg_tebd_sweep_(reg, g, C, delta, omega, dt)
        return reg
    # device PEPS: all nstep - 1 sweeps in ONE library call (tne_tebd_sweeps: no host round trip between the sweeps)
    P.apply_sweep_(reg, _sweep_plan_for(reg, g, C, delta, omega, dt), n_sweeps=nstep - 1)
    return reg


def rydberg_hamiltonian(g, C, delta, omega):
    """Synthetic benchmark Hamiltonian: ``C P1_i P1_j`` per edge as ``kron`` plus
    ``(omega/2) X + delta Z`` per site as a one-site ``put`` -- the terms of src/tebd.jl:3-13."""
    n = g.nv()
    blocks = []
    for (i, j) in g.edges():
        blocks.append(yao.kron(n, (i, C * yao.P1), (j, yao.P1)))
    for i in g.vertices():
        blocks.append(yao.put(n, i, xterm(omega) + zterm(delta)))
    return yao.Add(*blocks)
