"""numpy restatement of ``src/tebd.jl``.  TEST INFRASTRUCTURE ONLY (see oracle/__init__)."""
import numpy as np

from . import yao
from . import peps as P
from .graphs import SimpleGraph


def cterm(C):  # tebd.jl:3-5
    return C * yao.kron_mat(yao.P1, yao.P1)


def zterm(delta):  # tebd.jl:7-9  (Z, not n -- reference quirk)
    return delta * yao.Z


def xterm(omega):  # tebd.jl:11-13
    return omega / 2 * yao.X


def bond_hamiltonian(C, delta, omega, di, dj):
    """``hi`` of tebd.jl:19 as a 4x4 matrix (first site = least-significant bit)."""
    I = yao.I2
    return (cterm(C) + yao.kron_mat(xterm(omega / di), I) + yao.kron_mat(I, xterm(omega / dj))
            + yao.kron_mat(zterm(delta / di), I) + yao.kron_mat(I, zterm(delta / dj)))


def rydberg_tebd_sweep_(reg, g, C, delta, omega, dt):  # tebd.jl:15-23
    n = reg.n if isinstance(reg, yao.ArrayReg) else P.nsite(reg)
    for (src, dst) in g.edges():
        hi = bond_hamiltonian(C, delta, omega, g.degree(src), g.degree(dst))
        block = yao.put(n, (src, dst), yao.time_evolve(hi, dt))
        if isinstance(reg, yao.ArrayReg):
            reg.apply_(block)
        else:
            P.apply_block_(reg, block)
    return reg


def rydberg_tebd_(reg, g=None, *, t, C, delta, omega, nstep):  # tebd.jl:25-38
    if g is None:
        g = SimpleGraph(P.nsite(reg))
        for i, j in P.virtualbonds(reg):
            g.add_edge(i, j)
    dt = t / nstep
    for _ in range(nstep - 1):  # sic: nstep-1 sweeps (tebd.jl:27)
        rydberg_tebd_sweep_(reg, g, C, delta, omega, dt)
    return reg


def rydberg_hamiltonian(g, C, delta, omega):
    """The synthetic Hamiltonian of SURVEY.md 8(d): ``C P1_i P1_j`` per edge (as ``kron``) plus
    ``(omega/2) X + delta Z`` per site (one-site ``put``), built from tebd.jl:3-13's terms."""
    n = g.nv()
    blocks = []
    for (i, j) in g.edges():
        blocks.append(yao.kron(n, (i, C * yao.P1), (j, yao.P1)))
    for i in g.vertices():
        blocks.append(yao.put(n, i, xterm(omega) + zterm(delta)))
    return yao.Add(*blocks)
