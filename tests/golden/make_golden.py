"""Regenerates tests/golden/*.npz from the oracle (run from the repo root: python tests/golden/make_golden.py).

The reference (Julia) cannot run in this environment and stores no golden vectors of its own, so these fixtures are
outputs of the ORACLE, which is itself pinned on the reference's tests against dense linear algebra
(tests/test_oracle_*.py).  They freeze the numbers the device path is compared with, so that a silent change of the
oracle (or of the synthetic-input rule) shows up as a test failure."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import graphs as OG, peps as OP, tebd as OT, timeevolve as OTE, tensorad as AD  # noqa: E402
from helpers import oracle_peps, seed_for, sweep_order  # noqa: E402


def make(L, D):
    g = OG.grid([L, L])
    p = oracle_peps(g, D, seed_for(L, D), optimizer=sweep_order(L * L))
    h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    v = OP.variables(p)
    out = dict(variables=v, norm=np.asarray(OP.norm(p)), inner=np.asarray(OP.inner_product(p, p)),
               energy=np.asarray(OP.expect(h, p, p)), iloss2=np.asarray(OTE.iloss2(h, p, v)), fvec=OTE.fvec(p, h).data)
    x = np.random.default_rng(seed_for(L, D) + 1)
    x = (x.standard_normal(v.size) + 1j * x.standard_normal(v.size)) / np.sqrt(2.0)
    out["x"] = x
    out["smatrix_x"] = OTE.apply_smatrix(p, AD.DiffTensor(v.copy()), AD.DiffTensor(v.copy()), AD.DiffTensor(x, requires_grad=False)).data
    if L <= 3:
        q = oracle_peps(g, D, seed_for(L, D))
        OT.rydberg_tebd_(q, g, t=0.3, C=10.0, delta=0.2, omega=0.3, nstep=10)
        out["tebd_state"] = OP.vec(q)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "grid_%dx%d_D%d.npz" % (L, L, D)), **out)


if __name__ == "__main__":
    make(3, 2)
    make(4, 3)
