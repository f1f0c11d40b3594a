"""The product's host-side Yao / Graphs mirrors against LITERAL matrices and hand-computed values -- not against their
oracle twins (oracle/yao.py, oracle/graphs.py share most lines with them, so comparing the two proves nothing)."""
import numpy as np

import tne_b200 as T
from tne_b200 import yao as Y, tebd as TB


def test_kron_mat_is_little_endian():
    # first factor acts on the least-significant bit: M[o1 + 2 o2, i1 + 2 i2] = A[o1, i1] B[o2, i2]
    want = np.array([[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 0, -1], [0, 0, -1, 0]], dtype=complex)   # X (x) Z, written out
    assert np.array_equal(Y.kron_mat(Y.X, Y.Z), want)


def test_control_matrix_is_cnot_with_control_on_bit_0():
    want = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=complex)
    assert np.array_equal(Y.control(2, 1, 2, Y.X).two_site_matrix(), want)


def test_time_evolve_hand_computed():
    dt = 0.37
    # exp(-i dt X) = cos(dt) I - i sin(dt) X
    assert np.allclose(Y.time_evolve(Y.X, dt), np.array([[np.cos(dt), -1j * np.sin(dt)], [-1j * np.sin(dt), np.cos(dt)]]), atol=1e-15)
    # a diagonal 4 x 4 generator exponentiates entry by entry
    d = np.array([0.5, -1.25, 2.0, 3.5])
    assert np.allclose(Y.time_evolve(np.diag(d), dt), np.diag(np.exp(-1j * dt * d)), atol=1e-15)
    # 2 x 2 block: h = a Z + b X has eigenvalues +-r, exp(-i dt h) = cos(r dt) I - i sin(r dt) h / r
    a, b = 0.2, 0.15
    h = a * Y.Z + b * Y.X
    r = np.hypot(a, b)
    assert np.allclose(Y.time_evolve(h, dt), np.cos(r * dt) * np.eye(2) - 1j * np.sin(r * dt) * h / r, atol=1e-15)


def test_bond_hamiltonian_entry_by_entry():
    """src/tebd.jl:17-19 written out element by element: basis |q1 q2> with q1 the least-significant bit."""
    C, delta, omega, di, dj = 10.0, 0.2, 0.3, 3, 4
    want = np.zeros((4, 4), dtype=complex)
    for q1 in range(2):
        for q2 in range(2):
            k = q1 + 2 * q2
            want[k, k] = C * q1 * q2 + delta / di * (1 - 2 * q1) + delta / dj * (1 - 2 * q2)   # P1 P1, Z = diag(1, -1)
            want[(1 - q1) + 2 * q2, k] += omega / di / 2                                          # (Omega/d_i)/2 X on site 1
            want[q1 + 2 * (1 - q2), k] += omega / dj / 2
    assert np.allclose(TB.bond_hamiltonian(C, delta, omega, di, dj), want, atol=1e-15)


def test_grid_edges_and_neighbours_literal():
    g = T.grid([2, 3])            # Graphs.grid([2, 3]): vertex v = x + 2 (y - 1)
    assert g.edges() == [(1, 2), (1, 3), (2, 4), (3, 4), (3, 5), (4, 6), (5, 6)]
    assert g.neighbors(3) == [1, 4, 5] and g.degree(4) == 3 and g.nv() == 6 and g.ne() == 7


def test_array_reg_known_answers():
    """test/peps.jl:16-19: X on site 1 sets basis index 2 (1-based), CNOT(1 -> 2) then gives index 4."""
    psi = np.zeros(32, dtype=complex)
    psi[0] = 1
    reg = Y.ArrayReg(psi)
    reg.apply_(Y.put(5, 1, Y.X))
    assert np.argmax(np.abs(reg.state)) == 1
    reg.apply_(Y.control(5, 1, 2, Y.X))
    assert np.argmax(np.abs(reg.state)) == 3 and abs(reg.state[3] - 1) < 1e-15


def test_rydberg_hamiltonian_dense_on_a_path():
    """3-site path 1-2-3: <q|H|q> and the one-flip elements written out by hand (C P1 P1 per edge, Omega/2 X + Delta Z per site)."""
    g = T.graph_from_edges(3, [(1, 2), (2, 3)])
    C, delta, omega = 10.0, 0.2, 0.3
    H = Y.mat(TB.rydberg_hamiltonian(g, C, delta, omega))
    want = np.zeros((8, 8), dtype=complex)
    for k in range(8):
        q = [(k >> s) & 1 for s in range(3)]
        want[k, k] = C * (q[0] * q[1] + q[1] * q[2]) + delta * sum(1 - 2 * x for x in q)
        for s in range(3):
            want[k ^ (1 << s), k] += omega / 2
    assert np.allclose(H, want, atol=1e-14)
