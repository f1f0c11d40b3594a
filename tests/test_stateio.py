"""State I/O: the on-disk layout is the reference's flat ``variables`` vector (src/peps.jl:202-211)
plus label metadata.  Host-side format checks run on CPU; the device round trip is a GPU test."""
import os

import numpy as np
import pytest

from oracle import graphs as OG, peps as OP, tebd as OT
from helpers import device_peps_like, oracle_peps, rel, seed_for


def _stateio():
    import tne_b200
    return tne_b200.stateio


@pytest.mark.parametrize("kind", ["simple", "vidal"])
def test_pack_layout_is_the_variables_vector(kind, tmp_path):
    sio = _stateio()
    g = OG.test_graph()
    op = oracle_peps(g, 3, 7, kind)
    d = sio.pack_state(kind, op.vertex_labels, op.virtual_labels, op.alltensors(), op.Dmax, op.eps, time=0.25, step=4)
    # identical to vcat(vec.(alltensors(peps))...) of the oracle (src/peps.jl:202)
    assert np.array_equal(d["variables"], OP.variables(op))
    path = os.path.join(tmp_path, "state.npz")
    sio.write_state(path, d)
    meta, tensors = sio.unpack_state(sio.read_state(path))
    assert meta["kind"] == kind and meta["Dmax"] == op.Dmax and meta["eps"] == op.eps and meta["nflavor"] == 2
    assert meta["vertex_labels"] == [list(l) for l in op.vertex_labels] and meta["virtual_labels"] == list(op.virtual_labels)
    assert meta["time"] == 0.25 and meta["step"] == 4
    assert len(tensors) == len(op.alltensors())
    for a, b in zip(tensors, op.alltensors()):
        assert a.shape == np.shape(b) and np.array_equal(a, b)       # bit-exact


def test_ragged_and_empty_inputs_and_errors(tmp_path):
    sio = _stateio()
    # a single-vertex "PEPS" without bonds and ragged ranks on a path graph
    d = sio.pack_state("simple", [[1]], [], [np.array([1.0, 2.0j])], 1, 1e-12)
    meta, tensors = sio.unpack_state(d)
    assert meta["vertex_labels"] == [[1]] and tensors[0].shape == (2,)
    labels = [[1, 4], [2, 4, 5], [3, 5]]
    ts = [np.ones((2, 3)), np.ones((2, 3, 2)), np.ones((2, 2))]
    d = sio.pack_state("simple", labels, [4, 5], ts, 3, 1e-12)
    assert d["variables"].size == 6 + 12 + 4
    with pytest.raises(ValueError):
        sio.pack_state("simple", labels, [4, 5], ts[:2], 3, 1e-12)              # missing tensor
    with pytest.raises(ValueError):
        sio.pack_state("simple", labels, [4, 5], [ts[0], ts[1], np.ones(4)], 3, 1e-12)   # rank mismatch
    with pytest.raises(ValueError):
        sio.pack_state("vidal", labels, [4, 5], ts + [np.ones((2, 2)), np.ones(2)], 3, 1e-12)   # bond tensor not a vector
    bad = dict(d); bad["variables"] = d["variables"][:-1]
    with pytest.raises(ValueError):
        sio.unpack_state(bad)
    bad = dict(d); bad["format"] = np.array("something else")
    with pytest.raises(ValueError):
        sio.unpack_state(bad)
    bad = dict(d); bad["version"] = np.array(99)
    with pytest.raises(ValueError):
        sio.unpack_state(bad)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["simple", "vidal"])
def test_device_checkpoint_round_trip(tne, kind, tmp_path):
    """TEBD on the device, checkpoint, reload: bit-identical variables, same state vector as the oracle's evolution."""
    g = OG.test_graph()
    kw = dict(Dmax=6) if kind == "simple" else dict(Dmax=6, eps=1e-10)
    op = oracle_peps(g, 2, 43, kind, **kw)
    dp = device_peps_like(tne, op, g, 2, **kw)
    args = dict(t=0.3, C=10.0, delta=0.2, omega=0.3, nstep=4)
    gd = tne.graph_from_edges(5, g.edges())
    tne.rydberg_tebd_(dp, gd, **args)
    OT.rydberg_tebd_(op, g, **args)
    path = os.path.join(tmp_path, "ckpt.npz")
    tne.save_peps(path, dp, time=0.3, step=3)
    dq, meta = tne.load_peps(path)
    assert meta["time"] == 0.3 and meta["step"] == 3 and dq.kind == kind
    assert np.array_equal(tne.variables(dq).to_numpy(), tne.variables(dp).to_numpy())
    assert [t.shape for t in dq.alltensors()] == [t.shape for t in dp.alltensors()]
    assert rel(tne.vec(dq), OP.vec(op)) < 1e-8
    assert rel(tne.norm(dq).to_numpy(), OP.norm(op)) < 1e-8
    # the file is readable without a device, in the oracle's own layout
    m2, tensors = tne.stateio.unpack_state(tne.stateio.read_state(path))
    assert np.array_equal(np.concatenate([t.reshape(-1, order="F") for t in tensors]), tne.variables(dp).to_numpy())
