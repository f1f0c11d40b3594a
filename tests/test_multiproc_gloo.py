"""The N > 1 host logic on CPU (gloo, world_size 2): work items are dealt round-robin to the ranks exactly as the
device library does (index % world == rank: Hamiltonian terms in the Python host, segment-local slices inside
tne_tree.cu), every rank's partial energy / gradient is summed with an all-reduce, and the sum equals the
single-process result.  The arithmetic here is the oracle's (no GPU in this container); the device
counterpart is tools/gpu_multi_check.py / tests/test_gpu_multi.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import graphs as OG, peps as OP, tebd as OT, yao as OY, tensorad as AD, timeevolve as OTE
    from helpers import oracle_peps, seed_for, sweep_order
    L, D = 3, 2
    g = OG.grid([L, L])
    p = oracle_peps(g, D, seed_for(L, D), optimizer=sweep_order(L * L))
    h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    # term sharding: rank r keeps terms k with k % world == r (tne_b200.peps.expect(..., shard=(r, world)))
    mine = OY.Add(*[b for k, b in enumerate(h.blocks) if k % world == rank])
    v = OP.variables(p)

    def partial_loss(x):  # iloss2 with this rank's terms; the normalisation chain is replicated
        return OTE.iloss2(mine, OTE.wrap_difftensor(p), x)
    gpart = AD.gradient(partial_loss, AD.DiffTensor(v.copy()))[0].data
    ypart = OTE.iloss2(mine, p, v)
    t = torch.tensor(np.concatenate([[ypart.real, 0.0], gpart.real, gpart.imag]))
    dist.all_reduce(t)          # the role of tne_allreduce_sum / ncclAllReduce(sum)
    # slice sharding of a closed network: partial sums over the values of a sliced bond label
    q = oracle_peps(g, D, seed_for(L, D), optimizer=sweep_order(L * L), nslices=2)
    code = q.code_inner_product
    xs = [np.conj(x) for x in q.alltensors()] + list(q.alltensors())
    sd = {}
    for ix, x in zip(code.ixs, xs):
        for l, s in zip(ix, x.shape):
            sd[l] = s
    import itertools
    part = 0.0 + 0.0j
    for it, vals in enumerate(itertools.product(*[range(sd[l]) for l in code.slicing])):
        if it % world != rank:
            continue
        m = dict(zip(code.slicing, vals))
        sl = [x[tuple(slice(m[l], m[l] + 1) if l in m else slice(None) for l in ix)] for ix, x in zip(code.ixs, xs)]
        part += complex(code.eins(*sl))
    s = torch.tensor([part.real, part.imag], dtype=torch.float64)
    dist.all_reduce(s)
    if rank == 0:
        full_y = OTE.iloss2(h, p, v)
        full_g = 1j * OTE.fvec(p, h).data
        n = gpart.size
        got_g = t[2:2 + n].numpy() + 1j * t[2 + n:].numpy()
        np.save(out, np.array([abs(t[0].item() - full_y.real) / abs(full_y.real),
                               np.abs(got_g - full_g).max() / np.abs(full_g).max(),
                               abs(complex(s[0].item(), s[1].item()) - complex(OP.inner_product(p, p))) / abs(complex(OP.inner_product(p, p)))]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_partials_allreduce_to_the_full_result(tmp_path):
    out = str(tmp_path / "res.npy")
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    err = np.load(out)
    assert err.max() < 1e-10, err


def test_wavefronts_and_shard_deal_are_partitions():
    import tne_b200
    g = tne_b200.grid([8, 8])
    lv = tne_b200.tebd.wavefronts(g.edges())
    assert sum(len(b) for b in lv) == g.ne() == 112
    assert all(len({s for e in b for s in e}) == 2 * len(b) for b in lv)      # site-disjoint inside a wavefront
    order = {e: (i, k) for i, b in enumerate(lv) for k, e in enumerate(b)}
    for a_i, a in enumerate(g.edges()):                                         # order of site-sharing gates is kept
        for b in g.edges()[a_i + 1:]:
            if set(a) & set(b):
                assert order[a][0] < order[b][0]
    for world in (2, 3, 8):
        items = list(range(16))
        deal = [[i for i in items if i % world == r] for r in range(world)]
        assert sorted(sum(deal, [])) == items
