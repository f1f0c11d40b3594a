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


def _product_worker(rank, world, port, outdir):
    """PRODUCT host code under the planner context (TNE_PLAN_ONLY=1: everything is planned, nothing computed)."""
    os.environ["TNE_PLAN_ONLY"] = "1"
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    log = os.path.join(outdir, "plan_rank%d.log" % rank)
    fd = os.open(log, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    os.dup2(fd, 2)                                    # the planner dumps to the C stderr
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tne_b200 as T
    from tne_b200 import einsum as es, peps as P
    ctx = T.context()
    L, D = 4, 2
    g = T.grid([L, L])
    order, segs = T.sweep_plan(g, L)
    p = T.rand_simplepeps(g, D, np.random.default_rng(7), optimizer=order, segments=segs)
    h = T.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    # (1) Hamiltonian-term sharding of peps.expect(..., shard=(rank, world)): record what each rank would hand to the device
    seen = []
    real = es.expect_terms_device

    def recorder(code, xs, conj, terms, grad_leaves=(), seed=1.0 + 0j, shard=(0, 1)):
        seen.append([(complex(c), tuple((leaf, t.handle) for leaf, t in repl)) for c, repl in terms])
        return T.B200Array.from_numpy(np.zeros(())), []
    es.expect_terms_device = recorder
    try:
        T.expect(h, p, p, shard=(rank, world))
        mine = [(c, tuple(leaf for leaf, _ in repl)) for c, repl in seen[-1]]
        T.expect(h, p, p, shard=(0, 1))
        full = [(c, tuple(leaf for leaf, _ in repl)) for c, repl in seen[-1]]
        # a rank with no terms must not evaluate the bare overlap (tne_expect_terms with n_terms == 0 would)
        n0 = len(seen)
        z = T.expect(T.yao.Add(*h.blocks[:1]), p, p, shard=(1, 2))
        empty_ok = len(seen) == n0 and z.shape == ()
    finally:
        es.expect_terms_device = real
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # (2) segment-local slices dealt to the ranks: every rank must plan the SAME collectives in the same order
    p.code_inner_product.seg_shard = True
    ctx.set_option("debug", 2)
    ctx.set_option("rs_comm", world)
    leaves = [P._dev(t) for t in p.alltensors()] * 2
    n = len(p.alltensors())
    try:
        es.expect_terms_device(p.code_inner_product, leaves, [1] * n + [0] * n, [], list(range(n)), 1.0 + 0j)
        refused = False
    except T.TneError:
        refused = True
    os.fsync(2)
    plan = [l for l in open(log).read().splitlines() if l.startswith("[tne-plan]")]
    plans = [None] * world
    dist.all_gather_object(plans, plan)
    if rank == 0:
        import json
        json.dump(dict(gathered=[[list(map(str, t)) for t in r] for r in gathered], full=[list(map(str, t)) for t in full],
                       empty_ok=empty_ok, refused=refused, plans=plans), open(os.path.join(outdir, "res.json"), "w"))
    dist.barrier()
    dist.destroy_process_group()


def test_product_sharding_logic_under_the_planner_context(tmp_path):
    """World 2 over gloo, product code only: peps.expect deals the Hamiltonian terms k % world == rank into disjoint lists
    whose union is the full list; an empty share yields an explicit zero; the planner gives both ranks identical
    collective schedules for the slice-sharded programme (reduce-scatter / all-gather classification of checkpoints)."""
    import json
    port = 29950 + (os.getpid() % 40)
    mp.spawn(_product_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = json.load(open(tmp_path / "res.json"))
    full = [tuple(t) for t in res["full"]]
    parts = [[tuple(t) for t in r] for r in res["gathered"]]
    assert sorted(parts[0] + parts[1]) == sorted(full) and len(full) == 40
    assert parts[0] == full[0::2] and parts[1] == full[1::2]
    assert res["empty_ok"] and res["refused"]
    assert res["plans"][0] == res["plans"][1] and any("reduce-scatter" in l for l in res["plans"][0])
