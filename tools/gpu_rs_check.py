"""Multi-GPU check of the reduce-scatter / all-gather checkpoint exchange (option rs_comm; run under torchrun, one rank per GPU):
slice-sharded energy + gradient of 4x4 D=2 (4 slices per column) and 4x4 D=4 with a two-term H (16 slices per column) against the
numpy oracle, with rs_comm off (all-reduce) and on; also compares the two modes bit for bit on rank 0."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import torch.distributed as dist
import tne_b200 as T
from oracle import graphs as OG, peps as OP, tebd as OT, timeevolve as OTE, yao as OY
from helpers import *

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = T.context()
T.array.init_comm(ctx, dist, rank, world)
errs = []
for (L, D, full) in [(4, 2, True), (4, 4, False)]:
    g = OG.grid([L, L]); gd = T.grid([L, L])
    order, segs = T.sweep_plan(gd, L)
    op = oracle_peps(g, D, seed_for(L, D), optimizer=sweep_order(L * L))
    if full:
        h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    else:
        h = OY.Add(OY.kron(L * L, (6, 10.0 * OY.P1), (7, OY.P1)), OY.put(L * L, 11, OT.xterm(0.3) + OT.zterm(0.2)))
    hd = to_device_hamiltonian(T, h)
    fref = OTE.fvec(op, h).data; yref = OTE.iloss2(h, op, OP.variables(op)); nref = OP.norm(op)
    res = {}
    for mode in (0, 1):
        ctx.set_option("rs_comm", mode)
        dp = device_peps_like(T, op, g, D, optimizer=order, segments=segs)
        dp.code_inner_product.seg_shard = True
        t0 = time.time()
        y, f = T.energy_and_gradient(dp, hd)
        n = T.norm(dp).to_numpy()
        res[mode] = (y.to_numpy(), f.to_numpy(), n, time.time() - t0)
        errs += [rel(res[mode][1], fref), rel(res[mode][0], yref), rel(n, nref)]
    ctx.set_option("rs_comm", 0)
    if rank == 0:
        print("RS CHECK %dx%d D=%d world %d: all-reduce err %.2e, rs/ag err %.2e, modes differ by %.2e" % (
            L, L, D, world, max(rel(res[0][1], fref), rel(res[0][0], yref)), max(rel(res[1][1], fref), rel(res[1][0], yref)),
            rel(res[1][1], res[0][1])), flush=True)
t = torch.tensor([max(errs)], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("RS CHECK world %d: max rel err %.2e" % (world, t.item()))
    assert t.item() < 1e-10
dist.barrier()
dist.destroy_process_group()
