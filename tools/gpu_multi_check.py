"""Multi-GPU correctness check (run under torchrun, one rank per GPU): 4x4 D=3 energy + gradient with the
segment-local slices dealt to the ranks (tne_tree.seg_shard) and, separately, with Hamiltonian terms sharded,
against the numpy oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import torch.distributed as dist
import tne_b200 as T
from oracle import graphs as OG, peps as OP, tebd as OT, timeevolve as OTE
from helpers import *

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = T.context()
T.array.init_comm(ctx, dist, rank, world)
L, D = 4, 3
g = OG.grid([L, L]); gd = T.grid([L, L])
order, segs = T.sweep_plan(gd, L)
op = oracle_peps(g, D, seed_for(L, D), optimizer=sweep_order(L * L))
h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3); hd = to_device_hamiltonian(T, h)
fref = OTE.fvec(op, h).data; yref = OTE.iloss2(h, op, OP.variables(op))
# 1. slice sharding inside the library
dp = device_peps_like(T, op, g, D, optimizer=order, segments=segs)
dp.code_inner_product.seg_shard = True
y, f = T.energy_and_gradient(dp, hd)
e1 = max(rel(f.to_numpy(), fref), rel(y.to_numpy(), yref))
n1 = rel(T.norm(dp).to_numpy(), OP.norm(op))
# 2. term sharding + allreduce of the partial results
dq = device_peps_like(T, op, g, D, optimizer=order, segments=segs)
y2, f2 = T.energy_and_gradient(dq, hd, shard=(rank, world))
T.array.allreduce_sum(ctx, [y2.data, f2.data])
e2 = max(rel(f2.to_numpy(), fref), rel(y2.to_numpy(), yref))
# 3. BASELINE configs[4] at the validation size: 10x10, D=2 exact norm, bond slices dealt to the ranks
L2, D2 = 10, 2
g2 = OG.grid([L2, L2])
order2, segs2 = T.sweep_plan(T.grid([L2, L2]), L2)
op2 = oracle_peps(g2, D2, seed_for(L2, D2), optimizer=sweep_order(L2 * L2))
dp2 = device_peps_like(T, op2, g2, D2, optimizer=order2, segments=segs2)
dp2.code_inner_product.seg_shard = True
n10 = rel(T.norm(dp2).to_numpy(), OP.norm(op2))
t = torch.tensor([e1, n1, e2, n10], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("MULTI-GPU CHECK world %d: slice-sharded energy+grad rel err %.2e, norm %.2e; term-sharded %.2e; 10x10 D=2 sliced norm %.2e" % (world, *t.tolist()))
    assert t.max().item() < 1e-10
dist.barrier()
dist.destroy_process_group()
