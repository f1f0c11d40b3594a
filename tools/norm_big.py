"""Exact <psi|psi> of big square-lattice PEPS with the two-bond sliced sweep (sweep_plan(slice_first_row=True)):
golden checks at the validation sizes, then a timed run.  One rank or under torchrun (slices dealt to the ranks).

    python tools/norm_big.py check                 # 4x4 D=3, 5x5 D=4, 8x8 D=2, 8x8 D=3, 10x10 D=2 against the oracle goldens
    python tools/norm_big.py time 8 4 [reps]       # BASELINE configs[3]: exact 8x8 D=4 norm (64 GiB boundary)
    python tools/norm_big.py slices 10 4 4,5,6 2   # BASELINE configs[4]: a fixed subset of the globally sliced 10x10 D=4 sum
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import tne_b200 as T
from helpers import seed_for

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
ctx = T.context()
for kv in os.environ.get("TNE_OPTS", "").split(","):
    if "=" in kv:
        ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
if world > 1:
    T.array.init_comm(ctx, dist, rank, world)


def big_peps(L, D):
    g = T.grid([L, L])
    if os.environ.get("TNE_AUTO_PLAN") == "1":     # the same plan derived from the network alone (einsum.sweep_order / sweep_segments)
        p = T.rand_simplepeps(g, D, np.random.default_rng(seed_for(L, D)), optimizer="sweep", segments="auto2")
    else:
        order, segs = T.sweep_plan(g, L, slice_first_row=True)
        p = T.rand_simplepeps(g, D, np.random.default_rng(seed_for(L, D)), optimizer=order, segments=segs)
    if world > 1:
        p.code_inner_product.seg_shard = True
    return p


def say(*a):
    if rank == 0:
        print(*a, flush=True)


what = sys.argv[1]
if what == "check":
    worst = 0.0
    for L, D, name in ((4, 3, "grid_4x4_D3.npz"), (5, 4, "grid_5x5_D4.npz"), (8, 2, "grid_8x8_D2_norm.npz"), (8, 3, "grid_8x8_D3_norm.npz"),
                       (10, 2, "grid_10x10_D2_norm.npz")):
        path = os.path.join(ROOT, "tests", "golden", name)
        if not os.path.exists(path):
            say("missing", name)
            continue
        gold = np.load(path)
        p = big_peps(L, D)
        assert np.array_equal(T.variables(p).to_numpy(), gold["variables"])
        t0 = time.time()
        ip = complex(T.inner_product(p, p).to_numpy())
        err = abs(ip - complex(gold["inner"])) / abs(complex(gold["inner"]))
        worst = max(worst, err)
        st = ctx.program_stats()
        say("%dx%d D=%d  <psi|psi> rel err vs oracle %.2e  (%.3e flop, workspace %.2f GiB, %.2f s, world %d)" % (L, L, D, err, st["flops"], st["workspace_bytes"] / 2 ** 30, time.time() - t0, world))
    assert worst < 1e-10, worst
    say("NORM CHECK OK")
elif what == "slices":
    # 10x10 D=4: the horizontal bonds of the given lattice rows are sliced globally (bra and ket copies independently);
    # every term is a contraction with dimension-1 legs that fits one GPU; a fixed subset is dealt to the ranks and timed
    L, D = int(sys.argv[2]), int(sys.argv[3])
    rows = [int(r) for r in sys.argv[4].split(",")]
    per_rank = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    g = T.grid([L, L])
    order, segs = T.sweep_plan(g, L, slice_first_row=True)
    p = T.rand_simplepeps(g, D, np.random.default_rng(seed_for(L, D)), optimizer=order, segments=segs)
    n = L * L
    emap = {e: n + 1 + k for k, e in enumerate(g.edges())}
    bonds = [emap[(r + (y - 2) * L, r + (y - 1) * L)] for r in rows for y in range(2, L + 1)]
    rng = np.random.default_rng(7)
    subset = [(tuple(rng.integers(0, D, len(bonds))), tuple(rng.integers(0, D, len(bonds)))) for _ in range(per_rank * world)]
    T.inner_product_sliced(p, p, bonds, subset[:world], shard=(rank, world))      # warm-up (workspace, kernels)
    ctx.sync()
    if dist is not None:
        dist.barrier()
    t0 = time.time()
    part, cnt = T.inner_product_sliced(p, p, bonds, subset, shard=(rank, world))
    st = ctx.program_stats()
    acc = T.B200Array.from_numpy(np.asarray(part))
    if world > 1:
        T.array.allreduce_sum(ctx, [acc])           # NCCL all-reduce of the slice sums
    total = complex(acc.to_numpy())
    ctx.sync()
    if dist is not None:
        dist.barrier()
    dt = time.time() - t0
    nsl_total = float(D) ** (2 * len(bonds))
    rate = len(subset) / dt
    say(json.dumps({"workload": "%dx%d D=%d, horizontal bonds of rows %s sliced globally: %d of %.3e slice terms on %d GPU(s)" % (L, L, D, rows, len(subset), nsl_total, world),
                    "seconds": dt, "slices_per_s": rate, "flops_per_slice": st["flops"], "tflops_per_gpu": st["flops"] * cnt / dt / 1e12,
                    "workspace_gib": st["workspace_bytes"] / 2 ** 30, "partial_sum": [total.real, total.imag],
                    "extrapolated_seconds_all_slices": nsl_total / rate,
                    "unsliced_flops": 6.13e17, "slicing_overhead_flops": nsl_total * st["flops"] / 6.13e17}))
else:
    L, D = int(sys.argv[2]), int(sys.argv[3])
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    p = big_peps(L, D)
    out = []
    for r in range(reps + 1):
        ctx.sync()
        if dist is not None:
            dist.barrier()
        t0 = time.time()
        v = T.inner_product(p, p)
        ctx.sync()
        if dist is not None:
            dist.barrier()
        dt = time.time() - t0
        st = ctx.program_stats()
        out.append(dt)
        say("%dx%d D=%d exact norm: run %d  %.3f s  %.3e flop on this rank -> %.2f TFLOP/s per GPU, log10 <psi|psi> = %.12f, workspace %.1f GiB"
            % (L, L, D, r, dt, st["flops"], st["flops"] / dt / 1e12, np.log10(abs(complex(v.to_numpy()))), st["workspace_bytes"] / 2 ** 30))
    say(json.dumps({"workload": "%dx%d D=%d exact norm" % (L, L, D), "world": world, "seconds": out, "flops_per_rank": st["flops"]}))
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
