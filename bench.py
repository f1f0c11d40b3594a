#!/usr/bin/env python
"""Benchmark of the hot path: one step = one <H> + gradient evaluation (``energy_and_gradient``: the value
of src/timeevolve.jl's ``iloss2`` and the TDVP force ``fvec``) of a 6x6, D=4 SimplePEPS with the Rydberg
Hamiltonian (96 terms), on synthetic random tensors.

  python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--L 6 --D 4]

Prints ONE JSON line (rank 0).  ``value`` = evals/s with the PEPS resident in HBM; ``e2e`` = the same through
the public API with host buffers (pinned host -> device upload of every tensor and device -> host read of the
energy and the force inside the timed region).  ``--impl reference`` times the CPU restatement of the
reference's own execution (numpy TTGT = OMEinsum's algorithm; Julia is not available on the box) on a bounded
sample and scales it to the same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "6x6 D=4 PEPS <H>+grad evals/sec"  # BASELINE.json metric (the defaults --L 6 --D 4)
C_, DELTA_, OMEGA_ = 10.0, 0.2, 0.3


def seed_for(L, D):
    return 20260000 + 100 * L + D


def workload_string(L, D):
    """The same ``config.workload`` in both arms (the driver compares them)."""
    n, K = L * L, 2 * L * (L - 1) + L * L
    bulk, edge, corner = max(L - 2, 0) ** 2, 4 * max(L - 2, 0), 4 if L > 1 else 1
    nvars = 2 * (bulk * D ** 4 + edge * D ** 3 + corner * D ** 2) if L > 2 else None
    return "%dx%d D=%d SimplePEPS, Rydberg H (%d terms), energy + full gradient (%s complex variables)" % (L, L, D, K, nvars if nvars else "?")


def sweep_tree_flops(L, D):
    """flops(tree) of SURVEY.md 8(d) for the boundary-sweep tree (bra then ket per site), from the host mirror of OMEinsum's
    cost accounting; one reference-style <H>+grad evaluation is 3 (K + 2) of these."""
    from oracle import einsum as OE, graphs as OG, peps as OP
    g = OG.grid([L, L])
    n = g.nv()
    order = []
    for i in range(n):
        order += [i, n + i]
    p = OP.zero_simplepeps(g, D, optimizer=order)
    sizes = {}
    for vl, t in zip(p.vertex_labels, p.vertex_tensors):
        for l, s_ in zip(vl, t.shape):
            sizes[l] = s_
    for k, l in enumerate(p.virtual_labels):
        sizes[p.max_index + k + 1] = sizes[l]
    return OE.tree_cost(p.code_inner_product, sizes)[0]


def peaks(ctx=None, device=0):
    """Roofline denominators: HBM from the driver-written MEASURED_PEAKS.json; the FP64 DMMA issue peak is measured HERE, in
    this process, by the library (tne_measure_fp64_peak) with its own clocks record."""
    out = {"hbm_gbs": 6650.0, "hbm_source": "fallback", "fp64_dmma_tflops": 37.18, "fp64_source": "profiles/fp64_peaks_r01.json"}
    try:
        m = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        out["hbm_gbs"] = float(m["hbm_gbs"]); out["hbm_source"] = "MEASURED_PEAKS.json"
    except Exception:
        pass
    try:
        f = json.load(open(os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")))
        out["fp64_dmma_tflops"] = float(f["dmma_tflops"])
    except Exception:
        pass
    if ctx is not None:
        try:
            clk = Clocks(device)
            best = {"dmma_tflops": 0.0, "dfma_tflops": 0.0}
            for _ in range(8):          # ~0.3 s under load so that the clocks sampler sees it
                m = ctx.measure_fp64_peak()
                best = {k: max(best[k], m[k]) for k in best}
            out["fp64_dmma_tflops"] = best["dmma_tflops"]
            out["fp64_dfma_tflops"] = best["dfma_tflops"]
            out["fp64_source"] = "measured in this run: mma.sync.m8n8k4.f64 issue peak (tne_measure_fp64_peak)"
            out["fp64_clocks"] = clk.stop()
        except Exception as e:      # an old library: keep the committed number
            out["fp64_error"] = str(e)
    return out


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's execution restated (oracle), bounded sample scaled to evals/s
# ---------------------------------------------------------------------------------------------
def cpu_sample(L, D, budget_s=15.0):
    """Times whole norm-tree contractions as the reference executes them (OMEinsum: permutedims -> reshape ->
    zgemm -> permutedims per binary node, numpy/OpenBLAS on all host cores) on a lattice small enough to finish
    in seconds (5x5 at the same D for the 6x6 workload: identical K/N step shapes, 16x shorter M), and scales
    by algorithmic flops to one <H>+grad evaluation of the L x L workload = 3 (K + 2) tree contractions (K
    energy terms + 2 norms; forward + 2 pullback contractions per node: src/peps.jl:470-476,
    src/TensorAD/rules.jl:1-8)."""
    from oracle import einsum as OE, graphs as OG, peps as OP, tebd as OT
    cores = os.cpu_count() or 1
    try:    # torchrun exports OMP_NUM_THREADS=1: pin the BLAS pool explicitly to every host core and say so
        from threadpoolctl import threadpool_limits, threadpool_info
        _limit = threadpool_limits(limits=cores)
        blas_threads = max([int(i.get("num_threads", 1)) for i in threadpool_info()] or [1])
    except Exception:
        blas_threads = None

    def build(Lx):
        g = OG.grid([Lx, Lx])
        n = g.nv()
        order = []
        for i in range(n):
            order += [i, n + i]
        p = OP.rand_simplepeps(g, D, np.random.default_rng(seed_for(Lx, D)), optimizer=order)
        sd = {l: 2 for l in p.physical_labels}
        for ix in p.code_inner_product.leaves() if hasattr(p.code_inner_product, "leaves") else []:
            pass
        sizes = {}
        for vl, t in zip(p.vertex_labels, p.vertex_tensors):
            for l, s_ in zip(vl, t.shape):
                sizes[l] = s_
        mx = p.max_index
        for k, l in enumerate(p.virtual_labels):
            sizes[mx + k + 1] = sizes[l]
        return g, p, OE.tree_cost(p.code_inner_product, sizes)[0]
    g, p, flops_full = build(L)
    K = len(OT.rydberg_hamiltonian(g, C_, DELTA_, OMEGA_).blocks)
    Ls = min(L, 5)
    _, ps, flops_s = build(Ls)
    xs = [np.conj(t) for t in ps.alltensors()] + list(ps.alltensors())
    t0 = time.perf_counter()
    reps = 0
    while True:
        ps.code_inner_product(*xs)
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= 50:
            break
    dt = time.perf_counter() - t0
    rate = reps * flops_s / dt
    eval_time = 3.0 * (K + 2) * flops_full / rate
    return {
        "value": 1.0 / eval_time, "unit": "evals/s", "cores": cores, "blas_threads": blas_threads, "kind": "port",
        "reference_style_flops_per_eval": 3.0 * (K + 2) * flops_full,
        "sample": "%d x one %dx%d D=%d norm tree (%.2e flop each) through the numpy TTGT restatement of OMEinsum in %.1f s = %.1f GFLOP/s; "
                  "scaled by flops to the %dx%d tree (%.2e flop) x 3*(K+2)=%d trees per eval, K=%d terms; Julia is not installed on the box"
                  % (reps, Ls, Ls, D, flops_s, dt, rate / 1e9, L, L, flops_full, 3 * (K + 2), K),
    }, eval_time


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    base = None
    for _ in range(args.warmup + args.steps):
        base, eval_time = cpu_sample(args.L, args.D, budget_s=args.cpu_seconds)
        vals.append(eval_time)
    vals = vals[args.warmup:] or vals
    ms = 1e3 * float(np.mean(vals))
    v = 1e3 / ms
    base["value"] = v
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "c128 (f64)", "data": "synthetic",
        "config": {"workload": workload_string(args.L, args.D),
                   "note": "CPU restatement of the reference's execution (numpy TTGT = OMEinsum's algorithm); each step is a bounded "
                           "sample (whole 5x5 D=4 norm trees) scaled by algorithmic flops to 3 (K + 2) trees of the 6x6 network"},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if sm:
            out["sm_mhz"] = float(np.median(sm)); out["sm_max_mhz"] = float(max(mx)); out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu(args):
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ.pop("NCCL_DEBUG")        # keep stdout to the one JSON line (NCCL prints its version banner there)
    import torch
    import tne_b200 as T
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = T.context()
    if os.environ.get("TNE_VARIANT"):
        ctx.set_option("absorb_tma", int(os.environ["TNE_VARIANT"]))
    if os.environ.get("TNE_NO_PDL"):
        ctx.set_option("no_pdl", 1)
    rs_comm = world > 1 and os.environ.get("TNE_RS_COMM", "1") == "1"
    if rs_comm:   # reduce-scatter / all-gather of the column checkpoints (half the NVLink bytes of the all-reduce; TNE_RS_COMM=0 disables)
        ctx.set_option("rs_comm", 1)
    if os.environ.get("TNE_COMM_OVERLAP") == "0":   # all-gathers of the reverse pass on the compute stream instead of behind the next column's recompute
        ctx.set_option("comm_overlap", 0)
    if world > 1:
        T.array.init_comm(ctx, dist, rank, world)
    L, D = args.L, args.D
    g = T.grid([L, L])
    order, segs = T.sweep_plan(g, L)
    rng = np.random.default_rng(seed_for(L, D))
    p = T.rand_simplepeps(g, D, rng, optimizer=order, segments=segs)
    # N > 1: the D^2 segment-local slices of every column are dealt to the ranks; the checkpoints of each column
    # and the outputs are summed with NCCL all-reduces inside the library (tne_tree.seg_shard)
    p.code_inner_product.seg_shard = world > 1
    h = T.rydberg_hamiltonian(g, C_, DELTA_, OMEGA_)
    K = len(h.blocks)
    nvars = sum(t.size for t in p.alltensors())
    shard = (0, 1)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def step_resident():
        return T.energy_and_gradient(p, h, shard=shard)

    # pinned staging buffers for the end-to-end arm
    host = []
    for t in p.alltensors():
        buf = torch.empty(t.size, dtype=torch.complex128).pin_memory()
        arr = buf.numpy()
        arr[:] = t.to_numpy().reshape(-1, order="F")
        host.append((buf, arr.reshape(t.shape, order="F")))
    out_f = torch.empty(nvars, dtype=torch.complex128).pin_memory()

    def step_e2e():
        q = T.peps.from_numpy_tensors(p, [a for _, a in host])          # H2D of every tensor
        y, f = T.energy_and_gradient(q, h, shard=shard)
        fv = f.to_numpy()                                                # D2H (synchronises)
        out_f.numpy()[:] = fv
        return y.to_numpy(), fv

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        e0.record(stream)
        for _ in range(steps):
            r = fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, ctx.launch_count() - l0, r

    for _ in range(args.warmup):
        step_resident()
    clocks = Clocks(local) if rank == 0 else None
    ms, launches, (y, f) = timed(step_resident, args.steps)
    clk = clocks.stop() if clocks else {}
    value = args.steps / (ms / 1e3)
    energy = complex(y.to_numpy())
    fmax = float(np.abs(f.to_numpy()).max())

    # executed / algorithmic work of one step and the dominant kernel, measured with CUDA events around
    # every launch of the same step (library option "profile"); done outside the headline timing
    ctx.set_option("profile", 1)
    ctx.profile_collect()
    barrier()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # per-programme statistics of the same step (norm programme + energy programme): algorithmic vs recomputed flops
    progs = []
    real_etd = T.einsum.expect_terms_device

    def spy(*a, **k):
        r = real_etd(*a, **k)
        progs.append(ctx.program_stats())
        return r
    T.einsum.expect_terms_device = spy
    pe0.record(stream)
    step_resident()
    pe1.record(stream)
    T.einsum.expect_terms_device = real_etd
    prof = ctx.profile_collect()
    prof_ms = pe0.elapsed_time(pe1)
    ctx.set_option("profile", 0)
    tot_flops = sum(v["flops"] for v in prof.values())
    recompute_flops = sum(q["recompute_flops"] for q in progs)
    pk = peaks(ctx, local)
    dom = max(prof, key=lambda k: prof[k]["ms"])
    dv = prof[dom]
    roof = {
        "bound": "hbm", "kernel": dom,
        "achieved": dv["bytes"] / (dv["ms"] * 1e-3) / 1e9 if dv["ms"] else None, "peak": pk["hbm_gbs"], "unit": "GB/s",
        "frac": (dv["bytes"] / (dv["ms"] * 1e-3) / 1e9 / pk["hbm_gbs"]) if dv["ms"] else None, "traffic": None,
        "peak_source": pk["hbm_source"], "launches_per_step": dv["launches"],
        "algorithmic_bytes_per_launch": dv["bytes"] / max(1, dv["launches"]),
        "avg_launch_ms": dv["ms"] / max(1, dv["launches"]),
        "share_of_step": dv["ms"] / prof_ms if prof_ms else None,
        "fp64_tflops": dv["flops"] / (dv["ms"] * 1e-3) / 1e12 if dv["ms"] else None,
        "fp64_peak_tflops": pk["fp64_dmma_tflops"], "fp64_peak_source": pk["fp64_source"],
        "fp64_frac": (dv["flops"] / (dv["ms"] * 1e-3) / 1e12 / pk["fp64_dmma_tflops"]) if dv["ms"] else None,
        "note": "AI of the absorb step = 8*K*N/(16*(K+N)) = 5.33 flop/B, just under the ridge 37.2 TF / 6.46 TB/s = 5.75: HBM-bound",
    }
    for tf in ("traffic_r02.json", "traffic_r01.json"):     # dram__bytes of one `ncu --set full` capture of the dominant kernel
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", tf)))
            if dom in tr:
                roof["traffic"] = tr[dom].get("dram_bytes_read_plus_write_per_launch")
                roof["traffic_note"] = dict(tr[dom], source="profiles/" + tf + " (ncu capture, not measured in this run)")
                break
        except Exception:
            pass

    # parity carried by the bench line itself (every N): oracle golden values of THIS configuration, contracted through the
    # same plan / sharding as the timed step (tests/golden/make_golden_big.py; tests/test_gpu_round2.py has the full test)
    check = {"energy_re": energy.real, "force_max_abs": fmax}
    try:
        gpath = os.path.join(ROOT, "tests", "golden", "grid_%dx%d_D%d_partial.npz" % (L, L, D))
        fpath = os.path.join(ROOT, "tests", "golden", "grid_%dx%d_D%d.npz" % (L, L, D))
        epath = os.path.join(ROOT, "tests", "golden", "grid_%dx%d_D%d_energy.npz" % (L, L, D))
        if os.path.exists(epath) and len(np.load(epath)["energy_terms"]) > (len(np.load(gpath)["energy_terms"]) if os.path.exists(gpath) else 0):
            gpath = epath      # <psi|h_k|psi> of (up to) ALL terms of the benchmark's Hamiltonian (make_golden_big.py energy66)
        if os.path.exists(gpath):
            gold = np.load(gpath)
            inner = complex(T.inner_product(p, p).to_numpy())
            check["rel_err_inner_vs_golden"] = abs(inner - complex(gold["inner"])) / abs(complex(gold["inner"]))
            nE = len(gold["energy_terms"])
            if nE:
                ks = [int(k) for k in gold["term_index"]][:nE]
                sub = T.yao.Add(*[h.blocks[k] for k in ks])
                e_sub = complex(T.expect(sub, p, p).to_numpy())
                check["rel_err_vs_golden"] = abs(e_sub - complex(np.sum(gold["energy_terms"]))) / float(np.sum(np.abs(gold["energy_terms"])))
                check["golden"] = "oracle values of <psi|psi> and of %d of the %d Hamiltonian terms (sum |e_k| scale), same plan and sharding as the timed step" % (nE, K)
        elif os.path.exists(fpath):
            gold = np.load(fpath)
            check["rel_err_vs_golden"] = abs(energy.real - float(gold["iloss2"])) / abs(float(gold["iloss2"]))
            check["force_rel_err_vs_golden"] = float(np.abs(f.to_numpy() - gold["fvec"]).max() / np.abs(gold["fvec"]).max())
            check["golden"] = "oracle iloss2 and fvec of this configuration (all terms)"
        dpath = os.path.join(ROOT, "tests", "golden", "device_%dx%d_D%d_n1.npz" % (L, L, D))
        if os.path.exists(dpath):      # the 1-GPU result of this library (itself checked against the oracle as above)
            dev1 = np.load(dpath)
            check["force_max_abs_diff_vs_n1"] = float(np.abs(f.to_numpy() - dev1["fvec"]).max())
            check["energy_rel_diff_vs_n1"] = abs(energy.real - float(dev1["iloss2"])) / abs(float(dev1["iloss2"]))
        elif world == 1 and rank == 0 and os.environ.get("TNE_WRITE_N1"):
            np.savez_compressed(dpath, fvec=f.to_numpy(), iloss2=np.asarray(energy.real))
    except Exception as e:
        check["golden_error"] = repr(e)
    if world > 1:
        # multi-GPU parity inside the driver's own N-GPU run (tests/test_gpu_multi.py needs a multi-GPU lease): slice-sharded
        # energy + force of the 4x4 D=3 golden configuration and the slice-sharded 10x10 D=2 exact norm (BASELINE configs[4]
        # at its validation size), both against committed oracle values, with the same collectives as the timed step
        try:
            mg = {}
            for (L2, D2, name) in ((4, 3, "grid_4x4_D3.npz"), (10, 2, "grid_10x10_D2_norm.npz")):
                gp = os.path.join(ROOT, "tests", "golden", name)
                if not os.path.exists(gp):
                    continue
                gold = np.load(gp)
                g2 = T.grid([L2, L2])
                o2, s2 = T.sweep_plan(g2, L2)
                q2 = T.rand_simplepeps(g2, D2, np.random.default_rng(seed_for(L2, D2)), optimizer=o2, segments=s2)
                q2.code_inner_product.seg_shard = True
                ip = complex(T.inner_product(q2, q2).to_numpy())
                mg["%dx%d_D%d_inner_rel_err" % (L2, L2, D2)] = abs(ip - complex(gold["inner"])) / abs(complex(gold["inner"]))
                if "fvec" in gold:
                    y2, f2 = T.energy_and_gradient(q2, T.rydberg_hamiltonian(g2, C_, DELTA_, OMEGA_))
                    mg["%dx%d_D%d_iloss2_rel_err" % (L2, L2, D2)] = abs(float(np.real(y2.to_numpy())) - float(gold["iloss2"])) / abs(float(gold["iloss2"]))
                    mg["%dx%d_D%d_force_rel_err" % (L2, L2, D2)] = float(np.abs(f2.to_numpy() - gold["fvec"]).max() / np.abs(gold["fvec"]).max())
            check["multi_gpu_vs_oracle_golden"] = mg
        except Exception as e:
            check["multi_gpu_error"] = repr(e)

    # end-to-end arm
    for _ in range(min(args.warmup, 1)):
        step_e2e()
    ems, _, _ = timed(step_e2e, args.steps)
    e2e_value = args.steps / (ems / 1e3)

    # secondary measurements of the same path (not part of the metric): the metric-vector product of the TDVP
    # step (src/timeevolve.jl:15-52) on the same PEPS and one TEBD sweep of an 8x8, D=4 PEPS (BASELINE configs 3, 4)
    extras = {}
    if not args.no_extras:
        # one TDVP step (src/timeevolve.jl:1-4) at a FIXED Krylov size on N GPUs: fvec + `tdvp_matvecs` metric-vector products
        # (each one tne_expect_terms programme + reverse pass over the slice-sharded tree) + the GMRES algebra on the host
        import scipy.sparse.linalg as sla
        nmv = args.tdvp_matvecs
        vv = T.variables(p).to_numpy()
        nn = vv.size
        memo2, mv_ms = {}, []

        def matvec(xr):
            xc = xr[:nn] + 1j * xr[nn:]
            t0 = time.perf_counter()
            sx = T.apply_smatrix(p, T.DiffTensor(vv.copy()), T.DiffTensor(vv.copy()), T.DiffTensor(xc, requires_grad=False), memo=memo2).to_numpy()
            mv_ms.append(1e3 * (time.perf_counter() - t0))
            return np.concatenate([sx.real, sx.imag])

        def tdvp_step():
            b = -1j * T.fvec(p, h).to_numpy()
            op = sla.LinearOperator((2 * nn, 2 * nn), matvec=matvec, dtype=np.float64)
            sol, _ = sla.gmres(op, np.concatenate([b.real, b.imag]), x0=np.concatenate([vv.real, vv.imag]), rtol=1e-30, restart=nmv - 1, maxiter=1)
            return sol
        tms, _, _ = timed(tdvp_step, 1)
        extras["tdvp_step"] = {"ms": tms, "matvecs": len(mv_ms), "ms_per_matvec_host_clock": float(np.median(mv_ms[1:])) if len(mv_ms) > 1 else None,
                               "ms_first_matvec_host_clock": mv_ms[0] if mv_ms else None, "n_gpus": world,
                               "note": "sr_step of the same %dx%d D=%d PEPS cut at %d S.x products (GMRES restart %d, one cycle): fvec + matvecs, "
                                       "x-independent contractions memoised after the first product; slices dealt to the ranks like the headline step" % (L, L, D, nmv, nmv - 1)}
    if world == 1 and not args.no_extras:
        v = T.variables(p).to_numpy()
        xr = np.random.default_rng(seed_for(L, D) + 1)
        xx = (xr.standard_normal(v.size) + 1j * xr.standard_normal(v.size)) / np.sqrt(2.0)
        sfun = lambda: T.apply_smatrix(p, T.DiffTensor(v.copy()), T.DiffTensor(v.copy()), T.DiffTensor(xx, requires_grad=False))
        sfun()
        sms, _, sx = timed(sfun, 1)
        memo = {}
        mfun = lambda: T.apply_smatrix(p, T.DiffTensor(v.copy()), T.DiffTensor(v.copy()), T.DiffTensor(xx, requires_grad=False), memo=memo)
        mfun()
        mms, _, mx = timed(mfun, 1)
        extras["apply_smatrix"] = {"ms": sms, "ms_inside_gmres": mms, "max_abs": float(np.abs(sx.to_numpy()).max()),
                                   "memo_max_abs_diff": float(np.abs(mx.to_numpy() - sx.to_numpy()).max()),
                                   "note": "S.x of the same %dx%d D=%d PEPS, tree-level algorithm (one tne_expect_terms + reverse pass); "
                                           "ms_inside_gmres: with sr_step's memo of the x-independent contractions" % (L, L, D)}
        Lt = 8
        gt = T.grid([Lt, Lt])
        pt = T.rand_simplepeps(gt, 4, np.random.default_rng(seed_for(Lt, 4)))
        sweep = lambda: T.rydberg_tebd_sweep_(pt, gt, C_, DELTA_, OMEGA_, 0.03)
        sweep(); sweep()
        nsw = 5
        tms, tl, _ = timed(lambda: [sweep() for _ in range(nsw)], 1)
        tms /= nsw
        multi = lambda: T.rydberg_tebd_(pt, gt, t=0.03 * 11, C=C_, delta=DELTA_, omega=OMEGA_, nstep=11)
        multi()
        mms10, _, _ = timed(multi, 1)
        extras["tebd_10_sweeps_one_call_8x8_D4"] = {"ms_per_sweep": mms10 / 10, "note": "rydberg_tebd! with nstep = 11: its 10 sweeps enqueued back to back by one tne_tebd_sweeps call, ranks verified once"}
        extras["tebd_sweep_8x8_D4"] = {"ms": tms, "bonds": gt.ne(), "bonds_per_s": gt.ne() / (tms * 1e-3), "wavefronts": len(T.tebd.wavefronts(gt.edges())),
                                       "note": "one rydberg_tebd_sweep! (mean of %d): 112 bond gates in one tne_tebd_sweep call (28 wavefront launches, no host "
                                               "sync in between), QR-reduced Jacobi SVD truncation to D=4, latency-bound" % nsw}
        # the largest exact D=4 norm contraction of a square lattice that fits one B200 (8x8 D=4 needs a 64 GiB boundary
        # tensor per variant: see DESIGN.md section 10)
        Ln = 7
        gn = T.grid([Ln, Ln])
        on, sn = T.sweep_plan(gn, Ln)
        pn = T.rand_simplepeps(gn, 4, np.random.default_rng(seed_for(Ln, 4)), optimizer=on, segments=sn)
        T.norm(pn)
        nms, _, nv = timed(lambda: T.norm(pn), 1)
        st = ctx.program_stats()
        extras["norm_7x7_D4"] = {"ms": nms, "flops": st["flops"], "tflops": st["flops"] / (nms * 1e-3) / 1e12, "log10_norm": float(np.log10(abs(nv.item()))),
                                 "note": "exact <psi|psi> of a 7x7 D=4 PEPS (one contraction, forward only)"}
    if not args.no_extras:
        # TEBD does not shard inside one PEPS (replicas only): R independent 8x8 D=4 PEPS per GPU go through ONE
        # tne_tebd_sweep call per sweep (the R copies of a gate share a wavefront launch); N GPUs run N x R replicas, no communication
        try:
            gt2 = T.grid([8, 8])
            rep = {}
            for Rr in (8, 32):
                reps = [T.rand_simplepeps(gt2, 4, np.random.default_rng(seed_for(8, 4) + 1000 * rank + r)) for r in range(Rr)]
                nsw = 5     # rydberg_tebd! with nstep = 6: its 5 sweeps go out in ONE tne_tebd_sweeps call (no sync between the sweeps)
                bs = lambda: T.rydberg_tebd_batch_(reps, gt2, t=0.03 * (nsw + 1), C=C_, delta=DELTA_, omega=OMEGA_, nstep=nsw + 1)
                bs(); bs()       # the device pool reaches its steady state after two calls
                rms, _, _ = timed(lambda: [bs() for _ in range(2)], 1)
                rms /= 2 * nsw
                rep["R%d" % Rr] = {"ms_per_batched_sweep": rms, "peps_sweeps_per_s_all_gpus": world * Rr / (rms * 1e-3),
                                   "bond_updates_per_s_all_gpus": world * Rr * gt2.ne() / (rms * 1e-3)}
                del reps
            rep["note"] = ("replica-batched rydberg_tebd! of 8x8 D=4 PEPS: R replicas per GPU, 5 sweeps per library call (tne_tebd_sweeps: 5 x 28 wavefront launches "
                           "of R x 4 CTAs back to back), %d GPU(s) x R replicas, max over ranks; a single PEPS: see tebd_sweep_8x8_D4" % world)
            extras["tebd_replicas_8x8_D4"] = rep
        except Exception as e:
            extras["tebd_replicas_error"] = repr(e)
    if not args.no_extras and (world > 1 or os.environ.get("TNE_BENCH_BIG") == "1"):
        # BASELINE configs[3] / configs[4] on N GPUs: exact norms of big lattices with the two-bond sliced sweep
        # (sweep_plan(slice_first_row=True)): every column loops over the bond entering its last site and the bond leaving
        # its first site, the D^4 (column, slice) units are dealt to the ranks, the 64 GiB boundaries are combined with
        # ncclReduceScatter (each rank only needs the slices it will read); one un-warmed run each (they take seconds)
        def big_norm(Lb, Db):
            gb = T.grid([Lb, Lb])
            ob, sb = T.sweep_plan(gb, Lb, slice_first_row=True)
            pb = T.rand_simplepeps(gb, Db, np.random.default_rng(seed_for(Lb, Db)), optimizer=ob, segments=sb)
            pb.code_inner_product.seg_shard = world > 1
            bms, _, val = timed(lambda: T.inner_product(pb, pb), 1)
            st = ctx.program_stats()
            return pb, {"ms": bms, "n_gpus": world, "flops_per_gpu": st["flops"], "tflops_per_gpu": st["flops"] / (bms * 1e-3) / 1e12,
                        "workspace_gib_per_gpu": st["workspace_bytes"] / 2 ** 30, "log10_inner": float(np.log10(abs(complex(val.to_numpy()))))}
        try:
            for (Lb, Db, name) in ((8, 3, "grid_8x8_D3_norm.npz"), (10, 2, "grid_10x10_D2_norm.npz")):   # parity of this very path first
                gp = os.path.join(ROOT, "tests", "golden", name)
                if os.path.exists(gp):
                    gold = np.load(gp)
                    pb, rec = big_norm(Lb, Db)
                    ip = complex(T.inner_product(pb, pb).to_numpy())
                    check["two_bond_sweep_%dx%d_D%d_inner_rel_err_vs_oracle" % (Lb, Lb, Db)] = abs(ip - complex(gold["inner"])) / abs(complex(gold["inner"]))
            _, rec = big_norm(8, 4)
            rec["note"] = ("exact <psi|psi> of an 8x8 D=4 PEPS (BASELINE configs[3]): 1.37e15 flop, two 64 GiB boundaries + 12 GiB per GPU; "
                           "one B200 alone: 72.6 s (tools/norm_big.py)")
            extras["norm_8x8_D4"] = rec
            if world >= 4 or os.environ.get("TNE_BENCH_BIG") == "1":
                _, rec = big_norm(10, 3)
                rec["note"] = "exact <psi|psi> of a 10x10 D=3 PEPS (BASELINE configs[4] at the largest bond dimension whose 52 GiB boundary fits): 6.2e14 flop; one B200 alone: 82.5 s"
                extras["norm_10x10_D3"] = rec
            # configs[4] at D=4 cannot be contracted exactly (16 TiB boundary): a fixed subset of a globally sliced sum, timed
            Lb, Db, rows, per_rank = 10, 4, [4, 5, 6], 1
            gb = T.grid([Lb, Lb])
            ob, sb = T.sweep_plan(gb, Lb, slice_first_row=True)
            pb = T.rand_simplepeps(gb, Db, np.random.default_rng(seed_for(Lb, Db)), optimizer=ob, segments=sb)
            nb = Lb * Lb
            emap = {e: nb + 1 + k for k, e in enumerate(gb.edges())}
            bonds = [emap[(r + (yy - 2) * Lb, r + (yy - 1) * Lb)] for r in rows for yy in range(2, Lb + 1)]
            srng = np.random.default_rng(7)
            subset = [(tuple(srng.integers(0, Db, len(bonds))), tuple(srng.integers(0, Db, len(bonds)))) for _ in range(per_rank * world)]
            part = []

            def slice_job():
                sv, cnt = T.inner_product_sliced(pb, pb, bonds, subset, shard=(rank, world))
                acc = T.B200Array.from_numpy(np.asarray(sv))
                if world > 1:
                    T.array.allreduce_sum(ctx, [acc])      # NCCL all-reduce of the slice sums
                part.append(complex(acc.to_numpy()))
            sms, _, _ = timed(slice_job, 1)
            st = ctx.program_stats()
            nsl = float(Db) ** (2 * len(bonds))
            extras["slices_10x10_D4"] = {"ms": sms, "slices": len(subset), "slices_per_s": len(subset) / (sms * 1e-3), "flops_per_slice": st["flops"],
                                         "tflops_per_gpu": st["flops"] * per_rank / (sms * 1e-3) / 1e12, "total_slices": nsl,
                                         "extrapolated_seconds_all_slices": nsl / (len(subset) / (sms * 1e-3)), "unsliced_flops": 6.13e17,
                                         "unsliced_seconds_at_this_rate_if_memory_allowed": 6.13e17 / (world * st["flops"] * per_rank / (sms * 1e-3)),
                                         "partial_sum": [part[-1].real, part[-1].imag],
                                         "note": "10x10 D=4 (BASELINE configs[4]): the 27 horizontal bonds of lattice rows 4-6 sliced globally, bra and ket copies "
                                                 "independently (16^27 terms, each a contraction with dimension-1 legs, 8.75 GiB, 1.1e14 flop); one term per GPU timed, "
                                                 "partial sums all-reduced with NCCL; the exact contraction needs a 16 TiB boundary and is out of reach of 8 x 180 GB "
                                                 "by any slicing -- the extrapolation is printed, not claimed"}
        except Exception as e:
            extras["big_lattice_error"] = repr(e)
    if rank == 0:
        cpu, _ = cpu_sample(L, D, budget_s=args.cpu_seconds) if world == 1 and not args.no_cpu else (None, None)
        line = {
            "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "c128 (f64)", "data": "synthetic",
            "config": {"workload": workload_string(L, D),
                       "plan": "boundary sweep, one checkpointed segment per column, D^2 segment-local slices; forward tapes of as many (column, slice) units as fit in "
                               "the remaining HBM are kept for the reverse pass, the others recomputed (recompute_flops_per_step_per_gpu)",
                       "parallelism": "1 GPU" if world == 1 else ("segment-local bond slices dealt to %d GPUs, NCCL %s of column checkpoints (one group per column%s), all-reduce of the outputs"
                                                                       % (world, "reduce-scatter / all-gather" if rs_comm else "all-reduce",
                                                                          ", all-gathers on a second stream behind the next column's recompute" if rs_comm and os.environ.get("TNE_COMM_OVERLAP") != "0" else "")),
                       "l2": "no flush needed: a step streams a %.0f GiB workspace (>> 126 MB L2)" % (ctx.mem_info()["workspace"] / 2 ** 30),
                       "seed": seed_for(L, D)},
            "executed_tflops_per_gpu": tot_flops / (prof_ms * 1e-3) / 1e12 if prof_ms else None,
            "executed_flops_per_step_per_gpu": tot_flops,
            "fp64_peak_frac": (tot_flops / (prof_ms * 1e-3) / 1e12 / pk["fp64_dmma_tflops"]) if prof_ms else None,
            "recompute_flops_per_step_per_gpu": recompute_flops,
            "fp64_frac_excluding_recompute": ((tot_flops - recompute_flops) / (prof_ms * 1e-3) / 1e12 / pk["fp64_dmma_tflops"]) if prof_ms else None,
            "fp64_peak": {k: pk.get(k) for k in ("fp64_dmma_tflops", "fp64_dfma_tflops", "fp64_source", "fp64_clocks", "fp64_error") if k in pk},
            "reference_style_flops_per_eval": 3.0 * (K + 2) * sweep_tree_flops(L, D),
            "algorithmic_speedup_vs_reference_style": 3.0 * (K + 2) * sweep_tree_flops(L, D) / max(1.0, (tot_flops - recompute_flops) * world),
            "roofline": roof,
            "kernels": {k: v for k, v in prof.items() if v["launches"]},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": 16 * nvars, "d2h_bytes_per_step": 16 * nvars + 16,
                    "ms_per_step": ems / args.steps},
            "gpu_launches": launches,
            "clocks": clk,
            "check": check,
            "extras": extras,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--L", type=int, default=6)
    ap.add_argument("--D", type=int, default=4)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--tdvp-matvecs", type=int, default=3)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
