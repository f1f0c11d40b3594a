"""Times the tiled DMMA GEMM (tne_gemm.cu) on the shapes the general-tree path produces: python tools/gemm_time.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import tne_b200 as T
ctx = T.context()
rng = np.random.default_rng(0)
def crand(*s): return rng.standard_normal(s) + 1j * rng.standard_normal(s)
cases = [("paired tree M=2^16 N=K=256", [[1, 2], [2, 3]], [(65536, 256), (256, 256)], [1, 3]),
         ("paired tree, K fastest in A", [[2, 1], [2, 3]], [(256, 65536), (256, 256)], [1, 3]),
         ("square 4096^3", [[1, 2], [2, 3]], [(4096, 4096), (4096, 4096)], [1, 3]),
         ("batched 64 x (512 x 512 x 512)", [[1, 2, 4], [2, 3, 4]], [(512, 512, 64), (512, 512, 64)], [1, 3, 4]),
         ("N = 16: M=2^18 K=64", [[1, 2], [2, 3]], [(262144, 64), (64, 16)], [1, 3])]
import torch
for L, D, opt_name in ((5, 4, "greedy"), (6, 4, "paired")):
    g = T.grid([L, L]); n = L * L
    opt = "greedy" if opt_name == "greedy" else [(i, n + i) for i in range(n)]
    p = T.rand_simplepeps(g, D, np.random.default_rng(1), optimizer=opt)
    T.inner_product(p, p); ctx.sync()
    ctx.set_option("profile", 1); ctx.profile_collect()
    t0 = time.time(); v = T.inner_product(p, p).to_numpy(); dt = time.time() - t0
    prof = ctx.profile_collect(); ctx.set_option("profile", 0)
    st = ctx.program_stats()
    print("%dx%d D=%d %s tree <psi|psi>: %.1f ms, %.3e flop, %.2f TFLOP/s overall" % (L, L, D, opt_name, dt * 1e3, st["flops"], st["flops"] / dt / 1e12))
    tot = sum(v["ms"] for v in prof.values())
    for k, v in prof.items():
        if v["launches"]:
            print("    %-14s %5d launches %9.2f ms (%.1f %%) %7.2f TFLOP/s" % (k, v["launches"], v["ms"], 100 * v["ms"] / tot, v["flops"] / max(v["ms"], 1e-9) / 1e9))
    del p
for name, ixs, shapes, iy in cases:
    xs = [T.B200Array.from_numpy(crand(*s)) for s in shapes]
    for opt in (0, 3, 2, 1):   # 0: tiled DMMA GEMM (16-warp big tile), 3: 8-warp big tile, 2: small tiles only, 1: FMA kernel; 0: tiled DMMA GEMM (big tiles where they fill the chip), 2: small tiles only, 1: FMA kernel
        ctx.set_option("no_gemm", opt)
        T.einsum.einsum(ixs, xs, iy); ctx.sync()
        ctx.set_option("profile", 1); ctx.profile_collect()
        for _ in range(3):
            T.einsum.einsum(ixs, xs, iy)
        prof = ctx.profile_collect(); ctx.set_option("profile", 0)
        for k, v in prof.items():
            if v["launches"]:
                print("%-34s %-14s %8.3f ms/launch %7.2f TFLOP/s %8.1f GB/s" % (name, k, v["ms"] / v["launches"], v["flops"] / v["ms"] / 1e9, v["bytes"] / v["ms"] / 1e6))
    ctx.set_option("no_gemm", 0)
