"""HBM throughput of the tensor permutation kernel (tne_permute.cu) against the measured copy peak: python tools/permute_time.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import tne_b200 as T
ctx = T.context()
stream = torch.cuda.ExternalStream(ctx.stream())
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
cases = [("2^26 elements, 13 legs of 4, reverse order", (4,) * 13, list(range(13, 0, -1))),
         ("boundary-like: move two fast legs to the back", (4,) * 13, [3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 1, 2]),
         ("matrix transpose 8192 x 8192", (8192, 8192), [2, 1]),
         ("swap two middle legs (aligned fast leg)", (4,) * 13, [1, 2, 3, 8, 5, 6, 7, 4, 9, 10, 11, 12, 13]),
         ("D = 3: 3^15 elements, rotate by 7", (3,) * 15, [8, 9, 10, 11, 12, 13, 14, 15, 1, 2, 3, 4, 5, 6, 7])]
for name, shape, perm in cases:
    n = int(np.prod(shape))
    x = T.array.zeros((n,)).reshape(shape)
    labels = list(range(1, len(shape) + 1))
    for opt in (0, 1):
        ctx.set_option("no_permute", opt)
        T.einsum.einsum([labels], [x], perm); ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record(stream)
        for _ in range(reps):
            T.einsum.einsum([labels], [x], perm)
        e1.record(stream); ctx.sync()
        ms = e0.elapsed_time(e1) / reps
        gbs = 32.0 * n / (ms * 1e-3) / 1e9
        print("%-48s %-22s %8.3f ms  %7.1f GB/s  %.2f of the measured copy peak" % (name, "gett_generic" if opt else "permute_kernel", ms, gbs, gbs / peak))
    ctx.set_option("no_permute", 0)
