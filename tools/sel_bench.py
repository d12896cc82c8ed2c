"""Adjacency product + selection gather as two kernels vs the fused by-product (kgan_adjmix_fwd_sel): python tools/sel_bench.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import kgan_b200 as kgan
ops, G = kgan.ops, kgan.geometry
kgan.set_precision("tf32")
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
def adjacency(k, v, w, seed=0):
    rng = np.random.default_rng(seed); A = np.zeros((k, v, w), np.float32)
    for kk in range(k):
        for ww in range(w): A[kk, rng.integers(v), ww] = rng.standard_normal()
    return torch.from_numpy(A).cuda()
n = 8192
for c, t, v, w, keep in ((64, 64, 12, 5, [1, 4, 6, 9, 11]), (128, 32, 5, 5, [0, 1, 2, 3, 4]), (256, 16, 5, 1, [2])):
    x = torch.randn(n, c, t, v, device="cuda"); A = adjacency(3, v, w)
    sel = G.select_table(t, v, list(range(0, t, 2)), keep)
    a = timeit(lambda: ops.adjmix_fwd(x, A)); b = timeit(lambda: ops.plane_spmm(x, sel)); f = timeit(lambda: ops.adjmix_fwd(x, A, sel))
    print("%dx%dx%d -> w=%d: product %.1f us + gather %.1f us = %.1f us; fused %.1f us" % (c, t, v, w, a, b, a + b, f))
