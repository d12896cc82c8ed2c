"""Times of the adjacency kernels on the largest shapes of the training step at per-GPU batch 4096 (critic passes over 8192 samples):
python tools/adjmix_bench.py"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import kgan_b200 as kgan  # noqa: E402

ops, G = kgan.ops, kgan.geometry
kgan.set_precision("tf32")


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def adjacency(k, v, w, nnz_per_col=1, seed=0):
    rng = np.random.default_rng(seed)
    A = np.zeros((k, v, w), np.float32)
    for kk in range(k):
        for ww in range(w):
            for vv in rng.choice(v, size=min(v, nnz_per_col), replace=False):
                A[kk, vv, ww] = rng.standard_normal()
    return torch.from_numpy(A).cuda()


n = 8192
total = 0.0
for name, c, t, v, w in (("D1 fwd", 32, 64, 12, 12), ("D2 fwd", 64, 64, 12, 5), ("D3 fwd", 128, 32, 5, 5), ("D4 fwd", 256, 16, 5, 1)):
    x = torch.randn(n, c, t, v, device="cuda")
    A = adjacency(3, v, w)
    us = timeit(lambda: ops.adjmix_fwd(x, A))
    gb = (x.numel() + n * 3 * c * t * w) * 4 / 1e9
    total += us
    print("%-8s fwd   %dx%dx%dx%d -> w=%d: %7.1f us  %5.0f GB/s" % (name, n, c, t, v, w, us, gb / us * 1e6))
    g = torch.randn(n, 3 * c, t, w, device="cuda")
    src = torch.randn(n, c, t, v, device="cuda")
    add = torch.randn(n, c, t, v, device="cuda")
    us = timeit(lambda: ops.adjmix_bwd_x(g, A, add, src))
    gb = (g.numel() + 3 * x.numel()) * 4 / 1e9
    total += us
    print("%-8s bwd_x (+add, mask) %dx%dx%dx%d:      %7.1f us  %5.0f GB/s" % (name, n, 3 * c, t, w, us, gb / us * 1e6))
    us = timeit(lambda: ops.adjmix_bwd_a(x, g, 3, (A != 0).float()))
    total += us
    print("%-8s bwd_a                               %7.1f us  %5.0f GB/s" % (name, us, (g.numel() + x.numel()) * 4 / 1e9 / us * 1e6))
print("sum %.1f us" % total)
