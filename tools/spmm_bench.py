"""Times of the plane gather / scatter kernel on the shapes of the training step (critic passes over 8192 samples): python tools/spmm_bench.py"""
import sys
from pathlib import Path

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


n = 8192
tot = 0.0
cases = []
for c, t, v in ((128, 64, 5), (256, 32, 5), (512, 16, 1)):
    unf = G.UnfoldedTcnGeom(c, c, t, v, 3, 1, 1, 1, list(range(0, t, 2))).unfold
    cases.append(("fold   %dx%dx%d <- unfolded" % (c, t, v), unf.T, (n, c, unf.t_out, unf.v_out)))
    cases.append(("unfold %dx%dx%d" % (c, t, v), unf, (n, c, t, v)))
for c, t, v, keep in ((64, 64, 12, [1, 4, 6, 9, 11]), (128, 32, 5, [0, 1, 2, 3, 4]), (256, 16, 5, [2])):
    sel = G.select_table(t, v, list(range(0, t, 2)), keep)
    cases.append(("select %dx%dx%d -> %dx%d" % (c, t, v, sel.t_out, sel.v_out), sel, (n, c, t, v)))
for name, table, shape in cases:
    x = torch.randn(*shape, device="cuda")
    us = timeit(lambda: ops.plane_spmm(x, table))
    gb = (x.numel() + shape[0] * shape[1] * table.p_out) * 4 / 1e9
    tot += us
    print("%-34s %7.1f us  %5.0f GB/s" % (name, us, gb / us * 1e6))
    del x
print("sum %.1f us" % tot)
