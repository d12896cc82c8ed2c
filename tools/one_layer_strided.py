"""Stride-2 temporal convolution of a down-sampling critic block (D2.tcn: 128 channels, 64 x 5 -> 32 x 5): time-unfolded copy +
TMA-fed kernel (round 1) against the operand-building kernel reading the tensor as it is.  Forward and data gradient."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kgan_b200 as kgan  # noqa: E402

ops, G = kgan.ops, kgan.geometry
kgan.set_precision("tf32")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for c, T, V in ((128, 64, 5), (256, 32, 5), (512, 16, 1)):
    t_sel = list(range(0, T, 2))
    x = torch.randn(n, c, T, V, device="cuda")
    w = torch.randn(c, c, 3, 1, device="cuda") / (3 * c) ** 0.5
    go = torch.randn(n, c, T // 2, V, device="cuda")
    G.STAGED_POLICY = "fallback"
    unf = G.UnfoldedTcnGeom(c, c, T, V, 3, 1, 1, 1, t_sel)
    G.STAGED_POLICY = "all"
    direct = G.TapConvGeom(c, c, T, V, kt=3, pad=1, t_sel=t_sel)

    def timeit(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20 * 1e3

    a = timeit(lambda: ops.tapconv_fwd(ops.plane_spmm(x, unf.unfold), w, unf.fwd))
    b = timeit(lambda: ops.tapconv_fwd(x, w, direct.fwd))
    a2 = timeit(lambda: ops.plane_spmm(ops.tapconv_fwd(go, w, unf.dgrad), unf.unfold.T))
    b2 = timeit(lambda: ops.tapconv_fwd(go, w, direct.dgrad))
    ya = ops.tapconv_fwd(ops.plane_spmm(x, unf.unfold), w, unf.fwd)
    yb = ops.tapconv_fwd(x, w, direct.fwd)
    print("c=%d T=%d V=%d n=%d: fwd unfold+tma %.1f us, staged %.1f us | dgrad tma+fold %.1f us, staged %.1f us | max diff %.2e"
          % (c, T, V, n, a, b, a2, b2, (ya - yb).abs().max().item()))
