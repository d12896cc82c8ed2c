"""Per-layer microbenchmark of the critic's tap convolutions and adjacency kernels at the bench batch size
(CUDA events, inputs larger than L2 or flushed).  python tools/layer_bench.py [--batch 256] [--precision tf32]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kgan_b200 as kgan  # noqa: E402

ops, G = kgan.ops, kgan.geometry
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--precision", default="tf32")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--only", default="")
a = ap.parse_args()
kgan.set_precision(a.precision)
N = a.batch
dev = "cuda"
flush = torch.empty(160 * 1024 * 1024 // 4, device=dev)
keep1 = [0, 2, 5, 7, 9, 11, 13, 14, 17, 18, 20]
keep2 = [2, 4, 6, 8, 10]


def _span(body):
    torch.cuda.synchronize()
    torch.cuda._sleep(6_000_000)          # ~3 ms of GPU busy time: the CPU enqueues everything behind it (no launch gaps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        body()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.reps * 1e3      # us


def timeit(fn):
    """GPU time of fn(), launches back to back (no CPU launch gaps).  No L2 flush: layers whose working set is below
    the 126 MB L2 read warm data, as they do inside the training step (the producer layer just wrote it)."""
    fn()
    return _span(fn)


LAYERS = [   # the critic's tap convolutions as Discriminator.forward runs them (joints 11 -> 12 padded, strided tcn unfolded)
    ("D0.tcn", dict(c_in=32, c_out=32, t_in=64, v_in=12, kt=3, pad=1)),
    ("D1.gcn", dict(c_in=32, c_out=64, t_in=64, v_in=12, K=3)),
    ("D1.tcn", dict(c_in=64, c_out=64, t_in=64, v_in=12, kt=3, pad=1)),
    ("D1.res", dict(c_in=32, c_out=64, t_in=64, v_in=12)),
    ("D2.gcn", dict(c_in=64, c_out=128, t_in=64, v_in=5, K=3)),
    ("D2.tcn", dict(unfold=True, c_in=128, c_out=128, t_in=64, v_in=5, kt=3, pad=1, stride=1, dil=1, t_sel=list(range(0, 64, 2)))),
    ("D2.res", dict(c_in=64, c_out=128, t_in=32, v_in=5)),
    ("D3.gcn", dict(c_in=128, c_out=256, t_in=32, v_in=5, K=3)),
    ("D3.tcn", dict(unfold=True, c_in=256, c_out=256, t_in=32, v_in=5, kt=3, pad=1, stride=1, dil=1, t_sel=list(range(0, 32, 2)))),
    ("D3.res", dict(c_in=128, c_out=256, t_in=16, v_in=5)),
    ("D4.gcn", dict(c_in=256, c_out=512, t_in=16, v_in=1, K=3)),
    ("D4.tcn", dict(unfold=True, c_in=512, c_out=512, t_in=16, v_in=1, kt=3, pad=1, stride=1, dil=1, t_sel=list(range(0, 16, 2)))),
    ("D4.res", dict(c_in=256, c_out=512, t_in=8, v_in=1)),
    ("D5.gcn", dict(c_in=512, c_out=512, t_in=8, v_in=1, K=3)),
    ("D5.tcn", dict(unfold=True, c_in=512, c_out=512, t_in=8, v_in=1, kt=3, pad=1, stride=1, dil=1, t_sel=[0, 2, 4, 6])),
]
print("%-8s %-6s %9s %9s %9s   (batch %d, %s)" % ("layer", "op", "us", "GB/s", "TFLOP/s", N, a.precision))
for name, kw in LAYERS:
    if a.only and a.only not in name:
        continue
    kw = dict(kw)
    if kw.pop("unfold", False):
        g = G.UnfoldedTcnGeom(**kw)
        xr = torch.randn(N, g.c_in, g.t_in, g.v_in, device=dev)
        us = timeit(lambda: ops.plane_spmm(xr, g.unfold))
        print("%-8s %-6s %9.1f %9.0f" % (name, "unfold", us, 4.0 * xr.numel() * (1 + g.kt * g.t_out / g.t_in) / us / 1e3))
        x = torch.randn(N, g.c_in, g.kt * g.t_out, g.v_in, device=dev)
    else:
        g = G.TapConvGeom(**kw)
        x = torch.randn(N, g.K * g.c_in, g.t_in, g.v_in, device=dev)
    w = torch.randn(g.K * g.c_out, g.c_in, g.kt, 1, device=dev) * 0.05
    go = torch.randn(N, g.c_out, g.t_out, g.v_out, device=dev)
    bytes_io = 4.0 * (x.numel() + go.numel())
    flops = 2.0 * N * g.p_out * g.c_out * g.c_in * g.kt * g.K
    for op, fn in (("fwd", lambda: ops.tapconv_fwd(x, w, g.fwd)), ("dgrad", lambda: ops.tapconv_fwd(go, w, g.dgrad)),
                   ("wgrad", lambda: ops.tapconv_wgrad(x, go, g.fwd, tuple(w.shape)))):
        us = timeit(fn)
        print("%-8s %-6s %9.1f %9.0f %9.1f" % (name, op, us, bytes_io / us / 1e3, flops / us / 1e6))
for name, (c, t, v, vo) in (("D0", (3, 64, 25, 12)), ("D1", (32, 64, 12, 12)), ("D2", (64, 64, 12, 5)), ("D3", (128, 32, 5, 5)),
                            ("D4", (256, 16, 5, 1)), ("D5", (512, 8, 1, 1))):
    if a.only and a.only not in name:
        continue
    x = torch.randn(N, c, t, v, device=dev)
    A = ((torch.rand(3, v, v, device=dev) < 0.15).float() * torch.rand(3, v, v, device=dev) + torch.eye(v, device=dev))[:, :, :vo].contiguous()
    gx = torch.randn(N, 3 * c, t, vo, device=dev)
    bytes_io = 4.0 * (x.numel() + gx.numel())
    for op, fn in (("mix", lambda: ops.adjmix_fwd(x, A)), ("mix_dx", lambda: ops.adjmix_bwd_x(gx, A)), ("mix_dA", lambda: ops.adjmix_bwd_a(x, gx, 3, A))):
        us = timeit(fn)
        print("%-8s %-6s %9.1f %9.0f" % (name, op, us, bytes_io / us / 1e3))
