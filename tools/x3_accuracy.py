"""Accuracy of the three arithmetic modes on one temporal convolution (64 frames x 12 joints, 3 taps) as a function of the contraction
length: rel-L2 against float64 (torch, same device) of forward, data gradient and weight gradient.  Forward / data gradient contract over
3 * C channels, the weight gradient over n * 768 positions (split-K over one wave of CTAs).
usage: python tools/x3_accuracy.py [--out FILE]"""
import argparse
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import kgan_b200 as kgan  # noqa: E402

ops, G = kgan.ops, kgan.geometry


def rel(a, b):
    return ((a.double() - b).norm() / b.norm()).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    opt = ap.parse_args()
    lines = ["%-28s %-8s %10s %10s %10s" % ("case", "mode", "fwd", "dgrad", "wgrad")]
    gen = torch.Generator(device="cuda").manual_seed(0)
    for c, n in ((32, 64), (64, 64), (128, 64), (256, 32), (512, 16), (64, 512), (64, 2048), (32, 4096)):
        geom = G.TapConvGeom(c_in=c, c_out=c, t_in=64, v_in=12, kt=3, pad=1)
        x = torch.randn(n, c, 64, 12, device="cuda", generator=gen)
        w = torch.randn(c, c, 3, 1, device="cuda", generator=gen) / (3 * c) ** 0.5
        go = torch.randn(n, c, 64, 12, device="cuda", generator=gen)
        xd, wd, gd = x.double().requires_grad_(True), w.double().requires_grad_(True), go.double()
        y = F.conv2d(xd, wd, padding=(1, 0))
        gx, gw = torch.autograd.grad(y, (xd, wd), gd)
        for mode in ("fp32", "fp32x3", "tf32"):
            kgan.set_precision(mode)
            xx = ops.round_tf32(x) if mode == "tf32" else x
            gg = ops.round_tf32(go) if mode == "tf32" else go
            ref = (F.conv2d(xx.double(), wd, padding=(1, 0)), None, None) if mode == "tf32" else (y, gx, gw)
            e1 = rel(ops.tapconv_fwd(xx, w, geom.fwd), ref[0].detach())
            e2 = rel(ops.tapconv_fwd(gg, w, geom.dgrad), gx if mode != "tf32" else torch.autograd.grad(F.conv2d(xd, wd, padding=(1, 0)), xd, gg.double())[0])
            e3 = rel(ops.tapconv_wgrad(xx, gg, geom.fwd, tuple(w.shape)),
                     gw if mode != "tf32" else torch.autograd.grad(F.conv2d(xx.double(), wd, padding=(1, 0)), wd, gg.double())[0])
            lines.append("C=%-4d n=%-5d K=%-5d Kw=%-8d %-8s %10.2e %10.2e %10.2e" % (c, n, 3 * c, n * 768, mode, e1, e2, e3))
            print(lines[-1], flush=True)
        del x, w, go, xd, wd, gd, y, gx, gw
    kgan.set_precision("fp32")
    text = "\n".join(lines)
    if opt.out:
        Path(opt.out).write_text(text + "\n")


if __name__ == "__main__":
    main()
