"""One fp32x3 forward launch (for compute-sanitizer): python tools/x3_one.py C N"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import kgan_b200 as kgan  # noqa: E402

c, n = int(sys.argv[1]), int(sys.argv[2])
kgan.set_precision("fp32x3")
geom = kgan.geometry.TapConvGeom(c_in=c, c_out=c, t_in=64, v_in=12, kt=3, pad=1)
x = torch.randn(n, c, 64, 12, device="cuda")
w = torch.randn(c, c, 3, 1, device="cuda") / (3 * c) ** 0.5
y = kgan.ops.tapconv_fwd(x, w, geom.fwd)
torch.cuda.synchronize()
print("fwd ok", float(y.abs().mean()))
dw = kgan.ops.tapconv_wgrad(x, y, geom.fwd, tuple(w.shape))
torch.cuda.synchronize()
print("wgrad ok", float(dw.abs().mean()))
