"""Bring-up probe: the TMA-fed weight-gradient kernel on planes of 32 / 16 / 8 / 4 positions (boxes spanning 1 / 2 / 4 / 8 samples),
each case in its own process so that a device fault does not poison the others.
    python tools/probe_wgrad_tma.py            # runs every case
    python tools/probe_wgrad_tma.py <case>     # one case, in this process"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    # name: (c_in, c_out, t, v, K, n)
    "p32_n5": (64, 64, 8, 4, 3, 40),
    "p16_n64": (64, 64, 4, 4, 3, 64),
    "p16_n65": (64, 64, 4, 4, 3, 65),
    "p80_n15": (256, 512, 16, 5, 3, 15),
    "p8_n128": (64, 64, 2, 4, 3, 128),
    "p8_n130": (64, 64, 2, 4, 3, 130),
    "p4_n256": (64, 64, 1, 4, 3, 256),
    "p4_n301": (64, 64, 1, 4, 3, 301),
    "p20_n64": (256, 128, 4, 5, 3, 64),
}


def run(name):
    import torch

    import emu_backend as emu
    import kgan_b200 as kgan

    ci, co, t, v, K, n = CASES[name]
    kgan.set_precision("tf32")
    geom = kgan.geometry.TapConvGeom(ci, co, t, v, K=K)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, K * ci, t, v, generator=g)
    go = torch.randn(n, co, t, v, generator=g)
    lib = __import__("importlib").import_module("kinetic-gan_b200._lib").lib()
    ok = lib.kgan_tapconv_wgrad_tma_ok(geom.fwd.cstruct(n, 0, 1))
    got = kgan.ops.tapconv_wgrad(x.cuda(), go.cuda(), geom.fwd, (K * co, ci, 1, 1))
    torch.cuda.synchronize()
    ref = emu.tapconv_wgrad(x.double(), go.double(), geom.fwd, (K * co, ci, 1, 1))
    err = ((got.cpu().double() - ref).norm() / ref.norm()).item()
    print("%-10s tma_ok=%d rel=%.2e %s" % (name, ok, err, "OK" if err < 1e-3 else "WRONG"))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        for name in CASES:
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, env=dict(os.environ, CUDA_LAUNCH_BLOCKING="1"))
            out = [ln for ln in r.stdout.splitlines() if ln.startswith(name)]
            print(out[0] if out else "%-10s FAULT: %s" % (name, (r.stderr.strip().splitlines() or ["?"])[-1][:150]))
