"""Runs the UNMODIFIED reference modules (models/generator.py, models/discriminator.py of DegardinBruno/Kinetic-GAN) as a timed
baseline.  TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py): used by `bench.py --impl reference` (host cores, `kind:
"reference"`), by bench.py's `cpu_baseline` and `gpu_reference` legs, and by tests.  Never imported by the product.

Where the reference comes from: /root/reference in the build container, else baseline/_ref/ - the pip-installed copy that
tools/install_reference.py makes at build time and that travels to the GPU box (git-ignored, not gpurun-ignored).

What is the reference's and what is restated here:
  * `Generator`, `Discriminator` and everything below them (st_gcn, ConvTemporalGraphical, graph_ntu, ...): the reference's own
    files, imported unchanged.  The import shim only (a) registers an empty `matplotlib.pyplot` (imported but unused at
    graph_ntu.py:3 / graph_h36m.py:3, not installed in this image) and, for CPU runs, (b) makes `.cuda()` the identity and drops
    `device='cuda:0'` from `torch.randn` (generator.py:47,179, discriminator.py:19 hard-code the device).
  * the loop body: kinetic-gan.py cannot be imported (it is a script that parses argv, creates run directories and opens the
    author's dataset paths at import, kinetic-gan.py:17-53,68-74), so `RefTrainer.iteration` re-enacts lines 137-174 and
    `compute_gradient_penalty` lines 94-114 statement by statement around the imported modules, with torch.optim.Adam as at :77-78.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.environ.get("KGAN_REFERENCE_ROOT", "/root/reference"), os.path.join(ROOT, "baseline", "_ref"))


def reference_root():
    for r in CANDIDATES:
        if os.path.exists(os.path.join(r, "models", "generator.py")):
            return r
    return None


def available():
    return reference_root() is not None


def load(force_cpu=False):
    """-> (models.generator, models.discriminator) of the untouched reference."""
    import importlib

    import torch

    root = reference_root()
    if root is None:
        raise RuntimeError("reference not found (neither /root/reference nor baseline/_ref: run tools/install_reference.py in the build container)")
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if force_cpu or not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        if not getattr(torch.randn, "_kgan_shim", False):
            _randn = torch.randn

            def randn(*a, **k):
                k.pop("device", None)
                return _randn(*a, **k)

            randn._kgan_shim = True
            torch.randn = randn
        torch.cuda.FloatTensor = torch.FloatTensor
    if root not in sys.path:
        sys.path.insert(0, root)
    return importlib.import_module("models.generator"), importlib.import_module("models.discriminator")


class RefTrainer:
    """kinetic-gan.py:59-60 (model construction), :77-78 (optimizers), :94-114 (gradient penalty), :137-174 (loop body)."""

    def __init__(self, shape, device="cpu", lr=0.0002, b1=0.5, b2=0.999, n_critic=5, lambda_gp=10, latent_dim=512):
        import torch

        gen, dis = load(force_cpu=(device == "cpu"))
        self.torch, self.dev = torch, torch.device(device)
        self.latent_dim, self.n_critic, self.lambda_gp = latent_dim, n_critic, lambda_gp
        self.G = gen.Generator(latent_dim, shape["channels"], shape["n_classes"], shape["t_size"], mlp_dim=shape["mlp_dim"], dataset=shape["dataset"])
        self.D = dis.Discriminator(shape["channels"], shape["n_classes"], shape["t_size"], latent_dim, dataset=shape["dataset"])
        if device != "cpu":
            self.G.cuda()
            self.D.cuda()
        self.opt_g = torch.optim.Adam(self.G.parameters(), lr=lr, betas=(b1, b2))
        self.opt_d = torch.optim.Adam(self.D.parameters(), lr=lr, betas=(b1, b2))

    def gradient_penalty(self, real, fake, labels):
        torch = self.torch
        alpha = torch.as_tensor(np.random.random((real.size(0), 1, 1, 1)), dtype=real.dtype, device=real.device)       # :97
        inter = (alpha * real + ((1 - alpha) * fake)).requires_grad_(True)                                              # :100
        d_inter = self.D(inter, labels)                                                                                  # :101
        ones = torch.ones(real.shape[0], 1, dtype=real.dtype, device=real.device)                                        # :102
        grads = torch.autograd.grad(outputs=d_inter, inputs=inter, grad_outputs=ones, create_graph=True, retain_graph=True,
                                    only_inputs=True)[0]                                                                 # :104-111
        grads = grads.reshape(grads.size(0), -1)
        return ((grads.norm(2, dim=1) - 1) ** 2).mean()                                                                  # :112-113

    def iteration(self, i, real, labels):
        torch = self.torch
        self.opt_d.zero_grad()                                                                                           # :137
        z = torch.as_tensor(np.random.normal(0, 1, (real.size(0), self.latent_dim)), dtype=real.dtype, device=real.device)   # :140
        fake = self.G(z, labels)                                                                                         # :143
        real_v = self.D(real, labels)                                                                                    # :146
        fake_v = self.D(fake, labels)                                                                                    # :148
        gp = self.gradient_penalty(real.data, fake.data, labels.data)                                                    # :150
        d_loss = -torch.mean(real_v) + torch.mean(fake_v) + self.lambda_gp * gp                                          # :152
        d_loss.backward()                                                                                                # :154
        self.opt_d.step()                                                                                                # :155
        self.opt_g.zero_grad()                                                                                           # :157
        g_loss = None
        if i % self.n_critic == 0:                                                                                       # :160
            fake = self.G(z, labels)                                                                                     # :167
            g_loss = -torch.mean(self.D(fake, labels))                                                                   # :170-171
            g_loss.backward()                                                                                            # :173
            self.opt_g.step()                                                                                            # :174
        return d_loss, g_loss


def time_training(shape, batch, steps, warmup, device="cpu", tf32=False, first_index=0, threads=None):
    """samples/s of the unmodified reference loop body on `device`; -> (samples_per_s, ms_per_step, cores_or_device_name)."""
    import time

    import torch

    if device == "cpu":
        torch.set_num_threads(threads or os.cpu_count() or 1)
    else:
        torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
        torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.manual_seed(0)
    np.random.seed(0)
    tr = RefTrainer(shape, device)
    g = torch.Generator().manual_seed(0)
    dev = torch.device(device)

    def batch_fn():
        real = (torch.rand(batch, shape["channels"], shape["t_size"], shape["joints"], generator=g) * 2 - 1).to(dev)
        labels = torch.randint(0, shape["n_classes"], (batch,), generator=g).to(dev)
        return real, labels

    sync = (lambda: torch.cuda.synchronize()) if device != "cpu" else (lambda: None)
    i = first_index
    for _ in range(warmup):
        tr.iteration(i, *batch_fn())
        i += 1
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        d_loss, _ = tr.iteration(i, *batch_fn())
        i += 1
    d_loss.item()
    sync()
    dt = time.perf_counter() - t0
    who = torch.get_num_threads() if device == "cpu" else torch.cuda.get_device_name(dev)
    return batch * steps / dt, dt / steps * 1e3, who


def time_generate(shape, batch, steps, warmup, device="cpu", tf32=False, threads=None):
    """generate.py:93 `generator(z, labels)` in eval mode (generate.py:67), without no_grad (the reference has none)."""
    import time

    import torch

    if device == "cpu":
        torch.set_num_threads(threads or os.cpu_count() or 1)
    else:
        torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
        torch.backends.cudnn.allow_tf32 = bool(tf32)
    gen, _ = load(force_cpu=(device == "cpu"))
    torch.manual_seed(0)
    G = gen.Generator(512, shape["channels"], shape["n_classes"], shape["t_size"], mlp_dim=shape["mlp_dim"], dataset=shape["dataset"])
    if device != "cpu":
        G.cuda()
    G.eval()
    dev = torch.device(device)
    g = torch.Generator().manual_seed(0)
    sync = (lambda: torch.cuda.synchronize()) if device != "cpu" else (lambda: None)

    def once():
        z = torch.randn(batch, 512, generator=g).to(dev)
        labels = torch.randint(0, shape["n_classes"], (batch,), generator=g).to(dev)
        return G(z, labels).data.cpu()                       # generate.py:93,95

    for _ in range(warmup):
        once()
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    sync()
    dt = time.perf_counter() - t0
    who = torch.get_num_threads() if device == "cpu" else torch.cuda.get_device_name(dev)
    return batch * steps / dt, dt / steps * 1e3, who


def tf32_gradient_deviation(shape, batch=64):
    """On the GPU: rel-L2 between the reference critic's parameter gradients computed by torch with TF32 tensor cores allowed
    (cudnn / matmul allow_tf32) and with exact fp32 - the yardstick for what a TF32 evaluation of THIS network does to its gradients
    (LeakyReLU slopes flip for every pre-activation within the forward error of zero, tests/test_bench_parity_gpu.py)."""
    import torch

    _, dis = load()
    torch.manual_seed(0)
    D = dis.Discriminator(shape["channels"], shape["n_classes"], shape["t_size"], 512, dataset=shape["dataset"]).cuda()
    g = torch.Generator().manual_seed(1)
    real = (torch.rand(batch, shape["channels"], shape["t_size"], shape["joints"], generator=g) * 2 - 1).cuda()
    labels = torch.randint(0, shape["n_classes"], (batch,), generator=g).cuda()
    cot = torch.randn(batch, 1, generator=g).cuda()
    out = {}
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        D.zero_grad(set_to_none=True)
        v = D(real, labels)
        (v * cot).sum().backward()
        out[tf32] = (v.detach().double(), torch.cat([p.grad.reshape(-1).double() for p in D.parameters()]))
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    return {"output_rel_l2": rel(out[True][0], out[False][0]), "grad_rel_l2": rel(out[True][1], out[False][1]), "batch": batch}
