"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, imported
through oracle/ref_shim.py) in the build container.  The reference has no golden vectors of its
own, so these fixtures are what pins the oracle (tests/test_oracle_golden.py).

    python tests/golden/make_golden.py

Inputs and the ~7 M parameters are NOT stored: they are re-derived from (key, shape, seed) by
oracle.networks.synth_params / the seeded generators in `inputs()` below, so only outputs, small
gradients and strided subsamples of the large gradients are committed.

The training-loop section re-enacts kinetic-gan.py:137-174 and :94-114 around the reference
modules (the script itself cannot be imported: it needs dataset files and creates run dirs at
import); host RNG draws (z at :140, alpha at :97) are replaced by the seeded
tensors from `inputs()`.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle import networks as onet  # noqa: E402
from oracle.graph import SkeletonTables  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SUB = 211          # stride of the gradient / parameter subsamples

CASES = {
    "ntu_small": dict(cfg=onet.Config(dataset="ntu", n_classes=6, t_size=16, mlp_dim=2, channels=3), n=3),
    "h36m_small": dict(cfg=onet.Config(dataset="h36m", n_classes=4, t_size=32, mlp_dim=3, channels=2), n=2),
}


def inputs(cfg, n, seed, dtype=torch.float32):
    v = SkeletonTables(cfg.dataset).num_node[0]
    g = torch.Generator().manual_seed(1000 + seed)
    real = torch.rand(n, cfg.channels, cfg.t_size, v, generator=g, dtype=torch.float64) * 2 - 1
    z = torch.randn(n, cfg.latent_dim, generator=g, dtype=torch.float64)
    labels = torch.randint(0, cfg.n_classes, (n,), generator=g)
    alpha = torch.rand(n, 1, 1, 1, generator=g, dtype=torch.float64)
    cot_g = torch.randn(n, cfg.channels, cfg.t_size, v, generator=g, dtype=torch.float64)
    cot_d = torch.randn(n, 1, generator=g, dtype=torch.float64)
    return dict(real=real.to(dtype), z=z.to(dtype), labels=labels, alpha=alpha.to(dtype),
                cot_g=cot_g.to(dtype), cot_d=cot_d.to(dtype))


def draw_noises(cfg, n, seed, dtype=torch.float32):
    """Same tensors the reference draws at generator.py:179 after torch.manual_seed(seed)."""
    torch.manual_seed(seed)
    return [torch.randn(*s).to(dtype) for s in onet.noise_shapes(cfg, n)]


def sub(t):
    return t.detach().reshape(-1)[::SUB].double().numpy().copy()


def build_reference(cfg, dtype):
    gen, dis = ref_shim.load()
    G = gen.Generator(cfg.latent_dim, cfg.channels, cfg.n_classes, cfg.t_size, cfg.mlp_dim, dataset=cfg.dataset)
    D = dis.Discriminator(cfg.channels, cfg.n_classes, cfg.t_size, cfg.latent_dim, dataset=cfg.dataset)
    pg = onet.synth_params(onet.g_param_shapes(cfg), 1)
    pd = onet.synth_params(onet.d_param_shapes(cfg), 2)
    assert set(pg) == set(G.state_dict()) and set(pd) == set(D.state_dict())
    G.load_state_dict(pg)
    D.load_state_dict(pd)
    if dtype == torch.float64:
        G, D = G.double(), D.double()
        # the reference casts A to float32 at generator.py:47 / discriminator.py:19; ground truth keeps float64
        G.A = [torch.tensor(a, dtype=torch.float64) for a in G.graph.As]
        D.A = [torch.tensor(a, dtype=torch.float64) for a in D.graph.As]
    return G, D


def run_case(name, cfg, n):
    out = {}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        ref_shim.set_float_type(dtype)
        G, D = build_reference(cfg, dtype)
        x = inputs(cfg, n, 0, dtype)
        # ---- generator forward (training mode), per-block outputs, first-order grads
        blocks = []
        hooks = [m.register_forward_hook(lambda m, i, o: blocks.append(o[0].detach().clone()))
                 for m in G.st_gcn_networks]
        G.train()
        torch.manual_seed(11)
        fake = G(x["z"], x["labels"])
        for h in hooks:
            h.remove()
        out[tag + "/g_out"] = fake.detach().double().numpy()
        for i, b in enumerate(blocks):
            out[tag + "/g_block%d" % i] = b.double().numpy()
        G.zero_grad()
        (fake * x["cot_g"]).sum().backward()
        for k, p in G.named_parameters():
            out[tag + "/g_grad/" + k] = sub(p.grad)
        for k, b in G.named_buffers():
            if "running" in k:
                out[tag + "/g_bn_after/" + k] = b.detach().double().numpy()
        # ---- generator forward (eval mode; BN running stats) and W-space truncation (generator.py:97-108)
        G2, _ = build_reference(cfg, dtype)
        G2.eval()
        torch.manual_seed(12)
        out[tag + "/g_out_eval"] = G2(x["z"], x["labels"]).detach().double().numpy()
        np.random.seed(5)
        torch.manual_seed(13)
        out[tag + "/g_out_trunc"] = G2(x["z"], x["labels"], 0.95).detach().double().numpy()
        # ---- discriminator forward, per-block outputs, grads of <D(x), cot>
        blocks = []
        hooks = [m.register_forward_hook(lambda m, i, o: blocks.append(o[0].detach().clone()))
                 for m in D.st_gcn_networks]
        xr = x["real"].clone().requires_grad_(True)
        dv = D(xr, x["labels"])
        for h in hooks:
            h.remove()
        out[tag + "/d_out"] = dv.detach().double().numpy()
        for i, b in enumerate(blocks):
            out[tag + "/d_block%d" % i] = b.double().numpy()
        D.zero_grad()
        (dv * x["cot_d"]).sum().backward()
        out[tag + "/d_grad_x"] = xr.grad.double().numpy()
        for k, p in D.named_parameters():
            out[tag + "/d_grad/" + k] = sub(p.grad)
        # ---- gradient penalty (kinetic-gan.py:94-114) with supplied alpha, and its parameter grads
        D.zero_grad()
        real, fk = x["real"], fake.detach()
        inter = (x["alpha"] * real + ((1 - x["alpha"]) * fk)).requires_grad_(True)
        d_inter = D(inter, x["labels"])
        ones = torch.ones(real.shape[0], 1, dtype=dtype)
        grads = torch.autograd.grad(outputs=d_inter, inputs=inter, grad_outputs=ones, create_graph=True,
                                    retain_graph=True, only_inputs=True)[0]
        gp = ((grads.reshape(grads.size(0), -1).norm(2, dim=1) - 1) ** 2).mean()
        gp.backward()
        out[tag + "/gp"] = np.array(gp.item())
        out[tag + "/gp_grads_x"] = grads.detach().double().numpy()
        for k, p in D.named_parameters():
            out[tag + "/gp_grad/" + k] = sub(p.grad if p.grad is not None else torch.zeros_like(p))
        # ---- two iterations of the training loop body (kinetic-gan.py:137-174): i=0 (D+G step), i=1 (D step)
        G, D = build_reference(cfg, dtype)
        G.train()
        opt_g = torch.optim.Adam(G.parameters(), lr=cfg.lr, betas=(cfg.b1, cfg.b2))
        opt_d = torch.optim.Adam(D.parameters(), lr=cfg.lr, betas=(cfg.b1, cfg.b2))
        for i in range(2):
            xi = inputs(cfg, n, 10 + i, dtype)
            opt_d.zero_grad()
            torch.manual_seed(100 + 2 * i)
            fake_imgs = G(xi["z"], xi["labels"])
            real_validity = D(xi["real"], xi["labels"])
            fake_validity = D(fake_imgs, xi["labels"])
            inter = (xi["alpha"] * xi["real"].data + ((1 - xi["alpha"]) * fake_imgs.data)).requires_grad_(True)
            d_inter = D(inter, xi["labels"])
            grads = torch.autograd.grad(outputs=d_inter, inputs=inter, grad_outputs=torch.ones(n, 1, dtype=dtype),
                                        create_graph=True, retain_graph=True, only_inputs=True)[0]
            gp = ((grads.reshape(n, -1).norm(2, dim=1) - 1) ** 2).mean()
            d_loss = -torch.mean(real_validity) + torch.mean(fake_validity) + cfg.lambda_gp * gp
            d_loss.backward()
            opt_d.step()
            opt_g.zero_grad()
            out[tag + "/train/d_loss%d" % i] = np.array(d_loss.item())
            if i % cfg.n_critic == 0:
                torch.manual_seed(101 + 2 * i)
                fake_imgs = G(xi["z"], xi["labels"])
                g_loss = -torch.mean(D(fake_imgs, xi["labels"]))
                g_loss.backward()
                opt_g.step()
                out[tag + "/train/g_loss%d" % i] = np.array(g_loss.item())
        for k, v in G.state_dict().items():
            out[tag + "/train/g_after/" + k] = sub(v) if v.numel() > 4096 else v.detach().double().numpy()
        for k, v in D.state_dict().items():
            out[tag + "/train/d_after/" + k] = sub(v) if v.numel() > 4096 else v.detach().double().numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "->", len(out), "arrays", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


def graph_tables():
    gen, _ = ref_shim.load()
    for name, g in (("ntu", gen.graph_ntu()), ("h36m", gen.Graph_h36m())):
        out = {"num_node": np.array(g.num_node), "center": np.array(g.center)}
        for l in range(4):
            out["As%d" % l] = g.As[l]
            out["map%d" % l] = g.map[l]
            out["edge%d" % l] = np.asarray(g.edge[l])
        for l in range(3):
            out["mapping%d_len" % l] = np.array(len(g.mapping[l]))
            for j, h in enumerate(g.mapping[l]):
                out["mapping%d_%d" % (l, j)] = np.asarray(h)
        np.savez_compressed(os.path.join(HERE, "graph_%s.npz" % name), **out)


if __name__ == "__main__":
    torch.set_num_threads(8)
    graph_tables()
    for name, c in CASES.items():
        run_case(name, c["cfg"], c["n"])
