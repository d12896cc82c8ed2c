"""Oracle restatement of the Kinetic-GAN networks and WGAN-GP step in plain CPU PyTorch.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for the CUDA path and the timed
CPU baseline of bench.py.  Works in float32 and float64 (ground truth).

Each function cites the reference lines it follows (paths relative to /root/reference):
  conv_temporal_graphical   models/init_gan/tgcn.py:58-68
  upsample_s                models/generator.py:185-200
  g_block                   models/generator.py:168-182 (+ ctor :112-166)
  mapping / truncate        models/generator.py:22-37, :80-87, :97-108
  generator_forward         models/generator.py:78-95 (+ block table :56-64)
  d_block                   models/discriminator.py:125-142 (+ ctor :80-123)
  discriminator_forward     models/discriminator.py:52-74 (+ block table :28-35)
  gradient_penalty          kinetic-gan.py:94-114
  d_loss / g_loss / train_iteration   kinetic-gan.py:137-174, Adam at :77-78

Parameters are a flat dict keyed exactly like the reference `state_dict()`.
The arithmetic itself (conv2d, einsum, batch_norm, interpolate, Adam) lives in torch, which is a
third-party dependency of the reference (requirements.txt:24 pins torch==1.7.1; this image has
2.11.0) - the oracle calls the same torch operators in the same order as the reference.
"""
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

from .graph import SkeletonTables


@dataclass
class Config:
    dataset: str = "ntu"
    latent_dim: int = 512      # kinetic-gan.py:30 (G in_channels / D `latent`)
    channels: int = 3          # kinetic-gan.py:35
    n_classes: int = 60        # kinetic-gan.py:32
    t_size: int = 64           # kinetic-gan.py:33
    mlp_dim: int = 4           # kinetic-gan.py:31
    n_critic: int = 5          # kinetic-gan.py:36
    lambda_gp: float = 10.0    # kinetic-gan.py:37
    lr: float = 2e-4           # kinetic-gan.py:26
    b1: float = 0.5            # kinetic-gan.py:27
    b2: float = 0.999          # kinetic-gan.py:28


def g_block_table(cfg):
    """(c_in, c_out, lvl, bn, residual, up_s, up_t, tan)  - models/generator.py:56-64."""
    t, L, C = cfg.t_size, cfg.latent_dim + cfg.n_classes, cfg.channels
    return [
        (L, 512, 3, False, False, False, 1, False),
        (512, 256, 3, True, True, False, int(t / 16), False),
        (256, 128, 2, False, True, True, int(t / 16), False),
        (128, 64, 2, True, True, False, int(t / 8), False),
        (64, 32, 1, False, True, True, int(t / 4), False),
        (32, C, 1, True, True, False, int(t / 2), False),
        (C, C, 0, False, True, True, t, True),
    ]


def d_block_table(cfg):
    """(c_in, c_out, lvl, residual, dw_s, dw_t)  - models/discriminator.py:28-35."""
    t, C = cfg.t_size, cfg.channels
    return [
        (C + cfg.n_classes, 32, 0, False, True, t),
        (32, 64, 1, True, False, t),
        (64, 128, 1, True, True, int(t / 2)),
        (128, 256, 2, True, False, int(t / 4)),
        (256, 512, 2, True, True, int(t / 8)),
        (512, cfg.latent_dim, 3, True, False, int(t / 16)),
    ]


# ----------------------------------------------------------------------------------------------
# parameter dictionaries (shapes = reference state_dict)
# ----------------------------------------------------------------------------------------------
def g_param_shapes(cfg, tables=None):
    tables = tables or SkeletonTables(cfg.dataset)
    L = cfg.latent_dim + cfg.n_classes
    s = {}
    for i in range(cfg.mlp_dim):
        s["mlp.mlp.%d.weight" % (2 * i)] = (L, L)
        s["mlp.mlp.%d.bias" % (2 * i)] = (L,)
    for i, (ci, co, lvl, bn, res, up_s, up_t, tan) in enumerate(g_block_table(cfg)):
        p = "st_gcn_networks.%d." % i
        s[p + "gcn.conv.weight"] = (3 * co, ci, 1, 1)
        s[p + "tcn.0.weight"] = (co, co, 3, 1)
        s[p + "tcn.0.bias"] = (co,)
        if bn:
            for k in ("weight", "bias", "running_mean", "running_var"):
                s[p + "tcn.1." + k] = (co,)
            s[p + "tcn.1.num_batches_tracked"] = ()
        if res and ci != co:
            s[p + "residual.0.weight"] = (co, ci, 1, 1)
            s[p + "residual.0.bias"] = (co,)
            for k in ("weight", "bias", "running_mean", "running_var"):
                s[p + "residual.1." + k] = (co,)
            s[p + "residual.1.num_batches_tracked"] = ()
        s[p + "noise.weight"] = (1, co, 1, 1)
    for i, blk in enumerate(g_block_table(cfg)):
        v = tables.num_node[blk[2]]
        s["edge_importance.%d" % i] = (3, v, v)
    s["label_emb.weight"] = (cfg.n_classes, cfg.n_classes)
    return s


def d_param_shapes(cfg, tables=None):
    tables = tables or SkeletonTables(cfg.dataset)
    s = {}
    for i, (ci, co, lvl, res, dw_s, dw_t) in enumerate(d_block_table(cfg)):
        p = "st_gcn_networks.%d." % i
        s[p + "gcn.conv.weight"] = (3 * co, ci, 1, 1)
        s[p + "tcn.weight"] = (co, co, 3, 1)
        s[p + "tcn.bias"] = (co,)
        if res and ci != co:
            s[p + "residual.weight"] = (co, ci, 1, 1)
            s[p + "residual.bias"] = (co,)
    for i, blk in enumerate(d_block_table(cfg)):
        v = tables.num_node[blk[2]]
        s["edge_importance.%d" % i] = (3, v, v)
    s["label_emb.weight"] = (cfg.n_classes, cfg.n_classes)
    s["fcn.weight"] = (1, cfg.latent_dim)
    s["fcn.bias"] = (1,)
    return s


def synth_params(shapes, seed, dtype=torch.float32, reference_init=False):
    """Deterministic synthetic parameters, reproducible from (key, shape, seed) alone so that golden
    fixtures need not store the ~7 M weights.  Magnitudes keep every activation O(1).
    With reference_init=True the mapping weights use the reference's N(0,1) (models/generator.py:29)."""
    import zlib

    out = {}
    for key in sorted(shapes):
        shape = shapes[key]
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) & 0x7FFFFFFF)
        if key.endswith("num_batches_tracked"):
            out[key] = torch.zeros((), dtype=torch.int64)
            continue
        r = torch.randn(shape, generator=g, dtype=torch.float64)
        if key.endswith("running_var"):
            v = 1.0 + 0.2 * r.abs()
        elif key.endswith("running_mean"):
            v = 0.1 * r
        elif key.startswith("edge_importance"):
            v = 1.0 + 0.2 * r
        elif ".tcn.1." in key or ".residual.1." in key:        # BatchNorm affine
            v = 1.0 + 0.2 * r if key.endswith("weight") else 0.1 * r
        elif key.endswith("noise.weight"):
            v = 0.3 * r
        elif key == "label_emb.weight":
            v = r
        elif key.endswith("bias"):
            v = 0.1 * r
        elif key.startswith("mlp.") and reference_init:
            v = r
        else:                                                  # conv / linear weights
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            v = r * (1.2 / np.sqrt(max(fan_in, 1)))
        out[key] = v.to(dtype)
    return out


# ----------------------------------------------------------------------------------------------
# operators
# ----------------------------------------------------------------------------------------------
def conv_temporal_graphical(x, weight, A):
    """tgcn.py:58-68: 1x1 conv (no bias) to K*C_out channels, view (n,K,C_out,t,v), einsum with A."""
    K = A.size(0)
    assert weight.size(0) % K == 0
    y = F.conv2d(x, weight)
    n, kc, t, v = y.size()
    y = y.view(n, K, kc // K, t, v)
    return torch.einsum("nkctv,kvw->nctw", y, A).contiguous()


def nearest_t(x, t_out):
    """F.interpolate(x, size=(t_out, V)) default mode='nearest' (generator.py:172, discriminator.py:134)."""
    return F.interpolate(x, size=(t_out, x.size(-1)))


def upsample_s(x, hoods, halve):
    """generator.py:185-200: insert, at fine index hood[0], the mean of coarse joints hood[1:]
    (divided by 2 iff lvl == 2, :195), in list order."""
    means = []
    for hood in hoods:
        m = torch.stack([x[..., int(j)] for j in hood[1:]], -1).mean(-1)
        means.append((m / 2 if halve else m).unsqueeze(-1))
    for hood, m in zip(hoods, means):
        idx = int(hood[0])
        x = torch.cat([x[..., :idx], m, x[..., idx:]], -1)
    return x


def batch_norm(x, p, prefix, training, state_out=None, eps=1e-5, momentum=0.1):
    """nn.BatchNorm2d defaults (generator.py:142,160).  When training, running stats in `p` are
    updated functionally into state_out (biased var normalises, unbiased var feeds running_var)."""
    w, b = p[prefix + "weight"], p[prefix + "bias"]
    rm, rv = p[prefix + "running_mean"], p[prefix + "running_var"]
    if not training:
        return F.batch_norm(x, rm, rv, w, b, False, momentum, eps)
    rm2, rv2 = rm.detach().clone().to(x.dtype), rv.detach().clone().to(x.dtype)
    y = F.batch_norm(x, rm2, rv2, w, b, True, momentum, eps)
    if state_out is not None:
        state_out[prefix + "running_mean"] = rm2
        state_out[prefix + "running_var"] = rv2
        state_out[prefix + "num_batches_tracked"] = p[prefix + "num_batches_tracked"] + 1
    return y


def g_block(x, A_eff, p, prefix, spec, tables, noise, training, state_out=None):
    """generator.py:168-182."""
    ci, co, lvl, bn, res, up_s, up_t, tan = spec
    if up_s:
        x = upsample_s(x, tables.mapping[lvl], halve=(lvl == 2))
    x = nearest_t(x, up_t)
    if not res:
        r = 0
    elif ci == co:
        r = x
    else:
        r = F.conv2d(x, p[prefix + "residual.0.weight"], p[prefix + "residual.0.bias"])
        r = batch_norm(r, p, prefix + "residual.1.", training, state_out)
    y = conv_temporal_graphical(x, p[prefix + "gcn.conv.weight"], A_eff)
    y = F.conv2d(y, p[prefix + "tcn.0.weight"], p[prefix + "tcn.0.bias"], padding=(1, 0))
    if bn:
        y = batch_norm(y, p, prefix + "tcn.1.", training, state_out)
    y = y + r
    y = y + p[prefix + "noise.weight"] * noise
    return torch.tanh(y) if tan else F.leaky_relu(y, 0.2)


def mapping(p, x, mlp_dim, per_sample_loop=False):
    """generator.py:22-37 applied as at :84-85 (a per-sample Python loop in the reference)."""
    def run(v):
        for i in range(mlp_dim):
            v = F.leaky_relu(F.linear(v, p["mlp.mlp.%d.weight" % (2 * i)], p["mlp.mlp.%d.bias" % (2 * i)]), 0.2)
        return v

    if per_sample_loop:
        return torch.stack([run(v) for v in x], 0)
    return run(x)


def truncate(p, w, t_lat, truncation, mlp_dim):
    """generator.py:97-108 with the N(0,1) latents `t_lat` (mean_size, L) supplied by the caller."""
    m = mapping(p, t_lat, mlp_dim).mean(0, keepdim=True)
    return m + truncation * (w - m)


def noise_shapes(cfg, n, tables=None):
    """Shapes of the per-block noise tensors drawn at generator.py:179, block order 0..6."""
    tables = tables or SkeletonTables(cfg.dataset)
    out, v = [], 1
    for (ci, co, lvl, bn, res, up_s, up_t, tan) in g_block_table(cfg):
        v = tables.num_node[lvl]
        out.append((n, 1, up_t, v))
    return out


def generator_forward(p, z, labels, cfg, tables, noises, training=True, state_out=None, trunc=None,
                      trunc_latents=None, per_sample_loop=False, collect=None):
    """generator.py:78-95.  `noises` = list of 7 tensors (N,1,T_i,V_i)."""
    c = p["label_emb.weight"][labels]
    x = torch.cat((c, z), -1)
    w = mapping(p, x, cfg.mlp_dim, per_sample_loop)
    if trunc is not None:
        w = truncate(p, w, trunc_latents, trunc, cfg.mlp_dim)
    x = w.view(*w.shape, 1, 1)
    for i, spec in enumerate(g_block_table(cfg)):
        A_eff = torch.as_tensor(tables.As[spec[2]], dtype=x.dtype) * p["edge_importance.%d" % i]
        x = g_block(x, A_eff, p, "st_gcn_networks.%d." % i, spec, tables, noises[i], training, state_out)
        if collect is not None:
            collect.append(x)
    return x


def d_block(x, A_eff, p, prefix, spec, tables, mask=None):
    """discriminator.py:125-142.  `mask` (checker-only, never set by the reference path): the LeakyReLU slopes (1 / 0.2 per
    element) to apply INSTEAD of deriving them from the sign of this block's own pre-activation - see discriminator_forward."""
    ci, co, lvl, res, dw_s, dw_t = spec
    if not res:
        r = 0
    elif ci == co:
        r = x
    else:
        r = F.conv2d(x, p[prefix + "residual.weight"], p[prefix + "residual.bias"])
    y = conv_temporal_graphical(x, p[prefix + "gcn.conv.weight"], A_eff)
    y = F.conv2d(y, p[prefix + "tcn.weight"], p[prefix + "tcn.bias"], padding=(1, 0)) + r
    if dw_s:
        keep = torch.as_tensor(tables.map[lvl + 1][:, 1], dtype=torch.long)
        y = y[:, :, :, keep]
    y = nearest_t(y, dw_t)
    return F.leaky_relu(y, 0.2) if mask is None else y * mask


def discriminator_forward(p, x, labels, cfg, tables, collect=None, masks=None):
    """discriminator.py:52-74.
    `masks` (checker-only): one tensor of LeakyReLU slopes per block, taken from ANOTHER evaluation of the same network (the CUDA
    path's).  The critic is piecewise linear; with the activation pattern pinned, its gradients are compared arithmetic against
    arithmetic - otherwise every pre-activation within the forward error of zero flips its slope between 1 and 0.2 and the gradient
    rel-L2 is ~0.8 * sqrt(fraction flipped), which measures the conditioning of LeakyReLU at zero, not the kernels."""
    N, C, T, V = x.size()
    c = p["label_emb.weight"][labels]
    c = c.view(N, -1, 1, 1).repeat(1, 1, T, V)
    x = torch.cat((c, x), 1)
    for i, spec in enumerate(d_block_table(cfg)):
        A_eff = torch.as_tensor(tables.As[spec[2]], dtype=x.dtype) * p["edge_importance.%d" % i]
        x = d_block(x, A_eff, p, "st_gcn_networks.%d." % i, spec, tables, None if masks is None else masks[i])
        if collect is not None:
            collect.append(x)
    x = F.avg_pool2d(x, x.size()[2:]).view(N, -1)
    return F.linear(x, p["fcn.weight"], p["fcn.bias"])


def gradient_penalty(pd, real, fake, labels, alpha, cfg, tables, return_grad=False, masks=None):
    """kinetic-gan.py:94-114 with alpha (N,1,1,1) supplied by the caller (host RNG at :97).  `masks`: see discriminator_forward."""
    inter = (alpha * real + (1 - alpha) * fake).requires_grad_(True)
    d_inter = discriminator_forward(pd, inter, labels, cfg, tables, masks=masks)
    ones = torch.ones_like(d_inter)
    (grads,) = torch.autograd.grad(d_inter, inter, ones, create_graph=True, retain_graph=True, only_inputs=True)
    flat = grads.reshape(grads.size(0), -1)
    gp = ((flat.norm(2, dim=1) - 1) ** 2).mean()
    return (gp, grads) if return_grad else gp


def d_loss_fn(pg, pd, real, labels, z, alpha, noises, cfg, tables, state_out=None, per_sample_loop=False):
    """kinetic-gan.py:140-152 (G forward in training mode: BN batch statistics)."""
    fake = generator_forward(pg, z, labels, cfg, tables, noises, True, state_out, per_sample_loop=per_sample_loop)
    real_v = discriminator_forward(pd, real, labels, cfg, tables)
    fake_v = discriminator_forward(pd, fake, labels, cfg, tables)
    gp = gradient_penalty(pd, real.detach(), fake.detach(), labels, alpha, cfg, tables)
    return -real_v.mean() + fake_v.mean() + cfg.lambda_gp * gp, gp, fake


def g_loss_fn(pg, pd, labels, z, noises, cfg, tables, state_out=None, per_sample_loop=False):
    """kinetic-gan.py:167-171."""
    fake = generator_forward(pg, z, labels, cfg, tables, noises, True, state_out, per_sample_loop=per_sample_loop)
    return -discriminator_forward(pd, fake, labels, cfg, tables).mean()


def is_trainable(key):
    return not (key.endswith("running_mean") or key.endswith("running_var") or key.endswith("num_batches_tracked"))


class Adam:
    """torch.optim.Adam(lr, betas=(b1,b2)) defaults eps=1e-8, no weight decay (kinetic-gan.py:77-78),
    restated functionally over a parameter dict."""

    def __init__(self, params, lr, b1, b2, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps, self.t = lr, b1, b2, eps, 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items() if is_trainable(k)}
        self.v = {k: torch.zeros_like(v) for k, v in params.items() if is_trainable(k)}

    def step(self, params, grads):
        self.t += 1
        bc1, bc2 = 1 - self.b1 ** self.t, 1 - self.b2 ** self.t
        with torch.no_grad():
            for k, g in grads.items():
                if g is None:
                    continue
                self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
                self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                denom = (self.v[k].sqrt() / (bc2 ** 0.5)).add_(self.eps)
                params[k].addcdiv_(self.m[k], denom, value=-self.lr / bc1)


class Trainer:
    """Functional re-enactment of the loop body kinetic-gan.py:137-174 on parameter dicts."""

    def __init__(self, cfg, pg, pd, tables=None, per_sample_loop=False):
        self.cfg, self.tables = cfg, tables or SkeletonTables(cfg.dataset)
        self.pg = {k: (v.clone().requires_grad_(True) if is_trainable(k) else v.clone()) for k, v in pg.items()}
        self.pd = {k: v.clone().requires_grad_(True) for k, v in pd.items()}
        self.opt_g = Adam(self.pg, cfg.lr, cfg.b1, cfg.b2)
        self.opt_d = Adam(self.pd, cfg.lr, cfg.b1, cfg.b2)
        self.loop = per_sample_loop

    def _apply_state(self, st):
        for k, v in st.items():
            self.pg[k] = v

    def iteration(self, i, real, labels, z, alpha, noises_d, noises_g=None):
        """One pass of kinetic-gan.py:137-174; `i` is the batch index tested at :160."""
        cfg = self.cfg
        st = {}
        d_loss, gp, _ = d_loss_fn(self.pg, self.pd, real, labels, z, alpha, noises_d, cfg, self.tables, st, self.loop)
        # the reference's d_loss.backward() also back-propagates into G; those grads are zeroed at :157
        # before any optimizer reads them, so only D's grads are algorithmic
        kd = list(self.pd)
        gd = torch.autograd.grad(d_loss, [self.pd[k] for k in kd], allow_unused=True)
        self._apply_state(st)
        self.opt_d.step(self.pd, dict(zip(kd, gd)))
        g_loss = None
        if i % cfg.n_critic == 0:
            st = {}
            g_loss = g_loss_fn(self.pg, self.pd, labels, z, noises_g, cfg, self.tables, st, self.loop)
            kg = [k for k in self.pg if is_trainable(k)]
            gg = torch.autograd.grad(g_loss, [self.pg[k] for k in kg], allow_unused=True)
            self._apply_state(st)
            self.opt_g.step(self.pg, dict(zip(kg, gg)))
        return d_loss.detach(), (g_loss.detach() if g_loss is not None else None), gp.detach()
