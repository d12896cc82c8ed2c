"""Generator with the reference's module surface (models/generator.py): same class names, ctor and
forward signatures, parameter init and state_dict keys; the arithmetic runs on the libkgan sm_100a
kernels.  Differences allowed by the drop-in contract: adjacency tensors are
non-persistent buffers instead of a `.cuda()` list (generator.py:47), the per-block noise is drawn on
`x.device` instead of 'cuda:0' (generator.py:179), and the mapping network runs batched instead of
one sample at a time (generator.py:84-85; same maths)."""
import numpy as np
import torch
import torch.nn as nn

from .. import functional as KF
from ..geometry import TapConvGeom, resample_table, upsample_matrix
from .init_gan.graph_h36m import Graph_h36m
from .init_gan.graph_ntu import graph_ntu
from .init_gan.tgcn import ConvTemporalGraphical


class NoiseInjection(nn.Module):
    def __init__(self, channel):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1, channel, 1, 1))

    def forward(self, image, noise):
        return KF.NoiseAct.apply(image, None, noise, self.weight, KF.ACT_NONE)


class Mapping_Net(nn.Module):
    def __init__(self, latent=1024, mlp=4):
        super().__init__()
        layers = []
        for i in range(mlp):
            linear = nn.Linear(latent, latent)          # parameter container (generator.py:28-30)
            linear.weight.data.normal_()
            linear.bias.data.zero_()
            layers.append(linear)
            layers.append(nn.LeakyReLU(0.2))
        self.mlp = nn.Sequential(*layers)
        self._geom = TapConvGeom(latent, latent, 1, 1)

    def forward(self, x):
        """x: (L,) as in the reference's per-sample call, or a whole batch (N, L)."""
        single = x.dim() == 1
        h = x.reshape(-1, x.size(-1), 1, 1)
        for m in self.mlp:
            if isinstance(m, nn.Linear):
                h = KF.TapConvEp.apply(h, m.weight, m.bias, None, self._geom, KF.ACT_LRELU)
        h = h.view(h.size(0), -1)
        return h[0] if single else h


class Generator(nn.Module):
    def __init__(self, in_channels, out_channels, n_classes, t_size, mlp_dim=4, edge_importance_weighting=True,
                 dataset='ntu', **kwargs):
        super().__init__()
        self.graph = graph_ntu() if dataset == 'ntu' else Graph_h36m()
        for i, Al in enumerate(self.graph.As):
            self.register_buffer("_A%d" % i, torch.tensor(Al, dtype=torch.float32), persistent=False)
        spatial_kernel_size = [A.size(0) for A in self.A]
        temporal_kernel_size = [3 for i, _ in enumerate(self.A)]
        kernel_size = (temporal_kernel_size, spatial_kernel_size)
        self.t_size = t_size
        self.mlp = Mapping_Net(in_channels + n_classes, mlp_dim)
        self.st_gcn_networks = nn.ModuleList((
            st_gcn(in_channels + n_classes, 512, kernel_size, 1, graph=self.graph, lvl=3, bn=False, residual=False, up_s=False, up_t=1, **kwargs),
            st_gcn(512, 256, kernel_size, 1, graph=self.graph, lvl=3, up_s=False, up_t=int(t_size / 16), **kwargs),
            st_gcn(256, 128, kernel_size, 1, graph=self.graph, lvl=2, bn=False, up_s=True, up_t=int(t_size / 16), **kwargs),
            st_gcn(128, 64, kernel_size, 1, graph=self.graph, lvl=2, up_s=False, up_t=int(t_size / 8), **kwargs),
            st_gcn(64, 32, kernel_size, 1, graph=self.graph, lvl=1, bn=False, up_s=True, up_t=int(t_size / 4), **kwargs),
            st_gcn(32, out_channels, kernel_size, 1, graph=self.graph, lvl=1, up_s=False, up_t=int(t_size / 2), **kwargs),
            st_gcn(out_channels, out_channels, kernel_size, 1, graph=self.graph, lvl=0, bn=False, up_s=True, up_t=t_size, tan=True, **kwargs),
        ))
        if edge_importance_weighting:
            self.edge_importance = nn.ParameterList([nn.Parameter(torch.ones(self.A[i.lvl].size())) for i in self.st_gcn_networks])
        else:
            self.edge_importance = [1] * len(self.st_gcn_networks)
        self.label_emb = nn.Embedding(n_classes, n_classes)

    @property
    def A(self):
        return [getattr(self, "_A%d" % i) for i in range(self.graph.lvls)]

    def forward(self, x, labels, trunc=None, noises=None):
        """`noises` (optional, not in the reference): the 7 per-block noise tensors, for reproducible parity runs;
        by default each block draws torch.randn(N,1,T,V) on x.device in block order, as generator.py:179 does."""
        c = self.label_emb(labels)
        x = KF.RoundTF32.apply(torch.cat((c, x), -1))                   # identity in fp32 mode
        w = self.mlp(x)
        w = KF.RoundTF32.apply(self.truncate(w, 1000, trunc)) if trunc is not None else w   # Truncation trick on W
        x = w.view((*w.shape, 1, 1))
        A = self.A
        for i, (gcn, importance) in enumerate(zip(self.st_gcn_networks, self.edge_importance)):
            x, _ = gcn(x, A[gcn.lvl] * importance, noise=None if noises is None else noises[i])
        return x

    def fold_batchnorm(self):
        """Inference only: fold every eval-mode BatchNorm (generator.py:142,160 with running
        statistics) into the convolution in front of it - w' = w * gamma * rstd, b' = (b - mean) * gamma * rstd + beta - so
        that the eval-mode pass launches no BatchNorm kernel at all.  Used by generate.GeneratorRunner; the folds are
        dropped by `.train()`, `.to()` and `load_state_dict()`, and must be redone after the parameters change."""
        for blk in self.st_gcn_networks:
            blk.fold_batchnorm()
        return self

    def load_state_dict(self, *args, **kwargs):
        self.invalidate_derived()
        return super().load_state_dict(*args, **kwargs)

    def invalidate_derived(self):
        """Everything computed FROM the weights - eval-mode BatchNorm folds, the cached W-space mean of truncate() - is dropped, and
        `weights_version` moves on so that holders of derived state (generate.GeneratorRunner: folds, mean, captured CUDA graph)
        rebuild it.  Called by load_state_dict(), .to() / .float() (`_apply`) and .train()."""
        for blk in self.st_gcn_networks:
            blk._fold = None
        self._w_mean = None
        self.weights_version = getattr(self, "weights_version", 0) + 1

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if hasattr(self, "st_gcn_networks"):
            self.invalidate_derived()
        return out

    def train(self, mode=True):
        if mode and hasattr(self, "st_gcn_networks"):
            self.invalidate_derived()
        return super().train(mode)

    def truncate(self, w, mean, truncation):
        """generator.py:97-108: W-space truncation towards the mean of `mean` mapped N(0,1) latents (host RNG, as the
        reference); the 1000 per-sample MLP passes of the reference become one batched pass."""
        m = getattr(self, "_w_mean", None)          # generate.GeneratorRunner(cache_mean=True): one estimate for all calls
        if m is None:
            t = torch.as_tensor(np.random.normal(0, 1, (mean, *w.shape[1:])), dtype=w.dtype, device=w.device)
            m = self.mlp(KF.RoundTF32.apply(t)).mean(0, keepdim=True)
        return m + truncation * (w - m)

    def estimate_w_mean(self, mean=1000):
        """The W-space mean of generator.py:98-104 (host RNG draw of `mean` latents, batched mapping pass) as a tensor."""
        dev, dt = self.label_emb.weight.device, self.label_emb.weight.dtype
        with torch.no_grad():
            t = torch.as_tensor(np.random.normal(0, 1, (mean, self.mlp.mlp[0].in_features)), dtype=dt, device=dev)
            return self.mlp(KF.RoundTF32.apply(t)).mean(0, keepdim=True)


class st_gcn(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, graph=None, lvl=3, dropout=0, bn=True,
                 residual=True, up_s=False, up_t=64, tan=False):
        super().__init__()
        assert len(kernel_size) == 2
        assert kernel_size[0][lvl] % 2 == 1
        padding = ((kernel_size[0][lvl] - 1) // 2, 0)
        self.graph, self.lvl, self.up_s, self.up_t, self.tan = graph, lvl, up_s, up_t, tan
        self.gcn = ConvTemporalGraphical(in_channels, out_channels, kernel_size[1][lvl])
        tcn = [nn.Conv2d(out_channels, out_channels, (kernel_size[0][lvl], 1), (stride, 1), padding)]   # parameter containers
        if bn:
            tcn.append(nn.BatchNorm2d(out_channels))
        self.tcn = nn.Sequential(*tcn)
        self._kt, self._stride, self._pad, self._bn = kernel_size[0][lvl], stride, padding[0], bn
        if not residual:
            self._res, self.residual = "none", (lambda x: 0)
        elif (in_channels == out_channels) and (stride == 1):
            self._res, self.residual = "identity", (lambda x: x)
        else:
            self._res = "conv"
            self.residual = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=(stride, 1)),
                                          nn.BatchNorm2d(out_channels))
        self.noise = NoiseInjection(out_channels)
        self.l_relu = nn.LeakyReLU(0.2, inplace=True)
        self.tanh = nn.Tanh()
        self._plans = {}
        self._fold = None               # eval-mode (weight, bias) pairs with the BatchNorm folded in, see Generator.fold_batchnorm

    @staticmethod
    def _folded(conv, bn):
        sc = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
        return (conv.weight * sc.view(-1, 1, 1, 1)).contiguous(), ((conv.bias - bn.running_mean) * sc + bn.bias).contiguous()

    def fold_batchnorm(self):
        with torch.no_grad():
            fold = {}
            if self._bn:
                fold["tcn"] = self._folded(self.tcn[0], self.tcn[1])
            if self._res == "conv":
                fold["res"] = self._folded(self.residual[0], self.residual[1])
        self._fold = fold

    def train(self, mode=True):
        if mode:
            self._fold = None
        return super().train(mode)

    def _apply(self, fn, *args, **kwargs):
        self._fold = None
        return super()._apply(fn, *args, **kwargs)

    def _plan(self, T, V):
        p = self._plans.get((T, V))
        if p is None:
            U = upsample_matrix(self.graph.mapping[self.lvl], V, halve=(self.lvl == 2)) if self.up_s else None
            Vf = V if U is None else U.shape[1]
            up = None if (U is None and T == self.up_t) else resample_table(T, V, self.up_t, U)
            co, ci = self.tcn[0].out_channels, self.gcn.conv.in_channels
            tcn = TapConvGeom(co, co, self.up_t, Vf, kt=self._kt, pad=self._pad, stride=self._stride)
            res = TapConvGeom(ci, co, self.up_t, Vf, kt=1, stride=self._stride) if self._res == "conv" else None
            p = self._plans[(T, V)] = (up, tcn, res)
        return p

    def _support(self, device):
        """Constant support of this block's adjacency (functional.AdjMix): the base skeleton adjacency of its level."""
        s = self._plans.get(("support", str(device)))
        if s is None:
            s = self._plans[("support", str(device))] = torch.tensor(self.graph.As[self.lvl] != 0, dtype=torch.float32, device=device)
        return s

    def _batch_norm(self, bn, x):
        if bn.training:
            if bn.num_batches_tracked is not None:
                bn.num_batches_tracked.add_(1)
            return KF.BatchNormTrain.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, bn.momentum)
        return KF.BatchNormEval.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)

    def forward(self, x, A, noise=None):
        up, tcn, res = self._plan(x.size(2), x.size(3))
        fold = (self._fold or {}) if not self.training else {}
        x = x if up is None else KF.PlaneSpmm.apply(x, up)          # upsample_s + F.interpolate, one gather
        if self._res == "none":
            r = None
        elif self._res == "identity":
            r = x
        elif "res" in fold:
            r = KF.TapConvEp.apply(x, fold["res"][0], fold["res"][1], None, res, KF.ACT_NONE)
        else:
            r = KF.TapConvEp.apply(x, self.residual[0].weight, self.residual[0].bias, None, res, KF.ACT_NONE)
            r = self._batch_norm(self.residual[1], r)
        g, A = self.gcn(x, A, self._support(x.device))
        if "tcn" in fold and not torch.is_grad_enabled():
            # inference (BatchNorm folded, no autograd graph): temporal conv + bias + residual + noise + activation as ONE kernel where the
            # layer runs on the operand-building kernel (kgan_tapconv_fwd_tf32_noise); None: not eligible, the two passes below
            if noise is None:
                noise = torch.randn(g.size(0), 1, self.up_t, g.size(3), device=g.device)
            out = KF.ops.tapconv_fwd_noise(KF._c(g), fold["tcn"][0], tcn.fwd, KF._c(noise), KF._c(self.noise.weight.reshape(-1)), fold["tcn"][1],
                                           None if r is None else KF._c(r), KF.ACT_TANH if self.tan else KF.ACT_LRELU)
            if out is not None:
                return out, A
            z = KF.TapConvEp.apply(g, fold["tcn"][0], fold["tcn"][1], None, tcn, KF.ACT_NONE)
        elif "tcn" in fold:
            z = KF.TapConvEp.apply(g, fold["tcn"][0], fold["tcn"][1], None, tcn, KF.ACT_NONE)
        else:
            z = KF.TapConvEp.apply(g, self.tcn[0].weight, self.tcn[0].bias, None, tcn, KF.ACT_NONE)
            bn = self.tcn[1] if self._bn else None
            if bn is not None and bn.training:
                # batch statistics, then normalise + residual + noise + activation in one pass (functional.BnNoiseAct)
                if noise is None:
                    noise = torch.randn(z.size(0), 1, z.size(2), z.size(3), device=z.device)
                if bn.num_batches_tracked is not None:
                    bn.num_batches_tracked.add_(1)
                out = KF.BnNoiseAct.apply(z, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, bn.momentum, r, noise, self.noise.weight,
                                          KF.ACT_TANH if self.tan else KF.ACT_LRELU)
                return out, A
            if bn is not None:
                z = self._batch_norm(bn, z)
        if noise is None:
            noise = torch.randn(z.size(0), 1, z.size(2), z.size(3), device=z.device)
        out = KF.NoiseAct.apply(z, r, noise, self.noise.weight, KF.ACT_TANH if self.tan else KF.ACT_LRELU)
        return out, A

    def upsample_s(self, tensor):
        """Kept for API parity (generator.py:185-200); the forward pass folds it into one plane gather."""
        U = upsample_matrix(self.graph.mapping[self.lvl], tensor.size(3), halve=(self.lvl == 2))
        return KF.PlaneSpmm.apply(tensor, resample_table(tensor.size(2), tensor.size(3), tensor.size(2), U))
