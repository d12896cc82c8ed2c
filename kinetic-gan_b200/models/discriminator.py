"""Critic with the reference's module surface (models/discriminator.py): same class names, ctor and
forward signatures and state_dict keys; the arithmetic runs on the libkgan sm_100a kernels.

Per block (discriminator.py:125-136)   lrelu( interp_T( downsample_s( tcn(gcn(x, A)) + res(x) ) ) )
is evaluated as two fused launches plus the adjacency product:
    xa  = adjmix(x, A[:, :, keep])                      kgan_adjmix_fwd   (kept joints only: selection folded into A)
    g   = tapconv(xa, gcn.conv.weight)                  kgan_tapconv_fwd  (K channel-block taps)
    r   = tapconv(x, residual.weight) + residual.bias   kgan_tapconv_fwd  (only at kept frames / joints)
    out = lrelu(tapconv(g, tcn.weight) + tcn.bias + r)  kgan_tapconv_fwd  (3 temporal taps, only kept outputs)
Joint / frame selection commutes with the linear maps, so dropped outputs are never computed."""
import torch
import torch.nn as nn

from .. import functional as KF
from ..geometry import TapConvGeom, mean_table, nearest_src, select_table
from .init_gan.graph_h36m import Graph_h36m
from .init_gan.graph_ntu import graph_ntu
from .init_gan.tgcn import ConvTemporalGraphical


class Discriminator(nn.Module):
    def __init__(self, in_channels, n_classes, t_size, latent, edge_importance_weighting=True, dataset='ntu', **kwargs):
        super().__init__()
        self.graph = graph_ntu() if dataset == 'ntu' else Graph_h36m()
        # the reference keeps `.cuda()` tensors in a plain list (discriminator.py:19); non-persistent buffers
        # follow `.to()` / DDP and stay out of the state_dict, as in the reference
        for i, Al in enumerate(self.graph.As):
            self.register_buffer("_A%d" % i, torch.tensor(Al, dtype=torch.float32), persistent=False)
        spatial_kernel_size = [A.size(0) for A in self.A]
        temporal_kernel_size = [3 for _ in self.A]
        kernel_size = (temporal_kernel_size, spatial_kernel_size)
        self.t_size = t_size
        self.st_gcn_networks = nn.ModuleList((
            st_gcn(in_channels + n_classes, 32, kernel_size, 1, graph=self.graph, lvl=0, dw_s=True, dw_t=t_size, residual=False, **kwargs),
            st_gcn(32, 64, kernel_size, 1, graph=self.graph, lvl=1, dw_s=False, dw_t=t_size, **kwargs),
            st_gcn(64, 128, kernel_size, 1, graph=self.graph, lvl=1, dw_s=True, dw_t=int(t_size / 2), **kwargs),
            st_gcn(128, 256, kernel_size, 1, graph=self.graph, lvl=2, dw_s=False, dw_t=int(t_size / 4), **kwargs),
            st_gcn(256, 512, kernel_size, 1, graph=self.graph, lvl=2, dw_s=True, dw_t=int(t_size / 8), **kwargs),
            st_gcn(512, latent, kernel_size, 1, graph=self.graph, lvl=3, dw_s=False, dw_t=int(t_size / 16), **kwargs),
        ))
        if edge_importance_weighting:
            self.edge_importance = nn.ParameterList([nn.Parameter(torch.ones(self.A[i.lvl].size())) for i in self.st_gcn_networks])
        else:
            self.edge_importance = [1] * len(self.st_gcn_networks)
        self.label_emb = nn.Embedding(n_classes, n_classes)
        self.fcn = nn.Linear(latent, 1)
        self._head = {}

    @property
    def A(self):
        return [getattr(self, "_A%d" % i) for i in range(self.graph.lvls)]

    def forward(self, x, labels):
        N, C, T, V = x.size()
        c = self.label_emb(labels)                       # (N, n_cls); the (N, n_cls, T, V) planes are never built
        A = self.A
        for i, (gcn, importance) in enumerate(zip(self.st_gcn_networks, self.edge_importance)):
            if i == 0 and gcn._res == "none":
                x, _ = gcn(x, A[gcn.lvl] * importance, label_emb=c)      # label channels folded analytically (I3)
            else:
                x = KF.LabelConcat.apply(c, x) if i == 0 else x
                x, _ = gcn(x, A[gcn.lvl] * importance)
        # global pooling + prediction (discriminator.py:68-72)
        key = (x.size(1), x.size(2), x.size(3))
        if key not in self._head:
            self._head[key] = (mean_table(x.size(2), x.size(3)), TapConvGeom(x.size(1), 1, 1, 1))
        pool, geom = self._head[key]
        x = KF.PlaneSpmm.apply(x, pool)
        validity = KF.TapConvEp.apply(x, self.fcn.weight, self.fcn.bias, None, geom, KF.ACT_NONE)
        return validity.view(N, -1)


class st_gcn(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, graph=None, lvl=3, dropout=0, residual=True,
                 dw_s=False, dw_t=64):
        super().__init__()
        assert len(kernel_size) == 2
        assert kernel_size[0][lvl] % 2 == 1
        padding = ((kernel_size[0][lvl] - 1) // 2, 0)
        self.graph, self.lvl, self.dw_s, self.dw_t = graph, lvl, dw_s, dw_t
        self.gcn = ConvTemporalGraphical(in_channels, out_channels, kernel_size[1][lvl])
        self.tcn = nn.Conv2d(out_channels, out_channels, (kernel_size[0][lvl], 1), (stride, 1), padding)   # parameter container
        self._kt, self._stride, self._pad = kernel_size[0][lvl], stride, padding[0]
        if not residual:
            self._res, self.residual = "none", (lambda x: 0)
        elif (in_channels == out_channels) and (stride == 1):
            self._res, self.residual = "identity", (lambda x: x)
        else:
            self._res = "conv"
            self.residual = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=(stride, 1))       # parameter container
        self.l_relu = nn.LeakyReLU(0.2, inplace=True)
        self._plans = {}

    def _plan(self, T, V):
        p = self._plans.get((T, V))
        if p is None:
            co, ci = self.tcn.out_channels, self.gcn.conv.in_channels
            t_conv = (T + 2 * self._pad - (self._kt - 1) - 1) // self._stride + 1
            t_sel = nearest_src(t_conv, self.dw_t)
            v_keep = [int(v) for v in self.graph.map[self.lvl + 1][:, 1]] if self.dw_s else list(range(V))
            # joint selection commutes with everything between the graph conv's adjacency product and the block output
            # (the temporal conv, bias, residual add and activation act per joint), so it is folded into the adjacency:
            # the graph conv runs with A[:, :, keep] and only ever produces the kept joints
            tcn = TapConvGeom(co, co, T, len(v_keep), kt=self._kt, pad=self._pad, stride=self._stride, t_sel=t_sel)
            res = None
            if self._res == "conv":
                res = TapConvGeom(ci, co, T, V, kt=1, stride=self._stride, t_sel=t_sel, v_keep=v_keep)
            elif self._res == "identity" and (t_sel != list(range(T)) or len(v_keep) != V):
                res = select_table(T, V, t_sel, v_keep)
            keep = torch.tensor(v_keep, dtype=torch.long) if len(v_keep) != V else None
            p = self._plans[(T, V)] = (tcn, res, keep)
        return p

    def forward(self, x, A, label_emb=None):
        """`label_emb` (optional, not in the reference): (N, n_cls) label embedding standing for the first n_cls input
        channels, which the reference materialises as constant planes (discriminator.py:57-60); x then holds only the
        data channels.  Only valid for a block without residual branch (the critic's first block)."""
        tcn, res, keep = self._plan(x.size(2), A.size(2))
        A_in = A
        if keep is not None:
            if keep.device != A.device:
                keep = keep.to(A.device)
                self._plans[(x.size(2), A.size(2))] = (tcn, res, keep)
            A = A.index_select(2, keep)                    # (K, V, V_keep): dropped joints are never computed
        if label_emb is not None:
            assert self._res == "none"
            g, _ = self.gcn.forward_with_labels(x, A, label_emb)
            return KF.TapConvEp.apply(g, self.tcn.weight, self.tcn.bias, None, tcn, KF.ACT_LRELU), A_in
        if self._res == "none":
            r = None
        elif self._res == "identity":
            r = x if res is None else KF.PlaneSpmm.apply(x, res)
        else:
            r = KF.TapConvEp.apply(x, self.residual.weight, self.residual.bias, None, res, KF.ACT_NONE)
        g, _ = self.gcn(x, A)
        x = KF.TapConvEp.apply(g, self.tcn.weight, self.tcn.bias, r, tcn, KF.ACT_LRELU)
        return x, A_in

    def downsample_s(self, tensor):
        """Kept for API parity (discriminator.py:139-142); the forward pass folds it into the conv's position map."""
        keep = [int(v) for v in self.graph.map[self.lvl + 1][:, 1]]
        return KF.PlaneSpmm.apply(tensor, select_table(tensor.size(2), tensor.size(3), list(range(tensor.size(2))), keep))
