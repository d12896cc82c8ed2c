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
from ..geometry import GcnFusedGeom, TapConvGeom, UnfoldedTcnGeom, mean_table, nearest_src, select_table
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
        self._shared_adj = None          # see share_adjacency()

    def share_adjacency(self):
        """Context manager for a training step that calls the critic several times on UNCHANGED parameters and differentiates the
        calls in one backward sweep (kinetic-gan.py:146-154: D(real), D(fake), D(x_hat), then d_loss.backward()): the effective
        adjacencies `A[lvl] * edge_importance[i]` and their per-block joint selections are computed by the first call and shared by
        the others, instead of ~10 tiny elementwise / index launches per block and call (forward and backward)."""
        return _SharedAdjacency(self)

    def _effective_adjacency(self, i, A, importance):
        sh = self._shared_adj
        if sh is None:
            return A * importance
        if i not in sh:
            sh[i] = A * importance
        return sh[i]

    @property
    def A(self):
        return [getattr(self, "_A%d" % i) for i in range(self.graph.lvls)]

    def forward(self, x, labels):
        N, C, T, V = x.size()
        c = KF.RoundTF32.apply(self.label_emb(labels))   # (N, n_cls); the (N, n_cls, T, V) planes are never built (identity in fp32 mode)
        A = self.A
        last = len(self.st_gcn_networks) - 1
        # Blocks chained here hand the LeakyReLU slope of a block's output to the NEXT block's backward (functional.GcnRes /
        # TapConvEp(act_bwd=False)): possible when every later block has a residual branch (all of the reference's do) and each
        # block output has exactly one consumer - the next block.  The last block's output goes to the pooling: it keeps its own.
        chain = all(b._res != "none" for b in list(self.st_gcn_networks)[1:])
        for i, (gcn, importance) in enumerate(zip(self.st_gcn_networks, self.edge_importance)):
            kw = dict(pad_joints=i < last, mask_input=chain and i > 0, act_bwd=not (chain and i < last))
            if i == 0 and gcn._res == "none":
                x, _ = gcn(x, self._effective_adjacency(i, A[gcn.lvl], importance), label_emb=c, **kw)   # label channels folded analytically (I3)
            else:
                x = KF.LabelConcat.apply(c, x) if i == 0 else x
                x, _ = gcn(x, self._effective_adjacency(i, A[gcn.lvl], importance), **kw)
        # global pooling + prediction (discriminator.py:68-72)
        key = (x.size(1), x.size(2), x.size(3))
        if key not in self._head:
            self._head[key] = (mean_table(x.size(2), x.size(3)), TapConvGeom(x.size(1), 1, 1, 1))
        pool, geom = self._head[key]
        x = KF.PlaneSpmm.apply(x, pool)
        validity = KF.TapConvEp.apply(x, self.fcn.weight, self.fcn.bias, None, geom, KF.ACT_NONE)
        return validity.view(N, -1)


class _SharedAdjacency:
    def __init__(self, D):
        self.D = D

    def __enter__(self):
        self.prev, self.D._shared_adj = self.D._shared_adj, {}
        for b in self.D.st_gcn_networks:
            b._sel_cache = None
        return self

    def __exit__(self, *exc):
        self.D._shared_adj = self.prev
        for b in self.D.st_gcn_networks:
            b._sel_cache = None


# Layout policy of the tensor-core path (module-level switches so that benchmarks can A/B them).  The TMA-fed tap
# convolution (csrc/tapconv_tma.cu) needs every tap to be a position shift by a multiple of 4:
PAD_JOINTS = True          # inside the critic, carry round_up(V, 4) joints where that costs <= 15 % (11 -> 12): the dummy joint has
                           # zero adjacency rows / columns, never reaches a real joint and is dropped by the next joint selection
SELECT_THEN_CONV = True    # residual 1x1 conv of a down-sampling block: gather the kept frames / joints first, then a plain 1x1 conv
UNFOLD_STRIDED_TCN = True  # temporal conv of a down-sampling block: time-unfold the kept frames' operands (geometry.UnfoldedTcnGeom)


def _pad4(v):
    vp = (v + 3) // 4 * 4
    return vp if PAD_JOINTS and vp != v and vp <= 1.15 * v else v


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
        self._sel_cache = None           # (A object, plan key, selected A): reused while Discriminator.share_adjacency() hands in the same A

    def _plan(self, T, V, Vx, pad_out, device):
        """Geometry of one call: T frames, V graph joints, Vx >= V joints carried by the input tensor (the extra ones are
        dummies), output joints padded to a multiple of 4 when `pad_out` and cheap."""
        key = (T, V, Vx, pad_out, str(device))
        p = self._plans.get(key)
        if p is None:
            co, ci = self.tcn.out_channels, self.gcn.conv.in_channels
            t_conv = (T + 2 * self._pad - (self._kt - 1) - 1) // self._stride + 1
            t_sel = nearest_src(t_conv, self.dw_t)
            # joint selection commutes with everything between the graph conv's adjacency product and the block output
            # (the temporal conv, bias, residual add and activation act per joint), so it is folded into the adjacency:
            # the graph conv runs with A[:, :, keep] and only ever produces the kept joints
            v_keep = [int(v) for v in self.graph.map[self.lvl + 1][:, 1]] if self.dw_s else list(range(V))
            W = len(v_keep)
            Wp = _pad4(W) if pad_out else W
            # source joint of every output joint; a dummy output joint takes the input's dummy joint when there is one (its
            # value is irrelevant but finite), else nothing (-1: zero)
            v_src = v_keep + [V if Vx > V else -1] * (Wp - W)
            plain_t = t_sel == list(range(T)) and self._stride == 1
            if plain_t or not UNFOLD_STRIDED_TCN or (len(t_sel) * Wp) % 4:
                tcn = TapConvGeom(co, co, T, Wp, kt=self._kt, pad=self._pad, stride=self._stride, t_sel=t_sel)
            else:
                tcn = UnfoldedTcnGeom(co, co, T, Wp, self._kt, self._pad, self._stride, 1, t_sel)
            res = sel = None
            ident = plain_t and v_src == list(range(Vx))
            if self._res == "conv":
                if ident or not SELECT_THEN_CONV:
                    res = TapConvGeom(ci, co, T, Vx, kt=1, stride=self._stride, t_sel=t_sel, v_keep=v_src)
                else:
                    sel = select_table(T, Vx, [t * self._stride for t in t_sel], v_src)
                    res = TapConvGeom(ci, co, len(t_sel), Wp, kt=1)
            elif self._res == "identity" and not ident:
                sel = select_table(T, Vx, t_sel, v_src)
            # A (K, V, V) -> (K, Vx, Wp): kept columns, zero rows for the input's dummy joints, zero columns for the output's
            cols = torch.tensor(v_keep + [V] * (Wp - W), dtype=torch.long, device=device)
            rows = torch.tensor(list(range(V)) + [V] * (Vx - V), dtype=torch.long, device=device)
            amap = None if (W == V and Wp == V and Vx == V) else (rows, cols)
            # constant support of the adjacency this block multiplies with (base skeleton adjacency, same row / column selection):
            # d loss / d edge_importance = gA * A_base, so the adjacency gradient is only evaluated where A_base != 0
            base = torch.tensor(self.graph.As[self.lvl] != 0, dtype=torch.float32, device=device)
            if amap is not None:
                base = torch.nn.functional.pad(base, (0, 1, 0, 1)).index_select(1, rows).index_select(2, cols)
            p = self._plans[key] = (tcn, res, sel, amap, base.contiguous())
        return p

    def forward(self, x, A, label_emb=None, pad_joints=False, mask_input=False, act_bwd=True):
        """`label_emb` (optional, not in the reference): (N, n_cls) label embedding standing for the first n_cls input
        channels, which the reference materialises as constant planes (discriminator.py:57-60); x then holds only the
        data channels.  Only valid for a block without residual branch (the critic's first block).
        `pad_joints` (optional, not in the reference; used by Discriminator.forward): x may carry dummy joints beyond the
        graph's V and the output may be padded likewise - see PAD_JOINTS above.
        `mask_input`, `act_bwd` (optional, not in the reference; set by Discriminator.forward when it chains the blocks): this
        block's backward also applies the LeakyReLU slope of its INPUT (the previous block's output) / leaves the slope of its own
        OUTPUT to the next block - see functional.GcnRes.  Defaults: plain autograd semantics."""
        V = A.size(1)
        assert x.size(3) >= V and (pad_joints or x.size(3) == V)
        tcn, res, sel, amap, support = self._plan(x.size(2), V, x.size(3), pad_joints, x.device)
        A_in = A
        if amap is not None:
            c = self._sel_cache
            if c is not None and c[0] is A and c[1] is amap:
                A = c[2]
            else:
                A = torch.nn.functional.pad(A, (0, 1, 0, 1)).index_select(1, amap[0]).index_select(2, amap[1])
                self._sel_cache = (A_in, amap, A)
        if label_emb is not None:
            assert self._res == "none" and not mask_input
            g, _ = self.gcn.forward_with_labels(x, A, label_emb, support)
            r = None
        elif self._res == "none":
            assert not mask_input, "mask_input needs a residual branch (functional.GcnRes)"
            g, _ = self.gcn(x, A, support)
            r = None
        else:
            # graph conv and residual branch as one autograd node: their input gradients are joined (and, with mask_input, multiplied
            # by the LeakyReLU slope of x) inside the kernel that finishes the graph-conv branch (functional.GcnRes)
            assert self.gcn.conv.bias is None and self.gcn._t == (1, 1, 0, 1)
            # (the joint node hands on the - selected - block input `r`; a residual CONV runs inside the temporal conv's kernel below)
            # ... and, in front of a strided temporal conv, stores the graph conv's result directly in the time-unfolded layout that conv reads
            g, r = KF.GcnRes.apply(x, A, self.gcn.conv.weight, None, None, self.gcn._geom(x.size(2), A.size(2)), None, sel, support, mask_input,
                                   tcn.unfold if isinstance(tcn, UnfoldedTcnGeom) else None,
                                   self._fused_geom(x.size(2), x.size(3), A.size(2), support) if KF.wants_fused_gcn(self.gcn.conv.weight) else None)
        if self._res == "none" and isinstance(tcn, UnfoldedTcnGeom):
            g = KF.PlaneSpmm.apply(g, tcn.unfold)
        if self._res == "conv":
            # temporal conv + residual 1x1 conv + both biases + LeakyReLU in one launch: the residual is an extra K panel of the accumulator
            x = KF.TcnRes.apply(g, self.tcn.weight, self.tcn.bias, r, self.residual.weight, self.residual.bias, tcn, res, KF.ACT_LRELU, act_bwd)
        else:
            x = KF.TapConvEp.apply(g, self.tcn.weight, self.tcn.bias, r, tcn, KF.ACT_LRELU, act_bwd)
        return x, A_in

    def _fused_geom(self, T, Vx, W, support):
        """One-kernel graph conv (adjacency product inside the GEMM, geometry.GcnFusedGeom) for passes without weight gradients.
        Used where its 128-row tiles are full - output planes that are a multiple of 128 positions, or small planes tiled by whole
        samples; on planes like 160 or 320 positions the last tile of every plane is mostly padding and the two-kernel formulation
        measured faster (profiles/r2_gcn_fused_ab.txt): None there."""
        key = ("fused", T, Vx, W)
        if key not in self._plans:
            p_out = T * W
            g = None
            if p_out % 128 == 0 or p_out <= 128:
                nnz = int((support != 0).sum(1).max().item())
                g = GcnFusedGeom(self.gcn.conv.in_channels, self.tcn.out_channels, T, Vx, W, self.gcn.kernel_size, max(nnz, 1))
            self._plans[key] = g
        return self._plans[key]

    def downsample_s(self, tensor):
        """Kept for API parity (discriminator.py:139-142); the forward pass folds it into the adjacency."""
        keep = [int(v) for v in self.graph.map[self.lvl + 1][:, 1]]
        return KF.PlaneSpmm.apply(tensor, select_table(tensor.size(2), tensor.size(3), list(range(tensor.size(2))), keep))
