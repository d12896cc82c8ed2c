"""Graph-convolution operator with the reference's module surface (models/init_gan/tgcn.py:6-68).

Same ctor / forward signature, same `conv.weight` parameter (shape, init, state_dict key); the
arithmetic is refolded: because `conv` has no bias, "1x1 conv to K*C_out channels,
then einsum('nkctv,kvw->nctw')" equals ONE contraction over (k, c_in) of the adjacency-mixed input
XA[n,k,ci,t,w] = sum_v x[n,ci,t,v] A[k,v,w] - kgan_adjmix_fwd followed by kgan_tapconv_fwd with K
channel-block taps - so the K-times larger intermediate of the reference is never written."""
import torch
import torch.nn as nn

from ... import functional as KF
from ...geometry import TapConvGeom


class ConvTemporalGraphical(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, t_kernel_size=1, t_stride=1, t_padding=0, t_dilation=1,
                 bias=False):
        super().__init__()
        self.kernel_size = kernel_size
        # parameter container only (shape/init/keys of tgcn.py:48-55); the cuDNN conv is never called
        self.conv = nn.Conv2d(in_channels, out_channels * kernel_size, kernel_size=(t_kernel_size, 1),
                              padding=(t_padding, 0), stride=(t_stride, 1), dilation=(t_dilation, 1), bias=bias)
        self._t = (t_kernel_size, t_stride, t_padding, t_dilation)
        self._geoms = {}

    def _geom(self, t, w):
        g = self._geoms.get((t, w))
        if g is None:
            kt, st, pad, dil = self._t
            g = TapConvGeom(self.conv.in_channels, self.conv.out_channels // self.kernel_size, t, w, K=self.kernel_size,
                            kt=kt, pad=pad, stride=st, dil=dil)
            self._geoms[(t, w)] = g
        return g

    def _label_geoms(self, n_lab, c, t, w):
        key = ("lab", n_lab, c, t, w)
        g = self._geoms.get(key)
        if g is None:
            K, kc, cin = self.kernel_size, self.conv.out_channels, self.conv.in_channels
            assert self._t == (1, 1, 0, 1) and n_lab + c == cin
            g = (TapConvGeom(n_lab, kc, 1, 1, w_cin=cin, w_ic0=0),                         # label channels -> (N, K*C_out)
                 TapConvGeom(c, kc // K, t, w, K=K, w_cin=cin, w_ic0=n_lab))               # data channels
            self._geoms[key] = g
        return g

    def forward_with_labels(self, x, A, label_emb, support=None):
        """The critic's first layer (discriminator.py:57-64) without the label planes: the input of the reference is
        cat(label_emb tiled over (T, V), x).  The tiled channels are constant over (t, v), so their share of the output is
            L[n, c, w] = sum_k (W[kC+c, :n_cls] . e[n]) * sum_v A[k, v, w]
        - a per-sample bias times the column sums of A - added (broadcast along T) in the epilogue of the GEMM over the
        C data channels only.  Exact up to summation order."""
        assert A.size(0) == self.kernel_size
        K, n, n_lab = self.kernel_size, x.size(0), label_emb.size(1)
        g_lab, g_dat = self._label_geoms(n_lab, x.size(1), x.size(2), A.size(2))
        b = KF.TapConv.apply(label_emb.reshape(n, n_lab, 1, 1), self.conv.weight, g_lab)      # (N, K*C_out, 1, 1)
        b = b.view(n, K, -1).transpose(1, 2).reshape(n, -1, 1, K)                             # (N, C_out, 1, K)
        L = KF.AdjMix.apply(b, A.sum(1).unsqueeze(0))                                         # (N, C_out, 1, W); dense adjacency gradient (K*W entries)
        xa = KF.AdjMix.apply(x, A, support)
        out = KF.TapConvEp.apply(xa, self.conv.weight, None, L, g_dat, KF.ACT_NONE)
        return out, A

    def _conv_first_geom(self, t, v):
        key = ("cf", t, v)
        g = self._geoms.get(key)
        if g is None:
            kt, st, pad, dil = self._t
            g = self._geoms[key] = TapConvGeom(self.conv.in_channels, self.conv.out_channels, t, v, kt=kt, pad=pad, stride=st, dil=dil)
        return g

    def forward(self, x, A, support=None):
        """`support` (optional, not in the reference): constant mask of the entries of A that can ever be non-zero (the base
        adjacency behind `A_base * edge_importance`); the adjacency gradient is then evaluated there only.  Without it - the
        reference's signature, any dense / learnable A - every entry of dA is computed (functional.AdjMix)."""
        assert A.size(0) == self.kernel_size
        c_in, c_out = self.conv.in_channels, self.conv.out_channels // self.kernel_size
        # Both evaluation orders of tgcn.py:61-66 are the same bilinear map; they differ in the intermediate that crosses HBM:
        # K*C_in channels on W joints (adjacency first) or K*C_out channels on V joints (convolution first, the reference's
        # order).  The generator halves its channels in every block, the critic doubles them: pick the smaller one.
        if c_out * A.size(1) < c_in * A.size(2) and self._t == (1, 1, 0, 1):
            y = KF.TapConv.apply(x, self.conv.weight, self._conv_first_geom(x.size(2), x.size(3)))      # (N, K*C_out, T, V)
            # out[c, t, w] = sum_k sum_v y[k*C + c, t, v] A[k, v, w]: the adjoint-product member of the adjacency family
            out = KF.AdjMixDx.apply(y, A.transpose(1, 2), None if support is None else support.transpose(1, 2))
            return self._with_bias(out, A), A
        xa = KF.AdjMix.apply(x, A, support)                         # (N, K*C_in, T, W)
        out = KF.TapConv.apply(xa, self.conv.weight, self._geom(x.size(2), A.size(2)))
        return self._with_bias(out, A), A

    def _with_bias(self, out, A):
        if self.conv.bias is not None:
            # never used by Kinetic-GAN (bias=False at generator.py:132 / discriminator.py:96): the K*C_out conv
            # biases reach the output through the column sums of A; tiny host-side torch ops
            b = torch.einsum("kc,kw->cw", self.conv.bias.view(self.kernel_size, -1), A.sum(1))
            out = out + b.view(1, b.size(0), 1, b.size(1))
        return out
