"""Differentiable operators over the libkgan kernels.

Every operator family is closed under differentiation: the backward of each Function is itself
written with Functions of the same family, so `autograd.grad(..., create_graph=True)` followed by
`.backward()` - the WGAN-GP gradient penalty, kinetic-gan.py:104-113,154 - stays entirely on the
device kernels (what torch does for the reference through convolution_backward /
_convolution_double_backward / bmm backward).

  tap convolution   F(x, w)   |  Dgrad(g, w) = dF/dx^T g  |  Wgrad(x, g) = dF/dw^T g     (bilinear)
  adjacency product M(x, A)   |  Dx(g, A)                 |  DA(x, g)                    (bilinear)
  activation mask   ActGrad(g, y) = g * act'(y)            (linear in g; the mask is piecewise constant)
"""
import torch
from torch.autograd import Function

from . import ops
from .geometry import PlaneTable, mean_table  # noqa: F401
from .ops import ACT_LRELU, ACT_NONE, ACT_TANH  # noqa: F401


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


# Two rules keep a backward sweep from doing work nobody reads (torch's built-in convolution / bmm nodes follow both, custom
# Functions have to do it themselves):
#   * undefined incoming gradients are NOT materialised as zeros (`set_materialize_grads(False)`): in the gradient penalty
#     the forward nodes of D(x_hat) hang off the second-order graph only through LeakyReLU masks, whose derivative is zero -
#     with materialisation on, autograd would run their whole backward (data, weight and adjacency gradients) on zeros;
#   * inside `data_grads_only()` a backward computes the gradient of its data input only.  `ctx.needs_input_grad` says
#     whether an input REQUIRES grad, not whether this particular `autograd.grad(..., inputs=x)` call asked for it: the
#     first-order pass of the penalty (kinetic-gan.py:104) wants d/dx_hat alone, yet every node would also produce its
#     weight / bias / adjacency gradients only for the engine to drop them.
_data_only = False


class data_grads_only:
    """Context manager: Function backwards executed inside skip parameter-side gradients (weights, biases, adjacency,
    label embedding).  Used around the `autograd.grad(outputs, inputs=x_hat, create_graph=True)` call of the gradient
    penalty; the graph built there still carries every parameter dependence (TapConvDgrad / AdjMixDx nodes are created
    with the live weights), so the following `backward()` is unchanged."""

    def __enter__(self):
        global _data_only
        self.prev, _data_only = _data_only, True

    def __exit__(self, *exc):
        global _data_only
        _data_only = self.prev


def _want(ctx, i):
    """Parameter-side gradient i of this node wanted?"""
    return ctx.needs_input_grad[i] and not _data_only


# Weight gradients straight into `p.grad`.  The weight-gradient kernels accumulate with atomics anyway; when the trainer has seated
# every parameter's .grad as a view of its flat gradient buffer (wgan_gp.FlatParams) a backward sweep can let them add into that
# view (kgan_tapconv_wgrad `accumulate`) and hand autograd nothing - instead of a fresh dW tensor per convolution, a memset, and an
# `at::add_` launch by AccumulateGrad (10 % of the step's launches in round 1).  Opt-in (`param_grads_in_place()` around
# `loss.backward()`): torch.autograd.grad(..., inputs=[w]) callers want the tensor back.
_in_place = False


class param_grads_in_place:
    def __enter__(self):
        global _in_place
        self.prev, _in_place = _in_place, True

    def __exit__(self, *exc):
        global _in_place
        _in_place = self.prev


def _wgrad(x, go, geom, w):
    """dW of a tap convolution as a differentiable Function - or, inside param_grads_in_place() during a first-order sweep, added
    in place to w.grad (returns None: autograd has nothing left to accumulate)."""
    if (_in_place and not torch.is_grad_enabled() and isinstance(w, torch.nn.Parameter) and w.grad is not None and w.grad.is_contiguous()
            and w.grad.shape == w.shape):
        ops.tapconv_wgrad(_c(x), _c(go), geom.fwd, tuple(w.shape), out=w.grad)
        return None
    return TapConvWgrad.apply(x, go, geom, w.shape)


# Forward passes whose weight gradients will never be asked for (inference, a frozen critic, the interpolate pass of the gradient penalty -
# its forward nodes hang off the second-order graph only through LeakyReLU slopes): the graph conv may then run as ONE kernel with the
# adjacency product inside the GEMM (ops.gcn_fused_fwd) instead of materialising the mixed tensor that only a weight gradient would read.
# If a weight gradient is asked for after all, the backward recomputes the mixed tensor: correct, merely slower.
_no_wgrad_hint = False


class no_weight_grads_expected:
    def __enter__(self):
        global _no_wgrad_hint
        self.prev, _no_wgrad_hint = _no_wgrad_hint, True

    def __exit__(self, *exc):
        global _no_wgrad_hint
        _no_wgrad_hint = self.prev


def wants_fused_gcn(weight):
    """True where a graph conv's mixed tensor would have no reader: no autograd graph is being recorded, the weight is frozen, or the
    caller declared the pass free of weight gradients (no_weight_grads_expected)."""
    return _no_wgrad_hint or not torch.is_grad_enabled() or not weight.requires_grad


_SUM_T = {}


def sum_t_table(t, v):
    from .geometry import sum_t_table as make

    if (t, v) not in _SUM_T:
        _SUM_T[(t, v)] = make(t, v)
    return _SUM_T[(t, v)]


# ------------------------------------------------------------------------------------------------
# tap convolution family
# ------------------------------------------------------------------------------------------------
class TapConv(Function):
    @staticmethod
    def forward(ctx, x, w, geom):
        ctx.geom = geom
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, w)
        return ops.tapconv_fwd(_c(x), _c(w), geom.fwd)

    @staticmethod
    def backward(ctx, go):
        if go is None:
            return None, None, None
        x, w = ctx.saved_tensors
        go = _c(go)
        gx = TapConvDgrad.apply(go, w, ctx.geom) if ctx.needs_input_grad[0] else None
        gw = _wgrad(x, go, ctx.geom, w) if _want(ctx, 1) else None
        return gx, gw, None


class TapConvDgrad(Function):
    @staticmethod
    def forward(ctx, go, w, geom):
        ctx.geom = geom
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(go, w)
        return ops.tapconv_fwd(_c(go), _c(w), geom.dgrad)

    @staticmethod
    def backward(ctx, h):
        if h is None:
            return None, None, None
        go, w = ctx.saved_tensors
        h = _c(h)
        ggo = TapConv.apply(h, w, ctx.geom) if ctx.needs_input_grad[0] else None
        gw = _wgrad(h, go, ctx.geom, w) if _want(ctx, 1) else None
        return ggo, gw, None


class TapConvWgrad(Function):
    @staticmethod
    def forward(ctx, x, go, geom, w_shape):
        ctx.geom = geom
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, go)
        return ops.tapconv_wgrad(_c(x), _c(go), geom.fwd, tuple(w_shape))

    @staticmethod
    def backward(ctx, hw):
        if hw is None:
            return None, None, None, None
        x, go = ctx.saved_tensors
        hw = _c(hw)
        gx = TapConvDgrad.apply(go, hw, ctx.geom) if ctx.needs_input_grad[0] else None
        ggo = TapConv.apply(x, hw, ctx.geom) if ctx.needs_input_grad[1] else None
        return gx, ggo, None, None


class TapConvEp(Function):
    """out = act(F(x, w) + bias[c] + add): the convolution with its fused epilogue (one kernel).
    `act_bwd=False` (critic blocks chained by Discriminator.forward): the incoming gradient has ALREADY been multiplied by act'(out)
    - the consumer of `out` (GcnRes with mask_input=True) applies the slope where it produces that gradient - so this node's
    backward starts from it as is.  Only valid when that consumer is the sole user of `out`."""

    @staticmethod
    def forward(ctx, x, w, bias, add, geom, act, act_bwd=True):
        ctx.geom, ctx.act = geom, (act if act_bwd else ACT_NONE)
        ctx.add_bcast = add is not None and add.shape[2] == 1 and geom.t_out > 1
        ctx.set_materialize_grads(False)
        out = ops.tapconv_fwd(_c(x), _c(w), geom.fwd, bias, None if add is None else _c(add), act)
        ctx.save_for_backward(x, w, out)
        return out

    @staticmethod
    def backward(ctx, go):
        if go is None:
            return (None,) * len(ctx.needs_input_grad)
        x, w, out = ctx.saved_tensors
        go = _c(go)
        gz = ActGrad.apply(go, out, ctx.act) if ctx.act != ACT_NONE else go
        gx = TapConvDgrad.apply(gz, w, ctx.geom) if ctx.needs_input_grad[0] else None
        gw = _wgrad(x, gz, ctx.geom, w) if _want(ctx, 1) else None
        gb = ChanSum.apply(gz) if _want(ctx, 2) else None
        ga = None
        if ctx.needs_input_grad[3]:
            ga = SumT.apply(gz) if ctx.add_bcast else gz
        return (gx, gw, gb, ga, None, None, None)[:len(ctx.needs_input_grad)]


class TcnRes(Function):
    """out = act(F_tcn(g, w) + b + F_res(xs, w_res) + b_res): the temporal conv of a critic block with its residual 1x1 conv as an extra
    K panel of the same accumulator (discriminator.py:128-136; kgan_tapconv_fwd_tf32_res) - where the pair is not eligible the two
    convolutions run separately, the first as the `add` operand of the second.  The backward is the sum of the two TapConvEp backwards
    (one activation mask, one bias reduction for both biases).  `act_bwd`: see TapConvEp."""

    @staticmethod
    def forward(ctx, g, w, b, xs, w_res, b_res, geom, res_geom, act, act_bwd=True):
        ctx.geom, ctx.res_geom, ctx.act = geom, res_geom, (act if act_bwd else ACT_NONE)
        ctx.set_materialize_grads(False)
        g, xs, w, w_res = _c(g), _c(xs), _c(w), _c(w_res)
        out = ops.tapconv_fwd_res(g, w, geom.fwd, xs, w_res, res_geom.fwd, b, b_res, act)
        if out is None:
            r = ops.tapconv_fwd(xs, w_res, res_geom.fwd, b_res)
            out = ops.tapconv_fwd(g, w, geom.fwd, b, r, act)
        ctx.save_for_backward(g, w, xs, w_res, out)
        return out

    @staticmethod
    def backward(ctx, go):
        nig = ctx.needs_input_grad
        if go is None:
            return (None,) * len(nig)
        g, w, xs, w_res, out = ctx.saved_tensors
        go = _c(go)
        gz = ActGrad.apply(go, out, ctx.act) if ctx.act != ACT_NONE else go
        gg = TapConvDgrad.apply(gz, w, ctx.geom) if nig[0] else None
        gw = _wgrad(g, gz, ctx.geom, w) if _want(ctx, 1) else None
        gb = ChanSum.apply(gz) if (_want(ctx, 2) or _want(ctx, 5)) else None
        gxs = TapConvDgrad.apply(gz, w_res, ctx.res_geom) if nig[3] else None
        gwr = _wgrad(xs, gz, ctx.res_geom, w_res) if _want(ctx, 4) else None
        return (gg, gw, gb if _want(ctx, 2) else None, gxs, gwr, gb if _want(ctx, 5) else None, None, None, None, None)[:len(nig)]


# ------------------------------------------------------------------------------------------------
# adjacency product family
# ------------------------------------------------------------------------------------------------
class AdjMix(Function):
    """`support` (optional constant (K, V, W) tensor, not differentiated): the entries of A that can be non-zero for EVERY value
    of the parameters behind A - the st_gcn blocks pass the skeleton's base adjacency (`A_eff = A_base * edge_importance`:
    d loss / d edge_importance = gA * A_base, so gA is only ever read where A_base != 0).  The adjacency gradient is then
    evaluated on that support only (a few dozen dot products instead of K*V*W).  None (the default of a standalone
    ConvTemporalGraphical call with an arbitrary, possibly dense and learnable A): every entry of gA is computed."""

    @staticmethod
    def forward(ctx, x, A, support=None):
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, A)
        ctx.support = support
        return ops.adjmix_fwd(_c(x), _c(A))

    @staticmethod
    def backward(ctx, go):
        if go is None:
            return (None,) * len(ctx.needs_input_grad)
        x, A = ctx.saved_tensors
        go = _c(go)
        gx = AdjMixDx.apply(go, A, ctx.support) if ctx.needs_input_grad[0] else None
        gA = AdjMixDA.apply(x, go, A.shape[0], ctx.support) if _want(ctx, 1) else None
        return (gx, gA, None)[:len(ctx.needs_input_grad)]


class AdjMixDx(Function):
    """gx = Dx(g, A) [+ add] [* leaky_relu'(mask_src)] - the optional terms are the fused epilogue of kgan_adjmix_bwd_x_fused
    (see GcnRes); `mask_src` is a constant here (the slope is piecewise constant in it).  `add_sel` (a selection PlaneTable): `add` is
    given in the selection's compact layout and enters through the selection's adjoint (kgan_adjmix_bwd_x_fused_sel)."""

    @staticmethod
    def forward(ctx, g, A, support=None, add=None, mask_src=None, add_sel=None):
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(g, A, mask_src)
        ctx.support, ctx.add_sel = support, add_sel
        return ops.adjmix_bwd_x(_c(g), _c(A), None if add is None else _c(add), None if mask_src is None else _c(mask_src.detach()), add_sel)

    @staticmethod
    def backward(ctx, h):
        if h is None:
            return (None,) * len(ctx.needs_input_grad)
        g, A, mask_src = ctx.saved_tensors
        h = _c(h)
        if mask_src is not None:
            h = ActGrad.apply(h, mask_src, ACT_LRELU)
        gg = AdjMix.apply(h, A, ctx.support) if ctx.needs_input_grad[0] else None
        gA = AdjMixDA.apply(h, g, A.shape[0], ctx.support) if _want(ctx, 1) else None
        nig = ctx.needs_input_grad                                  # as long as the argument list of this apply() call
        gadd = None
        if len(nig) > 3 and nig[3]:
            gadd = h if ctx.add_sel is None else PlaneSpmm.apply(h, ctx.add_sel)
        return (gg, gA, None, gadd, None, None)[:len(nig)]


class AdjMixDA(Function):
    """gA = dM/dA^T g, evaluated on the non-zeros of the constant mask `support` only (see AdjMix; entries outside it are
    returned as 0 and a cotangent hA of this output is never read there); support None: all entries."""

    @staticmethod
    def forward(ctx, x, g, k, support=None):
        mask = None if support is None else _c(support.detach())
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, g)
        return ops.adjmix_bwd_a(_c(x), _c(g), k, mask)

    @staticmethod
    def backward(ctx, hA):
        if hA is None:
            return None, None, None, None
        x, g = ctx.saved_tensors
        hA = _c(hA)
        gx = AdjMixDx.apply(g, hA) if ctx.needs_input_grad[0] else None
        gg = AdjMix.apply(x, hA) if ctx.needs_input_grad[1] else None
        return gx, gg, None, None


class GcnRes(Function):
    """The two consumers of a critic block's input as ONE node (discriminator.py:128-130: `self.gcn(x, A)` and `self.residual(x)`):

        g = TapConv(AdjMix(x, A), w_gcn)                                   graph conv, adjacency first (tgcn.py:61-66 refolded)
        r = TapConvEp(select(x), w_res, b_res)   |   select(x)   |   x     residual: 1x1 conv / identity, at the kept frames / joints

    Why one node: with two, autograd adds their input gradients with an elementwise kernel and the producer of x then applies its
    LeakyReLU slope with another - three passes over the largest tensors of the step.  Here the backward hands the residual branch's
    gradient to the kernel that finishes the graph-conv branch (kgan_adjmix_bwd_x_fused: product + add), and with `mask_input` also the
    slope of x itself: x is the previous block's LeakyReLU output, sign(x) = sign(pre-activation), so what this node returns IS the
    gradient w.r.t. that pre-activation and the producer (TapConvEp(act_bwd=False)) skips its own mask kernel.
    The backward is composed of the differentiable members of the two operator families, so the node is as closed under
    differentiation as they are (gradient penalty); intermediates (AdjMix(x, A), select(x)) are kept from the forward pass for a
    first-order sweep and recomputed as graph nodes when the sweep itself is recorded (create_graph)."""

    @staticmethod
    def forward(ctx, x, A, w_gcn, w_res, b_res, gcn_geom, res_geom, sel, support, mask_input, out_table=None, fused_geom=None):
        """`out_table` (optional PlaneTable): g is returned gathered through it - the time-unfolded layout a strided temporal conv
        reads (geometry.UnfoldedTcnGeom.unfold); where the table is a pure gather the graph conv's epilogue stores that layout
        directly (ops.tapconv_fwd_scatter), otherwise a gather kernel follows."""
        ctx.set_materialize_grads(False)
        ctx.gcn_geom, ctx.res_geom, ctx.sel, ctx.support, ctx.mask_input = gcn_geom, res_geom, sel, support, mask_input
        ctx.out_table = out_table
        x, A = _c(x), _c(A)
        g = xa = None
        # `fused_geom` (optional geometry.GcnFusedGeom): where no weight gradient is expected (frozen / no_grad / hinted pass) the whole
        # graph conv is one kernel with the adjacency product inside the GEMM, and the mixed tensor is not materialised
        # (the caller decides - grad mode is always off INSIDE a Function's forward - and passes fused_geom only then: wants_fused_gcn())
        if fused_geom is not None:
            g = ops.gcn_fused_fwd(x, A, _c(w_gcn), fused_geom, out_table)
        xs = None
        if g is None:
            if sel is not None:
                xa, xs = ops.adjmix_fwd(x, A, sel)             # the residual branch's input as a by-product of the same kernel
            else:
                xa = ops.adjmix_fwd(x, A)
            if out_table is not None:
                g = ops.tapconv_fwd_scatter(xa, _c(w_gcn), gcn_geom.fwd, out_table)
            if g is None:
                g = ops.tapconv_fwd(xa, _c(w_gcn), gcn_geom.fwd)
                if out_table is not None:
                    g = ops.plane_spmm(g, out_table)
        if xs is None:
            xs = x if sel is None else ops.plane_spmm(x, sel)
        r = ops.tapconv_fwd(xs, _c(w_res), res_geom.fwd, b_res) if w_res is not None else xs
        ctx.save_for_backward(x, A, w_gcn, w_res)
        ctx.xa, ctx.xs = xa, (xs if w_res is not None else None)
        if w_res is None and sel is None:
            r = x.view_as(x)              # a distinct output object (autograd marks outputs, not inputs)
        return g, r

    @staticmethod
    def backward(ctx, gg, gr):
        x, A, w_gcn, w_res = ctx.saved_tensors
        nig = ctx.needs_input_grad
        gx = gA = gw_gcn = gw_res = gb = None
        recorded = torch.is_grad_enabled()                       # create_graph: intermediates must be graph nodes of (x, A)
        # residual branch first: its input gradient is an operand of the kernel that closes the graph-conv branch
        gx_r = add_sel = None
        if gr is not None:
            gr = _c(gr)
            if w_res is not None:
                if _want(ctx, 3) or _want(ctx, 4):
                    xs = ctx.xs if not recorded else (x if ctx.sel is None else PlaneSpmm.apply(x, ctx.sel))
                    gw_res = _wgrad(xs, gr, ctx.res_geom, w_res) if _want(ctx, 3) else None
                    gb = ChanSum.apply(gr) if _want(ctx, 4) else None
                gxs = TapConvDgrad.apply(gr, w_res, ctx.res_geom) if nig[0] else None
            else:
                gxs = gr if nig[0] else None
            if gxs is not None:
                if ctx.sel is not None and gg is not None and nig[0] and ctx.sel.inverse_gather() is not None:
                    gx_r, add_sel = gxs, ctx.sel       # stays compact: the selection's adjoint is taken by the kernel that joins the branches
                else:
                    gx_r = gxs if ctx.sel is None else PlaneSpmm.apply(gxs, ctx.sel.T)
        mask = x if ctx.mask_input else None
        if gg is not None:
            gg = _c(gg)
            if ctx.out_table is not None:
                gg = PlaneSpmm.apply(gg, ctx.out_table.T)        # fold the gathered layout back (sum of the copies)
            if _want(ctx, 2):
                xa = AdjMix.apply(x, A, ctx.support) if (recorded or ctx.xa is None) else ctx.xa     # (fused forward: recomputed on demand)
                gw_gcn = _wgrad(xa, gg, ctx.gcn_geom, w_gcn)
            if nig[0] or _want(ctx, 1):
                g_xa = TapConvDgrad.apply(gg, w_gcn, ctx.gcn_geom)
                gA = AdjMixDA.apply(x, g_xa, A.shape[0], ctx.support) if _want(ctx, 1) else None
                gx = AdjMixDx.apply(g_xa, A, ctx.support, gx_r, mask, add_sel) if nig[0] else None
        elif gx_r is not None:
            gx = ActGrad.apply(gx_r, x, ACT_LRELU) if ctx.mask_input else gx_r
        return (gx, gA, gw_gcn, gw_res, gb, None, None, None, None, None, None, None)[:len(nig)]


# ------------------------------------------------------------------------------------------------
# activation mask, channel reductions
# ------------------------------------------------------------------------------------------------
class ActGrad(Function):
    """gz = go * act'(y), y = the block OUTPUT (LeakyReLU keeps the sign, so sign(y) == sign(pre-activation))."""

    @staticmethod
    def forward(ctx, go, y, act):
        ctx.act = act
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(go, y)
        return ops.act_bwd(_c(go), y, act)

    @staticmethod
    def backward(ctx, h):
        if h is None:
            return None, None, None
        go, y = ctx.saved_tensors
        h = _c(h)
        ggo = ActGrad.apply(h, y, ctx.act) if ctx.needs_input_grad[0] else None
        gy = None
        if ctx.needs_input_grad[1] and ctx.act == ACT_TANH:
            gy = -2.0 * y * go * h            # d/dy [go (1 - y^2)]; never reached by the WGAN-GP step (G is first-order only)
        return ggo, gy, None


class ChanSum(Function):
    """out[c] = sum_{n,t,v} g[n,c,t,v]  (bias gradients)."""

    @staticmethod
    def forward(ctx, g):
        ctx.shape = g.shape
        ctx.set_materialize_grads(False)
        return ops.chan_reduce(_c(g))

    @staticmethod
    def backward(ctx, h):
        return None if h is None else h.view(1, -1, 1, 1).expand(ctx.shape)


class ChanSumMul(Function):
    """out[c] = sum_{n,t,v} g[n,c,t,v] * m[n,0,t,v]  (NoiseInjection.weight gradient)."""

    @staticmethod
    def forward(ctx, g, m):
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(g, m)
        return ops.chan_reduce(_c(g), _c(m))

    @staticmethod
    def backward(ctx, h):
        if h is None:
            return None, None
        g, m = ctx.saved_tensors
        hv = h.view(1, -1, 1, 1)
        gg = hv * m if ctx.needs_input_grad[0] else None
        gm = (hv * g).sum(1, keepdim=True) if ctx.needs_input_grad[1] else None
        return gg, gm


class NoiseAct(Function):
    """out = act(a + b + nw[c] * noise[n,0,t,v]): `tcn(x) + res`, NoiseInjection and the activation of a
    generator block in one pass (generator.py:176-182)."""

    @staticmethod
    def forward(ctx, a, b, noise, nw, act):
        ctx.act = act
        out = ops.epilogue_fwd(_c(a), None if b is None else _c(b), None, _c(nw.reshape(-1)), _c(noise), act)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(out, noise)
        ctx.nw_shape = nw.shape
        return out

    @staticmethod
    def backward(ctx, go):
        if go is None:
            return None, None, None, None, None
        out, noise = ctx.saved_tensors
        gz = ActGrad.apply(_c(go), out, ctx.act) if ctx.act != ACT_NONE else _c(go)
        ga = gz if ctx.needs_input_grad[0] else None
        gb = gz if ctx.needs_input_grad[1] else None
        gnw = ChanSumMul.apply(gz, noise).view(ctx.nw_shape) if _want(ctx, 3) else None
        return ga, gb, None, gnw, None


class RoundTF32(Function):
    """Entry point of the tf32 path for a tensor that no libkgan kernel produced (latent / label-embedding rows, a truncated
    W): stored tf32-rounded like every activation the kernels write (include/kgan.h `out_tf32`), so that the tensor cores read
    it exactly.  Identity in fp32 mode; straight-through gradient (the rounding is piecewise constant + identity on its grid)."""

    @staticmethod
    def forward(ctx, x):
        ctx.set_materialize_grads(False)
        return ops.round_tf32(_c(x))

    @staticmethod
    def backward(ctx, go):
        return go


# ------------------------------------------------------------------------------------------------
# plane gather / scatter, label planes
# ------------------------------------------------------------------------------------------------
class PlaneSpmm(Function):
    @staticmethod
    def forward(ctx, x, table):
        ctx.table = table
        ctx.set_materialize_grads(False)
        return ops.plane_spmm(_c(x), table)

    @staticmethod
    def backward(ctx, go):
        if go is None:
            return None, None
        return PlaneSpmm.apply(_c(go), ctx.table.T), None


class SumT(Function):
    """(N, C, T, V) -> (N, C, 1, V), sum over frames: the adjoint of a per-joint term broadcast along T (TapConvEp `add` with one
    frame).  Its own adjoint is the broadcast, a plane gather with the transposed sum table."""

    @staticmethod
    def forward(ctx, x):
        ctx.tv = (x.shape[2], x.shape[3])
        ctx.set_materialize_grads(False)
        return ops.plane_sum_t(_c(x)) if x.shape[3] <= 32 else ops.plane_spmm(_c(x), sum_t_table(*ctx.tv))

    @staticmethod
    def backward(ctx, h):
        return None if h is None else PlaneSpmm.apply(_c(h), sum_t_table(*ctx.tv).T)


class LabelConcat(Function):
    """cat(label planes, x) along channels (discriminator.py:57-60) without the (N,n_cls,T,V) repeat."""

    @staticmethod
    def forward(ctx, e, x):
        ctx.ncls = e.shape[1]
        ctx.set_materialize_grads(False)
        return ops.label_concat(_c(e), _c(x))

    @staticmethod
    def backward(ctx, go):
        if go is None:
            return None, None
        ge, gx = LabelSplit.apply(_c(go), ctx.ncls)
        return (ge if _want(ctx, 0) else None), (gx if ctx.needs_input_grad[1] else None)


class LabelSplit(Function):
    @staticmethod
    def forward(ctx, g, ncls):
        ge, gx = ops.label_split(_c(g), ncls)
        return ge, gx

    @staticmethod
    def backward(ctx, he, hx):
        return LabelConcat.apply(_c(he), _c(hx)), None


# ------------------------------------------------------------------------------------------------
# BatchNorm2d
# ------------------------------------------------------------------------------------------------
class BatchNormTrain(Function):
    """nn.BatchNorm2d in training mode (generator.py:142,160); running stats are updated in place by the
    statistics kernel.  First-order backward only (nothing in the WGAN-GP step differentiates G twice)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, momentum):
        x = _c(x)
        mean, rstd = ops.bn_stats(x, running_mean, running_var, eps, momentum)
        ctx.save_for_backward(x, mean, rstd, gamma)
        return ops.bn_apply(x, mean, rstd, gamma, beta)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, mean, rstd, gamma = ctx.saved_tensors
        gx, gg, gb = ops.bn_bwd(_c(gy), x, mean, rstd, gamma)
        return gx, gg, gb, None, None, None, None


class BnNoiseAct(Function):
    """act(BatchNorm_train(z) + r + nw[c] * noise): normalisation with the batch statistics, residual add, NoiseInjection and the
    activation of a generator block (generator.py:160,176-182) as the statistics kernel + ONE pass over the tensor
    (kgan_bn_epilogue_fwd) instead of a normalise pass and a noise / activation pass.  First-order backward (as BatchNormTrain)."""

    @staticmethod
    def forward(ctx, z, gamma, beta, running_mean, running_var, eps, momentum, r, noise, nw, act):
        z = _c(z)
        mean, rstd = ops.bn_stats(z, running_mean, running_var, eps, momentum)
        out = ops.bn_epilogue_fwd(z, mean, rstd, gamma, beta, None if r is None else _c(r), _c(nw.reshape(-1)), _c(noise), act)
        ctx.save_for_backward(z, mean, rstd, gamma, out, noise)
        ctx.act, ctx.nw_shape = act, nw.shape
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        z, mean, rstd, gamma, out, noise = ctx.saved_tensors
        g = ops.act_bwd(_c(go), out, ctx.act) if ctx.act != ACT_NONE else _c(go)
        nig = ctx.needs_input_grad
        gz = gg = gb = None
        if nig[0] or nig[1] or nig[2]:
            gz, gg, gb = ops.bn_bwd(g, z, mean, rstd, gamma)
        gnw = ops.chan_reduce(g, noise).view(ctx.nw_shape) if nig[9] else None
        return gz, gg, gb, None, None, None, None, (g if nig[7] else None), None, gnw, None


class BatchNormEval(Function):
    """nn.BatchNorm2d in eval mode (generate.py:67): affine map with the running statistics."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps):
        rstd = torch.rsqrt(running_var + eps)
        ctx.save_for_backward(x, gamma, running_mean, rstd)
        return ops.bn_apply(_c(x), running_mean, rstd, gamma, beta)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, gamma, mean, rstd = ctx.saved_tensors
        sc = (gamma * rstd).view(1, -1, 1, 1)
        xh = (x - mean.view(1, -1, 1, 1)) * rstd.view(1, -1, 1, 1)
        return gy * sc, (gy * xh).sum((0, 2, 3)), gy.sum((0, 2, 3)), None, None, None
