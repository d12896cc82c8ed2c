"""TEST INFRASTRUCTURE: a torch-CPU emulation of every C-ABI primitive (include/kgan.h), descriptor
for descriptor.  It lets the CPU test-suite exercise all host logic of the product - geometry
tables, descriptors, the autograd Function families and their double-backward closure, the module
surface, the trainer and the DDP plumbing - without a GPU.  It is installed by monkeypatching
`kinetic-gan_b200.ops` inside tests only; the product has no such path (ops.py raises without CUDA).
The `-m gpu` tests then check the real kernels against these same semantics and against the oracle."""
import numpy as np
import torch

import kgan_b200 as kgan

ops = kgan.ops


def _act(v, act):
    if act == ops.ACT_LRELU:
        return torch.where(v > 0, v, 0.2 * v)
    if act == ops.ACT_TANH:
        return torch.tanh(v)
    return v


def _w_index(desc, g, tap):
    oc = torch.arange(desc.co)[:, None]
    ic = torch.arange(desc.ck)[None, :]
    return g * desc.g_w + desc.tap_w_off[tap] + oc * desc.w_oc + ic * desc.w_ic


def _gathered(xin, desc, g, tap, pmap):
    row = pmap[desc.tap_row[tap]]
    valid = (row >= 0).to(xin.dtype)
    ch0 = g * desc.g_in + desc.tap_in_ch[tap]
    return xin[:, ch0:ch0 + desc.ck, :][:, :, row.clamp(min=0)] * valid


def tapconv_fwd(x, w, desc, bias=None, add=None, act=0):
    n = x.shape[0]
    assert x.shape[1] == desc.c_in_total and x.shape[2] * x.shape[3] == desc.p_in
    xin = x.reshape(n, desc.c_in_total, desc.p_in)
    wf = w.reshape(-1)
    pmap = torch.from_numpy(desc.pmap).long()
    plane = getattr(desc, "p_out_plane", 0) or desc.p_out          # position-block groups (include/kgan.h p_out_plane / g_pout)
    g_pout = getattr(desc, "g_pout", 0)
    out = torch.zeros(n, desc.c_out_total, plane, dtype=x.dtype)
    for g in range(desc.groups):
        acc = torch.zeros(n, desc.co, desc.p_out, dtype=x.dtype)
        for tap in range(desc.ntap):
            acc = acc + torch.einsum("oi,nip->nop", wf[_w_index(desc, g, tap)], _gathered(xin, desc, g, tap, pmap))
        o0, q0 = g * desc.g_out, g * g_pout
        if bias is not None:
            acc = acc + bias[o0:o0 + desc.co].view(1, -1, 1)
        if add is not None:
            a = add
            if a.shape[2] == 1 and desc.t_out > 1:       # broadcast along T (add_period = V)
                a = a.expand(n, desc.c_out_total, desc.t_out, desc.v_out)
            acc = acc + a.reshape(n, desc.c_out_total, plane)[:, o0:o0 + desc.co, q0:q0 + desc.p_out]
        out[:, o0:o0 + desc.co, q0:q0 + desc.p_out] = _act(acc, act)
    return out.view(n, desc.c_out_total, desc.t_out, desc.v_out)


def tapconv_fwd_noise(x, w, desc, noise, nw, bias=None, add=None, act=0):
    return None          # the emulation always takes the two-pass formulation (same arithmetic)


def tapconv_fwd_res(x, w, desc, x2, w2, desc2, bias=None, bias2=None, act=0):
    return tapconv_fwd(x, w, desc, bias, tapconv_fwd(x2, w2, desc2, bias2), act)


def gcn_fused_fwd(x, A, w, fused, table=None):
    return None          # the emulation always takes the two-kernel formulation (same arithmetic)


def tapconv_fwd_scatter(x, w, desc, table):
    return None


def tapconv_wgrad(x, gout, desc, w_shape, out=None):
    if out is not None:                       # accumulate into the caller's buffer (kgan_tapconv_wgrad `accumulate`)
        with torch.no_grad():
            out.add_(tapconv_wgrad(x, gout, desc, w_shape))
        return out
    n = x.shape[0]
    xin = x.reshape(n, desc.c_in_total, desc.p_in)
    go = gout.reshape(n, desc.c_out_total, desc.p_out)
    pmap = torch.from_numpy(desc.pmap).long()
    dw = torch.zeros(int(np.prod(w_shape)), dtype=x.dtype)
    for g in range(desc.groups):
        o0 = g * desc.g_out
        for tap in range(desc.ntap):
            d = torch.einsum("nop,nip->oi", go[:, o0:o0 + desc.co], _gathered(xin, desc, g, tap, pmap))
            dw.index_put_((_w_index(desc, g, tap).reshape(-1),), d.reshape(-1), accumulate=True)
    return dw.view(w_shape)


def adjmix_fwd(x, A, sel=None):
    n, c, t, v = x.shape
    k, _, w = A.shape
    out = torch.einsum("nctv,kvw->nkctw", x, A).reshape(n, k * c, t, w)
    return out if sel is None else (out, plane_spmm(x, sel))


def adjmix_bwd_x(g, A, add=None, mask_src=None, add_sel=None):
    k, v, w = A.shape
    n, kc, t, _ = g.shape
    out = torch.einsum("nkctw,kvw->nctv", g.reshape(n, k, kc // k, t, w), A)
    if add is not None and add_sel is not None:                 # compact residual gradient through the selection's adjoint
        assert add_sel.inverse_gather() is not None
        add = plane_spmm(add, add_sel.T)
    if add is not None:
        out = out + add
    if mask_src is not None:
        out = out * torch.where(mask_src > 0, torch.ones_like(mask_src), torch.full_like(mask_src, 0.2))
    return out


def adjmix_bwd_a(x, g, k, mask=None):
    n, c, t, v = x.shape
    gA = torch.einsum("nctv,nkctw->kvw", x, g.reshape(n, k, c, t, g.shape[3]))
    return gA if mask is None else gA * (mask != 0).to(gA.dtype)


def epilogue_fwd(a, b=None, bias=None, nw=None, noise=None, act=0):
    v = a
    if b is not None:
        v = v + b
    if bias is not None:
        v = v + bias.view(1, -1, 1, 1)
    if nw is not None:
        v = v + nw.view(1, -1, 1, 1) * noise
    return _act(v, act)


def act_bwd(gout, out, act):
    if act == ops.ACT_LRELU:
        return gout * torch.where(out > 0, torch.ones_like(out), torch.full_like(out, 0.2))
    if act == ops.ACT_TANH:
        return gout * (1 - out * out)
    return gout.clone()


def chan_reduce(g, mul=None):
    return (g if mul is None else g * mul).sum((0, 2, 3))


def plane_spmm(x, table):
    n, c, t, v = x.shape
    idx = torch.from_numpy(table.idx).long()
    wgt = torch.from_numpy(table.wgt).to(x.dtype)
    # float64 runs (gradcheck) use the exact dense weights rather than their float32 rounding
    if x.dtype == torch.float64:
        dense = torch.from_numpy(table.dense)
        return torch.einsum("qp,ncp->ncq", dense, x.reshape(n, c, -1)).reshape(n, c, table.t_out, table.v_out)
    xs = x.reshape(n, c, -1)[:, :, idx.clamp(min=0)] * ((idx >= 0).to(x.dtype) * wgt)
    return xs.sum(-1).reshape(n, c, table.t_out, table.v_out)


def plane_sum_t(x):
    return x.sum(2, keepdim=True)


def label_concat(e, x):
    n, c, t, v = x.shape
    return torch.cat((e.view(n, -1, 1, 1).expand(n, e.shape[1], t, v), x), 1).contiguous()


def label_split(g, ncls, need_e=True, need_x=True):
    return (g[:, :ncls].sum((2, 3)) if need_e else None), (g[:, ncls:].contiguous() if need_x else None)


def bn_stats(x, running_mean=None, running_var=None, eps=1e-5, momentum=0.1):
    mean = x.mean((0, 2, 3))
    var = x.var((0, 2, 3), unbiased=False)
    cnt = x.numel() // x.shape[1]
    with torch.no_grad():
        if running_mean is not None:
            running_mean.mul_(1 - momentum).add_(momentum * mean)
        if running_var is not None:
            running_var.mul_(1 - momentum).add_(momentum * var * cnt / max(cnt - 1, 1))
    return mean, torch.rsqrt(var + eps)


def bn_apply(x, mean, rstd, gamma, beta):
    v = lambda t: t.view(1, -1, 1, 1)
    return (x - v(mean)) * v(rstd) * v(gamma) + v(beta)


def bn_epilogue_fwd(x, mean, rstd, gamma, beta, b=None, nw=None, noise=None, act=0):
    return epilogue_fwd(bn_apply(x, mean, rstd, gamma, beta), b, None, nw, noise, act)


def bn_bwd(gy, x, mean, rstd, gamma):
    v = lambda t: t.view(1, -1, 1, 1)
    xh = (x - v(mean)) * v(rstd)
    cnt = x.numel() // x.shape[1]
    s1, s2 = gy.sum((0, 2, 3)), (gy * xh).sum((0, 2, 3))
    gx = v(gamma * rstd) * (gy - v(s1) / cnt - xh * v(s2) / cnt)
    return gx, s2, s1


def adam_step(p, g, m, v, lr, b1, b2, eps, step, grad_scale=1.0):
    with torch.no_grad():
        gr = g * grad_scale
        m.mul_(b1).add_(gr, alpha=1 - b1)
        v.mul_(b2).addcmul_(gr, gr, value=1 - b2)
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        p.addcdiv_(m, v.sqrt() / (bc2 ** 0.5) + eps, value=-lr / bc1)


def interpolate(alpha, x, y):
    a = alpha.view(-1, 1, 1, 1)
    return a * x + (1 - a) * y


NAMES = ["tapconv_fwd", "tapconv_fwd_noise", "tapconv_fwd_res", "gcn_fused_fwd", "tapconv_fwd_scatter", "tapconv_wgrad", "adjmix_fwd", "adjmix_bwd_x", "adjmix_bwd_a", "epilogue_fwd", "act_bwd",
         "chan_reduce", "plane_spmm", "plane_sum_t", "label_concat", "label_split", "bn_stats", "bn_apply", "bn_epilogue_fwd", "bn_bwd", "adam_step",
         "interpolate"]


def install(monkeypatch):
    g = globals()
    for name in NAMES:
        monkeypatch.setattr(ops, name, g[name])
