"""Tensor-level wrappers of the C ABI (include/kgan.h): allocate outputs with torch, pass raw device
pointers and the current CUDA stream.  No arithmetic happens here, and there is no fallback - every
function ends in a libkgan.so kernel launch."""
import torch

from . import _lib
from ._lib import ACT_LRELU, ACT_NONE, ACT_TANH, PREC_FP32, PREC_TF32, PREC_X3  # noqa: F401

_precision = PREC_FP32
launches = 0       # number of libkgan kernels launched by this process (bench.py reports it)


def set_precision(name):
    """'fp32' = exact SIMT FMA path (rel-L2 <= 1e-5 vs the fp32 reference);
    'fp32x3' = the same accuracy class on the tensor cores: operands split hi + lo inside the TMA-fed kernels, three tcgen05
    kind::tf32 MMAs per product (KGAN_PREC_TF32X3), fp32 activations in HBM; shapes without a TMA-fed plan run the FMA kernels;
    'tf32' = tcgen05 kind::tf32 tensor-core path with fp32 accumulation (<= 1e-3)."""
    global _precision
    _precision = _PRECISIONS[name]


_PRECISIONS = {"fp32": PREC_FP32, "tf32": PREC_TF32, "fp32x3": PREC_X3}


def get_precision():
    return {v: k for k, v in _PRECISIONS.items()}[_precision]


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _rnd():
    """`out_tf32` of include/kgan.h: in tf32 mode every kernel stores the activations it produces tf32-rounded, so the tensor
    core's 19-bit operand read is exact (no truncation bias, exact on tf32-representable data)."""
    return 1 if _precision == PREC_TF32 else 0


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _chk(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), "kgan ops need contiguous float32 CUDA tensors"


def _count(n=1):
    global launches
    launches += n


# ---- optional per-launch timing (bench.py roofline pass): CUDA events on the launching stream around each call
_prof = None


def profile_start():
    global _prof
    _prof = []
    return _prof


def profile_stop(prof):
    """-> {family: {"ms": total, "n": launches, "flops": total algorithmic flops}} (synchronises)."""
    global _prof
    _prof = None
    torch.cuda.synchronize()
    out, sites = {}, {}
    for name, flops, e0, e1, sig, nbytes in prof:
        ms = e0.elapsed_time(e1)
        for table, key in ((out, name), (sites, name + " " + sig)):
            d = table.setdefault(key, {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0})
            d["ms"] += ms
            d["n"] += 1
            d["flops"] += flops
            d["bytes"] += nbytes
    out["_sites"] = sites
    return out


_sig = ""          # shape signature of the launch being issued (profiling only)
_nbytes = 0.0      # algorithmic bytes of the launch being issued: every operand tensor read or written once (profiling only)


def _io(*ts):
    global _nbytes
    if _prof is not None:
        _nbytes = float(sum(t.numel() * t.element_size() for t in ts if t is not None))


def _run(family, flops, fn, *args):
    """One kernel launch through the C ABI (+ CUDA events around it when profiling)."""
    global _sig, _nbytes
    if _prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.check(fn(*args), fn.__name__)
    _count()
    if _prof is not None:
        e1.record()
        _prof.append((family, flops, e0, e1, _sig, _nbytes))
        _sig, _nbytes = "", 0.0


def _tap_flops(desc, n):
    global _sig
    if _prof is not None:
        _sig = "n%d ck%d co%d tap%d g%d pin%d pout%d tma%d" % (n, desc.ck, desc.co, desc.ntap, desc.groups, desc.p_in, desc.p_out, desc.tma_mode)
    return 2.0 * n * desc.p_out * desc.co * desc.ck * desc.ntap * desc.groups


def _shape_sig(*ts):
    global _sig
    if _prof is not None:
        _sig = " ".join("x".join(str(d) for d in t.shape) for t in ts)


# ---- tf32 tensor-core path: packed weight images ----------------------------------------------------------------------
# Parameters keep a persistent packed image per (storage, geometry, batch class): it is refreshed by ONE batched launch after the
# optimizer step of ITS network (repack_weights) - inside a captured CUDA graph the tap convolutions then simply read it.  Any other
# tensor used as a "weight" (the second-order terms of the gradient penalty) is packed on the spot, cached only until the next
# optimizer step.
#   * freshness is tracked per flat parameter buffer (`_epochs`): the critic's Adam step does not make the generator's images stale;
#   * an image is re-packed by the batched launch only if it was used since the previous one, or was touched while a CUDA graph was
#     being captured (a replay reads it without coming through here); the others are marked stale and re-packed lazily;
#   * entries hold their parameter weakly: images of a model that is gone are dropped (no leak across models / tests).
import weakref

weights_epoch = 0          # bumped by every fused Adam step (kept for callers that key caches on "the weights changed")
_epochs = {}               # flat-buffer pointer -> epoch; None: parameters outside any registered flat buffer
_flats = []                # registered flat parameter buffers: (lo, hi) address ranges
_packed = {}               # temporaries: key -> (wp, desc, w)
_persist = {}              # parameters:  key -> _PackedParam
_batches = {}              # flat-buffer id -> _PackBatch


class _PackedParam:
    __slots__ = ("wref", "ptr", "desc", "cs", "wp", "epoch", "version", "flat", "used", "pinned")

    @property
    def w(self):
        return self.wref()


class _PackBatch:
    __slots__ = ("entries", "table", "uploaded", "arrays")


def register_flat(flat):
    """A trainer's flat parameter buffer (wgan_gp.FlatParams): its parameters' packed images follow this buffer's optimizer steps."""
    lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
    if (lo, hi) not in _flats:
        _flats.append((lo, hi))
    _epochs.setdefault(lo, 0)


def _flat_of(ptr):
    for lo, hi in _flats:
        if lo <= ptr < hi:
            return lo
    return None


def invalidate_packed_weights(flat=None):
    """The weights in `flat` (a registered flat buffer; None: all weights) were rewritten behind autograd's version counter."""
    global weights_epoch
    weights_epoch += 1
    if flat is None:
        for k in list(_epochs):
            _epochs[k] += 1
        _epochs[None] = _epochs.get(None, 0) + 1
    else:
        lo = flat.data_ptr()
        _epochs[lo] = _epochs.get(lo, 0) + 1
    _packed.clear()


def clear_temporary_packs():
    _packed.clear()


def _drop_dead():
    dead = [k for k, e in _persist.items() if e.wref() is None]
    for k in dead:
        del _persist[k]
    if dead:
        _batches.clear()


def repack_weights(flat):
    """Refresh, with one launch, the persistent packed images of the weights that live in the flat parameter buffer `flat` and were
    used since the last refresh (or are read by a captured CUDA graph); called by FlatParams.adam right after the optimizer kernel."""
    _drop_dead()
    lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
    mine = [e for e in _persist.values() if lo <= e.ptr < hi]
    ents = [e for e in mine if e.used or e.pinned]
    for e in mine:
        if not (e.used or e.pinned):
            e.epoch = -1                       # stale: re-packed on its next use
    if not ents:
        return
    b = _batches.get(lo)
    if b is None:
        b = _batches[lo] = _PackBatch()
        b.entries, b.table, b.uploaded, b.arrays = [], None, False, None
    if len(b.entries) != len(ents) or any(x is not y for x, y in zip(b.entries, ents)):
        import ctypes as C

        n = len(ents)
        descs = (_lib.TapConvDesc * n)(*[e.cs for e in ents])
        wv = (C.c_void_p * n)(*[e.ptr for e in ents])
        wpv = (C.c_void_p * n)(*[e.wp.data_ptr() for e in ents])
        b.entries, b.arrays, b.uploaded = ents, (descs, wv, wpv), False
        b.table = torch.empty(n * int(_lib.lib().kgan_tapconv_pack_item_bytes()), device=flat.device, dtype=torch.uint8)
    descs, wv, wpv = b.arrays
    _run('tapconv_pack', 0.0, _lib.lib().kgan_tapconv_pack_tf32_batched, len(ents), descs, wv, wpv, b.table.data_ptr(), 0 if b.uploaded else 1,
         _stream())
    b.uploaded = True
    ep = _epochs.get(lo, 0)
    for e in ents:
        w = e.wref()
        e.epoch, e.version, e.used = ep, (w._version if w is not None else -1), False


def _packed_weights(w, desc, cs, l, cache_image=True):
    """`cache_image=False`: `w` is not a weight but a tensor of the step (see _linear_wgrad): always packed afresh - the address of a recycled
    activation buffer says nothing about its contents."""
    cache = desc.__dict__.setdefault("_tf32_numel", {})       # eligibility / image size depend on the batch size too
    numel = cache.get((cs.n, cs.precision))
    if numel is None:
        numel = cache[(cs.n, cs.precision)] = int(l.kgan_tapconv_tf32_workspace(cs))
    if numel <= 0:
        return None
    if not cache_image:
        wp = torch.empty(numel, device=w.device, dtype=torch.float32)
        _run('tapconv_pack', 0.0, l.kgan_tapconv_pack_tf32, cs, w.data_ptr(), wp.data_ptr(), _stream())
        return wp
    if isinstance(w, torch.nn.Parameter):
        key = (w.data_ptr(), id(desc), numel)
        e = _persist.get(key)
        if e is None or e.desc is not desc or e.wref() is not w:
            e = _PackedParam()
            e.wref, e.ptr, e.desc, e.cs, e.epoch, e.version = weakref.ref(w), w.data_ptr(), desc, cs, -1, -1
            e.flat, e.used, e.pinned = _flat_of(w.data_ptr()), False, False
            e.wp = torch.empty(numel, device=w.device, dtype=torch.float32)
            _persist[key] = e
        e.used = True
        if torch.cuda.is_current_stream_capturing():
            e.pinned = True                    # a graph replay reads this image without passing through here: always refresh it
        if e.epoch != _epochs.get(e.flat, 0) or e.version != w._version:
            _run('tapconv_pack', 0.0, l.kgan_tapconv_pack_tf32, cs, w.data_ptr(), e.wp.data_ptr(), _stream())
            e.epoch, e.version = _epochs.get(e.flat, 0), w._version
        return e.wp
    key = (w.data_ptr(), w._version, id(desc), numel)
    hit = _packed.get(key)
    if hit is not None and hit[1] is desc:
        return hit[0]
    wp = torch.empty(numel, device=w.device, dtype=torch.float32)
    _run('tapconv_pack', 0.0, l.kgan_tapconv_pack_tf32, cs, w.data_ptr(), wp.data_ptr(), _stream())
    if w.is_leaf:                             # reuse until the next optimizer step
        if len(_packed) > 256:
            _packed.clear()
        _packed[key] = (wp, desc, w)          # keeps `w` alive so data_ptr cannot be recycled under the key
    return wp


def tapconv_fwd(x, w, desc, bias=None, add=None, act=ACT_NONE):
    _chk(x, w, bias, add)
    n = x.shape[0]
    assert x.shape[1] == desc.c_in_total and x.shape[2] * x.shape[3] == desc.p_in, (tuple(x.shape), desc.c_in_total, desc.p_in)
    out = torch.empty((n, desc.c_out_total, desc.t_out, desc.v_out), device=x.device, dtype=torch.float32)
    add_period = 0
    if add is not None and add.shape != out.shape:      # per-joint term broadcast along T: (N, C, 1, V)
        assert add.shape == (n, desc.c_out_total, 1, desc.v_out), (tuple(add.shape), tuple(out.shape))
        add_period = desc.v_out
    l = _lib.lib()
    cs = desc.cstruct(n, act, _precision, add_period)
    if _precision != PREC_FP32:
        wp = _packed_weights(w, desc, cs, l)
        if wp is not None:
            _io(x, w, bias, add, out)
            _run('tapconv_fwd_tf32' if _precision == PREC_TF32 else 'tapconv_fwd_x3', _tap_flops(desc, n), l.kgan_tapconv_fwd_tf32, cs, x.data_ptr(), wp.data_ptr(),
                 desc.pmap_on(x.device).data_ptr(), _ptr(bias), _ptr(add), out.data_ptr(), _stream())
            return out
    _io(x, w, bias, add, out)
    _run('tapconv_fwd', _tap_flops(desc, n), l.kgan_tapconv_fwd, cs, x.data_ptr(), w.data_ptr(), desc.pmap_on(x.device).data_ptr(),
         _ptr(bias), _ptr(add), out.data_ptr(), _stream())
    return out


def gcn_fused_fwd(x, A, w, fused, table=None):
    """tapconv_fwd(adjmix_fwd(x, A), w) - the refolded graph convolution of tgcn.py:61-66 - as ONE kernel (kgan_gcn_fwd_tf32): the adjacency
    product is taken in shared memory by the GEMM's operand builders, the K*C-channel mixed tensor is never written.  `fused`: a
    geometry.GcnFusedGeom; `table` (optional pure-gather PlaneTable): the result is stored through it (see tapconv_fwd_scatter).
    Returns None when not eligible (fp32 mode, shapes outside the staged plan)."""
    if _precision != PREC_TF32 or fused.fwd.mix[0] == 0 or (table is not None and table.scatter_map() is None):
        return None
    _chk(x, A, w)
    desc = fused.fwd
    n = x.shape[0]
    assert x.shape[1] == desc.c_in_total and x.shape[2] * x.shape[3] == desc.p_in and tuple(A.shape) == (desc.ntap, desc.mix[0], desc.mix[1])
    l = _lib.lib()
    cs = desc.cstruct(n, ACT_NONE, _precision, 0, 0 if table is None else table.p_out)
    ok = desc.__dict__.setdefault("_fused_ok", {})
    key = (n, 0 if table is None else table.p_out)
    if key not in ok:
        ok[key] = bool(l.kgan_gcn_fused_ok(cs))
    if not ok[key]:
        return None
    wp = _packed_weights(w, desc, desc.cstruct(n, ACT_NONE, _precision), l)
    if wp is None:
        return None
    shape = (n, desc.c_out_total, desc.t_out, desc.v_out) if table is None else (n, desc.c_out_total, table.t_out, table.v_out)
    out = torch.empty(shape, device=x.device, dtype=torch.float32)
    _io(x, A, w, out)
    _run('gcn_fused_tf32', _tap_flops(desc, n), l.kgan_gcn_fwd_tf32, cs, x.data_ptr(), wp.data_ptr(), A.data_ptr(), 0, 0,
         0 if table is None else table.scatter_on(x.device).data_ptr(), out.data_ptr(), _stream())
    return out


def tapconv_fwd_scatter(x, w, desc, table):
    """plane_spmm(tapconv_fwd(x, w, desc), table) for a pure-gather `table` (geometry.PlaneTable.scatter_map) in ONE launch: the
    convolution's epilogue stores every result at its (one or two) positions of the gathered layout and zero-fills the empty slots
    (kgan_tapconv_fwd_tf32_scatter).  Returns None when not eligible (fp32 mode, no TMA-fed plan, table not a gather)."""
    if _precision != PREC_TF32 or table.scatter_map() is None:
        return None
    _chk(x, w)
    n = x.shape[0]
    assert table.p_in == desc.p_out
    l = _lib.lib()
    cs = desc.cstruct(n, ACT_NONE, _precision, 0, table.p_out)
    ok = desc.__dict__.setdefault("_scatter_ok", {})
    key = (n, table.p_out)
    if key not in ok:
        ok[key] = bool(l.kgan_tapconv_scatter_ok(cs))
    if not ok[key]:
        return None
    wp = _packed_weights(w, desc, desc.cstruct(n, ACT_NONE, _precision), l)      # the packed image does not depend on the output layout
    if wp is None:
        return None
    out = torch.empty((n, desc.c_out_total, table.t_out, table.v_out), device=x.device, dtype=torch.float32)
    _io(x, w, out)
    _run('tapconv_fwd_tf32', _tap_flops(desc, n), l.kgan_tapconv_fwd_tf32_scatter, cs, x.data_ptr(), wp.data_ptr(), table.scatter_on(x.device).data_ptr(),
         0, out.data_ptr(), _stream())
    return out


def tapconv_fwd_noise(x, w, desc, noise, nw, bias=None, add=None, act=ACT_NONE):
    """act(conv_desc(x, w) + bias + add + nw[c] * noise) as one kernel (kgan_tapconv_fwd_tf32_noise) - the eval-mode generator block.
    Returns None when not eligible (fp32 modes, layers the TMA-fed kernel takes): the caller then runs convolution and noise pass separately."""
    if _precision != PREC_TF32:
        return None
    _chk(x, w, noise, nw, bias, add)
    n = x.shape[0]
    l = _lib.lib()
    cs = desc.cstruct(n, act, _precision)
    ok = desc.__dict__.setdefault("_noise_ok", {})
    if n not in ok:
        ok[n] = bool(l.kgan_tapconv_noise_ok(cs))
    if not ok[n]:
        return None
    wp = _packed_weights(w, desc, cs, l)
    if wp is None:
        return None
    out = torch.empty((n, desc.c_out_total, desc.t_out, desc.v_out), device=x.device, dtype=torch.float32)
    assert tuple(noise.shape) == (n, 1, desc.t_out, desc.v_out) and nw.numel() == desc.c_out_total and (add is None or add.shape == out.shape)
    _io(x, w, bias, add, noise, out)
    _run('tapconv_fwd_tf32', _tap_flops(desc, n), l.kgan_tapconv_fwd_tf32_noise, cs, x.data_ptr(), wp.data_ptr(), desc.pmap_on(x.device).data_ptr(),
         _ptr(bias), _ptr(add), noise.data_ptr(), nw.data_ptr(), out.data_ptr(), _stream())
    return out


def tapconv_fwd_res(x, w, desc, x2, w2, desc2, bias=None, bias2=None, act=ACT_NONE):
    """act(conv_desc(x, w) + bias + conv_desc2(x2, w2) + bias2) - a tap convolution with a fused residual 1x1 convolution of a second
    tensor (kgan_tapconv_fwd_tf32_res: one accumulator, the residual never visits HBM).  Returns None when the pair is not eligible
    (fp32 mode, shapes outside the TMA-fed plans): the caller then runs the two convolutions separately."""
    if _precision != PREC_TF32:
        return None
    _chk(x, w, x2, w2, bias, bias2)
    n = x.shape[0]
    l = _lib.lib()
    cs, cs2 = desc.cstruct(n, act, _precision), desc2.cstruct(n, ACT_NONE, _precision)
    ok = desc.__dict__.setdefault("_res_ok", {})
    key = (n, id(desc2))
    if key not in ok:
        ok[key] = bool(l.kgan_tapconv_res_ok(cs, cs2))
    if not ok[key]:
        return None
    wp, wp2 = _packed_weights(w, desc, cs, l), _packed_weights(w2, desc2, cs2, l)
    if wp is None or wp2 is None:
        return None
    out = torch.empty((n, desc.c_out_total, desc.t_out, desc.v_out), device=x.device, dtype=torch.float32)
    _io(x, w, x2, w2, bias, bias2, out)
    flops = _tap_flops(desc, n) + 2.0 * n * desc2.p_out * desc2.co * desc2.ck
    _run('tapconv_fwd_tf32', flops, l.kgan_tapconv_fwd_tf32_res, cs, x.data_ptr(), wp.data_ptr(), cs2, x2.data_ptr(), wp2.data_ptr(), _ptr(bias),
         _ptr(bias2), out.data_ptr(), _stream())
    return out


_LINEAR_T = {}


def _linear_wgrad(x, gout, desc, dw, acc):
    """Weight gradient of a Linear layer (one-position planes), dW[oc, ic] = sum_n gout[n, oc] * x[n, ic], on the tensor-core FORWARD kernel:
    it is the tap convolution of the one-sample tensor x viewed as (1, N channels, C_in positions) with gout as the weight matrix
    (W[oc][n] = gout[n, oc]: strides (1, C_out)), whose output (1, C_out, C_in positions) IS dW; `acc`: dW is also the `add` operand.
    (The weight-gradient kernels contract over positions, of which these layers have one: they ran the FMA kernel at 0.1 TB/s.)
    None: not eligible."""
    from .geometry import TapDesc
    import numpy as np
    n, c_in, c_out = x.shape[0], desc.ck, desc.co
    key = (n, c_in, c_out, desc.c_out_total)
    dt = _LINEAR_T.get(key)
    if dt is None:
        dt = _LINEAR_T[key] = TapDesc(c_in_total=n, p_in=c_in, c_out_total=c_out, p_out=c_in, ntap=1, ck=n, co=c_out, groups=1, g_in=0, g_out=0, g_w=0,
                                      w_oc=1, w_ic=desc.c_out_total, tap_in_ch=[0], tap_w_off=[0], tap_row=[0],
                                      pmap=np.arange(c_in, dtype=np.int32).reshape(1, c_in), t_out=1, v_out=c_in)
    l = _lib.lib()
    cs = dt.cstruct(1, ACT_NONE, _precision)
    wp = _packed_weights(gout, dt, cs, l, cache_image=False)
    if wp is None:
        return None
    _io(x, gout, dw)
    _run('tapconv_wgrad_tf32', 2.0 * n * c_in * c_out, l.kgan_tapconv_fwd_tf32, cs, x.data_ptr(), wp.data_ptr(), dt.pmap_on(x.device).data_ptr(), 0,
         dw.data_ptr() if acc else 0, dw.data_ptr(), _stream())
    return dw


def tapconv_wgrad(x, gout, desc, w_shape, out=None):
    """`out` (optional): a contiguous fp32 tensor of the weight's shape that the result is ADDED to (the flat gradient view of
    the parameter) instead of being returned in a fresh tensor - the kernels accumulate with atomics anyway."""
    _chk(x, gout, out)
    n = x.shape[0]
    acc = 0 if out is None else 1
    dw = torch.empty(w_shape, device=x.device, dtype=torch.float32) if out is None else out
    assert tuple(dw.shape) == tuple(w_shape)
    l = _lib.lib()
    cs = desc.cstruct(n, ACT_NONE, _precision)
    if (_precision == PREC_TF32 and desc.p_in == 1 and desc.p_out == 1 and desc.ntap == 1 and desc.groups == 1 and desc.tap_in_ch[0] == 0
            and desc.ck == desc.c_in_total and desc.co == desc.c_out_total and desc.ck % 4 == 0 and desc.w_oc == desc.ck and desc.w_ic == 1
            and desc.tap_w_off[0] == 0 and n >= 512 and desc.co >= 16):
        r = _linear_wgrad(x, gout, desc, dw, acc)
        if r is not None:
            return r
    _io(x, gout, dw)
    if _precision != PREC_FP32:
        ok = desc.__dict__.setdefault("_tf32_wgrad_ok", {})
        if (n, _precision) not in ok:
            ok[(n, _precision)] = bool(l.kgan_tapconv_wgrad_tf32_ok(cs))
        if ok[(n, _precision)]:
            _run('tapconv_wgrad_tf32' if _precision == PREC_TF32 else 'tapconv_wgrad_x3', _tap_flops(desc, n), l.kgan_tapconv_wgrad_tf32, cs, x.data_ptr(), gout.data_ptr(),
                 desc.pmap_on(x.device).data_ptr(), dw.data_ptr(), dw.numel(), acc, _stream())
            return dw
    _run('tapconv_wgrad', _tap_flops(desc, n), l.kgan_tapconv_wgrad, cs, x.data_ptr(), gout.data_ptr(),
         desc.pmap_on(x.device).data_ptr(), dw.data_ptr(), dw.numel(), acc, _stream())
    return dw


def adjmix_fwd(x, A, sel=None):
    """`sel` (optional selection PlaneTable with inverse_gather()): also returns plane_spmm(x, sel) - written by the same kernel from the
    tile it has staged (kgan_adjmix_fwd_sel) - as a second result; (out, None) when that by-product has no plan for the shape."""
    _chk(x, A)
    n, c, t, v = x.shape
    k, v2, w = A.shape
    assert v2 == v
    out = torch.empty((n, k * c, t, w), device=x.device, dtype=torch.float32)
    _shape_sig(x, A)
    if sel is not None:
        l = _lib.lib()
        ok = sel.__dict__.setdefault("_fwd_sel_ok", {})
        key = (n, c, t, v, w, k, x.data_ptr() & 15)
        if key not in ok:
            ok[key] = sel.inverse_gather() is not None and sel.p_in == t * v and bool(l.kgan_adjmix_fwd_sel_ok(x.data_ptr(), n, c, t, v, w, k))
        if not ok[key]:
            return adjmix_fwd(x, A), None
        xs = torch.empty((n, c, sel.t_out, sel.v_out), device=x.device, dtype=torch.float32)
        _io(x, A, out, xs)
        _run('adjmix', 0.0, l.kgan_adjmix_fwd_sel, x.data_ptr(), A.data_ptr(), sel.on(x.device)[0].data_ptr(), sel.p_out, out.data_ptr(), xs.data_ptr(),
             n, c, t, v, w, k, _rnd(), _stream())
        return out, xs
    _io(x, A, out)
    _run('adjmix', 0.0, _lib.lib().kgan_adjmix_fwd, x.data_ptr(), A.data_ptr(), out.data_ptr(), n, c, t, v, w, k, _rnd(), _stream())
    return out


def adjmix_bwd_x(g, A, add=None, mask_src=None, add_sel=None):
    """`add`, `mask_src` (optional, shaped like the result): gx = (product + add) * leaky_relu'(mask_src) in the same kernel.
    `add_sel` (a selection PlaneTable with inverse_gather()): `add` is COMPACT - shaped like the selection's output - and enters at the
    positions the selection reads (the adjoint of the selection, taken inside the kernel: kgan_adjmix_bwd_x_fused_sel)."""
    _chk(g, A, add, mask_src)
    k, v, w = A.shape
    n, kc, t, w2 = g.shape
    assert w2 == w and kc % k == 0
    c = kc // k
    gx = torch.empty((n, c, t, v), device=g.device, dtype=torch.float32)
    assert mask_src is None or mask_src.shape == gx.shape
    _shape_sig(g, A)
    _io(g, A, add, mask_src, gx)
    if add_sel is not None:
        assert add is not None and add_sel.p_in == t * v and tuple(add.shape) == (n, c, add_sel.t_out, add_sel.v_out), (tuple(add.shape), t, v)
        _run('adjmix', 0.0, _lib.lib().kgan_adjmix_bwd_x_fused_sel, g.data_ptr(), A.data_ptr(), add.data_ptr(), add_sel.inverse_on(g.device).data_ptr(),
             add_sel.p_out, _ptr(mask_src), gx.data_ptr(), n, c, t, v, w, k, _rnd(), _stream())
        return gx
    assert add is None or add.shape == gx.shape
    if add is None and mask_src is None:
        _run('adjmix', 0.0, _lib.lib().kgan_adjmix_bwd_x, g.data_ptr(), A.data_ptr(), gx.data_ptr(), n, c, t, v, w, k, _rnd(), _stream())
    else:
        _run('adjmix', 0.0, _lib.lib().kgan_adjmix_bwd_x_fused, g.data_ptr(), A.data_ptr(), _ptr(add), _ptr(mask_src), gx.data_ptr(), n, c, t, v, w,
             k, _rnd(), _stream())
    return gx


def adjmix_bwd_a(x, g, k, mask=None):
    """gA (k, v, w); with `mask` (k, v, w) only the entries where mask != 0 are computed, the others are 0."""
    _chk(x, g, mask)
    n, c, t, v = x.shape
    w = g.shape[3]
    assert g.shape[0] == n and g.shape[1] == k * c and g.shape[2] == t
    assert mask is None or tuple(mask.shape) == (k, v, w)
    gA = torch.empty((k, v, w), device=x.device, dtype=torch.float32)
    _shape_sig(x, g)
    _io(x, g, gA)
    _run('adjmix_bwd_a', 0.0, _lib.lib().kgan_adjmix_bwd_a_masked, x.data_ptr(), g.data_ptr(), _ptr(mask), gA.data_ptr(), n, c, t, v, w, k, _stream())
    return gA


def epilogue_fwd(a, b=None, bias=None, nw=None, noise=None, act=ACT_NONE):
    _chk(a, b, bias, nw, noise)
    n, c, t, v = a.shape
    out = torch.empty_like(a)
    _io(a, b, noise, out)
    _run('pointwise', 0.0, _lib.lib().kgan_epilogue_fwd, a.data_ptr(), _ptr(b), _ptr(bias), _ptr(nw), _ptr(noise), out.data_ptr(), n, c, t * v, act,
                                            _rnd(), _stream())
    return out


def act_bwd(gout, out, act):
    _chk(gout, out)
    gz = torch.empty_like(out)
    _shape_sig(out)
    _io(gout, out, gz)
    _run('pointwise', 0.0, _lib.lib().kgan_act_bwd, gout.data_ptr(), out.data_ptr(), gz.data_ptr(), out.numel(), act, _rnd(), _stream())
    return gz


def chan_reduce(g, mul=None):
    _chk(g, mul)
    n, c, t, v = g.shape
    out = torch.empty((c,), device=g.device, dtype=torch.float32)
    _shape_sig(g)
    _io(g, mul)
    _run('reduce', 0.0, _lib.lib().kgan_chan_reduce, g.data_ptr(), _ptr(mul), out.data_ptr(), n, c, t * v, _stream())
    return out


def plane_spmm(x, table):
    _chk(x)
    n, c, t, v = x.shape
    assert t * v == table.p_in, (tuple(x.shape), table.p_in)
    idx, wgt = table.on(x.device)
    out = torch.empty((n, c, table.t_out, table.v_out), device=x.device, dtype=torch.float32)
    _shape_sig(x, out)
    _io(x, out)
    _run('plane_spmm', 0.0, _lib.lib().kgan_plane_spmm, x.data_ptr(), idx.data_ptr(), wgt.data_ptr(), out.data_ptr(), n * c, table.p_in, table.p_out,
                                          table.J, _rnd(), _stream())
    return out


def plane_sum_t(x):
    """(N, C, T, V) -> (N, C, 1, V): sum over frames."""
    _chk(x)
    n, c, t, v = x.shape
    out = torch.empty((n, c, 1, v), device=x.device, dtype=torch.float32)
    _shape_sig(x, out)
    _io(x, out)
    _run('plane_spmm', 0.0, _lib.lib().kgan_plane_sum_t, x.data_ptr(), out.data_ptr(), n * c, t, v, _rnd(), _stream())
    return out


def label_concat(e, x):
    _chk(e, x)
    n, c, t, v = x.shape
    ncls = e.shape[1]
    out = torch.empty((n, ncls + c, t, v), device=x.device, dtype=torch.float32)
    _run('label', 0.0, _lib.lib().kgan_label_concat, e.data_ptr(), x.data_ptr(), out.data_ptr(), n, ncls, c, t * v, _rnd(), _stream())
    return out


def label_split(g, ncls, need_e=True, need_x=True):
    _chk(g)
    n, ct, t, v = g.shape
    c = ct - ncls
    ge = torch.empty((n, ncls), device=g.device, dtype=torch.float32) if need_e else None
    gx = torch.empty((n, c, t, v), device=g.device, dtype=torch.float32) if need_x else None
    _run('label', 0.0, _lib.lib().kgan_label_split, g.data_ptr(), _ptr(ge), _ptr(gx), n, ncls, c, t * v, _rnd(), _stream())
    return ge, gx


def bn_stats(x, running_mean=None, running_var=None, eps=1e-5, momentum=0.1):
    _chk(x, running_mean, running_var)
    n, c, t, v = x.shape
    mean = torch.empty((c,), device=x.device, dtype=torch.float32)
    rstd = torch.empty((c,), device=x.device, dtype=torch.float32)
    ws = torch.empty((int(_lib.lib().kgan_bn_workspace(n, c)),), device=x.device, dtype=torch.float32)
    _io(x)
    _run('batchnorm', 0.0, _lib.lib().kgan_bn_stats, x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _ptr(running_mean), _ptr(running_var), n, c, t * v,
                                        eps, momentum, ws.data_ptr(), _stream())
    return mean, rstd


def bn_apply(x, mean, rstd, gamma, beta):
    _chk(x, mean, rstd, gamma, beta)
    n, c, t, v = x.shape
    y = torch.empty_like(x)
    _io(x, y)
    _run('batchnorm', 0.0, _lib.lib().kgan_bn_apply, x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), n, c,
                                        t * v, _rnd(), _stream())
    return y


def bn_epilogue_fwd(x, mean, rstd, gamma, beta, b=None, nw=None, noise=None, act=ACT_NONE):
    """act(bn_apply(x, ...) + b + nw[c] * noise) in one pass (kgan_bn_epilogue_fwd)."""
    _chk(x, mean, rstd, gamma, beta, b, nw, noise)
    n, c, t, v = x.shape
    out = torch.empty_like(x)
    _io(x, b, noise, out)
    _run('batchnorm', 0.0, _lib.lib().kgan_bn_epilogue_fwd, x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(b),
         _ptr(nw), _ptr(noise), out.data_ptr(), n, c, t * v, act, _rnd(), _stream())
    return out


def bn_bwd(gy, x, mean, rstd, gamma):
    _chk(gy, x, mean, rstd, gamma)
    n, c, t, v = x.shape
    gx = torch.empty_like(x)
    gg = torch.empty((c,), device=x.device, dtype=torch.float32)
    gb = torch.empty((c,), device=x.device, dtype=torch.float32)
    ws = torch.empty((int(_lib.lib().kgan_bn_workspace(n, c)),), device=x.device, dtype=torch.float32)
    _io(gy, x, gy, x, gx)          # two passes over (gy, x): sums, then the elementwise pass
    _run('batchnorm', 0.0, _lib.lib().kgan_bn_bwd, gy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), gx.data_ptr(),
                                      gg.data_ptr(), gb.data_ptr(), n, c, t * v, _rnd(), ws.data_ptr(), _stream())
    return gx, gg, gb


def adam_step(p, g, m, v, lr, b1, b2, eps, step, grad_scale=1.0):
    _chk(p, g, m, v)
    _io(p, g, m, v, p, m, v)
    _run('adam', 0.0, _lib.lib().kgan_adam_step, p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, b1, b2, eps, step,
                                         grad_scale, _stream())


def interpolate(alpha, x, y):
    _chk(alpha, x, y)
    n = x.shape[0]
    out = torch.empty_like(x)
    _run('pointwise', 0.0, _lib.lib().kgan_interpolate, alpha.data_ptr(), x.data_ptr(), y.data_ptr(), out.data_ptr(), n, x.numel() // n, _rnd(), _stream())
    return out


def round_tf32(x):
    """x rounded to tf32 (no-op copy-free pass-through in fp32 mode): for tensors entering the tf32 path from outside the library."""
    if _precision != PREC_TF32:
        return x
    _chk(x)
    out = torch.empty_like(x)
    _io(x, out)
    _run('pointwise', 0.0, _lib.lib().kgan_round_tf32, x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    return out
