"""Host-side geometry tables for the tap-convolution and plane-gather kernels.

Everything the reference does with shapes and indices around its convolutions - temporal zero
padding (generator.py:129, discriminator.py:94), `downsample_s` joint selection
(discriminator.py:139-142), nearest-neighbour `F.interpolate` along T (generator.py:172,
discriminator.py:134), `upsample_s` joint insertion (generator.py:185-200), global average pooling
(discriminator.py:68) - becomes a small int32/float32 table that the kernels consume, so that no
resampled tensor is ever materialised on the discriminator side and the generator side needs one
gather.  Tables are built once per module (ctor time) with numpy and cached per device.
"""
import numpy as np
import torch

from ._lib import MAX_TAPS

# Which tap convolutions take the operand-building tensor-core kernel (csrc/tapconv_build.cu) in tf32 mode:
#   "fallback": only those the TMA-fed kernel cannot serve (unaligned shifts, stride / selection maps);
#   "tcn":      also every multi-tap convolution whose taps read the same channels (one staged tile serves all taps);
#   "all":      every eligible one.
STAGED_POLICY = "fallback"


def nearest_src(t_in, t_out):
    """Index rule of F.interpolate(mode='nearest'): src = floor(dst * in / out)."""
    return [min(int(np.floor(d * (t_in / t_out))), t_in - 1) for d in range(t_out)]


class _DeviceCache:
    def __init__(self):
        self._cache = {}

    def get(self, key, device, make):
        k = (key, str(device))
        if k not in self._cache:
            self._cache[k] = make().to(device)
        return self._cache[k]


class TapDesc:
    """Static part of a `kgan_tapconv_desc` plus its position map (see include/kgan.h)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)
        assert self.ntap <= MAX_TAPS
        assert self.pmap.dtype == np.int32 and self.pmap.shape[1] == self.p_out
        self._dev = _DeviceCache()
        self._structs = {}
        self.pmap_vec_mask = self._vec_mask()
        self.tma_mode, self.tap_shift = self._shift_form()
        self.stage_span = self._stage_span()
        self.prefer_staged = int(self.stage_span > 0 and STAGED_POLICY != "fallback" and
                                 (STAGED_POLICY == "all" or (self.ntap > 1 and len(set(self.tap_in_ch)) == 1)))

    def _vec_mask(self):
        """Rows of pmap whose aligned groups of 4 output positions map to 4 consecutive, aligned inputs (or all to -1)."""
        if self.p_in % 4 or self.p_out % 4:
            return 0
        mask = 0
        for r in range(min(self.pmap.shape[0], 31)):
            q = self.pmap[r].reshape(-1, 4).astype(np.int64)
            hole = (q < 0).all(1)
            run = (q[:, 0] >= 0) & (q[:, 0] % 4 == 0) & (q == q[:, :1] + np.arange(4)).all(1)
            if (hole | run).all():
                mask |= 1 << r
        return mask

    def _shift_form(self):
        """tma_mode 1 (include/kgan.h): every tap reads input position p + shift (zero outside the input plane)."""
        none = (0, [0] * self.ntap)
        p = np.arange(self.p_out, dtype=np.int64)
        shifts = []
        for t in range(self.ntap):
            row = self.pmap[self.tap_row[t]].astype(np.int64)
            hit = np.nonzero(row >= 0)[0]
            if len(hit) == 0:
                return none
            sh = int(row[hit[0]] - hit[0])
            want = np.where((p + sh >= 0) & (p + sh < self.p_in), p + sh, -1)
            if not (row == want).all():
                return none
            shifts.append(sh)
        return 1, shifts

    def _stage_span(self):
        """include/kgan.h `stage_span`: input positions a 128-row output tile can touch through the taps of one channel block, from a
        4-aligned first position (planes of at most 128 output positions: the whole input plane)."""
        if self.p_in % 4:
            return 0
        if self.p_out <= 128:
            return self.p_in if self.p_in <= 512 else 0
        blocks = {}
        for t in range(self.ntap):
            blocks.setdefault(self.tap_in_ch[t], []).append(self.tap_row[t])
        span = 4
        for rows in blocks.values():
            m = self.pmap[sorted(set(rows))].astype(np.int64)
            for r0 in range(0, self.p_out, 128):
                v = m[:, r0:r0 + 128]
                v = v[v >= 0]
                if v.size:
                    span = max(span, int(v.max()) + 1 - (int(v.min()) & ~3))
        span = (span + 3) // 4 * 4
        return span if span <= 512 else 0

    def pmap_on(self, device):
        return self._dev.get("pmap", device, lambda: torch.from_numpy(self.pmap))

    def cstruct(self, n, act, precision, add_period=0, out_plane=0):
        """`out_plane` > 0: the result is stored through a scatter table into planes of that many positions (ops.tapconv_fwd_scatter)."""
        from ._lib import TapConvDesc

        key = (n, act, precision, add_period, out_plane)
        s = self._structs.get(key)
        if s is None:
            s = TapConvDesc()
            s.n, s.c_in_total, s.p_in, s.c_out_total, s.p_out = n, self.c_in_total, self.p_in, self.c_out_total, self.p_out
            s.ntap, s.ck, s.co, s.groups = self.ntap, self.ck, self.co, self.groups
            s.g_in, s.g_out, s.g_w = self.g_in, self.g_out, self.g_w
            s.w_oc, s.w_ic, s.w_oc_blk, s.w_ocblk = self.w_oc, self.w_ic, 0, 0
            for i in range(self.ntap):
                s.tap_in_ch[i], s.tap_w_off[i], s.tap_row[i] = self.tap_in_ch[i], self.tap_w_off[i], self.tap_row[i]
            s.pmap_vec_mask, s.add_period, s.act, s.precision = self.pmap_vec_mask, add_period, act, precision
            s.tma_mode = self.tma_mode
            for i in range(self.ntap):
                s.tap_shift[i] = self.tap_shift[i]
            s.p_out_plane, s.g_pout = (out_plane, 0) if out_plane else (getattr(self, "p_out_plane", 0), getattr(self, "g_pout", 0))
            s.mix_v, s.mix_w, s.mix_l = getattr(self, "mix", (0, 0, 0))
            s.stage_span, s.prefer_staged = (0, 0) if (s.p_out_plane and not s.mix_v) else (self.stage_span, self.prefer_staged)
            self._structs[key] = s
        return s


class TapConvGeom:
    """One convolution site: K channel blocks (graph partitions; 1 for ordinary convs) x kt temporal taps.

    natural weight shape (K*c_out, c_in, kt, 1) - exactly nn.Conv2d's (tgcn.py:48, generator.py:134,155,
    discriminator.py:99,115) - or (c_out, c_in) for nn.Linear (kt = 1).
    input  (N, K*c_in, t_in, v_in)   [for K > 1 the input is the adjacency-mixed tensor, kgan_adjmix_fwd]
    output (N, c_out, len(t_sel), len(v_keep)): only the frames / joints that survive the block's
    down-sampling are computed (SURVEY.md §7 I2 - selection commutes with the linear map).
    """

    def __init__(self, c_in, c_out, t_in, v_in, K=1, kt=1, pad=0, stride=1, dil=1, t_sel=None, v_keep=None, w_cin=None,
                 w_ic0=0):
        """w_cin / w_ic0: the weight tensor has w_cin >= c_in input channels and this site contracts over the slice
        [w_ic0, w_ic0 + c_in) only (used to split the critic's first layer into label and data channels)."""
        self.c_in, self.c_out, self.t_in, self.v_in, self.K, self.kt = c_in, c_out, t_in, v_in, K, kt
        w_cin = c_in if w_cin is None else w_cin
        assert w_ic0 + c_in <= w_cin
        t_conv = (t_in + 2 * pad - dil * (kt - 1) - 1) // stride + 1
        assert t_conv >= 1, "temporal kernel larger than the padded input"
        self.t_sel = list(range(t_conv)) if t_sel is None else [int(t) for t in t_sel]
        self.v_keep = list(range(v_in)) if v_keep is None else [int(v) for v in v_keep]
        assert all(0 <= t < t_conv for t in self.t_sel) and all(-1 <= v < v_in for v in self.v_keep)   # -1: padded (dummy) joint, reads zero
        self.t_out, self.v_out = len(self.t_sel), len(self.v_keep)
        self.p_in, self.p_out = t_in * v_in, self.t_out * self.v_out
        self.w_numel = K * c_out * w_cin * kt

        pmap = np.full((kt, self.p_out), -1, np.int32)
        inv = np.full((kt, self.p_in), -1, np.int32)
        for dt in range(kt):
            for a, tau in enumerate(self.t_sel):
                t = tau * stride + dt * dil - pad
                if 0 <= t < t_in:
                    for b, v in enumerate(self.v_keep):
                        if v < 0:
                            continue
                        q, p = a * self.v_out + b, t * v_in + v
                        pmap[dt, q] = p
                        assert inv[dt, p] == -1, "position map must be injective per tap"
                        inv[dt, p] = q
        # temporal taps that only ever read zero padding (a 3-tap kernel on a 1-frame plane: the generator's first block) are
        # dropped from the contraction; their weight gradients are exactly zero (the weight-gradient kernels clear dw first)
        live = [dt for dt in range(kt) if (pmap[dt] >= 0).any()]
        taps = [(k, dt) for k in range(K) for dt in live]
        self.fwd = TapDesc(
            c_in_total=K * c_in, p_in=self.p_in, c_out_total=c_out, p_out=self.p_out, ntap=len(taps), ck=c_in, co=c_out,
            groups=1, g_in=0, g_out=0, g_w=0, w_oc=w_cin * kt, w_ic=kt,
            tap_in_ch=[k * c_in for k, dt in taps], tap_w_off=[(k * c_out * w_cin + w_ic0) * kt + dt for k, dt in taps],
            tap_row=[dt for k, dt in taps], pmap=pmap, t_out=self.t_out, v_out=self.v_out)
        # data gradient: same kernel, roles of (oc, ic) swapped, inverse map, one group per channel block
        self.dgrad = TapDesc(
            c_in_total=c_out, p_in=self.p_out, c_out_total=K * c_in, p_out=self.p_in, ntap=len(live), ck=c_out, co=c_in,
            groups=K, g_in=0, g_out=c_in, g_w=c_out * w_cin * kt, w_oc=kt, w_ic=w_cin * kt,
            tap_in_ch=[0] * len(live), tap_w_off=[w_ic0 * kt + dt for dt in live], tap_row=list(live), pmap=inv, t_out=t_in,
            v_out=v_in)


class GcnFusedGeom:
    """The graph convolution of tgcn.py:61-66 as ONE kernel (include/kgan.h kgan_gcn_fwd_tf32): x (N, C_in, T, V) and the effective adjacency
    A (K, V, W) in, (N, C_out, T, W) out; the adjacency product is taken inside the GEMM's operand builder.  `nnz_max`: the largest number
    of non-zeros in a column of any partition of the (constant) base adjacency."""

    def __init__(self, c_in, c_out, t, v, w, K, nnz_max):
        self.c_in, self.c_out, self.t, self.v, self.w, self.K = c_in, c_out, t, v, w, K
        p_in, p_out = t * v, t * w
        if p_out <= 128:
            span = p_in
        else:
            span = 4
            for r0 in range(0, p_out, 128):
                t0, t1 = r0 // w, (min(r0 + 128, p_out) - 1) // w
                span = max(span, (t1 + 1) * v - ((t0 * v) & ~3))
        span = (span + 3) // 4 * 4
        ok = p_in % 4 == 0 and span <= 512 and 1 <= nnz_max <= 8
        self.fwd = TapDesc(
            c_in_total=c_in, p_in=p_in, c_out_total=c_out, p_out=p_out, ntap=K, ck=c_in, co=c_out, groups=1, g_in=0, g_out=0, g_w=0,
            w_oc=c_in, w_ic=1, tap_in_ch=[0] * K, tap_w_off=[k * c_out * c_in for k in range(K)], tap_row=[0] * K,
            pmap=np.zeros((1, p_out), np.int32), t_out=t, v_out=w)
        self.fwd.tma_mode, self.fwd.stage_span, self.fwd.prefer_staged = 0, (span if ok else 0), 0
        self.fwd.mix = (v, w, int(nnz_max)) if ok else (0, 0, 0)
        self.fwd._structs.clear()


class UnfoldedTcnGeom:
    """Temporal convolution with frame selection (stride / nearest-T resampling) over a TIME-UNFOLDED copy of its input.

    The temporal conv of a down-sampling critic block (discriminator.py:99-105 followed by F.interpolate at :134) reads,
    for output frame a and tap d, input frame t_sel[a]*stride + d*dil - pad.  `unfold` (a PlaneTable, applied with
    kgan_plane_spmm) gathers exactly those frames into kt consecutive blocks of H = len(t_sel)*V positions - block d holds
    the operand of tap d (zero rows where the tap falls into the temporal padding) - so that the convolution itself becomes
    out[q] = sum_d W_d . u[d*H + q]: every tap is a pure position shift by a multiple of H.  With H a multiple of 4 this is
    the form the TMA-fed tensor-core kernel accepts (tapconv_tma.cu); the unaligned shifts by V = 5 or 1 positions of the
    direct formulation are not.  For stride 2 the unfolded copy is 1.5x the input.
    Exposes the same attributes as TapConvGeom (.fwd / .dgrad descriptors) so the TapConv* Functions apply unchanged."""

    def __init__(self, c_in, c_out, t_in, v_in, kt, pad, stride, dil, t_sel):
        self.c_in, self.c_out, self.t_in, self.v_in, self.K, self.kt = c_in, c_out, t_in, v_in, 1, kt
        self.t_sel = [int(t) for t in t_sel]
        self.t_out, self.v_out = len(self.t_sel), v_in
        H = self.t_out * v_in
        self.p_in, self.p_out = kt * H, H
        self.w_numel = c_out * c_in * kt
        dense = np.zeros((kt * H, t_in * v_in))
        for d in range(kt):
            for a, tau in enumerate(self.t_sel):
                t = tau * stride + d * dil - pad
                if 0 <= t < t_in:
                    dense[d * H + a * v_in + np.arange(v_in), t * v_in + np.arange(v_in)] = 1.0
        self.unfold = PlaneTable(dense, kt * self.t_out, v_in, t_in, v_in)
        q = np.arange(H, dtype=np.int32)
        pmap = np.stack([q + d * H for d in range(kt)]).astype(np.int32)
        self.fwd = TapDesc(
            c_in_total=c_in, p_in=kt * H, c_out_total=c_out, p_out=H, ntap=kt, ck=c_in, co=c_out, groups=1, g_in=0, g_out=0, g_w=0,
            w_oc=c_in * kt, w_ic=kt, tap_in_ch=[0] * kt, tap_w_off=list(range(kt)), tap_row=list(range(kt)), pmap=pmap,
            t_out=self.t_out, v_out=v_in)
        # data gradient: block d of the unfolded gradient is W_d^T g - kt position-block groups with ONE tap each (include/kgan.h
        # p_out_plane / g_pout) instead of kt taps over the whole 3H plane of which kt - 1 would read nothing
        self.dgrad = TapDesc(
            c_in_total=c_out, p_in=H, c_out_total=c_in, p_out=H, ntap=1, ck=c_out, co=c_in, groups=kt, g_in=0, g_out=0, g_w=1,
            w_oc=kt, w_ic=c_in * kt, tap_in_ch=[0], tap_w_off=[0], tap_row=[0], pmap=q.reshape(1, H).astype(np.int32),
            t_out=kt * self.t_out, v_out=v_in, p_out_plane=kt * H, g_pout=H)


class PlaneTable:
    """Sparse (p_out x p_in) matrix applied to every (n, c) plane: out[q] = sum_j wgt[q,j] * x[idx[q,j]]."""

    def __init__(self, dense, t_out, v_out, t_in, v_in):
        dense = np.asarray(dense, np.float64)                # (p_out, p_in)
        self.p_out, self.p_in = dense.shape
        self.t_out, self.v_out, self.t_in, self.v_in = t_out, v_out, t_in, v_in
        assert self.p_out == t_out * v_out and self.p_in == t_in * v_in
        self.dense = dense
        self.J = max(1, int((dense != 0).sum(1).max()))
        self.idx = np.full((self.p_out, self.J), -1, np.int32)
        self.wgt = np.zeros((self.p_out, self.J), np.float32)
        for q in range(self.p_out):
            nz = np.nonzero(dense[q])[0]
            self.idx[q, :len(nz)] = nz
            self.wgt[q, :len(nz)] = dense[q, nz]
        self._dev = _DeviceCache()
        self._T = None

    @property
    def T(self):
        if self._T is None:
            self._T = PlaneTable(self.dense.T, self.t_in, self.v_in, self.t_out, self.v_out)
            self._T._T = self
        return self._T

    def on(self, device):
        return (self._dev.get("idx", device, lambda: torch.from_numpy(self.idx)),
                self._dev.get("wgt", device, lambda: torch.from_numpy(self.wgt)))

    def scatter_map(self):
        """The table as a STORE pattern (include/kgan.h kgan_tapconv_fwd_tf32_scatter): for a pure gather table (every output
        position copies one input position with weight 1, or is zero) in which an input position is copied at most twice, an int32
        (p_in, 3) array - the two destinations of every source position (-1: none) and one zero slot it is responsible for.
        None if the table is not of that form."""
        if hasattr(self, "_scatter"):
            return self._scatter
        self._scatter = None
        if self.J != 1 or not np.all((self.wgt[:, 0] == 1.0) | (self.idx[:, 0] < 0)):
            return None
        m = np.full((self.p_in, 3), -1, np.int32)
        zeros = []
        for q in range(self.p_out):
            p = int(self.idx[q, 0])
            if p < 0:
                zeros.append(q)
            elif m[p, 0] < 0:
                m[p, 0] = q
            elif m[p, 1] < 0:
                m[p, 1] = q
            else:
                return None
        if (m[:, 0] < 0).any() or len(zeros) > self.p_in:        # every source must have a destination; one zero slot per source at most
            return None
        for i, q in enumerate(zeros):
            m[i, 2] = q
        self._scatter = m
        return m

    def inverse_gather(self):
        """For a selection table (every output position copies ONE input position with weight 1, no input position copied twice): the
        int32 array inv[p_in] = output position that copies p, -1 if none - the form kgan_adjmix_bwd_x_fused_sel takes the adjoint of
        the selection in.  None if the table is not of that form."""
        if not hasattr(self, "_inv"):
            self._inv = None
            if self.J == 1 and np.all(self.idx[:, 0] >= 0) and np.all(self.wgt[:, 0] == 1.0) and len(np.unique(self.idx[:, 0])) == self.p_out:
                inv = np.full(self.p_in, -1, np.int32)
                inv[self.idx[:, 0]] = np.arange(self.p_out, dtype=np.int32)
                self._inv = inv
        return self._inv

    def inverse_on(self, device):
        return self._dev.get("inv", device, lambda: torch.from_numpy(self.inverse_gather()))

    def scatter_on(self, device):
        return self._dev.get("scatter", device, lambda: torch.from_numpy(self.scatter_map()))


def upsample_matrix(hoods, v_coarse, halve):
    """U (v_coarse x v_fine) with  upsample_s(x) == x @ U  (generator.py:185-200): every hood
    [fine_idx, coarse...] inserts, at fine_idx, the mean of the listed coarse joints (/2 iff lvl == 2)."""
    cols = [np.eye(v_coarse)[:, j] for j in range(v_coarse)]
    new = []
    for hood in hoods:
        col = np.zeros(v_coarse)
        for j in hood[1:]:
            col[int(j)] += 1.0 / (len(hood) - 1)
        new.append(col / 2 if halve else col)
    for hood, col in zip(hoods, new):
        cols.insert(int(hood[0]), col)
    return np.stack(cols, 1)


def resample_table(t_in, v_in, t_out, U=None):
    """Plane table of  F.interpolate(upsample_s(x), size=(t_out, V))  (generator.py:170-172)."""
    U = np.eye(v_in) if U is None else np.asarray(U)
    v_out = U.shape[1]
    src = nearest_src(t_in, t_out)
    dense = np.zeros((t_out * v_out, t_in * v_in))
    for t in range(t_out):
        dense[t * v_out:(t + 1) * v_out, src[t] * v_in:(src[t] + 1) * v_in] = U.T
    return PlaneTable(dense, t_out, v_out, t_in, v_in)


def select_table(t_in, v_in, t_sel, v_keep):
    """Plane table of  F.interpolate(x[..., keep], size=(t_out, V'))  on the identity residual (discriminator.py:128-134)."""
    dense = np.zeros((len(t_sel) * len(v_keep), t_in * v_in))
    for a, t in enumerate(t_sel):
        for b, v in enumerate(v_keep):
            if v >= 0:                                       # -1: padded (dummy) joint, stays zero
                dense[a * len(v_keep) + b, t * v_in + v] = 1.0
    return PlaneTable(dense, len(t_sel), len(v_keep), t_in, v_in)


def sum_t_table(t_in, v_in):
    """Plane table summing over frames: (T, V) -> (1, V); the adjoint of broadcasting a per-joint term along T."""
    dense = np.zeros((v_in, t_in * v_in))
    for t in range(t_in):
        dense[np.arange(v_in), t * v_in + np.arange(v_in)] = 1.0
    return PlaneTable(dense, 1, v_in, t_in, v_in)


def mean_table(t_in, v_in):
    """Plane table of F.avg_pool2d(x, (T, V)) (discriminator.py:68)."""
    return PlaneTable(np.full((1, t_in * v_in), 1.0 / (t_in * v_in)), 1, 1, t_in, v_in)
