/* kgan.h - C ABI of libkgan.so: the B200 (sm_100a) kernels behind the Kinetic-GAN ST-GCN hot path.
 *
 * The reference (DegardinBruno/Kinetic-GAN) has no FFI of its own: every arithmetic call site is a
 * PyTorch operator.  Each entry point below names the reference call sites it replaces (paths are
 * relative to the reference root).  Conventions for ALL entry points:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer owned by the caller
 *     (torch's allocator); the library borrows it for the duration of the call and keeps nothing;
 *   - tensors are float32, dense, NCHW-contiguous (N, C, T, V); a "plane" is the T*V block of one
 *     (n, c) pair, P = T*V;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and re-entrant per
 *     stream.  The library reads no environment variables and keeps no mutable state besides the
 *     per-thread last-error string and a per-(kernel, device) "shared-memory attribute set" bit;
 *   - `out_tf32` (and kgan_tapconv_desc.precision == KGAN_PREC_TF32 for the tap convolution): the
 *     activations a kernel PRODUCES are stored rounded to tf32 (round to nearest, ties away; fp32
 *     container).  The tensor-core kernels feed raw fp32 words to tcgen05.mma kind::tf32, which
 *     reads the upper 19 bits: on tf32-rounded tensors that read is exact, so the tf32 mode
 *     computes RN(x) * RN(w) with fp32 accumulation - unbiased, and exact on tf32-representable
 *     data.  A tensor that did not come from a libkgan kernel is truncated instead (one-sided,
 *     < 2^-10 relative); pass it through kgan_round_tf32 first if that matters;
 *   - return value: 0 = ok, non-zero = error (message via kgan_last_error()); nothing throws
 *     across the boundary; there is NO CPU fallback.
 */
#ifndef KGAN_H_
#define KGAN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KGAN_MAX_TAPS 16

enum { KGAN_ACT_NONE = 0, KGAN_ACT_LRELU = 1 /* slope 0.2 */, KGAN_ACT_TANH = 2 };
enum {
    KGAN_PREC_FP32 = 0,   /* SIMT fp32 FMA */
    KGAN_PREC_TF32 = 1,   /* tcgen05 kind::tf32, fp32 accumulate, operands rounded to tf32 */
    KGAN_PREC_TF32X3 = 2  /* fp32-accurate tensor-core mode: every operand is split x = hi + lo inside the kernel (hi = the 19 bits the
                             tensor core reads, lo = x - hi, exact) and the product is taken as hi*hi + lo*hi + hi*lo on
                             tcgen05 kind::tf32 with fp32 accumulation (the dropped lo*lo term is ~2^-22 relative).  Activations stay
                             full fp32 in HBM.  Entry points: the *_tf32 ones; eligible shapes: the TMA-fed plans (forward / data
                             gradient / weight gradient), everything else reports "not eligible" and runs on the fp32 FMA kernels. */
};

/* Geometry of one "tap convolution": the single GEMM-shaped primitive that the graph conv, the
 * temporal conv, the residual 1x1 conv, the mapping MLP and the critic head all reduce to, in
 * forward, data-gradient and weight-gradient form.
 *
 *   out[n, out_ch0 + oc, p] = act( bias[oc] + add[n, oc, p (or p % add_period)]
 *        + sum_{tap < ntap} sum_{ic < ck}  W[w_base + tap_w_off[tap] + woff(oc) + ic*w_ic]
 *                                        * in[n, in_ch0 + tap_in_ch[tap] + ic, pmap[tap_row[tap]*p_out + p]] )
 *   woff(oc) = w_oc_blk ? (oc / w_oc_blk) * w_ocblk + (oc % w_oc_blk) * w_oc : oc * w_oc
 *   pmap entries < 0 contribute zero (temporal zero padding, dropped joints / frames).
 * `groups` > 1 repeats the whole operation with in_ch0 += g_in, out_ch0 += g_out, w_base += g_w.
 */
typedef struct kgan_tapconv_desc {
    int32_t n;                 /* samples */
    int32_t c_in_total;        /* channels of the `in` tensor  */
    int32_t p_in;              /* plane size of `in`  (T_in * V_in)  */
    int32_t c_out_total;       /* channels of the `out` tensor */
    int32_t p_out;             /* plane size of `out` (T_out * V_out) */
    int32_t ntap;              /* number of taps (<= KGAN_MAX_TAPS) */
    int32_t ck;                /* contraction channels per tap */
    int32_t co;                /* output channels computed per group */
    int32_t groups;            /* >= 1 */
    int32_t g_in, g_out;       /* per-group channel steps */
    int64_t g_w;               /* per-group weight step (elements) */
    int64_t w_oc, w_ic;        /* weight strides (elements) */
    int32_t w_oc_blk;          /* 0 = single-level oc addressing */
    int64_t w_ocblk;
    int32_t tap_in_ch[KGAN_MAX_TAPS];
    int64_t tap_w_off[KGAN_MAX_TAPS];
    int32_t tap_row[KGAN_MAX_TAPS];  /* row of pmap used by the tap */
    int32_t pmap_vec_mask;     /* bit r set: pmap row r sends every aligned group of 4 output positions either to 4 consecutive,
                                  16-byte aligned input positions or entirely to -1, and p_in, p_out are multiples of 4 (the
                                  tensor-core weight-gradient kernel then stages 16 bytes per cp.async); 0 is always valid */
    int32_t add_period;        /* 0: `add` has the shape of out; else add is (N, c_out_total, add_period) and is read at
                                  p % add_period (a term that is constant over frames, broadcast along T) */
    int32_t act;               /* KGAN_ACT_* applied by kgan_tapconv_fwd only */
    int32_t precision;         /* KGAN_PREC_TF32: the output is stored tf32-rounded (see `out_tf32` above), by every tapconv entry point */
    /* Shift form of the position maps, used by the TMA-fed tensor-core kernel (kgan_tapconv_fwd_tf32); the gather form
     * above stays authoritative for every other entry point.
     *   tma_mode 0: no shift form.
     *   tma_mode 1: pmap[tap_row[t]][p] == p + tap_shift[t] when that lies in [0, p_in), else -1 (p_in may differ from p_out).
     * The TMA kernel additionally needs p_in, p_out and every shift to be multiples of 4 (16-byte box origins), or one-position
     * planes with a channel count that is a multiple of 4 (see kgan_tapconv_tma_ok). */
    int32_t tma_mode;
    int32_t tap_shift[KGAN_MAX_TAPS];
    /* Position-block groups (kgan_tapconv_fwd / _fwd_tf32 only; 0, 0 = off): the output planes hold p_out_plane >= p_out positions
     * and group g writes its p_out results at plane offset g * g_pout.  The data gradient of a time-unfolded temporal conv
     * (geometry.UnfoldedTcnGeom) is kt such groups with ONE tap each instead of kt taps of which kt - 1 read nothing. */
    int32_t p_out_plane;
    int32_t g_pout;
    /* Staged form, used by the operand-building tensor-core kernel (kgan_tapconv_fwd_tf32, csrc/tapconv_build.cu): the number of
     * input positions one 128-row output tile can touch through the taps of one channel block (taps with equal tap_in_ch), counted
     * from a 4-aligned first position and rounded up to 4; for planes of at most 128 output positions: p_in (whole planes are staged).
     * 0 = not available (p_in not a multiple of 4, or more than 512).  Computed from pmap by the caller (geometry.py).
     * prefer_staged != 0: take that kernel even where the TMA-fed one is eligible (same-channel multi-tap convolutions: one staged
     * tile serves all taps). */
    int32_t stage_span;
    int32_t prefer_staged;
    /* Fused graph convolution (kgan_gcn_fwd_tf32 only; 0, 0, 0 elsewhere): the adjacency product is taken inside the kernel.  `in` is the
     * block input x (c_in_total = ck channels, planes of T * mix_v positions), the output planes hold T * mix_w positions, the ntap taps
     * are the K partitions (tap_in_ch all 0; tap_w_off = the partitions' weight blocks) and tap k contracts the channels of
     * x * A[k] (A: (ntap, mix_v, mix_w)).  mix_l >= the largest number of non-zeros in a column of any A[k] (<= 8). */
    int32_t mix_v, mix_w, mix_l;
} kgan_tapconv_desc;

/* Version / diagnostics. */
int kgan_version(void);
const char* kgan_last_error(void);
/* 1 if the device behind the current context is sm_100 (B200); kernels refuse to run otherwise. */
int kgan_device_ok(void);

/* ---- tap convolution ---------------------------------------------------------------------------
 * Replaces: nn.Conv2d 1x1 of the graph conv (models/init_gan/tgcn.py:48,61) once the adjacency has been
 * applied by kgan_adjmix_fwd; the temporal conv (models/generator.py:134-140, models/discriminator.py:99-105);
 * the residual 1x1 conv (generator.py:155-159, discriminator.py:115-120); nn.Linear of the mapping network
 * (generator.py:28, applied at :84-85) and of the critic head (discriminator.py:50,72); the fused epilogue
 * replaces the bias add, `tcn(x) + res` (generator.py:176, discriminator.py:130), downsample_s
 * (discriminator.py:139-142), F.interpolate nearest (discriminator.py:134) and LeakyReLU (discriminator.py:136).
 * The data gradient (convolution_backward w.r.t. input) is the same entry point called with the transposed
 * weight strides and the inverse position map; see kinetic-gan_b200/geometry.py.
 * `bias` (co floats) and `add` (same shape as out) may be NULL. */
int kgan_tapconv_fwd(const kgan_tapconv_desc* d, const float* in, const float* w, const int32_t* pmap,
                     const float* bias, const float* add, float* out, void* stream);

/* Tensor-core path of kgan_tapconv_fwd: tcgen05.mma kind::tf32 (weights rounded to tf32 by the packing kernel, activations
 * expected tf32-rounded - see `out_tf32` in the conventions; fp32 accumulation in TMEM; rel-L2 ~3e-4 per layer).  The weights are first packed into the shared-memory image the
 * kernel streams with bulk copies:
 *   kgan_tapconv_tf32_workspace(d)  -> number of floats of the packed image; 0 if the shape is not eligible
 *                                      (ck < 16, co < 16 or fewer than 256 output positions: use kgan_tapconv_fwd)
 *   kgan_tapconv_pack_tf32(d, w, wp) -> fills wp (caller-owned, 16-byte aligned) from the natural weights
 *   kgan_tapconv_fwd_tf32(d, in, wp, ...) -> same semantics as kgan_tapconv_fwd
 * A packed image depends on the descriptor's geometry but not on d->n / act / add_period. */
int64_t kgan_tapconv_tf32_workspace(const kgan_tapconv_desc* d);
int kgan_tapconv_pack_tf32(const kgan_tapconv_desc* d, const float* w, float* wp, void* stream);
int kgan_tapconv_fwd_tf32(const kgan_tapconv_desc* d, const float* in, const float* wp, const int32_t* pmap,
                          const float* bias, const float* add, float* out, void* stream);
/* One launch that (re)packs `count` weights - what a trainer calls after its optimizer step instead of `count` launches of
 * kgan_tapconv_pack_tf32.  descs / w / wp are HOST arrays (w[i], wp[i] device pointers as above); `items` is a caller-owned
 * device workspace of count * kgan_tapconv_pack_item_bytes() bytes that holds the item table: pass upload = 1 whenever the
 * set of items changed since the last call with this workspace (the table is rebuilt and copied), 0 to reuse it. */
int64_t kgan_tapconv_pack_item_bytes(void);
int kgan_tapconv_pack_tf32_batched(int count, const kgan_tapconv_desc* descs, const float* const* w, float* const* wp, void* items,
                                   int upload, void* stream);

/* Graph convolution as ONE kernel (tgcn.py:61-66: conv1x1 to K*C_out channels, then einsum('nkctv,kvw->nctw'), refolded):
 *     out[n, oc, t, w] = act( bias[oc] + add[...] + sum_k sum_ic W[k*C_out + oc, ic] * sum_v x[n, ic, t, v] * A[k, v, w] )
 * The x tile is staged by TMA, the (A (.) edge_importance) product is taken over the non-zeros of A's columns by the operand-building
 * warps, written tf32-rounded into the swizzled shared-memory operand image, and the K partitions are K taps of one tcgen05 GEMM - the
 * K*C-channel mixed tensor of kgan_adjmix_fwd + kgan_tapconv_fwd_tf32 never exists.  `d`: see mix_v / mix_w / mix_l; wp: the packed
 * image of d.  The callers use it where nothing else reads the mixed tensor (no weight gradient will be taken: inference, the critic
 * pass of the generator update, the interpolate pass of the gradient penalty); kgan_gcn_fused_ok(d) == 1 when eligible.
 * `omap` (optional): scatter store of the result as in kgan_tapconv_fwd_tf32_scatter (d->p_out_plane = positions of the output plane; `add` NULL). */
int kgan_gcn_fused_ok(const kgan_tapconv_desc* d);
int kgan_gcn_fwd_tf32(const kgan_tapconv_desc* d, const float* x, const float* wp, const float* A, const float* bias, const float* add,
                      const int32_t* omap, float* out, void* stream);

/* Tap convolution stored through a scatter table: the p_out computed positions of a plane go to an output plane of d->p_out_plane
 * positions - position p to omap[3p] and (if >= 0) also to omap[3p + 1]; omap[3p + 2] (if >= 0) is a slot that position keeps ZERO.
 * Every slot of the output plane must be covered by exactly one of the three roles.  Replaces "graph conv, then gather its output into the
 * time-unfolded layout of the strided temporal conv" (geometry.UnfoldedTcnGeom.unfold via kgan_plane_spmm): the unfolded tensor is written
 * by the convolution's epilogue, the intermediate and the copy kernel disappear.  d->groups == 1, no `add`.
 * kgan_tapconv_scatter_ok(d) == 1 when eligible (TMA-fed plan). */
/* The tap convolution with the generator's NoiseInjection in its epilogue - the eval-mode generator block (generator.py:168-182 with the
 * BatchNorm folded into the weights) as ONE kernel:
 *     out = act( conv_d(in) + bias + add + nw[oc] * noise[n, 0, p] )        noise: (n, 1, T, V), nw: (C_out)
 * Runs on the operand-building kernel (the generator's planes of 5 / 11 / 25 joints); kgan_tapconv_noise_ok(d) == 1 when eligible
 * (tf32 mode, groups == 1, no TMA-fed plan for d). */
int kgan_tapconv_noise_ok(const kgan_tapconv_desc* d);
int kgan_tapconv_fwd_tf32_noise(const kgan_tapconv_desc* d, const float* in, const float* wp, const int32_t* pmap, const float* bias, const float* add,
                                const float* noise, const float* nw, float* out, void* stream);
int kgan_tapconv_scatter_ok(const kgan_tapconv_desc* d);
int kgan_tapconv_fwd_tf32_scatter(const kgan_tapconv_desc* d, const float* in, const float* wp, const int32_t* omap, const float* bias,
                                  float* out, void* stream);

/* Tap convolution with a fused residual branch (the critic block's `tcn(gcn(x)) + residual(x)` then LeakyReLU, discriminator.py:128-136):
 *     out = act( conv_d(in) + bias + conv_d2(in2) + bias2 )
 * d2 is a 1x1 convolution (one tap, shift 0, groups 1) of a SECOND tensor with the same samples, output channels and plane as d's output;
 * its K panel is accumulated into the same TMEM accumulator, so the residual never goes to HBM and back.  wp / wp2 are the
 * packed images of d / d2 (kgan_tapconv_pack_tf32).  kgan_tapconv_res_ok(d, d2) == 1 when the pair is eligible (both TMA-fed plans, same
 * channel tiling); otherwise run kgan_tapconv_fwd_tf32 twice with the first result as `add`. */
int kgan_tapconv_res_ok(const kgan_tapconv_desc* d, const kgan_tapconv_desc* d2);
int kgan_tapconv_fwd_tf32_res(const kgan_tapconv_desc* d, const float* in, const float* wp, const kgan_tapconv_desc* d2, const float* in2,
                              const float* wp2, const float* bias, const float* bias2, float* out, void* stream);

/* 1 if kgan_tapconv_fwd_tf32 will run the TMA-fed kernel for this descriptor: activations then reach shared memory by
 * cp.async.bulk.tensor instead of per-thread gathers.  Eligible: tma_mode != 0, tensor-core eligible, and either
 *   - planes of a multiple of 4 positions with every tap shift a multiple of 4 positions (activations as an MN-major operand:
 *     boxes of 32 positions x 32 channels, a temporal tap is a coordinate shift), or
 *   - one-position planes (p_in == p_out == 1: nn.Linear) with all shifts 0 and c_in_total a multiple of 4 (the (N, C)
 *     activation matrix as a K-major operand: boxes of 128 samples x 32 channels). */
int kgan_tapconv_tma_ok(const kgan_tapconv_desc* d);
/* 1 if kgan_tapconv_fwd_tf32 will run the operand-building kernel for this descriptor (stage_span > 0, tensor-core eligible, and
 * either prefer_staged or not eligible for the TMA-fed kernel): raw tiles by cp.async.bulk.tensor, tap operands gathered in shared
 * memory through pmap - any injective map, no alignment rule on shifts. */
int kgan_tapconv_staged_ok(const kgan_tapconv_desc* d);

/* Tensor-core path of kgan_tapconv_wgrad (tcgen05.mma kind::tf32, split-K over CTAs, fp32 atomics into dw).
 * kgan_tapconv_wgrad_tf32_ok(d) -> 1 if the shape is eligible (else use kgan_tapconv_wgrad). */
int kgan_tapconv_wgrad_tf32_ok(const kgan_tapconv_desc* d);
/* 1 if kgan_tapconv_wgrad_tf32 will feed both operands with cp.async.bulk.tensor (shift form, output plane a multiple of 32
 * positions, shifts multiples of 4) instead of per-thread cp.async copies. */
int kgan_tapconv_wgrad_tma_ok(const kgan_tapconv_desc* d);
int kgan_tapconv_wgrad_tf32(const kgan_tapconv_desc* d, const float* in, const float* gout, const int32_t* pmap,
                            float* dw, int64_t dw_numel, int accumulate, void* stream);

/* dW[...same addressing as W...] = sum_{n,p} gout[n, out_ch0+oc, p] * in[n, in_ch0+tap_in_ch+ic, pmap[..]]
 * Replaces convolution_backward w.r.t. weight for the same call sites, and (called with swapped roles)
 * the weight terms of _convolution_double_backward used by the gradient penalty (kinetic-gan.py:104-113,154).
 * `dw` must hold `dw_numel` floats.  accumulate == 0: dw is overwritten (zeroed, then accumulated with fp32 atomics);
 * accumulate != 0: the result is ADDED to what dw holds - what autograd's AccumulateGrad does with a second launch
 * (p.grad += dw) when dw is a view of the flat gradient buffer (kinetic-gan.py:154,173 `backward()`). */
int kgan_tapconv_wgrad(const kgan_tapconv_desc* d, const float* in, const float* gout, const int32_t* pmap,
                       float* dw, int64_t dw_numel, int accumulate, void* stream);

/* ---- adjacency product ---------------------------------------------------------------------------
 * out[r, k, w] = sum_v x[r, v] * A[k, v, w]      rows r = (n, c, t); written as (N, K*C, T, W)
 * Replaces torch.einsum('nkctv,kvw->nctw') at tgcn.py:66 (applied before the 1x1 conv, which commutes with it
 * because gcn.conv has bias=False, tgcn.py:44,55).  A is (K, V, W) - rectangular so that upsample_s
 * (generator.py:185-200) can be folded into it. */
int kgan_adjmix_fwd(const float* x, const float* A, float* out, int n, int c, int t, int v, int w, int k, int out_tf32, void* stream);
/* The same product with a by-product: out2[n, c, q] = x[n, c, sidx[q]], q < pc - x gathered at the plane positions sidx (int32, t * v
 * indexed) - the input of the critic block's residual branch (x[..., keep] at every other frame, discriminator.py:132-134), written from
 * the tile the product has staged in shared memory instead of by a gather kernel that reads x from HBM again.
 * kgan_adjmix_fwd_sel_ok(...) == 1 when the pipelined plan takes the shape with whole channel planes per tile. */
int kgan_adjmix_fwd_sel_ok(const float* x, int n, int c, int t, int v, int w, int k);
int kgan_adjmix_fwd_sel(const float* x, const float* A, const int32_t* sidx, int pc, float* out, float* out2, int n, int c, int t, int v, int w,
                        int k, int out_tf32, void* stream);
/* gx[r, v] = sum_k sum_w gout[r, k, w] * A[k, v, w] */
int kgan_adjmix_bwd_x(const float* gout, const float* A, float* gx, int n, int c, int t, int v, int w, int k, int out_tf32, void* stream);
/* Same with a fused epilogue - the join of a critic block's backward (discriminator.py:128-136 differentiated):
 *   gx[r, v] = ( sum_k sum_w gout[r, k, w] * A[k, v, w]  +  add[r, v] ) * slope(ysrc[r, v]),   slope(y) = y > 0 ? 1 : 0.2
 * `add` (the residual branch's input gradient) and `ysrc` (the block input = the previous block's LeakyReLU output, whose
 * activation gradient is applied here) are optional, shaped like gx.  Replaces an elementwise add (autograd's accumulation of the
 * two branches) and leaky_relu_backward over the same tensor. */
int kgan_adjmix_bwd_x_fused(const float* gout, const float* A, const float* add, const float* ysrc, float* gx, int n, int c, int t, int v,
                            int w, int k, int out_tf32, void* stream);
/* Same, with the residual branch's gradient given where that branch lives: `add_c` is COMPACT, shaped (n, c, pc) - the frames / joints
 * the residual branch kept (discriminator.py:132-134: x[..., keep] at every other frame) - and inv[p] (int32, t * v entries) is the
 * compact position of plane position p, -1 where the branch did not read x:
 *   gx[r, p] = ( mix[r, p] + (inv[p] >= 0 ? add_c[r, inv[p]] : 0) ) * slope(ysrc[r, p])
 * The adjoint of the selection is taken here instead of by a scatter kernel writing a mostly-zero full-size tensor. */
int kgan_adjmix_bwd_x_fused_sel(const float* gout, const float* A, const float* add_c, const int32_t* inv, int pc, const float* ysrc, float* gx,
                                int n, int c, int t, int v, int w, int k, int out_tf32, void* stream);
/* gA[k, v, w] = sum_r x[r, v] * gout[r, k, w]  (gA overwritten) */
int kgan_adjmix_bwd_a(const float* x, const float* gout, float* gA, int n, int c, int t, int v, int w, int k, void* stream);
/* Same, restricted to the support of `mask` (K, V, W): gA[k, v, w] = 0 where mask[k, v, w] == 0.  The callers pass the
 * adjacency A (.) edge_importance itself (generator.py:93, discriminator.py:64): its gradient only flows on into
 * edge_importance through the product with A, so entries outside the skeleton's support are never used; mask == NULL
 * computes every entry. */
int kgan_adjmix_bwd_a_masked(const float* x, const float* gout, const float* mask, float* gA, int n, int c, int t, int v, int w, int k,
                             void* stream);

/* ---- pointwise epilogues -------------------------------------------------------------------------
 * out[n,c,p] = act( a[n,c,p] + b[n,c,p] + bias[c] + nw[c] * noise[n,p] ); b, bias, (nw,noise) may be NULL.
 * Replaces `tcn(x) + res`, NoiseInjection (generator.py:12-19,179-180) and LeakyReLU/Tanh (generator.py:182). */
int kgan_epilogue_fwd(const float* a, const float* b, const float* bias, const float* nw, const float* noise,
                      float* out, int n, int c, int p, int act, int out_tf32, void* stream);
/* gz = gout * act'(out)   (LeakyReLU: out > 0 ? 1 : 0.2;  tanh: 1 - out^2).  Replaces leaky_relu_backward /
 * tanh_backward; also the second-order use of the same mask in the gradient penalty. */
int kgan_act_bwd(const float* gout, const float* out, float* gz, int64_t numel, int act, int out_tf32, void* stream);
/* out[c] = sum_{n,p} g[n,c,p] * (mul ? mul[n,p] : 1).  Bias gradients and NoiseInjection.weight gradient. */
int kgan_chan_reduce(const float* g, const float* mul, float* out, int n, int c, int p, void* stream);

/* ---- plane gather / scatter ----------------------------------------------------------------------
 * out[r, q] = sum_{j<J} wgt[q*J+j] * x[r, idx[q*J+j]]   (idx < 0 skipped); r over N*C planes.
 * Replaces upsample_s + F.interpolate in G (generator.py:170-172,185-200), downsample_s + F.interpolate on the
 * identity residual in D (discriminator.py:128-134), avg_pool2d (discriminator.py:68), and their backward passes
 * (the transposed table). */
int kgan_plane_spmm(const float* x, const int32_t* idx, const float* wgt, float* out, int64_t rows, int p_in, int p_out,
                    int j, int out_tf32, void* stream);

/* out[r, v] = sum_t x[r, t, v] over the N*C planes r (V <= 32): the adjoint of broadcasting a per-joint term along T (the label term of
 * the critic's first layer, added with add_period in the tap convolution's epilogue). */
int kgan_plane_sum_t(const float* x, float* out, int64_t rows, int t, int v, int out_tf32, void* stream);

/* ---- label planes (discriminator.py:57-60) -------------------------------------------------------
 * out[n, c, p] = c < n_cls ? e[n, c] : x[n, c - n_cls, p] */
int kgan_label_concat(const float* e, const float* x, float* out, int n, int n_cls, int c, int p, int out_tf32, void* stream);
/* ge[n, c] = sum_p g[n, c, p] (c < n_cls);  gx[n, c, p] = g[n, n_cls + c, p].  Either output may be NULL. */
int kgan_label_split(const float* g, float* ge, float* gx, int n, int n_cls, int c, int p, int out_tf32, void* stream);

/* ---- BatchNorm2d, training mode (generator.py:142,160) ---------------------------------------------
 * stats: mean[c], rstd[c] from biased variance over (n, p); running stats updated in place with `momentum`
 * (unbiased variance), exactly nn.BatchNorm2d defaults.  running_* may be NULL.
 * `workspace`: kgan_bn_workspace(n, c) floats owned by the caller (per-chunk partial sums; they are added in a fixed order,
 * so statistics and gradients are bit-reproducible - no atomics). */
int64_t kgan_bn_workspace(int n, int c);
int kgan_bn_stats(const float* x, float* mean, float* rstd, float* running_mean, float* running_var, int n, int c, int p,
                  float eps, float momentum, float* workspace, void* stream);
/* y = (x - mean[c]) * rstd[c] * gamma[c] + beta[c]  (also eval mode with running stats folded by the caller) */
int kgan_bn_apply(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, float* y,
                  int n, int c, int p, int out_tf32, void* stream);
/* Training-mode BatchNorm, residual add, NoiseInjection and activation of a generator block in ONE pass (generator.py:160,176-182):
 *   out = act( (a - mean[c]) * rstd[c] * gamma[c] + beta[c]  +  b  +  nw[c] * noise[n, 0, p] ),   mean / rstd from kgan_bn_stats;
 * b, (nw, noise) optional.  Replaces kgan_bn_apply followed by kgan_epilogue_fwd (one read and one write of the tensor less). */
int kgan_bn_epilogue_fwd(const float* a, const float* mean, const float* rstd, const float* gamma, const float* beta, const float* b,
                         const float* nw, const float* noise, float* out, int n, int c, int p, int act, int out_tf32, void* stream);
/* gx, ggamma[c], gbeta[c] of training-mode BN */
int kgan_bn_bwd(const float* gy, const float* x, const float* mean, const float* rstd, const float* gamma, float* gx,
                float* ggamma, float* gbeta, int n, int c, int p, int out_tf32, float* workspace, void* stream);

/* ---- fused Adam over a flat parameter buffer (torch.optim.Adam, kinetic-gan.py:77-78,155,174) ----
 * g is multiplied by grad_scale first (1/world_size after the DDP sum all-reduce). `step` is 1-based. */
int kgan_adam_step(float* p, const float* g, float* m, float* v, int64_t numel, float lr, float b1, float b2, float eps,
                   int step, float grad_scale, void* stream);

/* out = alpha * x + (1 - alpha) * y with alpha per sample (kinetic-gan.py:100); per_sample = C*T*V */
int kgan_interpolate(const float* alpha, const float* x, const float* y, float* out, int n, int64_t per_sample, int out_tf32, void* stream);

/* out = x rounded to tf32 (round to nearest, ties away; in place when out == x): for tensors that enter the tf32 path from
 * outside the library (the latent / label-embedding rows fed to the mapping network, generator.py:80-85). */
int kgan_round_tf32(const float* x, float* out, int64_t numel, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KGAN_H_ */
