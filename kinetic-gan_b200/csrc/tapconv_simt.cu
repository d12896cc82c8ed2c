// Tap convolution, fp32 SIMT path (KGAN_PREC_FP32): exact-fp32 FMA GEMM used for the <=1e-5 parity
// path, for the small / ragged layers (C <= 9 tail of G, D0's 3 input channels, critic head) and as the
// on-device cross-check of the tcgen05 path.  See include/kgan.h for the operator definition.
//
// GEMM view: rows = flattened output positions (n, p), columns = output channels,
// contraction = (tap, input channel).  Positions are contiguous in memory for a fixed channel (NCHW), so
// lanes run along positions for every global load/store (coalesced 128 B segments).
#include "common.cuh"

namespace kgan {

constexpr int BM = 128;   // positions per CTA
constexpr int BK = 16;    // contraction slice
constexpr int TM = 8;     // positions per thread
constexpr int NT = 256;   // threads per CTA

template <int TN>   // output channels per thread; BN = 16 * TN
__global__ void __launch_bounds__(NT) tapconv_fwd_simt(const __grid_constant__ kgan_tapconv_desc d, const float* __restrict__ in,
                                                       const float* __restrict__ w, const int32_t* __restrict__ pmap,
                                                       const float* __restrict__ bias, const float* __restrict__ add,
                                                       float* __restrict__ out) {
    constexpr int BN = 16 * TN;
    constexpr int WS = BN + 4;
    __shared__ float Xs[BK][BM];
    __shared__ __align__(16) float Ws[BK][WS];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int g = blockIdx.z;
    const int in_ch0 = g * d.g_in, out_ch0 = g * d.g_out;
    const float* wg = w + (int64_t)g * d.g_w;
    const int64_t total_pos = (int64_t)d.n * d.p_out;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int oc0 = blockIdx.y * BN;

    // X loader role: one position, 8 contraction rows
    const int lm = tid & (BM - 1), lk0 = tid >> 7;             // rows lk0 + 2*j
    const int64_t lpos = m0 + lm;
    const bool lvalid = lpos < total_pos;
    const int ln = lvalid ? (int)(lpos / d.p_out) : 0;
    const int lp = lvalid ? (int)(lpos % d.p_out) : 0;
    // W loader role
    const bool ic_fast = d.w_ic <= d.w_oc;
    const int wk = ic_fast ? (tid & 15) : (tid >> 4);          // contraction row
    const int wo = ic_fast ? (tid >> 4) : (tid & 15);          // first column; columns wo + 16*j

    const int nk = ceil_div(d.ck, BK);
    const int iters = d.ntap * nk;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float xr[8], wr[TN];
    auto load = [&](int it) {
        const int tap = it / nk, ic0 = (it - tap * nk) * BK;
        const int src = lvalid ? pmap[(int64_t)d.tap_row[tap] * d.p_out + lp] : -1;
        const float* xb = in + ((int64_t)ln * d.c_in_total + in_ch0 + d.tap_in_ch[tap] + ic0) * d.p_in + src;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int kk = lk0 + 2 * j;
            xr[j] = (src >= 0 && ic0 + kk < d.ck) ? __ldg(xb + (int64_t)kk * d.p_in) : 0.f;
        }
        const float* wb = wg + d.tap_w_off[tap];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int oc = oc0 + wo + 16 * j;
            wr[j] = (oc < d.co && ic0 + wk < d.ck) ? __ldg(wb + w_oc_offset(d, oc) + (int64_t)(ic0 + wk) * d.w_ic) : 0.f;
        }
    };

    load(0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) Xs[lk0 + 2 * j][lm] = xr[j];
#pragma unroll
        for (int j = 0; j < TN; ++j) Ws[wk][wo + 16 * j] = wr[j];
        __syncthreads();
        if (it + 1 < iters) load(it + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float xv[TM], wv[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) xv[i] = Xs[kk][tx + 16 * i];
#pragma unroll
            for (int j = 0; j < TN; ++j) wv[j] = Ws[kk][ty * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t pos = m0 + tx + 16 * i;
        if (pos >= total_pos) continue;
        const int nn = (int)(pos / d.p_out), p = (int)(pos % d.p_out);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int oc = oc0 + ty * TN + j;
            if (oc >= d.co) continue;
            const int pg = p + g * d.g_pout;                      // position inside the output plane (position-block groups)
            const int64_t o = ((int64_t)nn * d.c_out_total + out_ch0 + oc) * out_plane(d) + pg;
            float v = acc[i][j];
            if (bias) v += __ldg(bias + out_ch0 + oc);
            if (add) v += __ldg(add + (d.add_period ? ((int64_t)nn * d.c_out_total + out_ch0 + oc) * d.add_period + pg % d.add_period : o));
            out[o] = tf32_out(apply_act(v, d.act), d.precision == KGAN_PREC_TF32);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// weight gradient: dW[tap, oc, ic] = sum_pos gout[oc, pos] * in[(tap, ic), pmap(pos)]
// CTA tile 64 oc x 64 ic for one tap, contraction over a chunk of positions; fp32 atomics merge chunks.
// ---------------------------------------------------------------------------------------------
constexpr int WB = 64, WK = 32, WSTR = WB + 1;

__global__ void __launch_bounds__(NT) tapconv_wgrad_simt(const __grid_constant__ kgan_tapconv_desc d, const float* __restrict__ in,
                                                         const float* __restrict__ gout, const int32_t* __restrict__ pmap,
                                                         float* __restrict__ dw, int nchunks, int64_t chunk) {
    __shared__ float Gs[WK][WSTR];
    __shared__ float Xs[WK][WSTR];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int ictiles = ceil_div(d.ck, WB);
    const int tap = blockIdx.x / ictiles, ic0 = (blockIdx.x % ictiles) * WB;
    const int oc0 = blockIdx.y * WB;
    const int g = blockIdx.z / nchunks, ch = blockIdx.z % nchunks;
    const int in_ch0 = g * d.g_in + d.tap_in_ch[tap], out_ch0 = g * d.g_out;
    const int64_t total_pos = (int64_t)d.n * d.p_out;
    const int64_t pbeg = (int64_t)ch * chunk, pend = min(total_pos, pbeg + chunk);
    const int lane = tid & 31, r0 = tid >> 5;                    // channel rows r0 + 8*j
    const int32_t* prow = pmap + (int64_t)d.tap_row[tap] * d.p_out;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int64_t k0 = pbeg; k0 < pend; k0 += WK) {
        const int64_t pos = k0 + lane;
        const bool valid = pos < pend;
        const int nn = valid ? (int)(pos / d.p_out) : 0, p = valid ? (int)(pos % d.p_out) : 0;
        const int src = valid ? prow[p] : -1;
        const float* gb = gout + ((int64_t)nn * d.c_out_total + out_ch0 + oc0) * d.p_out + p;
        const float* xb = in + ((int64_t)nn * d.c_in_total + in_ch0 + ic0) * d.p_in + src;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = r0 + 8 * j;
            Gs[lane][c] = (valid && oc0 + c < d.co) ? __ldg(gb + (int64_t)c * d.p_out) : 0.f;
            Xs[lane][c] = (src >= 0 && ic0 + c < d.ck) ? __ldg(xb + (int64_t)c * d.p_in) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < WK; ++kk) {
            float gv[4], xv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) gv[i] = Gs[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) xv[j] = Xs[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gv[i], xv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* wb = dw + (int64_t)g * d.g_w + d.tap_w_off[tap];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int oc = oc0 + ty * 4 + i;
        if (oc >= d.co) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ic = ic0 + tx + 16 * j;
            if (ic >= d.ck) continue;
            atomicAdd(wb + w_oc_offset(d, oc) + (int64_t)ic * d.w_ic, acc[i][j]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// "Thin" layers: ck * ntap * co <= 1024 and co <= 32 (the 3-channel tail of the generator, the critic's first layer on the 3
// data channels, their data gradients).  As GEMMs they would pad a contraction of 3..32 to a 128 x 16 x 32 tile per tap and
// spend 20-50x the memory time; here a thread owns one output position, streams the ck * ntap inputs it needs once (lanes run
// along positions: coalesced), keeps CO accumulators in registers and reads the weights as shared-memory broadcasts.
// Exact fp32 FMA arithmetic: used in both precision modes.
// ---------------------------------------------------------------------------------------------
constexpr int THIN_MAX_W = 1024;

// groups that read the same input channels (g_in == 0: the K partitions of a graph conv's data gradient) are merged into one pass:
// accumulator a = (group, oc), so the input is streamed once instead of once per group
static inline int thin_merge(const kgan_tapconv_desc& d) { return (d.groups > 1 && d.g_in == 0 && d.g_pout == 0 && d.groups * d.co <= 16) ? d.groups : 1; }

template <int CO>   // accumulators per thread (ng * co <= CO), CO in {4, 8, 16}
__global__ void __launch_bounds__(NT, 4) tapconv_fwd_thin(const __grid_constant__ kgan_tapconv_desc d, const float* __restrict__ in,
                                                       const float* __restrict__ w, const int32_t* __restrict__ pmap,
                                                       const float* __restrict__ bias, const float* __restrict__ add, float* __restrict__ out,
                                                       int ng) {
    __shared__ __align__(16) float ws[4096];     // ntap * ck * CO <= 4096 (tapconv_is_thin)
    __shared__ int och[CO];                      // output channel of accumulator a, -1: unused
    const int g0 = blockIdx.y * ng, na = ng * d.co;
    // weights -> ws[(tap * ck + ic) * CO + a], zero for a >= na
    for (int i = threadIdx.x; i < d.ntap * d.ck * CO; i += NT) {
        const int a = i % CO, r = i / CO, ic = r % d.ck, tap = r / d.ck;
        const int gg = a / d.co, oc = a - gg * d.co;
        ws[i] = a < na ? __ldg(w + (int64_t)(g0 + gg) * d.g_w + d.tap_w_off[tap] + w_oc_offset(d, oc) + (int64_t)ic * d.w_ic) : 0.f;
    }
    if (threadIdx.x < CO) och[threadIdx.x] = (int)threadIdx.x < na ? (g0 + threadIdx.x / d.co) * d.g_out + threadIdx.x % d.co : -1;
    __syncthreads();
    const int in_ch0 = g0 * d.g_in;
    const int64_t total = (int64_t)d.n * d.p_out;
    for (int64_t pos = (int64_t)blockIdx.x * NT + threadIdx.x; pos < total; pos += (int64_t)gridDim.x * NT) {
        // (a 64-bit division per position is ~150 instructions - a fifth of this loop; positions below 2^31 divide in 32 bits)
        const int nn = total < (1ll << 31) ? (int)pos / d.p_out : (int)(pos / d.p_out), p = (int)(pos - (int64_t)nn * d.p_out);
        float acc[CO];
#pragma unroll
        for (int j = 0; j < CO; ++j) acc[j] = 0.f;
        const float* xn = in + ((int64_t)nn * d.c_in_total + in_ch0) * d.p_in;
        for (int tap = 0; tap < d.ntap; ++tap) {
            const int src = __ldg(pmap + (int64_t)d.tap_row[tap] * d.p_out + p);
            if (src < 0) continue;
            const float* xb = xn + (int64_t)d.tap_in_ch[tap] * d.p_in + src;
            const float* wt = ws + tap * d.ck * CO;
            // eight channel loads in flight per thread before their FMAs, four blocks per SM (<= 64 registers): with four loads in flight and
            // three resident blocks the kernel was bound by its bytes in flight - 465 -> 367 us on the 32-channel data gradient of the critic's
            // first graph conv (same-box A/B, tools/thin_bench.py)
            int ic = 0;
            for (; ic + 8 <= d.ck; ic += 8) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = __ldg(xb + (int64_t)(ic + u) * d.p_in);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
#pragma unroll
                    for (int j4 = 0; j4 < CO / 4; ++j4) {
                        const float4 q = *reinterpret_cast<const float4*>(wt + (ic + u) * CO + 4 * j4);
                        acc[4 * j4 + 0] = fmaf(v[u], q.x, acc[4 * j4 + 0]);
                        acc[4 * j4 + 1] = fmaf(v[u], q.y, acc[4 * j4 + 1]);
                        acc[4 * j4 + 2] = fmaf(v[u], q.z, acc[4 * j4 + 2]);
                        acc[4 * j4 + 3] = fmaf(v[u], q.w, acc[4 * j4 + 3]);
                    }
                }
            }
            for (; ic < d.ck; ++ic) {
                const float v = __ldg(xb + (int64_t)ic * d.p_in);
#pragma unroll
                for (int j4 = 0; j4 < CO / 4; ++j4) {
                    const float4 q = *reinterpret_cast<const float4*>(wt + ic * CO + 4 * j4);
                    acc[4 * j4 + 0] = fmaf(v, q.x, acc[4 * j4 + 0]);
                    acc[4 * j4 + 1] = fmaf(v, q.y, acc[4 * j4 + 1]);
                    acc[4 * j4 + 2] = fmaf(v, q.z, acc[4 * j4 + 2]);
                    acc[4 * j4 + 3] = fmaf(v, q.w, acc[4 * j4 + 3]);
                }
            }
        }
        const int pg = p + (ng == 1 ? g0 * d.g_pout : 0);           // position-block groups are never merged (thin_merge)
        const int pa = d.add_period ? pg % d.add_period : 0;
#pragma unroll
        for (int j = 0; j < CO; ++j) {
            const int c = och[j];
            if (c >= 0) {
                const int64_t o = ((int64_t)nn * d.c_out_total + c) * out_plane(d) + pg;
                float v = acc[j];
                if (bias) v += __ldg(bias + c);
                if (add) v += __ldg(add + (d.add_period ? ((int64_t)nn * d.c_out_total + c) * d.add_period + pa : o));
                out[o] = tf32_out(apply_act(v, d.act), d.precision == KGAN_PREC_TF32);
            }
        }
    }
}

// The same, four consecutive output positions per thread (planes that are multiples of 4 positions, 32-bit
// indexing): one 16-byte load of the position map per tap, 16-byte input loads where the four sources are contiguous and aligned (the
// centre tap; a shift by an odd number of joints falls back to four scalar loads that hit L1), 16-byte stores - a quarter of the index
// arithmetic and of the memory instructions per output.  (The one-position kernel ran the 3-channel tail of the generator at 1.2-1.8 TB/s.)
template <int CO>
__global__ void __launch_bounds__(NT) tapconv_fwd_thin4(const __grid_constant__ kgan_tapconv_desc d, const float* __restrict__ in,
                                                        const float* __restrict__ w, const int32_t* __restrict__ pmap,
                                                        const float* __restrict__ bias, const float* __restrict__ add, float* __restrict__ out,
                                                        int ng, int in_vec) {
    __shared__ __align__(16) float ws[4096];
    __shared__ int och[CO];
    const int g0 = blockIdx.y * ng, na = ng * d.co;
    for (int i = threadIdx.x; i < d.ntap * d.ck * CO; i += NT) {
        const int a = i % CO, r = i / CO, ic = r % d.ck, tap = r / d.ck;
        const int gg = a / d.co, oc = a - gg * d.co;
        ws[i] = a < na ? __ldg(w + (int64_t)(g0 + gg) * d.g_w + d.tap_w_off[tap] + w_oc_offset(d, oc) + (int64_t)ic * d.w_ic) : 0.f;
    }
    if (threadIdx.x < CO) och[threadIdx.x] = (int)threadIdx.x < na ? (g0 + threadIdx.x / d.co) * d.g_out + threadIdx.x % d.co : -1;
    __syncthreads();
    const int in_ch0 = g0 * d.g_in;
    const int quads = d.p_out >> 2, total = d.n * quads, plane = out_plane(d);
    const int rnd = d.precision == KGAN_PREC_TF32;
    for (int qd = blockIdx.x * NT + threadIdx.x; qd < total; qd += gridDim.x * NT) {
        const int nn = qd / quads, p = (qd - nn * quads) << 2;
        float acc[CO][4];
#pragma unroll
        for (int j = 0; j < CO; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        const float* xn = in + (nn * d.c_in_total + in_ch0) * d.p_in;
        for (int tap = 0; tap < d.ntap; ++tap) {
            const int4 s4 = __ldg(reinterpret_cast<const int4*>(pmap + d.tap_row[tap] * d.p_out + p));
            if ((s4.x & s4.y & s4.z & s4.w) < 0) continue;                        // all four sources are padding
            const float* xb = xn + d.tap_in_ch[tap] * d.p_in;
            const float* wt = ws + tap * d.ck * CO;
            const bool contig = in_vec && s4.x >= 0 && !(s4.x & 3) && s4.y == s4.x + 1 && s4.z == s4.x + 2 && s4.w == s4.x + 3;
#pragma unroll 2
            for (int ic = 0; ic < d.ck; ++ic) {
                const float* xc = xb + ic * d.p_in;
                float4 v;
                if (contig) {
                    v = __ldg(reinterpret_cast<const float4*>(xc + s4.x));
                } else {
                    v.x = s4.x >= 0 ? __ldg(xc + s4.x) : 0.f;
                    v.y = s4.y >= 0 ? __ldg(xc + s4.y) : 0.f;
                    v.z = s4.z >= 0 ? __ldg(xc + s4.z) : 0.f;
                    v.w = s4.w >= 0 ? __ldg(xc + s4.w) : 0.f;
                }
#pragma unroll
                for (int j4 = 0; j4 < CO / 4; ++j4) {
                    const float4 q = *reinterpret_cast<const float4*>(wt + ic * CO + 4 * j4);
                    const float qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        acc[4 * j4 + e][0] = fmaf(v.x, qq[e], acc[4 * j4 + e][0]);
                        acc[4 * j4 + e][1] = fmaf(v.y, qq[e], acc[4 * j4 + e][1]);
                        acc[4 * j4 + e][2] = fmaf(v.z, qq[e], acc[4 * j4 + e][2]);
                        acc[4 * j4 + e][3] = fmaf(v.w, qq[e], acc[4 * j4 + e][3]);
                    }
                }
            }
        }
        const int pg = p + (ng == 1 ? g0 * d.g_pout : 0);
#pragma unroll
        for (int j = 0; j < CO; ++j) {
            const int c = och[j];
            if (c < 0) continue;
            const int o = (nn * d.c_out_total + c) * plane + pg;
            float r[4] = {acc[j][0], acc[j][1], acc[j][2], acc[j][3]};
            const float b = bias ? __ldg(bias + c) : 0.f;
            if (add && d.add_period == 0) {
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(add + o));
                r[0] += a4.x, r[1] += a4.y, r[2] += a4.z, r[3] += a4.w;
            } else if (add) {
                const int ab = (nn * d.c_out_total + c) * d.add_period;
#pragma unroll
                for (int e = 0; e < 4; ++e) r[e] += __ldg(add + ab + (pg + e) % d.add_period);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) r[e] = tf32_out(apply_act(r[e] + b, d.act), rnd);
            *reinterpret_cast<float4*>(out + o) = make_float4(r[0], r[1], r[2], r[3]);
        }
    }
}

// eligibility of the four-position kernel: vector-addressable planes, 32-bit element indices
static bool thin4_ok(const kgan_tapconv_desc& d, const void* pmap, const void* add, const void* out) {
    const int na = thin_merge(d) * d.co, plane = out_plane(d);
    if (na > 8 || (d.p_out & 3) || (plane & 3) || (d.g_pout & 3)) return false;      // (16 accumulators x 4 positions: 127 registers, measured slower)
    if ((reinterpret_cast<uintptr_t>(pmap) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(add)) & 15) return false;
    return (int64_t)d.n * d.c_in_total * d.p_in < (1ll << 31) && (int64_t)d.n * d.c_out_total * plane < (1ll << 31) &&
           (int64_t)KGAN_MAX_TAPS * d.p_out < (1ll << 31);
}

bool tapconv_is_thin(const kgan_tapconv_desc& d) {
    const int na = thin_merge(d) * d.co;
    if (na > 16 || d.w_oc_blk != 0 || (int64_t)d.n * d.p_out < 4096) return false;
    const int CO = na <= 4 ? 4 : na <= 8 ? 8 : 16;
    return (int64_t)d.ck * d.ntap * na <= THIN_MAX_W && (int64_t)d.ck * d.ntap * CO <= 4096;
}

template <int CO>
static void launch_thin(const kgan_tapconv_desc& d, const float* in, const float* w, const int32_t* pmap, const float* bias, const float* add,
                        float* out, cudaStream_t s) {
    const int ng = thin_merge(d), gy = d.groups / ng;
    const int64_t total = (int64_t)d.n * d.p_out;
    const int64_t cap = ceil_div64(8 * kNumSMs, gy);
    if (CO <= 8 && thin4_ok(d, pmap, add, out)) {
        int64_t gx = ceil_div64(total / 4, NT);
        if (gx > 4 * cap) gx = 4 * cap;              // a thread's work is 4 positions: an uneven second pass over the grid would cost up to 2x
        const int in_vec = (d.p_in & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
        tapconv_fwd_thin4<CO <= 8 ? CO : 8><<<dim3((unsigned)gx, gy), NT, 0, s>>>(d, in, w, pmap, bias, add, out, ng, in_vec);
        return;
    }
    int64_t gx = ceil_div64(total, NT);
    if (gx > cap) gx = cap;
    tapconv_fwd_thin<CO><<<dim3((unsigned)gx, gy), NT, 0, s>>>(d, in, w, pmap, bias, add, out, ng);
}

// Weight gradient of the thinnest layers (ntap * ck * co <= 32: the 3 -> 3 channel tail of the generator).  As a GEMM the tensor-core
// kernel pads M = 3 output channels to 128 and ran this layer at 0.2 TB/s; here a thread walks positions (lanes along positions:
// coalesced), keeps all ntap * co * ck partial sums in registers, and the block reduces them once at the end (shuffles, shared memory,
// one atomic per weight and block).  Exact fp32 FMA arithmetic in every precision mode.
constexpr int TW_MAX = 32;
static bool wgrad_is_thin(const kgan_tapconv_desc& d) {
    return d.groups == 1 && d.w_oc_blk == 0 && d.ntap * d.ck * d.co <= TW_MAX && (int64_t)d.n * d.p_out >= 4096 &&
           (int64_t)d.n * d.c_in_total * d.p_in < (1ll << 31) && (int64_t)d.n * d.c_out_total * d.p_out < (1ll << 31);
}
// TAPS / CO / CK > 0: the shape as compile-time constants (loops fully unrolled, accumulators indexed by constants); 0: any shape within
// TW_MAX, accumulators addressed by a select chain
template <int TAPS, int CO, int CK>
__global__ void __launch_bounds__(NT) tapconv_wgrad_thin(const __grid_constant__ kgan_tapconv_desc d, const float* __restrict__ in,
                                                         const float* __restrict__ gout, const int32_t* __restrict__ pmap, float* __restrict__ dw) {
    __shared__ float red[TW_MAX];
    if (threadIdx.x < TW_MAX) red[threadIdx.x] = 0.f;
    __syncthreads();
    float acc[TW_MAX];
#pragma unroll
    for (int i = 0; i < TW_MAX; ++i) acc[i] = 0.f;
    const int total = d.n * d.p_out, cc = d.co * d.ck;
    for (int pos = blockIdx.x * NT + threadIdx.x; pos < total; pos += gridDim.x * NT) {
        const int nn = pos / d.p_out, p = pos - nn * d.p_out;
        const float* gp = gout + nn * d.c_out_total * d.p_out + p;
        const float* xn = in + nn * d.c_in_total * d.p_in;
        if (TAPS > 0) {
            float g[CO > 0 ? CO : 1];
#pragma unroll
            for (int oc = 0; oc < CO; ++oc) g[oc] = __ldg(gp + oc * d.p_out);
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
                const int src = __ldg(pmap + d.tap_row[tap] * d.p_out + p);
                const float* xb = xn + d.tap_in_ch[tap] * d.p_in + (src < 0 ? 0 : src);
#pragma unroll
                for (int ic = 0; ic < CK; ++ic) {
                    const float x = src < 0 ? 0.f : __ldg(xb + ic * d.p_in);
#pragma unroll
                    for (int oc = 0; oc < CO; ++oc) acc[(tap * CO + oc) * CK + ic] = fmaf(g[oc], x, acc[(tap * CO + oc) * CK + ic]);
                }
            }
            continue;
        }
        for (int tap = 0; tap < d.ntap; ++tap) {
            const int src = __ldg(pmap + d.tap_row[tap] * d.p_out + p);
            if (src < 0) continue;
            const float* xb = xn + d.tap_in_ch[tap] * d.p_in + src;
            for (int oc = 0; oc < d.co; ++oc) {
                const float g = __ldg(gp + oc * d.p_out);
                for (int ic = 0; ic < d.ck; ++ic) {
                    const float v = g * __ldg(xb + ic * d.p_in);
                    const int a = tap * cc + oc * d.ck + ic;
                    // (registers are indexed by constants only after full unrolling: a select chain over the <= 32 accumulators)
#pragma unroll
                    for (int i = 0; i < TW_MAX; ++i) acc[i] += (i == a) ? v : 0.f;
                }
            }
        }
    }
    const int na = d.ntap * cc;
#pragma unroll
    for (int i = 0; i < TW_MAX; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && i < na) atomicAdd(red + i, v);
    }
    __syncthreads();
    if ((int)threadIdx.x < na) {
        const int a = threadIdx.x, tap = a / cc, oc = (a - tap * cc) / d.ck, ic = a - tap * cc - oc * d.ck;
        atomicAdd(dw + d.tap_w_off[tap] + w_oc_offset(d, oc) + (int64_t)ic * d.w_ic, red[a]);
    }
}
static int launch_wgrad_thin(const kgan_tapconv_desc& d, const float* in, const float* gout, const int32_t* pmap, float* dw, int64_t dw_numel,
                             int accumulate, cudaStream_t s) {
    if (!accumulate && cudaMemsetAsync(dw, 0, sizeof(float) * dw_numel, s) != cudaSuccess) return check_launch("tapconv_wgrad memset");
    int64_t gx = ceil_div64((int64_t)d.n * d.p_out, NT);
    if (gx > 4 * kNumSMs) gx = 4 * kNumSMs;
    if (d.ntap == 3 && d.co == 3 && d.ck == 3) tapconv_wgrad_thin<3, 3, 3><<<(unsigned)gx, NT, 0, s>>>(d, in, gout, pmap, dw);
    else tapconv_wgrad_thin<0, 0, 0><<<(unsigned)gx, NT, 0, s>>>(d, in, gout, pmap, dw);
    return check_launch("tapconv_wgrad (thin)");
}

static int validate(const kgan_tapconv_desc* d) {
    KGAN_REQUIRE(d != nullptr, "tapconv: null descriptor");
    KGAN_REQUIRE(d->n > 0 && d->p_in > 0 && d->p_out > 0 && d->ck > 0 && d->co > 0, "tapconv: empty dimension");
    KGAN_REQUIRE(d->ntap >= 1 && d->ntap <= KGAN_MAX_TAPS, "tapconv: ntap=%d out of range", d->ntap);
    KGAN_REQUIRE(d->groups >= 1 && d->groups <= 65535, "tapconv: groups=%d out of range", d->groups);
    KGAN_REQUIRE(d->act >= KGAN_ACT_NONE && d->act <= KGAN_ACT_TANH, "tapconv: bad act %d", d->act);
    KGAN_REQUIRE(d->add_period >= 0 && d->add_period <= d->p_out, "tapconv: bad add_period %d", d->add_period);
    for (int t = 0; t < d->ntap; ++t)
        KGAN_REQUIRE(d->tap_in_ch[t] >= 0 && d->tap_in_ch[t] + d->ck + (d->groups - 1) * d->g_in <= d->c_in_total,
                     "tapconv: tap %d reads channels beyond c_in_total", t);
    KGAN_REQUIRE(d->co + (d->groups - 1) * d->g_out <= d->c_out_total, "tapconv: writes channels beyond c_out_total");
    KGAN_REQUIRE(d->p_out_plane >= 0 && d->g_pout >= 0 && (d->p_out_plane == 0 ? d->g_pout == 0 : (d->g_pout == 0 || d->p_out + (d->groups - 1) * d->g_pout <= d->p_out_plane)),
                 "tapconv: position-block groups write beyond the output plane");
    KGAN_REQUIRE(d->p_out_plane == 0 || d->add_period == 0 || d->g_pout % d->add_period == 0, "tapconv: add_period does not divide g_pout");
    return 0;
}

}  // namespace kgan

using namespace kgan;

extern "C" int kgan_tapconv_fwd(const kgan_tapconv_desc* d, const float* in, const float* w, const int32_t* pmap,
                                const float* bias, const float* add, float* out, void* stream) {
    if (int e = validate(d)) return e;
    KGAN_REQUIRE(in && w && pmap && out, "tapconv_fwd: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t total = (int64_t)d->n * d->p_out;
    const int64_t gx = ceil_div64(total, BM);
    KGAN_REQUIRE(gx < (1ll << 31), "tapconv_fwd: too many positions");
    if (tapconv_is_thin(*d)) {
        const int na = thin_merge(*d) * d->co;
        if (na <= 4) launch_thin<4>(*d, in, w, pmap, bias, add, out, s);
        else if (na <= 8) launch_thin<8>(*d, in, w, pmap, bias, add, out, s);
        else launch_thin<16>(*d, in, w, pmap, bias, add, out, s);
        return check_launch("tapconv_fwd");
    }
    if (d->co <= 32) {
        dim3 grid((unsigned)gx, ceil_div(d->co, 32), d->groups);
        tapconv_fwd_simt<2><<<grid, NT, 0, s>>>(*d, in, w, pmap, bias, add, out);
    } else if (d->co <= 64 || gx * ceil_div(d->co, 128) < 2 * kNumSMs) {
        dim3 grid((unsigned)gx, ceil_div(d->co, 64), d->groups);
        tapconv_fwd_simt<4><<<grid, NT, 0, s>>>(*d, in, w, pmap, bias, add, out);
    } else {
        dim3 grid((unsigned)gx, ceil_div(d->co, 128), d->groups);
        tapconv_fwd_simt<8><<<grid, NT, 0, s>>>(*d, in, w, pmap, bias, add, out);
    }
    return check_launch("tapconv_fwd");
}

extern "C" int kgan_tapconv_wgrad(const kgan_tapconv_desc* d, const float* in, const float* gout, const int32_t* pmap,
                                  float* dw, int64_t dw_numel, int accumulate, void* stream) {
    if (int e = validate(d)) return e;
    KGAN_REQUIRE(in && gout && pmap && dw && dw_numel > 0, "tapconv_wgrad: null pointer");
    KGAN_REQUIRE(d->p_out_plane == 0, "tapconv_wgrad: position-block groups are a forward / data-gradient feature");
    cudaStream_t s = (cudaStream_t)stream;
    if (wgrad_is_thin(*d)) return launch_wgrad_thin(*d, in, gout, pmap, dw, dw_numel, accumulate, s);
    if (!accumulate && cudaMemsetAsync(dw, 0, sizeof(float) * dw_numel, s) != cudaSuccess) return check_launch("tapconv_wgrad memset");
    const int64_t total = (int64_t)d->n * d->p_out;
    const int tiles = d->ntap * ceil_div(d->ck, WB) * ceil_div(d->co, WB) * d->groups;
    int64_t nchunks = ceil_div64(4 * kNumSMs, tiles);
    const int64_t max_chunks = ceil_div64(total, 8 * WK);
    if (nchunks > max_chunks) nchunks = max_chunks;
    if (nchunks < 1) nchunks = 1;
    int64_t chunk = ceil_div64(ceil_div64(total, nchunks), WK) * WK;
    nchunks = ceil_div64(total, chunk);
    KGAN_REQUIRE((int64_t)d->groups * nchunks <= 65535, "tapconv_wgrad: grid too large");
    dim3 grid(d->ntap * ceil_div(d->ck, WB), ceil_div(d->co, WB), (unsigned)(d->groups * nchunks));
    tapconv_wgrad_simt<<<grid, NT, 0, s>>>(*d, in, gout, pmap, dw, (int)nchunks, chunk);
    return check_launch("tapconv_wgrad");
}

// Groups that read the same input channels and write adjacent output channel blocks (the K partitions of a graph conv's data
// gradient: g_in == 0, g_out == co) are ONE GEMM over groups * co output channels whose weight rows are addressed block-wise
// (w_oc_blk): the activation tile is fetched once instead of once per group.  Applied to every descriptor entering the tensor-core
// forward path (workspace size, packing, launch), so the packed image and the kernel always agree.
static kgan_tapconv_desc merge_groups(const kgan_tapconv_desc& d) {
    if (d.groups <= 1 || d.g_in != 0 || d.g_out != d.co || d.w_oc_blk != 0 || d.g_pout != 0) return d;
    kgan_tapconv_desc m = d;
    m.w_oc_blk = d.co;
    m.w_ocblk = d.g_w;
    m.co = d.groups * d.co;
    m.groups = 1;
    m.g_out = 0;
    m.g_w = 0;
    return m;
}

extern "C" int64_t kgan_tapconv_tf32_workspace(const kgan_tapconv_desc* d) {
    if (validate(d)) return -1;
    if (tapconv_is_thin(*d)) return 0;         // small contractions run on the exact streaming kernel in both precision modes
    const kgan_tapconv_desc m = merge_groups(*d);
    if (m.precision == KGAN_PREC_TF32X3 && !(m.tma_mode != 0 && tapconv_tma_eligible(m))) return 0;      // x3: the TMA-fed kernel only
    return tapconv_tf32_packed_numel(m);
}

extern "C" int kgan_tapconv_pack_tf32(const kgan_tapconv_desc* d, const float* w, float* wp, void* stream) {
    if (int e = validate(d)) return e;
    KGAN_REQUIRE(w && wp, "tapconv_pack_tf32: null pointer");
    return tapconv_pack_tf32(merge_groups(*d), w, wp, (cudaStream_t)stream);
}

extern "C" int64_t kgan_tapconv_pack_item_bytes(void) { return tapconv_pack_item_bytes(); }

extern "C" int kgan_tapconv_pack_tf32_batched(int count, const kgan_tapconv_desc* descs, const float* const* w, float* const* wp, void* items,
                                              int upload, void* stream) {
    KGAN_REQUIRE(count > 0 && count <= 65535 && descs && w && wp && items, "tapconv_pack_tf32_batched: bad argument");
    if (!upload) return tapconv_pack_tf32_batched(count, descs, w, wp, items, 0, (cudaStream_t)stream);
    kgan_tapconv_desc* merged = new kgan_tapconv_desc[count];
    for (int i = 0; i < count; ++i) merged[i] = merge_groups(descs[i]);
    const int rc = tapconv_pack_tf32_batched(count, merged, w, wp, items, 1, (cudaStream_t)stream);
    delete[] merged;
    return rc;
}

extern "C" int kgan_tapconv_fwd_tf32(const kgan_tapconv_desc* d, const float* in, const float* wp, const int32_t* pmap,
                                     const float* bias, const float* add, float* out, void* stream) {
    if (int e = validate(d)) return e;
    KGAN_REQUIRE(in && wp && pmap && out, "tapconv_fwd_tf32: null pointer");
    const kgan_tapconv_desc m = merge_groups(*d);
    if (m.precision == KGAN_PREC_TF32X3) {
        const int rt = m.tma_mode != 0 ? tapconv_fwd_tma(m, in, wp, pmap, bias, add, out, (cudaStream_t)stream) : -1;
        if (rt == -1) {
            set_error("tapconv_fwd_tf32: shape not eligible for the 3xTF32 path (kgan_tapconv_tf32_workspace() == 0)");
            return 1;
        }
        return rt;
    }
    if (m.prefer_staged) {
        const int rb = tapconv_fwd_build(m, in, wp, pmap, bias, add, out, (cudaStream_t)stream);
        if (rb != -1) return rb;
    }
    if (m.tma_mode != 0) {
        const int rt = tapconv_fwd_tma(m, in, wp, pmap, bias, add, out, (cudaStream_t)stream);
        if (rt != -1) return rt;
    }
    if (!m.prefer_staged) {
        const int rb = tapconv_fwd_build(m, in, wp, pmap, bias, add, out, (cudaStream_t)stream);
        if (rb != -1) return rb;
    }
    int r = tapconv_fwd_tf32(m, in, wp, pmap, bias, add, out, (cudaStream_t)stream);
    if (r == -1) {
        set_error("tapconv_fwd_tf32: shape not eligible for the tensor-core path (kgan_tapconv_tf32_workspace() == 0)");
        return 1;
    }
    return r;
}

extern "C" int kgan_tapconv_noise_ok(const kgan_tapconv_desc* d) {
    if (validate(d) || tapconv_is_thin(*d) || d->precision != KGAN_PREC_TF32 || d->groups != 1 || d->add_period != 0 || d->p_out_plane != 0) return 0;
    if (d->tma_mode != 0 && tapconv_tma_eligible(*d)) return 0;      // the TMA-fed kernel + a separate noise pass is the faster pair there
    return tapconv_tf32_packed_numel(*d) > 0 && tapconv_build_eligible(*d);
}

extern "C" int kgan_tapconv_fwd_tf32_noise(const kgan_tapconv_desc* d, const float* in, const float* wp, const int32_t* pmap, const float* bias,
                                           const float* add, const float* noise, const float* nw, float* out, void* stream) {
    if (int e = validate(d)) return e;
    KGAN_REQUIRE(in && wp && pmap && noise && nw && out, "tapconv_fwd_tf32_noise: null pointer");
    KGAN_REQUIRE(d->groups == 1, "tapconv_fwd_tf32_noise: groups");
    const int r = tapconv_fwd_build_noise(*d, in, wp, pmap, bias, add, noise, nw, out, (cudaStream_t)stream);
    if (r == -1) {
        set_error("tapconv_fwd_tf32_noise: not eligible (kgan_tapconv_noise_ok() == 0)");
        return 1;
    }
    return r;
}

extern "C" int kgan_gcn_fused_ok(const kgan_tapconv_desc* d) {
    if (validate(d) || d->mix_v <= 0) return 0;
    return tapconv_tf32_packed_numel(*d) > 0 && tapconv_build_eligible(*d);
}

extern "C" int kgan_gcn_fwd_tf32(const kgan_tapconv_desc* d, const float* x, const float* wp, const float* A, const float* bias, const float* add,
                                 const int32_t* omap, float* out, void* stream) {
    if (int e = validate(d)) return e;
    KGAN_REQUIRE(x && wp && A && out, "gcn_fwd_tf32: null pointer");
    const int r = gcn_fwd_fused(*d, x, wp, A, bias, add, out, omap, (cudaStream_t)stream);
    if (r == -1) {
        set_error("gcn_fwd_tf32: not eligible (kgan_gcn_fused_ok() == 0)");
        return 1;
    }
    return r;
}

extern "C" int kgan_tapconv_scatter_ok(const kgan_tapconv_desc* d) {
    if (validate(d) || tapconv_is_thin(*d) || tapconv_tf32_packed_numel(*d) <= 0) return 0;
    return tapconv_tma_scatter_eligible(*d);
}

extern "C" int kgan_tapconv_fwd_tf32_scatter(const kgan_tapconv_desc* d, const float* in, const float* wp, const int32_t* omap, const float* bias,
                                             float* out, void* stream) {
    if (int e = validate(d)) return e;
    KGAN_REQUIRE(in && wp && omap && out, "tapconv_fwd_tf32_scatter: null pointer");
    const int r = tapconv_fwd_tma_scatter(*d, in, wp, omap, bias, out, (cudaStream_t)stream);
    if (r == -1) {
        set_error("tapconv_fwd_tf32_scatter: not eligible (kgan_tapconv_scatter_ok() == 0)");
        return 1;
    }
    return r;
}

extern "C" int kgan_tapconv_res_ok(const kgan_tapconv_desc* d, const kgan_tapconv_desc* d2) {
    if (validate(d) || validate(d2) || tapconv_is_thin(*d) || tapconv_is_thin(*d2)) return 0;
    if (tapconv_tf32_packed_numel(*d) <= 0 || tapconv_tf32_packed_numel(*d2) <= 0) return 0;
    return tapconv_tma_res_eligible(*d, *d2);
}

extern "C" int kgan_tapconv_fwd_tf32_res(const kgan_tapconv_desc* d, const float* in, const float* wp, const kgan_tapconv_desc* d2, const float* in2,
                                         const float* wp2, const float* bias, const float* bias2, float* out, void* stream) {
    if (int e = validate(d)) return e;
    if (int e = validate(d2)) return e;
    KGAN_REQUIRE(in && wp && in2 && wp2 && out, "tapconv_fwd_tf32_res: null pointer");
    const int r = tapconv_fwd_tma_res(*d, *d2, in, wp, in2, wp2, bias, bias2, out, (cudaStream_t)stream);
    if (r == -1) {
        set_error("tapconv_fwd_tf32_res: not eligible (kgan_tapconv_res_ok() == 0): run the two convolutions separately");
        return 1;
    }
    return r;
}

extern "C" int kgan_tapconv_tma_ok(const kgan_tapconv_desc* d) {
    if (validate(d) || tapconv_is_thin(*d)) return 0;
    const kgan_tapconv_desc m = merge_groups(*d);
    return tapconv_tf32_packed_numel(m) > 0 && tapconv_tma_eligible(m);
}

extern "C" int kgan_tapconv_staged_ok(const kgan_tapconv_desc* d) {
    if (validate(d) || tapconv_is_thin(*d)) return 0;
    const kgan_tapconv_desc m = merge_groups(*d);
    if (tapconv_tf32_packed_numel(m) <= 0 || !tapconv_build_eligible(m)) return 0;
    return m.prefer_staged || !(m.tma_mode != 0 && tapconv_tma_eligible(m));
}

extern "C" int kgan_tapconv_wgrad_tma_ok(const kgan_tapconv_desc* d) {
    if (validate(d)) return 0;
    return tapconv_wgrad_tma_eligible(*d);
}

extern "C" int kgan_tapconv_wgrad_tf32_ok(const kgan_tapconv_desc* d) {
    if (validate(d)) return 0;
    if (wgrad_is_thin(*d)) return 1;
    return tapconv_wgrad_tf32_eligible(*d);
}

extern "C" int kgan_tapconv_wgrad_tf32(const kgan_tapconv_desc* d, const float* in, const float* gout, const int32_t* pmap,
                                       float* dw, int64_t dw_numel, int accumulate, void* stream) {
    if (int e = validate(d)) return e;
    KGAN_REQUIRE(in && gout && pmap && dw && dw_numel > 0, "tapconv_wgrad_tf32: null pointer");
    KGAN_REQUIRE(d->p_out_plane == 0, "tapconv_wgrad_tf32: position-block groups are a forward / data-gradient feature");
    if (wgrad_is_thin(*d)) return launch_wgrad_thin(*d, in, gout, pmap, dw, dw_numel, accumulate, (cudaStream_t)stream);      // exact, and far faster than M = 128 tiles
    int r = tapconv_wgrad_tf32(*d, in, gout, pmap, dw, dw_numel, accumulate, (cudaStream_t)stream);
    if (r == -1) {
        set_error("tapconv_wgrad_tf32: shape not eligible for the tensor-core path (kgan_tapconv_wgrad_tf32_ok() == 0)");
        return 1;
    }
    return r;
}
