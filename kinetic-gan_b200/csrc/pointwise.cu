// Memory-bound kernels of the ST-GCN hot path (the adjacency product lives in adjmix.cu): fused epilogues, plane gathers,
// label planes, BatchNorm, Adam.  All are HBM-bound streaming kernels: lanes run along the contiguous
// (t, v) plane axis, grids are sized in multiples of the SM count.  See include/kgan.h for semantics.
#include "common.cuh"

namespace kgan {

constexpr int PT = 256;

static inline int grid_for(int64_t work, int per_block = PT, int max_waves = 16) {
    int64_t b = ceil_div64(work, per_block);
    int64_t cap = (int64_t)kNumSMs * max_waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum, result valid in every thread; `red` holds >= 32 floats
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : 0.f;
    r = warp_sum(r);
    return r;
}

// ------------------------------------------------------------------------------------------------
// pointwise epilogues
// ------------------------------------------------------------------------------------------------
// one warp per (n, c) plane: per-plane constants, 16-byte accesses when the plane allows, no per-element index arithmetic
// `bn` (kgan_bn_epilogue_fwd): a is normalised first - a' = (a - mean[c]) * rstd[c] * gamma[c] + beta[c], training-mode BatchNorm with the
// batch statistics of kgan_bn_stats - so that BatchNorm, residual add, noise injection and activation of a generator block
// (generator.py:160,176-182) are ONE pass over the tensor
struct BnArgs {
    const float *mean, *rstd, *gamma, *beta;
};
__global__ void __launch_bounds__(PT) epilogue_fwd_k(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ bias,
                                                      const float* __restrict__ nw, const float* __restrict__ noise, float* __restrict__ out,
                                                      int n, int c, int p, int act, int vec, int rnd, BnArgs bn = BnArgs{nullptr, nullptr, nullptr, nullptr}) {
    const int lane = threadIdx.x & 31;
    const int64_t planes = (int64_t)n * c, tw = (int64_t)gridDim.x * (PT / 32);
    for (int64_t pl = (int64_t)blockIdx.x * (PT / 32) + (threadIdx.x >> 5); pl < planes; pl += tw) {
        const int64_t nn = pl / c;
        const int cc = (int)(pl - nn * c);
        float bs = bias ? __ldg(bias + cc) : 0.f;
        const float wn = nw ? __ldg(nw + cc) : 0.f;
        float sc = 1.f, mu = 0.f;
        if (bn.mean) {                           // same expression as kgan_bn_apply: (a - mean) * (rstd * gamma) + beta
            mu = __ldg(bn.mean + cc);
            sc = __ldg(bn.rstd + cc) * __ldg(bn.gamma + cc);
            bs += __ldg(bn.beta + cc);
        }
        const int64_t o = pl * p, on = nn * p;
        if (vec) {
            const float4* ap = reinterpret_cast<const float4*>(a + o);
            const float4* bp = reinterpret_cast<const float4*>(b + o);
            const float4* zp = reinterpret_cast<const float4*>(noise + on);
            float4* op = reinterpret_cast<float4*>(out + o);
            for (int i = lane; i < (p >> 2); i += 32) {
                float4 v = ap[i];
                v.x = fmaf(v.x - mu, sc, bs), v.y = fmaf(v.y - mu, sc, bs), v.z = fmaf(v.z - mu, sc, bs), v.w = fmaf(v.w - mu, sc, bs);
                if (b) {
                    const float4 t = bp[i];
                    v.x += t.x, v.y += t.y, v.z += t.z, v.w += t.w;
                }
                if (nw) {
                    const float4 z = __ldg(zp + i);
                    v.x = fmaf(wn, z.x, v.x), v.y = fmaf(wn, z.y, v.y), v.z = fmaf(wn, z.z, v.z), v.w = fmaf(wn, z.w, v.w);
                }
                op[i] = make_float4(tf32_out(apply_act(v.x, act), rnd), tf32_out(apply_act(v.y, act), rnd), tf32_out(apply_act(v.z, act), rnd),
                                    tf32_out(apply_act(v.w, act), rnd));
            }
        } else {
            for (int i = lane; i < p; i += 32) {
                float v = fmaf(a[o + i] - mu, sc, bs);
                if (b) v += b[o + i];
                if (nw) v = fmaf(wn, __ldg(noise + on + i), v);
                out[o + i] = tf32_out(apply_act(v, act), rnd);
            }
        }
    }
}

// planes below one warp's worth of positions (the generator's first blocks: 1, 4 or 20 positions per plane): one thread per
// element, consecutive threads on consecutive addresses; the warp-per-plane kernel above would leave 31 / 28 / 12 lanes idle
__global__ void __launch_bounds__(PT) epilogue_fwd_small_k(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ bias,
                                                            const float* __restrict__ nw, const float* __restrict__ noise, float* __restrict__ out,
                                                            unsigned total, unsigned c, unsigned p, int act, int rnd,
                                                            BnArgs bn = BnArgs{nullptr, nullptr, nullptr, nullptr}) {
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const unsigned pl = e / p, pp = e - pl * p;
        const unsigned nn = pl / c, cc = pl - nn * c;
        float v = a[e];
        if (bn.mean) v = fmaf(v - __ldg(bn.mean + cc), __ldg(bn.rstd + cc) * __ldg(bn.gamma + cc), __ldg(bn.beta + cc));
        if (b) v += b[e];
        if (bias) v += __ldg(bias + cc);
        if (nw) v = fmaf(__ldg(nw + cc), __ldg(noise + nn * p + pp), v);
        out[e] = tf32_out(apply_act(v, act), rnd);
    }
}

// 16-byte vector variant (numel % 4 == 0, 16-byte aligned pointers)
__global__ void __launch_bounds__(PT) act_bwd4_k(const float4* __restrict__ go, const float4* __restrict__ o, float4* __restrict__ gz, int64_t n4,
                                                  int act, int rnd) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 y = o[i], g = go[i];
        float4 r;
        if (act == KGAN_ACT_LRELU) {
            r.x = g.x * (y.x > 0.f ? 1.f : 0.2f);
            r.y = g.y * (y.y > 0.f ? 1.f : 0.2f);
            r.z = g.z * (y.z > 0.f ? 1.f : 0.2f);
            r.w = g.w * (y.w > 0.f ? 1.f : 0.2f);
        } else if (act == KGAN_ACT_TANH) {
            r.x = g.x * (1.f - y.x * y.x);
            r.y = g.y * (1.f - y.y * y.y);
            r.z = g.z * (1.f - y.z * y.z);
            r.w = g.w * (1.f - y.w * y.w);
        } else {
            r = g;
        }
        gz[i] = make_float4(tf32_out(r.x, rnd), tf32_out(r.y, rnd), tf32_out(r.z, rnd), tf32_out(r.w, rnd));
    }
}

__global__ void __launch_bounds__(PT) act_bwd_k(const float* __restrict__ go, const float* __restrict__ o, float* __restrict__ gz, int64_t numel,
                                                 int act, int rnd) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
        const float y = o[i];
        float m = 1.f;
        if (act == KGAN_ACT_LRELU) m = y > 0.f ? 1.f : 0.2f;
        else if (act == KGAN_ACT_TANH) m = 1.f - y * y;
        gz[i] = tf32_out(go[i] * m, rnd);
    }
}

// out[c] += sum over a slice of samples; grid (c, slices)
__global__ void __launch_bounds__(PT) chan_reduce_k(const float* __restrict__ g, const float* __restrict__ mul, float* __restrict__ out, int n,
                                                     int c, int p, int n_per_slice) {
    __shared__ float red[32];
    const int cc = blockIdx.x;
    const int nbeg = blockIdx.y * n_per_slice, nend = min(n, nbeg + n_per_slice);
    float acc = 0.f;
    const int64_t cnt = (int64_t)(nend - nbeg) * p;
    for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) {
        const int nn = nbeg + (int)(i / p), pp = (int)(i % p);
        float v = __ldg(g + ((int64_t)nn * c + cc) * p + pp);
        if (mul) v *= __ldg(mul + (int64_t)nn * p + pp);
        acc += v;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(out + cc, acc);
}

// 16-byte vector variant without multiplier (bias gradients): p % 4 == 0, 16-byte aligned g
__global__ void __launch_bounds__(PT) chan_reduce4_k(const float* __restrict__ g, float* __restrict__ out, int n, int c, int p4, int n_per_slice) {
    __shared__ float red[32];
    const int cc = blockIdx.x;
    const int nbeg = blockIdx.y * n_per_slice, nend = min(n, nbeg + n_per_slice);
    const unsigned cnt = (unsigned)(nend - nbeg) * (unsigned)p4;          // < 2^31 by the launcher's check
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    auto ld = [&](unsigned i) {
        const unsigned nn = i / (unsigned)p4, q = i - nn * (unsigned)p4;
        return __ldg(reinterpret_cast<const float4*>(g + ((int64_t)(nbeg + nn) * c + cc) * (int64_t)(4 * p4)) + q);
    };
    unsigned i = threadIdx.x;
    for (; i + 3 * blockDim.x < cnt; i += 4 * blockDim.x) {               // 4 independent 16-byte loads in flight per thread
        const float4 v0 = ld(i), v1 = ld(i + blockDim.x), v2 = ld(i + 2 * blockDim.x), v3 = ld(i + 3 * blockDim.x);
        a0 += (v0.x + v1.x) + (v2.x + v3.x);
        a1 += (v0.y + v1.y) + (v2.y + v3.y);
        a2 += (v0.z + v1.z) + (v2.z + v3.z);
        a3 += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; i < cnt; i += blockDim.x) {
        const float4 v = ld(i);
        a0 += v.x;
        a1 += v.y;
        a2 += v.z;
        a3 += v.w;
    }
    const float acc = block_sum((a0 + a1) + (a2 + a3), red);
    if (threadIdx.x == 0) atomicAdd(out + cc, acc);
}

// Row-blocked variant for small tables (J <= 4) and planes of at least 64 positions: a thread owns ONE output position, keeps
// its table entries in registers and walks over the (n, c) planes - no index arithmetic and no table loads per element.
template <int J>
__global__ void __launch_bounds__(PT) plane_spmm_rows_k(const float* __restrict__ x, const int32_t* __restrict__ idx, const float* __restrict__ wgt,
                                                         float* __restrict__ out, int64_t rows, int p_in, int p_out, int rnd) {
    const int q = blockIdx.y * blockDim.x + threadIdx.x;
    if (q >= p_out) return;
    int id[J];
    float w[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int s = __ldg(idx + q * J + j);
        id[j] = s >= 0 ? s : 0;
        w[j] = s >= 0 ? __ldg(wgt + q * J + j) : 0.f;
    }
    // U planes in flight per thread: with 4-byte accesses the kernel is bound by bytes in flight / loaded DRAM latency
    // (2048 threads x 4 loads x 4 B = 32 KB per SM gave ~3.3 TB/s), so single-entry tables keep 8 loads outstanding
    constexpr int U = J <= 2 ? 8 : 4;
    int64_t r = blockIdx.x;
    for (; r + (U - 1) * (int64_t)gridDim.x < rows; r += U * (int64_t)gridDim.x) {
        float acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float* xr = x + (r + u * (int64_t)gridDim.x) * p_in;
            acc[u] = 0.f;
#pragma unroll
            for (int j = 0; j < J; ++j) acc[u] = fmaf(w[j], __ldg(xr + id[j]), acc[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) out[(r + u * (int64_t)gridDim.x) * p_out + q] = tf32_out(acc[u], rnd);
    }
    for (; r < rows; r += gridDim.x) {
        const float* xr = x + r * p_in;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j) acc = fmaf(w[j], __ldg(xr + id[j]), acc);
        out[r * p_out + q] = tf32_out(acc, rnd);
    }
}

// Small planes (p_out <= 64) with small tables (J <= 4) - the one-joint / five-joint ends of both networks: a CTA works on
// PT / tpp planes at once (tpp = threads per plane, a power of two >= p_out), every thread owns one output position of one plane
// lane with its table entries in registers, 4 planes in flight.  Consecutive planes are contiguous, so a warp still touches
// consecutive addresses.  (The generic kernel below pays a 64-bit division and J table loads per element.)
template <int J>
__global__ void __launch_bounds__(PT) plane_spmm_small_k(const float* __restrict__ x, const int32_t* __restrict__ idx, const float* __restrict__ wgt,
                                                          float* __restrict__ out, int64_t rows, int p_in, int p_out, int tpp, int rnd) {
    const int q = threadIdx.x & (tpp - 1);
    if (q >= p_out) return;
    const int ppb = PT / tpp;
    int id[J];
    float w[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int s = __ldg(idx + q * J + j);
        id[j] = s >= 0 ? s : 0;
        w[j] = s >= 0 ? __ldg(wgt + q * J + j) : 0.f;
    }
    const int64_t stride = (int64_t)gridDim.x * ppb;
    int64_t r = (int64_t)blockIdx.x * ppb + threadIdx.x / tpp;
    for (; r + 3 * stride < rows; r += 4 * stride) {
        float acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float* xr = x + (r + u * stride) * p_in;
            acc[u] = 0.f;
#pragma unroll
            for (int j = 0; j < J; ++j) acc[u] = fmaf(w[j], __ldg(xr + id[j]), acc[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) out[(r + u * stride) * p_out + q] = tf32_out(acc[u], rnd);
    }
    for (; r < rows; r += stride) {
        const float* xr = x + r * p_in;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j) acc = fmaf(w[j], __ldg(xr + id[j]), acc);
        out[r * p_out + q] = tf32_out(acc, rnd);
    }
}

// generic tables (long lists: frame sums, pooling): one thread per output element, four independent partial sums
__global__ void __launch_bounds__(PT) plane_spmm_k(const float* __restrict__ x, const int32_t* __restrict__ idx, const float* __restrict__ wgt,
                                                    float* __restrict__ out, int64_t rows, int p_in, int p_out, int jn, int rnd) {
    const int64_t total = rows * p_out;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / p_out;
        const int q = (int)(i - r * p_out);
        const float* xr = x + r * p_in;
        const int32_t* iq = idx + q * jn;
        const float* wq = wgt + q * jn;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int j = 0;
        for (; j + 3 < jn; j += 4) {
            const int s0 = __ldg(iq + j), s1 = __ldg(iq + j + 1), s2 = __ldg(iq + j + 2), s3 = __ldg(iq + j + 3);
            const float v0 = s0 >= 0 ? __ldg(xr + s0) : 0.f, v1 = s1 >= 0 ? __ldg(xr + s1) : 0.f;
            const float v2 = s2 >= 0 ? __ldg(xr + s2) : 0.f, v3 = s3 >= 0 ? __ldg(xr + s3) : 0.f;
            a0 = fmaf(__ldg(wq + j), v0, a0);
            a1 = fmaf(__ldg(wq + j + 1), v1, a1);
            a2 = fmaf(__ldg(wq + j + 2), v2, a2);
            a3 = fmaf(__ldg(wq + j + 3), v3, a3);
        }
        for (; j < jn; ++j) {
            const int s0 = __ldg(iq + j);
            if (s0 >= 0) a0 = fmaf(__ldg(wq + j), __ldg(xr + s0), a0);
        }
        out[i] = tf32_out((a0 + a1) + (a2 + a3), rnd);
    }
}

__global__ void __launch_bounds__(PT) label_concat_k(const float* __restrict__ e, const float* __restrict__ x, float* __restrict__ out, int n,
                                                      int ncls, int c, int p, int rnd) {
    const int ct = ncls + c;
    const int64_t total = (int64_t)n * ct * p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int pp = (int)(i % p);
        const int64_t nc = i / p;
        const int cc = (int)(nc % ct);
        const int64_t nn = nc / ct;
        out[i] = tf32_out(cc < ncls ? __ldg(e + nn * ncls + cc) : __ldg(x + (nn * c + (cc - ncls)) * p + pp), rnd);
    }
}

// one warp per (n, channel) plane of the label part; remaining threads copy the data part
__global__ void __launch_bounds__(PT) label_split_k(const float* __restrict__ g, float* __restrict__ ge, float* __restrict__ gx, int n, int ncls,
                                                     int c, int p, int rnd) {
    const int ct = ncls + c;
    if (ge) {
        const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
        const int lane = threadIdx.x & 31;
        for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < (int64_t)n * ncls; r += warps) {
            const int64_t nn = r / ncls;
            const int cc = (int)(r % ncls);
            const float* gp = g + (nn * ct + cc) * p;
            float acc = 0.f;
            for (int i = lane; i < p; i += 32) acc += __ldg(gp + i);
            acc = warp_sum(acc);
            if (lane == 0) ge[r] = acc;
        }
    }
    if (gx) {
        const int64_t total = (int64_t)n * c * p;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            const int pp = (int)(i % p);
            const int64_t nc = i / p;
            const int cc = (int)(nc % c);
            const int64_t nn = nc / c;
            gx[i] = tf32_out(__ldg(g + (nn * ct + ncls + cc) * p + pp), rnd);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm2d (training).  Statistics: grid (sample chunks, channels) - every warp walks whole (n, c) planes of its channel with
// 16-byte loads, the CTA's partial sums go to a (chunk, channel) workspace and a small kernel adds the chunks IN ORDER (double):
// the statistics - and with them everything downstream of the generator - are bit-reproducible from run to run.  (Round 1 merged
// the CTAs with fp32 atomics: the summation order varied, and a 1e-7 wobble of a channel mean, amplified by |mean| / std and by
// the tf32 rounding of the normalised activations, made two identical generator passes differ by 3e-4.)
// Sums are taken around a per-channel shift (the channel's first element), so E[d^2] - E[d]^2 does not cancel.  (The first
// version used ONE CTA per channel with a 64-bit division per element: 3 CTAs for the generator's 3-channel layers.)
// Elementwise passes: one warp per (n, c) plane, per-plane constants, no per-element index arithmetic.
// ------------------------------------------------------------------------------------------------
constexpr int BNW = PT / 32;      // warps per CTA

template <bool BWD>
__global__ void __launch_bounds__(PT) bn_partial_k(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ mean,
                                                    const float* __restrict__ rstd, float* __restrict__ part, int n, int c,
                                                    int p, int vec) {
    __shared__ float red[32];
    const int cc = blockIdx.y, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * BNW + (threadIdx.x >> 5), tw = gridDim.x * BNW;
    const float sh = BWD ? __ldg(mean + cc) : __ldg(x + (int64_t)cc * p);
    const float rs = BWD ? __ldg(rstd + cc) : 1.f;
    float s1 = 0.f, s2 = 0.f;
    for (int nn = gw; nn < n; nn += tw) {
        const int64_t o = ((int64_t)nn * c + cc) * p;
        if (vec) {
            const float4* xp = reinterpret_cast<const float4*>(x + o);
            const float4* gp = reinterpret_cast<const float4*>(gy + o);
            for (int i = lane; i < (p >> 2); i += 32) {
                const float4 v = __ldg(xp + i);
                if (BWD) {
                    const float4 g = __ldg(gp + i);
                    s1 += (g.x + g.y) + (g.z + g.w);
                    s2 = fmaf(g.x, (v.x - sh) * rs, s2);
                    s2 = fmaf(g.y, (v.y - sh) * rs, s2);
                    s2 = fmaf(g.z, (v.z - sh) * rs, s2);
                    s2 = fmaf(g.w, (v.w - sh) * rs, s2);
                } else {
                    const float d0 = v.x - sh, d1 = v.y - sh, d2 = v.z - sh, d3 = v.w - sh;
                    s1 += (d0 + d1) + (d2 + d3);
                    s2 = fmaf(d0, d0, s2);
                    s2 = fmaf(d1, d1, s2);
                    s2 = fmaf(d2, d2, s2);
                    s2 = fmaf(d3, d3, s2);
                }
            }
        } else {
            for (int i = lane; i < p; i += 32) {
                const float v = __ldg(x + o + i);
                if (BWD) {
                    const float g = __ldg(gy + o + i);
                    s1 += g;
                    s2 = fmaf(g, (v - sh) * rs, s2);
                } else {
                    const float d0 = v - sh;
                    s1 += d0;
                    s2 = fmaf(d0, d0, s2);
                }
            }
        }
    }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) {              // part[chunk][channel][2]
        float* dst = part + ((int64_t)blockIdx.x * c + cc) * 2;
        dst[0] = s1;
        dst[1] = s2;
    }
}

// out1[c] = sum_chunks part[chunk][c][0], out2[c] likewise: fixed order, double accumulation
__global__ void bn_reduce_k(const float* __restrict__ part, int chunks, int c, float* __restrict__ out1, float* __restrict__ out2) {
    const int cc = blockIdx.x * blockDim.x + threadIdx.x;
    if (cc >= c) return;
    double a = 0.0, b = 0.0;
    for (int s = 0; s < chunks; ++s) {
        a += (double)part[((int64_t)s * c + cc) * 2];
        b += (double)part[((int64_t)s * c + cc) * 2 + 1];
    }
    out1[cc] = (float)a;
    out2[cc] = (float)b;
}

// part holds per-chunk sum(d), sum(d^2) (d = x - first element of the channel): added in chunk order in double
__global__ void bn_finalize_k(const float* __restrict__ x, const float* __restrict__ part, int chunks, float* __restrict__ mean,
                              float* __restrict__ rstd, float* __restrict__ rm, float* __restrict__ rv, int n, int c, int p, float eps, float mom) {
    const int cc = blockIdx.x * blockDim.x + threadIdx.x;
    if (cc >= c) return;
    double a = 0.0, b = 0.0;
    for (int s = 0; s < chunks; ++s) {
        a += (double)part[((int64_t)s * c + cc) * 2];
        b += (double)part[((int64_t)s * c + cc) * 2 + 1];
    }
    const double cntd = (double)((int64_t)n * p);
    const float cnt = (float)cntd;
    const float sh = __ldg(x + (int64_t)cc * p);
    const double md = a / cntd;
    const float m = (float)md;
    const float var = fmaxf((float)(b / cntd - md * md), 0.f);
    const float mu = sh + m;
    mean[cc] = mu;
    rstd[cc] = rsqrtf(var + eps);
    if (rm) rm[cc] = (1.f - mom) * rm[cc] + mom * mu;
    if (rv) rv[cc] = (1.f - mom) * rv[cc] + mom * var * (cnt / fmaxf(1.f, cnt - 1.f));
}

// MODE 0: y = (x - mean) * rstd * gamma + beta.   MODE 1: gx = gamma * rstd * (gy - mean(gy) - xhat * mean(gy * xhat)), with the two
// sums in s1 (= gbeta) and s2 (= ggamma).
template <int MODE>
__global__ void __launch_bounds__(PT) bn_elem_k(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ mean,
                                                 const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                 const float* __restrict__ s1, const float* __restrict__ s2, float* __restrict__ out, int n, int c, int p,
                                                 int vec, int rnd) {
    const int lane = threadIdx.x & 31;
    const int64_t planes = (int64_t)n * c, tw = (int64_t)gridDim.x * BNW;
    const float inv_cnt = 1.f / (float)((int64_t)n * p);
    for (int64_t pl = (int64_t)blockIdx.x * BNW + (threadIdx.x >> 5); pl < planes; pl += tw) {
        const int cc = (int)(pl % c);
        const float mu = __ldg(mean + cc), rs = __ldg(rstd + cc), sc = rs * __ldg(gamma + cc);
        const float b0 = MODE == 0 ? __ldg(beta + cc) : 0.f;
        const float m1 = MODE == 1 ? __ldg(s1 + cc) * inv_cnt : 0.f, m2 = MODE == 1 ? __ldg(s2 + cc) * inv_cnt : 0.f;
        const int64_t o = pl * p;
        auto f = [&](float xv, float gv) { return tf32_out(MODE == 0 ? fmaf(xv - mu, sc, b0) : sc * (gv - m1 - (xv - mu) * rs * m2), rnd); };
        if (vec) {
            const float4* xp = reinterpret_cast<const float4*>(x + o);
            const float4* gp = reinterpret_cast<const float4*>(gy + o);
            float4* op = reinterpret_cast<float4*>(out + o);
            for (int i = lane; i < (p >> 2); i += 32) {
                const float4 v = __ldg(xp + i);
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                if (MODE == 1) g = __ldg(gp + i);
                op[i] = make_float4(f(v.x, g.x), f(v.y, g.y), f(v.z, g.z), f(v.w, g.w));
            }
        } else {
            for (int i = lane; i < p; i += 32) out[o + i] = f(__ldg(x + o + i), MODE == 1 ? __ldg(gy + o + i) : 0.f);
        }
    }
}

static inline int bn_vec_ok(int p, const void* a, const void* b, const void* c2) {
    return (p & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c2)) & 15) == 0;
}
static inline dim3 bn_partial_grid(int n, int c) {
    int s = ceil_div(4 * kNumSMs, c);
    const int smax = ceil_div(n, BNW);
    if (s > smax) s = smax;
    if (s < 1) s = 1;
    return dim3((unsigned)s, (unsigned)c);
}

__global__ void __launch_bounds__(PT) adam_k(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                              int64_t numel, float lr_over_bc1, float b1, float b2, float eps, float inv_sqrt_bc2, float gs) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
        const float gr = g[i] * gs;
        const float mi = b1 * m[i] + (1.f - b1) * gr;
        const float vi = b2 * v[i] + (1.f - b2) * gr * gr;
        m[i] = mi;
        v[i] = vi;
        p[i] -= lr_over_bc1 * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    }
}

__global__ void __launch_bounds__(PT) interpolate_k(const float* __restrict__ alpha, const float* __restrict__ x, const float* __restrict__ y,
                                                     float* __restrict__ out, int n, int64_t per, int rnd) {
    const int64_t total = (int64_t)n * per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const float a = __ldg(alpha + i / per);
        out[i] = tf32_out(a * x[i] + (1.f - a) * y[i], rnd);
    }
}

// out[r, v] = sum_t x[r, t, v]: one warp per (T, V) plane.  Lane = (frame slot, joint): the first floor(32 / V) * V lanes read that many
// consecutive frames per step (contiguous addresses), every lane sums ONE joint in a register, four steps in flight; the frame slots are
// merged through shared memory at the end.  (The generic gather kernel did this at 1.2 TB/s: one thread per output walking T strided
// elements; a first version with shared-memory atomics per element reached 1.6 TB/s.)
__global__ void __launch_bounds__(PT) plane_sum_t_k(const float* __restrict__ x, float* __restrict__ out, int64_t rows, int t, int v, int rnd) {
    __shared__ float bins[PT / 32][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int fpw = 32 / v;                                   // frames per warp step
    const int sub = lane / v, col = lane - sub * v;
    const bool active = sub < fpw;
    const int p = t * v, stepe = fpw * v;
    const int64_t tw = (int64_t)gridDim.x * (PT / 32);
    for (int64_t r = (int64_t)blockIdx.x * (PT / 32) + wid; r < rows; r += tw) {
        const float* xr = x + r * p + lane;                   // element sub * v + col == lane for active lanes
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (active) {
            int e = lane;
            for (; e + 3 * stepe < p; e += 4 * stepe, xr += 4 * stepe) {
                a0 += __ldg(xr);
                a1 += __ldg(xr + stepe);
                a2 += __ldg(xr + 2 * stepe);
                a3 += __ldg(xr + 3 * stepe);
            }
            for (; e < p; e += stepe, xr += stepe) a0 += __ldg(xr);
        }
        bins[wid][lane] = (a0 + a1) + (a2 + a3);
        __syncwarp();
        if (lane < v) {
            float acc = 0.f;
            for (int s2 = 0; s2 < fpw; ++s2) acc += bins[wid][s2 * v + lane];
            out[r * v + lane] = tf32_out(acc, rnd);
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(PT) round_tf32_k(const float* __restrict__ x, float* __restrict__ out, int64_t numel) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) out[i] = tf32_out(x[i], 1);
}

}  // namespace kgan

using namespace kgan;

extern "C" int kgan_plane_sum_t(const float* x, float* out, int64_t rows, int t, int v, int out_tf32, void* stream) {
    KGAN_REQUIRE(x && out && rows > 0 && t > 0 && v > 0 && v <= 32, "plane_sum_t: bad argument (1 <= V <= 32)");
    plane_sum_t_k<<<grid_for(rows, PT / 32, 8), PT, 0, (cudaStream_t)stream>>>(x, out, rows, t, v, out_tf32);
    return check_launch("plane_sum_t");
}

extern "C" int kgan_round_tf32(const float* x, float* out, int64_t numel, void* stream) {
    KGAN_REQUIRE(x && out && numel > 0, "round_tf32: bad argument");
    round_tf32_k<<<grid_for(numel), PT, 0, (cudaStream_t)stream>>>(x, out, numel);
    return check_launch("round_tf32");
}

extern "C" int kgan_epilogue_fwd(const float* a, const float* b, const float* bias, const float* nw, const float* noise, float* out,
                                 int n, int c, int p, int act, int out_tf32, void* stream) {
    KGAN_REQUIRE(a && out && n > 0 && c > 0 && p > 0, "epilogue_fwd: bad argument");
    KGAN_REQUIRE((nw == nullptr) == (noise == nullptr), "epilogue_fwd: nw and noise go together");
    const int vec = (p & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(noise) |
                                      reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (p < 32 && (int64_t)n * c * p < (1ll << 31))
        epilogue_fwd_small_k<<<grid_for((int64_t)n * c * p), PT, 0, (cudaStream_t)stream>>>(a, b, bias, nw, noise, out, (unsigned)((int64_t)n * c * p),
                                                                                          (unsigned)c, (unsigned)p, act, out_tf32);
    else
        epilogue_fwd_k<<<grid_for((int64_t)n * c, PT / 32, 8), PT, 0, (cudaStream_t)stream>>>(a, b, bias, nw, noise, out, n, c, p, act, vec, out_tf32);
    return check_launch("epilogue_fwd");
}

extern "C" int kgan_bn_epilogue_fwd(const float* a, const float* mean, const float* rstd, const float* gamma, const float* beta, const float* b,
                                    const float* nw, const float* noise, float* out, int n, int c, int p, int act, int out_tf32, void* stream) {
    KGAN_REQUIRE(a && mean && rstd && gamma && beta && out && n > 0 && c > 0 && p > 0, "bn_epilogue_fwd: bad argument");
    KGAN_REQUIRE((nw == nullptr) == (noise == nullptr), "bn_epilogue_fwd: nw and noise go together");
    const BnArgs bn{mean, rstd, gamma, beta};
    const int vec = (p & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(noise) |
                                      reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (p < 32 && (int64_t)n * c * p < (1ll << 31))
        epilogue_fwd_small_k<<<grid_for((int64_t)n * c * p), PT, 0, (cudaStream_t)stream>>>(a, b, nullptr, nw, noise, out, (unsigned)((int64_t)n * c * p),
                                                                                          (unsigned)c, (unsigned)p, act, out_tf32, bn);
    else
        epilogue_fwd_k<<<grid_for((int64_t)n * c, PT / 32, 8), PT, 0, (cudaStream_t)stream>>>(a, b, nullptr, nw, noise, out, n, c, p, act, vec, out_tf32, bn);
    return check_launch("bn_epilogue_fwd");
}

extern "C" int kgan_act_bwd(const float* gout, const float* out, float* gz, int64_t numel, int act, int out_tf32, void* stream) {
    KGAN_REQUIRE(gout && out && gz && numel > 0, "act_bwd: bad argument");
    if ((numel & 3) == 0 && ((reinterpret_cast<uintptr_t>(gout) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gz)) & 15) == 0)
        act_bwd4_k<<<grid_for(numel / 4), PT, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(gout), reinterpret_cast<const float4*>(out),
                                                                        reinterpret_cast<float4*>(gz), numel / 4, act, out_tf32);
    else
        act_bwd_k<<<grid_for(numel), PT, 0, (cudaStream_t)stream>>>(gout, out, gz, numel, act, out_tf32);
    return check_launch("act_bwd");
}

extern "C" int kgan_chan_reduce(const float* g, const float* mul, float* out, int n, int c, int p, void* stream) {
    KGAN_REQUIRE(g && out && n > 0 && c > 0 && p > 0, "chan_reduce: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(out, 0, sizeof(float) * c, s) != cudaSuccess) return check_launch("chan_reduce memset");
    int slices = ceil_div(4 * kNumSMs, c);
    if (slices > n) slices = n;
    if (slices > 65535) slices = 65535;
    const int per = ceil_div(n, slices);
    slices = ceil_div(n, per);
    if (!mul && (p & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 && (int64_t)per * (p / 4) < (1ll << 31))
        chan_reduce4_k<<<dim3(c, slices), PT, 0, s>>>(g, out, n, c, p / 4, per);
    else
        chan_reduce_k<<<dim3(c, slices), PT, 0, s>>>(g, mul, out, n, c, p, per);
    return check_launch("chan_reduce");
}

extern "C" int kgan_plane_spmm(const float* x, const int32_t* idx, const float* wgt, float* out, int64_t rows, int p_in, int p_out, int j,
                               int out_tf32, void* stream) {
    KGAN_REQUIRE(x && idx && wgt && out && rows > 0 && p_in > 0 && p_out > 0 && j > 0, "plane_spmm: bad argument");
    if (j <= 4 && p_out >= 64 && rows >= 64) {
        const int chunks = ceil_div(p_out, PT);
        int64_t gx = (int64_t)kNumSMs * 8 / chunks;
        if (gx > rows) gx = rows;
        if (gx < 1) gx = 1;
        const dim3 grid((unsigned)gx, (unsigned)chunks);
        cudaStream_t s = (cudaStream_t)stream;
        if (j == 1) plane_spmm_rows_k<1><<<grid, PT, 0, s>>>(x, idx, wgt, out, rows, p_in, p_out, out_tf32);
        else if (j == 2) plane_spmm_rows_k<2><<<grid, PT, 0, s>>>(x, idx, wgt, out, rows, p_in, p_out, out_tf32);
        else if (j == 3) plane_spmm_rows_k<3><<<grid, PT, 0, s>>>(x, idx, wgt, out, rows, p_in, p_out, out_tf32);
        else plane_spmm_rows_k<4><<<grid, PT, 0, s>>>(x, idx, wgt, out, rows, p_in, p_out, out_tf32);
        return check_launch("plane_spmm");
    }
    if (j <= 4 && p_out <= 64) {
        int tpp = 1;
        while (tpp < p_out) tpp *= 2;
        const int ppb = PT / tpp;
        int64_t gx = ceil_div64(rows, ppb);
        if (gx > (int64_t)kNumSMs * 8) gx = (int64_t)kNumSMs * 8;
        cudaStream_t s = (cudaStream_t)stream;
        if (j == 1) plane_spmm_small_k<1><<<(unsigned)gx, PT, 0, s>>>(x, idx, wgt, out, rows, p_in, p_out, tpp, out_tf32);
        else if (j == 2) plane_spmm_small_k<2><<<(unsigned)gx, PT, 0, s>>>(x, idx, wgt, out, rows, p_in, p_out, tpp, out_tf32);
        else if (j == 3) plane_spmm_small_k<3><<<(unsigned)gx, PT, 0, s>>>(x, idx, wgt, out, rows, p_in, p_out, tpp, out_tf32);
        else plane_spmm_small_k<4><<<(unsigned)gx, PT, 0, s>>>(x, idx, wgt, out, rows, p_in, p_out, tpp, out_tf32);
        return check_launch("plane_spmm");
    }
    plane_spmm_k<<<grid_for(rows * p_out), PT, 0, (cudaStream_t)stream>>>(x, idx, wgt, out, rows, p_in, p_out, j, out_tf32);
    return check_launch("plane_spmm");
}

extern "C" int kgan_label_concat(const float* e, const float* x, float* out, int n, int n_cls, int c, int p, int out_tf32, void* stream) {
    KGAN_REQUIRE(e && x && out && n > 0 && n_cls > 0 && c > 0 && p > 0, "label_concat: bad argument");
    label_concat_k<<<grid_for((int64_t)n * (n_cls + c) * p), PT, 0, (cudaStream_t)stream>>>(e, x, out, n, n_cls, c, p, out_tf32);
    return check_launch("label_concat");
}

extern "C" int kgan_label_split(const float* g, float* ge, float* gx, int n, int n_cls, int c, int p, int out_tf32, void* stream) {
    KGAN_REQUIRE(g && (ge || gx) && n > 0 && n_cls > 0 && c > 0 && p > 0, "label_split: bad argument");
    const int64_t work = (int64_t)n * (n_cls * 32 > c * p ? n_cls * 32 : c * p);
    label_split_k<<<grid_for(work), PT, 0, (cudaStream_t)stream>>>(g, ge, gx, n, n_cls, c, p, out_tf32);
    return check_launch("label_split");
}

extern "C" int64_t kgan_bn_workspace(int n, int c) {
    if (n <= 0 || c <= 0 || c > 65535) return 0;
    return (int64_t)bn_partial_grid(n, c).x * c * 2;
}

extern "C" int kgan_bn_stats(const float* x, float* mean, float* rstd, float* running_mean, float* running_var, int n, int c, int p, float eps,
                             float momentum, float* workspace, void* stream) {
    KGAN_REQUIRE(x && mean && rstd && workspace && n > 0 && c > 0 && p > 0, "bn_stats: bad argument");
    KGAN_REQUIRE(c <= 65535, "bn_stats: too many channels");
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 grid = bn_partial_grid(n, c);
    bn_partial_k<false><<<grid, PT, 0, s>>>(x, nullptr, nullptr, nullptr, workspace, n, c, p, bn_vec_ok(p, x, nullptr, nullptr));
    bn_finalize_k<<<ceil_div(c, 128), 128, 0, s>>>(x, workspace, (int)grid.x, mean, rstd, running_mean, running_var, n, c, p, eps, momentum);
    return check_launch("bn_stats");
}

extern "C" int kgan_bn_apply(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, float* y, int n, int c,
                             int p, int out_tf32, void* stream) {
    KGAN_REQUIRE(x && mean && rstd && gamma && beta && y && n > 0 && c > 0 && p > 0, "bn_apply: bad argument");
    bn_elem_k<0><<<grid_for((int64_t)n * c, BNW, 8), PT, 0, (cudaStream_t)stream>>>(x, nullptr, mean, rstd, gamma, beta, nullptr, nullptr, y, n, c, p,
                                                                                  bn_vec_ok(p, x, y, nullptr), out_tf32);
    return check_launch("bn_apply");
}

extern "C" int kgan_bn_bwd(const float* gy, const float* x, const float* mean, const float* rstd, const float* gamma, float* gx, float* ggamma,
                           float* gbeta, int n, int c, int p, int out_tf32, float* workspace, void* stream) {
    KGAN_REQUIRE(gy && x && mean && rstd && gamma && gx && ggamma && gbeta && workspace && n > 0 && c > 0 && p > 0, "bn_bwd: bad argument");
    KGAN_REQUIRE(c <= 65535, "bn_bwd: too many channels");
    cudaStream_t s = (cudaStream_t)stream;
    const int vec = bn_vec_ok(p, x, gy, gx);
    const dim3 grid = bn_partial_grid(n, c);
    bn_partial_k<true><<<grid, PT, 0, s>>>(x, gy, mean, rstd, workspace, n, c, p, vec);
    bn_reduce_k<<<ceil_div(c, 128), 128, 0, s>>>(workspace, (int)grid.x, c, gbeta, ggamma);
    bn_elem_k<1><<<grid_for((int64_t)n * c, BNW, 8), PT, 0, s>>>(x, gy, mean, rstd, gamma, nullptr, gbeta, ggamma, gx, n, c, p, vec, out_tf32);
    return check_launch("bn_bwd");
}

extern "C" int kgan_adam_step(float* p, const float* g, float* m, float* v, int64_t numel, float lr, float b1, float b2, float eps, int step,
                              float grad_scale, void* stream) {
    KGAN_REQUIRE(p && g && m && v && numel > 0 && step >= 1, "adam_step: bad argument");
    const double bc1 = 1.0 - pow((double)b1, step), bc2 = 1.0 - pow((double)b2, step);
    adam_k<<<grid_for(numel), PT, 0, (cudaStream_t)stream>>>(p, g, m, v, numel, (float)(lr / bc1), b1, b2, eps, (float)(1.0 / sqrt(bc2)),
                                                           grad_scale);
    return check_launch("adam_step");
}

extern "C" int kgan_interpolate(const float* alpha, const float* x, const float* y, float* out, int n, int64_t per_sample, int out_tf32,
                                void* stream) {
    KGAN_REQUIRE(alpha && x && y && out && n > 0 && per_sample > 0, "interpolate: bad argument");
    interpolate_k<<<grid_for((int64_t)n * per_sample), PT, 0, (cudaStream_t)stream>>>(alpha, x, y, out, n, per_sample, out_tf32);
    return check_launch("interpolate");
}
