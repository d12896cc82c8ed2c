// Adjacency product of the graph convolution and its two adjoints (see include/kgan.h):
//   fwd    out[n, k*C + c, t, w] = sum_v x[n, c, t, v] * A[k, v, w]
//   bwd_x  gx[n, c, t, v]        = sum_k sum_w g[n, k*C + c, t, w] * A[k, v, w]
//   bwd_a  gA[k, v, w]           = sum_{n,c,t} x[n, c, t, v] * g[n, k*C + c, t, w]
// HBM-bound streaming kernels.  A "row" is one (n, c, t): V contiguous inputs, W contiguous outputs per partition k;
// rows q = c*T + t of one sample are contiguous in x and, for every k, in out/g.  A CTA therefore handles a block of RB
// rows of one sample: it stages the block with straight coalesced (16-byte when aligned) copies into shared memory,
// computes with one thread per row out of shared memory, and writes each partition back as one contiguous copy.
// All index arithmetic is per block, not per element.
// A_eff = A (.) edge_importance keeps the skeleton's sparsity (73 of 1875 entries at 25 joints): the forward and dx
// kernels compact the non-zeros of A in shared memory and touch only those (exact).
#include "common.cuh"

namespace kgan {

constexpr int AT = 256;        // threads per CTA
constexpr int RB = 128;        // rows per block

// contiguous global -> shared copy (and back); 16-byte vectors when both ends allow it
__device__ __forceinline__ void copy_in(float* dst, const float* __restrict__ src, int count) {
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0 && (count & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = threadIdx.x; i < (count >> 2); i += blockDim.x) d4[i] = __ldg(s4 + i);
    } else {
        for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = __ldg(src + i);
    }
}
__device__ __forceinline__ void copy_out(float* __restrict__ dst, const float* src, int count) {
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0 && (count & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = threadIdx.x; i < (count >> 2); i += blockDim.x) d4[i] = s4[i];
    } else {
        for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = src[i];
    }
}

// shared-memory carve-up shared by the three kernels (all segments 16-byte aligned)
struct Carve {
    float* As;     // [k][v][w]
    int* nz;       // compacted non-zero lists
    int* nzc;      // list lengths
    float* xs;     // [RB][v]
    float* gs;     // [k][RB][w]
};
__host__ __device__ inline int align4(int n) { return (n + 3) & ~3; }
__host__ __device__ inline size_t carve_floats(int k, int v, int w, int lists, int list_len) {
    return (size_t)align4(k * v * w) + align4(lists * list_len) + align4(lists) + align4(RB * v) + (size_t)k * align4(RB * w);
}
__device__ __forceinline__ Carve carve(float* sm, int k, int v, int w, int lists, int list_len) {
    Carve c;
    c.As = sm;
    c.nz = reinterpret_cast<int*>(sm + align4(k * v * w));
    c.nzc = c.nz + align4(lists * list_len);
    c.xs = reinterpret_cast<float*>(c.nzc + align4(lists));
    c.gs = c.xs + align4(RB * v);
    return c;
}

// Forward and dx share one kernel.  Both are a small sparse mix along the joint axis of every row q = (c, t):
//     out[n, ko, q, wo] = sum_j coef_j * in[n, kin_j, q, vi_j]
//   fwd (MODE 0): in = x (one block per sample), out blocks ko = k:   entries of (k, w) = { (0, v, A[k,v,w]) : A[k,v,w] != 0 }
//   dx  (MODE 1): in = g (K blocks per sample), one out block:        entries of v      = { (k, w, A[k,v,w]) : A[k,v,w] != 0 }
// Every CTA first compacts A into per-output entry lists in shared memory (exact: zero entries contribute nothing), then a
// thread produces 4 CONSECUTIVE outputs of one (n, ko) block - one 16-byte store, one integer division per 4 outputs - and
// reads its inputs straight from global memory (a row of V <= 32 floats is shared by the W threads next to it: L1 hits).
// Compared with the previous staged version (one thread per row looping over shared memory: 78 % SM-busy at 6 % of HBM
// peak) this removes the per-element index arithmetic and the three block-wide barriers per 128 rows.
struct MixEntry {
    int off;       // element offset of the input relative to the sample's first block at row q = 0
    float coef;
};

template <int MODE>
__global__ void __launch_bounds__(AT) adjmix_rowmix_k(const float* __restrict__ in, const float* __restrict__ A, float* __restrict__ out, int n,
                                                       int ct, int v, int w, int k, int chunks_per_block, int64_t total_items) {
    extern __shared__ __align__(16) float sm[];
    const int ko = MODE == 0 ? k : 1, wo = MODE == 0 ? w : v;           // output blocks per sample / outputs per row
    const int vi = MODE == 0 ? v : w, ki = MODE == 0 ? 1 : k;           // inputs per row / input blocks per sample
    const int maxj = MODE == 0 ? v : k * w;
    const int nlists = ko * wo;
    int* cnt = reinterpret_cast<int*>(sm);
    MixEntry* ent = reinterpret_cast<MixEntry*>(sm + ((nlists + 3) & ~3));
    for (int o = threadIdx.x; o < nlists; o += blockDim.x) {
        int c = 0;
        if (MODE == 0) {
            const int kk = o / w, ww = o - kk * w;
            for (int vv = 0; vv < v; ++vv) {
                const float a = __ldg(A + (kk * v + vv) * w + ww);
                if (a != 0.f) ent[o * maxj + c++] = MixEntry{vv, a};
            }
        } else {
            for (int kk = 0; kk < k; ++kk)
                for (int ww = 0; ww < w; ++ww) {
                    const float a = __ldg(A + (kk * v + o) * w + ww);
                    if (a != 0.f) ent[o * maxj + c++] = MixEntry{kk * ct * w + ww, a};
                }
        }
        cnt[o] = c;
    }
    __syncthreads();
    const unsigned plane = (unsigned)ct * (unsigned)wo;                  // outputs of one (n, ko) block, contiguous
    const bool vec = (plane & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    for (int64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int chunk = (int)(item % chunks_per_block);
        const int64_t nb = item / chunks_per_block;                      // n * ko + kb
        const int kb = (int)(nb % ko);
        const int64_t nn = nb / ko;
        const unsigned e0 = ((unsigned)chunk * AT + threadIdx.x) * 4u;
        if (e0 >= plane) continue;
        const float* in_n = in + nn * (int64_t)ki * ct * vi;
        unsigned q = e0 / (unsigned)wo, ww = e0 - q * (unsigned)wo;
        float r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float acc = 0.f;
            if (e0 + u < plane) {
                const int o = kb * wo + (int)ww;
                const MixEntry* el = ent + o * maxj;
                const float* xq = in_n + (int64_t)q * vi;
                const int c = cnt[o];
                for (int j = 0; j < c; ++j) acc = fmaf(el[j].coef, __ldg(xq + el[j].off), acc);
            }
            r[u] = acc;
            if (++ww == (unsigned)wo) {
                ww = 0;
                ++q;
            }
        }
        float* op = out + nb * (int64_t)plane + e0;
        if (vec) {
            *reinterpret_cast<float4*>(op) = make_float4(r[0], r[1], r[2], r[3]);
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (e0 + u < plane) op[u] = r[u];
        }
    }
}

// gA: a (k*w) x v GEMM with a very long contraction (all rows).  Each CTA walks over many row blocks; every thread owns
// one 4x4 (v, w) register tile of one partition k and a strided subset of the staged rows.  The CTA's partial result is
// first merged in shared memory, then added to gA with one atomic per output (k*v*w can be as few as 3 addresses).
__global__ void __launch_bounds__(AT) adjmix_bwd_a_k(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ gA,
                                                      int ct, int v, int w, int k, int blocks_per_sample, int total_blocks) {
    extern __shared__ __align__(16) float sm[];
    const Carve s = carve(sm, k, v, w, 0, 0);
    const int vt = (v + 3) >> 2, wt = (w + 3) >> 2, ntile = k * vt * wt;
    const int rgroups = max(1, AT / ntile);
    const int tile = threadIdx.x % ntile, grp = threadIdx.x / ntile;
    const bool active = grp < rgroups;
    const int kk = tile / (vt * wt), v0 = ((tile / wt) % vt) * 4, w0 = (tile % wt) * 4;
    const int gstride = align4(RB * w);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int blk = blockIdx.x; blk < total_blocks; blk += gridDim.x) {
        const int nn = blk / blocks_per_sample, q0 = (blk % blocks_per_sample) * RB;
        const int rows = min(RB, ct - q0);
        copy_in(s.xs, x + ((int64_t)nn * ct + q0) * v, rows * v);
        for (int k2 = 0; k2 < k; ++k2) copy_in(s.gs + k2 * gstride, g + (((int64_t)nn * k + k2) * ct + q0) * w, rows * w);
        __syncthreads();
        if (active) {
            const float* gk = s.gs + kk * gstride;
            for (int rr = grp; rr < rows; rr += rgroups) {
                float xa[4], ga[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) xa[i] = (v0 + i < v) ? s.xs[rr * v + v0 + i] : 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) ga[j] = (w0 + j < w) ? gk[rr * w + w0 + j] : 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], ga[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
    const int nout = k * v * w;
    for (int i = threadIdx.x; i < nout; i += blockDim.x) s.As[i] = 0.f;
    __syncthreads();
    if (active) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (v0 + i < v && w0 + j < w) atomicAdd(s.As + (kk * v + v0 + i) * w + w0 + j, acc[i][j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nout; i += blockDim.x) atomicAdd(gA + i, s.As[i]);
}

static int check_shape(const char* what, int n, int c, int t, int v, int w, int k, size_t smem_floats) {
    KGAN_REQUIRE(n > 0 && c > 0 && t > 0 && v > 0 && w > 0 && k > 0, "%s: empty dimension", what);
    KGAN_REQUIRE((int64_t)c * t < (1ll << 31) && (int64_t)n * ceil_div64((int64_t)c * t, RB) < (1ll << 31), "%s: too many rows", what);
    KGAN_REQUIRE(smem_floats * 4 <= 200 * 1024, "%s: V=%d, W=%d, K=%d do not fit in shared memory", what, v, w, k);
    return 0;
}

template <typename Kern>
static int set_smem(Kern kern, size_t bytes, bool& done) {
    if (!done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return check_launch("adjmix attribute");
        done = true;
    }
    (void)bytes;
    return 0;
}

}  // namespace kgan

using namespace kgan;

template <int MODE>
static int launch_rowmix(const char* what, const float* in, const float* A, float* out, int n, int c, int t, int v, int w, int k, void* stream) {
    KGAN_REQUIRE(n > 0 && c > 0 && t > 0 && v > 0 && w > 0 && k > 0, "%s: empty dimension", what);
    const int64_t ct64 = (int64_t)c * t;
    KGAN_REQUIRE(ct64 * (MODE == 0 ? w : v) < (1ll << 31) && ct64 * k * w < (1ll << 31), "%s: plane too large", what);
    const int ct = (int)ct64, ko = MODE == 0 ? k : 1, wo = MODE == 0 ? w : v;
    const int nlists = ko * wo;
    const size_t smem = (size_t)((nlists + 3) & ~3) * 4 + (size_t)k * v * w * sizeof(MixEntry);
    KGAN_REQUIRE(smem <= 200 * 1024, "%s: V=%d, W=%d, K=%d do not fit in shared memory", what, v, w, k);
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(adjmix_rowmix_k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return check_launch("adjmix attribute");
        attr = true;
    }
    const int chunks = (int)ceil_div64((int64_t)ct * wo, AT * 4);
    const int64_t items = (int64_t)n * ko * chunks;
    const int64_t grid = items < 8 * kNumSMs ? items : 8 * kNumSMs;
    adjmix_rowmix_k<MODE><<<(unsigned)grid, AT, smem, (cudaStream_t)stream>>>(in, A, out, n, ct, v, w, k, chunks, items);
    return check_launch(what);
}

extern "C" int kgan_adjmix_fwd(const float* x, const float* A, float* out, int n, int c, int t, int v, int w, int k, void* stream) {
    KGAN_REQUIRE(x && A && out, "adjmix_fwd: null pointer");
    return launch_rowmix<0>("adjmix_fwd", x, A, out, n, c, t, v, w, k, stream);
}

extern "C" int kgan_adjmix_bwd_x(const float* g, const float* A, float* gx, int n, int c, int t, int v, int w, int k, void* stream) {
    KGAN_REQUIRE(g && A && gx, "adjmix_bwd_x: null pointer");
    return launch_rowmix<1>("adjmix_bwd_x", g, A, gx, n, c, t, v, w, k, stream);
}

extern "C" int kgan_adjmix_bwd_a(const float* x, const float* g, float* gA, int n, int c, int t, int v, int w, int k, void* stream) {
    KGAN_REQUIRE(x && g && gA, "adjmix_bwd_a: null pointer");
    KGAN_REQUIRE(k * ((v + 3) / 4) * ((w + 3) / 4) <= AT, "adjmix_bwd_a: k*v*w too large");
    const size_t fl = carve_floats(k, v, w, 0, 0);
    if (int e = check_shape("adjmix_bwd_a", n, c, t, v, w, k, fl)) return e;
    static bool attr = false;
    if (int e = set_smem(adjmix_bwd_a_k, fl * 4, attr)) return e;
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(gA, 0, sizeof(float) * k * v * w, s) != cudaSuccess) return check_launch("adjmix_bwd_a memset");
    const int ct = c * t, bps = ceil_div(ct, RB);
    const int64_t total = (int64_t)n * bps;
    const int64_t grid = total < 4 * kNumSMs ? total : 4 * kNumSMs;
    adjmix_bwd_a_k<<<(unsigned)grid, AT, fl * 4, s>>>(x, g, gA, ct, v, w, k, bps, (int)total);
    return check_launch("adjmix_bwd_a");
}
