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

// Fused epilogue of the data-adjoint product (kgan_adjmix_bwd_x_fused): out = (mix + add) * act'(ysrc), both optional and laid out
// like `out`.  It is the join of a critic block's backward: the gradient of the residual branch (`add`) and the LeakyReLU slope of
// the block INPUT (`ysrc` = the previous block's activated output, slope 0.2 where it is <= 0) are applied where the graph-conv
// branch's input gradient is produced, instead of by an elementwise add and a mask kernel over the same tensor.
// `inv` (kgan_adjmix_bwd_x_fused_sel): `add` is a COMPACT tensor (n, c, pc) - the residual branch's gradient at the frames / joints that
// branch kept - and inv[p] is the compact position of plane position p = t * V + v (-1: not kept): the adjoint of the selection is taken
// here instead of by a scatter kernel that writes a mostly-zero full-size tensor for this kernel to read back.
struct MixEpi {
    const float* add;
    const float* ysrc;
    int rnd;
    const int32_t* inv = nullptr;
    int pc = 0, frames = 0;      // compact plane size; frames per channel (rows of the output are (channel, frame) pairs)
    // forward product with a selection by-product (kgan_adjmix_fwd_sel, EPI == 3): out2[n, c, q] = x[n, c, sidx[q]], q < pc - the input of the
    // block's residual branch (x at every other frame / the kept joints), written from the staged tile instead of by a gather kernel
    // that reads x from HBM a second time.  Tiles hold whole channel planes (R % frames == 0).
    const int32_t* sidx = nullptr;
    float* out2 = nullptr;
};
__device__ __forceinline__ float mix_fin(float v, const MixEpi& e, int64_t idx, int wo = 0) {
    if (e.add) {
        if (e.inv) {             // idx = (row r = (n * C + c) * T + t) * wo + joint  (MODE 1: one output block per sample)
            const int64_t r = idx / wo;
            const int j = (int)(idx - r * wo);
            const int64_t nc = r / e.frames;
            const int ci = __ldg(e.inv + (int)(r - nc * e.frames) * wo + j);
            if (ci >= 0) v += __ldg(e.add + nc * e.pc + ci);
        } else {
            v += __ldg(e.add + idx);
        }
    }
    if (e.ysrc) v *= (__ldg(e.ysrc + idx) > 0.f ? 1.f : 0.2f);
    return tf32_out(v, e.rnd);
}
// the same for the pipelined kernel: EPI = false compiles the epilogue operands (and their address arithmetic) away; with EPI the two
// operand pointers advance like the output pointer
template <int EPI>
__device__ __forceinline__ float mix_fin2(float v, const float* ap, const float* yp, int off, int rnd) {
    if (EPI) {
        if (EPI == 1 && ap) v += __ldg(ap + off);
        if (yp) v *= (__ldg(yp + off) > 0.f ? 1.f : 0.2f);
    }
    return tf32_out(v, rnd);
}
// EPI == 2: the compact `add` through the inverse selection map.  (c, t) = channel / frame of the thread's current row, advanced with the row
struct SelPos {
    int c, t;
};
__device__ __forceinline__ float sel_add(const MixEpi& e, int64_t nc0, SelPos ps, int wo, int col) {
    if (col < 0) return 0.f;                      // this thread's joint is not kept by the selection at any frame
    const int ci = __ldg(e.inv + ps.t * wo + col);
    return ci >= 0 ? __ldg(e.add + (nc0 + ps.c) * e.pc + ci) : 0.f;
}
// rows advance by a constant stride: (dc, dt) = (stride / frames, stride % frames), one carry at most
__device__ __forceinline__ SelPos sel_next(SelPos ps, int dc, int dt, int frames) {
    ps.c += dc;
    ps.t += dt;
    if (ps.t >= frames) {
        ps.t -= frames;
        ++ps.c;
    }
    return ps;
}

template <int MODE>
__global__ void __launch_bounds__(AT) adjmix_rowmix_k(const float* __restrict__ in, const float* __restrict__ A, float* __restrict__ out, int n,
                                                       int ct, int v, int w, int k, int chunks_per_block, int64_t total_items, MixEpi epi) {
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
            r[u] = (e0 + u < plane) ? mix_fin(acc, epi, nb * (int64_t)plane + e0 + u, wo) : 0.f;
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


// ---------------------------------------------------------------------------------------------------------------
// Bulk-copy pipelined versions (the default whenever a row block is 16-byte addressable).
//
// Rows q = (c, t) of one sample are contiguous, so a block of R rows is ONE contiguous range per input partition:
// an elected thread fetches it with cp.async.bulk (no registers, no per-element instructions) into a double-buffered
// shared-memory tile while the other buffer is being consumed; an mbarrier carries the completion.
//
// rowmix2: consecutive lanes produce CONSECUTIVE outputs (coalesced 4-byte stores, conflict-free shared-memory reads for
// the self-loop partition, near conflict-free for the neighbour partitions).  The compacted non-zero lists of A are padded
// with zero coefficients to one length per output partition, so the inner loop has a warp-uniform trip count: no
// divergence, and the four independent outputs of a thread give the loads instruction-level parallelism.  (The previous
// kernel ran variable-length dependent load chains per lane: 18-25 % of HBM peak, issue-bound.)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t am_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void am_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void am_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void am_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "AM_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra AM_WAIT_DONE;\n\t"
        "bra AM_WAIT_LOOP;\n\t"
        "AM_WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void am_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}

struct Mix2Plan {
    int R;              // rows per tile
    int tiles_per_n;    // ceil(ct / R)
    int64_t tiles;      // n * tiles_per_n
    int in_floats;      // floats of one staged tile: ki * R * vi
    int lc_max;         // upper bound of the padded list length (stride of the final entry table)
    uint32_t magic;     // ceil(2^32 / wo): e / wo == umulhi(e, magic) for e < 2^20
    int smem_bytes;
};

// rows q, q + rs, ... of one output column: LT = (padded) length of the column's non-zero list, held in registers
template <int LT, int EPI>
__device__ __forceinline__ void mix_rows(const float* __restrict__ xs, const MixEntry* __restrict__ el, float* __restrict__ ob, int q, int rs, int rows,
                                         int vi, int wo, const MixEpi& epi, int64_t obase, int64_t nc0 = 0, int row0 = 0, int col = 0) {
    float cf[LT > 0 ? LT : 1];
    int of[LT > 0 ? LT : 1];
#pragma unroll
    for (int j = 0; j < LT; ++j) {
        const MixEntry en = el[j];
        cf[j] = en.coef;
        of[j] = en.off;
    }
    const float* xp = xs + q * vi;
    float* op = ob + (size_t)q * wo;
    const float* ap = (EPI == 1 && epi.add) ? epi.add + obase + (int64_t)q * wo : nullptr;
    const float* yp = (EPI && epi.ysrc) ? epi.ysrc + obase + (int64_t)q * wo : nullptr;
    const int rnd = epi.rnd;
    const int xstep = rs * vi, ostep = rs * wo;
    SelPos ps{0, 0};
    int sdc = 0, sdt = 0;
    if (EPI == 2) {
        ps.c = (row0 + q) / epi.frames;
        ps.t = (row0 + q) - ps.c * epi.frames;
        sdc = rs / epi.frames;
        sdt = rs - sdc * epi.frames;
    }
    for (; q + 3 * rs < rows; q += 4 * rs, xp += 4 * xstep, op += 4 * ostep) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (EPI == 2) {          // the residual branch's gradient at this thread's four rows (issued before the shared-memory sums)
            a0 = sel_add(epi, nc0, ps, wo, col);
            ps = sel_next(ps, sdc, sdt, epi.frames);
            a1 = sel_add(epi, nc0, ps, wo, col);
            ps = sel_next(ps, sdc, sdt, epi.frames);
            a2 = sel_add(epi, nc0, ps, wo, col);
            ps = sel_next(ps, sdc, sdt, epi.frames);
            a3 = sel_add(epi, nc0, ps, wo, col);
            ps = sel_next(ps, sdc, sdt, epi.frames);
        }
#pragma unroll
        for (int j = 0; j < LT; ++j) {
            a0 = fmaf(cf[j], xp[of[j]], a0);
            a1 = fmaf(cf[j], xp[xstep + of[j]], a1);
            a2 = fmaf(cf[j], xp[2 * xstep + of[j]], a2);
            a3 = fmaf(cf[j], xp[3 * xstep + of[j]], a3);
        }
        op[0] = mix_fin2<EPI>(a0, ap, yp, 0, rnd);
        op[ostep] = mix_fin2<EPI>(a1, ap, yp, ostep, rnd);
        op[2 * ostep] = mix_fin2<EPI>(a2, ap, yp, 2 * ostep, rnd);
        op[3 * ostep] = mix_fin2<EPI>(a3, ap, yp, 3 * ostep, rnd);
        if (EPI) {
            if (ap) ap += 4 * ostep;
            if (yp) yp += 4 * ostep;
        }
    }
    for (; q < rows; q += rs, xp += xstep, op += ostep) {
        float a0 = 0.f;
        if (EPI == 2) {
            a0 = sel_add(epi, nc0, ps, wo, col);
            ps = sel_next(ps, sdc, sdt, epi.frames);
        }
#pragma unroll
        for (int j = 0; j < LT; ++j) a0 = fmaf(cf[j], xp[of[j]], a0);
        op[0] = mix_fin2<EPI>(a0, ap, yp, 0, rnd);
        if (EPI) {
            if (ap) ap += ostep;
            if (yp) yp += ostep;
        }
    }
}
__device__ __noinline__ void mix_rows_any(const float* __restrict__ xs, const MixEntry* __restrict__ el, int L, float* __restrict__ ob, int q, int rs,
                                          int rows, int vi, int wo, const MixEpi& epi, int64_t obase) {
    for (; q < rows; q += rs) {
        float a0 = 0.f;
        for (int j = 0; j < L; ++j) a0 = fmaf(el[j].coef, xs[q * vi + el[j].off], a0);
        ob[(size_t)q * wo] = mix_fin(a0, epi, obase + (int64_t)q * wo, wo);
    }
}

template <int MODE, int EPI>
__global__ void __launch_bounds__(AT, 4) adjmix_rowmix2_k(const float* __restrict__ in, const float* __restrict__ A, float* __restrict__ out, int ct, int v,
                                                        int w, int k, const __grid_constant__ Mix2Plan pl, MixEpi epi) {
    extern __shared__ __align__(128) float sm[];
    const int ko = MODE == 0 ? k : 1, wo = MODE == 0 ? w : v;
    const int vi = MODE == 0 ? v : w, ki = MODE == 0 ? 1 : k;
    const int maxj = MODE == 0 ? v : k * w;
    const int nlists = ko * wo;
    const int R = pl.R;
    // carve: [2 x in tile][bars 16 B][Lk ko ints (padded to 4)][cnt nlists][ent nlists * lc_max]
    float* tile0 = sm;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 2 * pl.in_floats);
    int* Lk = reinterpret_cast<int*>(bars + 2);
    int* cnt = Lk + 4;
    MixEntry* ent = reinterpret_cast<MixEntry*>(cnt + ((nlists + 3) & ~3));
    const uint32_t bar0 = am_smem_u32(bars);

    if (threadIdx.x == 0) {
        am_mbar_init(bar0, 1);
        am_mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 4) Lk[threadIdx.x] = 0;
    __syncthreads();

    auto issue = [&](int64_t tile, int buf) {      // one thread: ki contiguous ranges of this tile -> shared memory
        const int64_t nn = (int)tile / pl.tiles_per_n;       // (tiles < 2^31: checked by the launcher; a 64-bit division per tile and thread is ~150 instructions)
        const int q0 = (int)(tile - nn * pl.tiles_per_n) * R;
        const int rows = min(R, ct - q0);
        const uint32_t bytes = (uint32_t)rows * vi * 4u;
        const uint32_t bar = bar0 + 8 * buf;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic reads of this buffer vs the async writes
        am_mbar_expect_tx(bar, bytes * ki);
        for (int kin = 0; kin < ki; ++kin)
            am_bulk_g2s(am_smem_u32(tile0 + buf * pl.in_floats + kin * R * vi), in + ((nn * ki + kin) * (int64_t)ct + q0) * vi, bytes, bar);
    };
    if (threadIdx.x == 0 && (int64_t)blockIdx.x < pl.tiles) issue(blockIdx.x, 0);       // overlaps the list construction below

    // ---- compacted, padded non-zero lists: entry (offset of the input inside the staged tile relative to its row, coefficient)
    const int lc = pl.lc_max;
    for (int o = threadIdx.x; o < nlists; o += blockDim.x) {
        int c = 0;
        if (MODE == 0) {
            const int kk = o / w, ww = o - kk * w;
            for (int vv = 0; vv < v; ++vv) {
                const float a = __ldg(A + (kk * v + vv) * w + ww);
                if (a != 0.f) ent[o * lc + c++] = MixEntry{vv, a};
            }
            atomicMax(Lk + kk, c);
        } else {
            for (int kk = 0; kk < k; ++kk)
                for (int ww = 0; ww < w; ++ww) {
                    const float a = __ldg(A + (kk * v + o) * w + ww);
                    if (a != 0.f) ent[o * lc + c++] = MixEntry{kk * R * w + ww, a};
                }
            atomicMax(Lk, c);
        }
        cnt[o] = c;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < nlists; o += blockDim.x) {
        const int L = Lk[MODE == 0 ? o / w : 0];
        for (int j = cnt[o]; j < L; ++j) ent[o * lc + j] = MixEntry{0, 0.f};          // padding: 0 * x[row][0]
    }
    __syncthreads();
    (void)maxj;
    // thread -> (row slot, output column): the first (AT / wo) * wo threads each own ONE output column, so a thread's non-zero
    // list is loop-invariant (held in registers) and its outputs are rows my_q, my_q + rs, ...; consecutive lanes still write
    // consecutive addresses
    const int rs = wo <= AT ? AT / wo : 0;
    const bool t_active = (int)threadIdx.x < rs * wo;
    const int my_q = t_active ? (int)threadIdx.x / wo : 0, my_w = t_active ? (int)threadIdx.x - my_q * wo : 0;
    int sel_col = my_w;                            // EPI == 2: -1 if the selection keeps this thread's joint at no frame (most joints of a coarsened graph)
    if (EPI == 2) {
        bool any = false;
        for (int tt = 0; tt < epi.frames; ++tt) any = any || __ldg(epi.inv + tt * wo + my_w) >= 0;
        if (!any) sel_col = -1;
    }

    constexpr int ME = EPI == 3 ? 0 : EPI;          // the selection by-product (EPI 3) is a loop of its own: the product itself has no epilogue
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < pl.tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const int64_t next = tile + gridDim.x;
        if (threadIdx.x == 0 && next < pl.tiles) issue(next, buf ^ 1);                  // buffer buf^1 was released by the barrier below
        const int64_t nn = (int)tile / pl.tiles_per_n;       // (tiles < 2^31: checked by the launcher; a 64-bit division per tile and thread is ~150 instructions)
        const int q0 = (int)(tile - nn * pl.tiles_per_n) * R;
        const int rows = min(R, ct - q0);
        am_mbar_wait(bar0 + 8 * buf, (uint32_t)(it >> 1) & 1u);
        const float* xs = tile0 + buf * pl.in_floats;
        if (t_active) {
            for (int kb = 0; kb < ko; ++kb) {
                const int L = Lk[kb];
                const MixEntry* el = ent + ((size_t)kb * wo + my_w) * lc;          // this thread's list: same output column for every row
                const int64_t obase = ((nn * ko + kb) * (int64_t)ct + q0) * wo + my_w;       // element index of ob[0] in out / add / ysrc
                float* ob = out + obase;
                const int64_t nc0 = EPI == 2 ? nn * (ct / epi.frames) : 0;           // (sample, channel 0) row of the compact `add`
                switch (L) {
                    case 0: mix_rows<0, ME>(xs, el, ob, my_q, rs, rows, vi, wo, epi, obase, nc0, q0, sel_col); break;
                    case 1: mix_rows<1, ME>(xs, el, ob, my_q, rs, rows, vi, wo, epi, obase, nc0, q0, sel_col); break;
                    case 2: mix_rows<2, ME>(xs, el, ob, my_q, rs, rows, vi, wo, epi, obase, nc0, q0, sel_col); break;
                    case 3: mix_rows<3, ME>(xs, el, ob, my_q, rs, rows, vi, wo, epi, obase, nc0, q0, sel_col); break;
                    case 4: mix_rows<4, ME>(xs, el, ob, my_q, rs, rows, vi, wo, epi, obase, nc0, q0, sel_col); break;
                    case 5: mix_rows<5, ME>(xs, el, ob, my_q, rs, rows, vi, wo, epi, obase, nc0, q0, sel_col); break;
                    case 6: mix_rows<6, ME>(xs, el, ob, my_q, rs, rows, vi, wo, epi, obase, nc0, q0, sel_col); break;
                    case 7: mix_rows<7, ME>(xs, el, ob, my_q, rs, rows, vi, wo, epi, obase, nc0, q0, sel_col); break;
                    case 8: mix_rows<8, ME>(xs, el, ob, my_q, rs, rows, vi, wo, epi, obase, nc0, q0, sel_col); break;
                    default: mix_rows_any(xs, el, L, ob, my_q, rs, rows, vi, wo, epi, obase); break;
                }
            }
        }
        if (EPI == 3) {
            const int c0 = q0 / epi.frames, total = (rows / epi.frames) * epi.pc;       // whole channel planes: rows % frames == 0
            float* o2 = epi.out2 + (nn * (ct / epi.frames) + c0) * (int64_t)epi.pc;
            for (int i = threadIdx.x; i < total; i += AT) {
                const int cl = i / epi.pc, qc = i - cl * epi.pc;
                o2[i] = xs[cl * epi.frames * vi + __ldg(epi.sidx + qc)];
            }
        }
        __syncthreads();                                                                // every read of xs[buf] is done
    }
}

// gA restricted to the non-zeros of a mask (the adjacency itself: d loss / d edge_importance = gA (.) A, so entries where
// A == 0 are never used).  lane = (non-zero entry, row group): for every staged row one shared-memory read of x, one of g and
// one FMA - a few dozen dot products instead of the dense (k*w) x v product, which made the dense kernel FMA/LDS-bound at
// ~20 % of HBM peak.  Tiles are fetched with cp.async.bulk like above.
struct DA2Plan {
    int R, tiles_per_n;
    int64_t tiles;
    int x_floats, g_floats;     // staged floats per tile: R * v, k * g_stride
    int g_stride;               // floats between the partitions of g in shared memory: R * w + pad (12 banks apart)
    int cap;                    // entry capacity: k * v * w
    int smem_bytes;
};
struct DAEntry {
    int xoff, goff, oidx;
};
constexpr int DA_EPT = 8;       // entries per thread at most: k * v * w <= DA_EPT * AT

__global__ void __launch_bounds__(AT) adjmix_bwd_a2_k(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ mask,
                                                       float* __restrict__ gA, int ct, int v, int w, int k, const __grid_constant__ DA2Plan pl) {
    extern __shared__ __align__(128) float sm[];
    const int R = pl.R, stage = pl.x_floats + pl.g_floats;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 2 * stage);
    int* nnz_s = reinterpret_cast<int*>(bars + 2);
    DAEntry* ent = reinterpret_cast<DAEntry*>(nnz_s + 4);
    float* red = reinterpret_cast<float*>(ent + pl.cap);
    const uint32_t bar0 = am_smem_u32(bars);
    if (threadIdx.x == 0) {
        am_mbar_init(bar0, 1);
        am_mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *nnz_s = 0;
    }
    __syncthreads();
    auto issue = [&](int64_t tile, int buf) {
        const int64_t nn = (int)tile / pl.tiles_per_n;       // (tiles < 2^31: checked by the launcher; a 64-bit division per tile and thread is ~150 instructions)
        const int q0 = (int)(tile - nn * pl.tiles_per_n) * R;
        const int rows = min(R, ct - q0);
        const uint32_t bar = bar0 + 8 * buf;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        am_mbar_expect_tx(bar, (uint32_t)rows * (v + k * w) * 4u);
        float* xs = sm + buf * stage;
        am_bulk_g2s(am_smem_u32(xs), x + (nn * (int64_t)ct + q0) * v, (uint32_t)rows * v * 4u, bar);
        for (int kk = 0; kk < k; ++kk)
            am_bulk_g2s(am_smem_u32(xs + pl.x_floats + kk * pl.g_stride), g + ((nn * k + kk) * (int64_t)ct + q0) * w, (uint32_t)rows * w * 4u, bar);
    };
    if (threadIdx.x == 0 && (int64_t)blockIdx.x < pl.tiles) issue(blockIdx.x, 0);
    // entry list in (k, v, w) order: warp 1 compacts the mask with ballots (k*v*w <= a few thousand)
    if (threadIdx.x >= 32 && threadIdx.x < 64) {
        const int lane = threadIdx.x - 32, total = k * v * w;
        int c = 0;
        for (int i0 = 0; i0 < total; i0 += 32) {
            const int i = i0 + lane;
            const bool nzp = i < total && __ldg(mask + i) != 0.f;
            const unsigned bal = __ballot_sync(0xffffffffu, nzp);
            if (nzp) {
                const int kk = i / (v * w), r = i - kk * v * w, vv = r / w, ww = r - vv * w;
                ent[c + __popc(bal & ((1u << lane) - 1u))] = DAEntry{vv, kk * pl.g_stride + ww, i};
            }
            c += __popc(bal);
        }
        if (lane == 0) *nnz_s = c;
    }
    __syncthreads();
    const int nnz = *nnz_s;
    // nnz <= AT: lane = (entry, row group), the rows of a tile are split over AT / nnz groups.  Otherwise (dense masks): one
    // row group, up to DA_EPT entries per thread.
    const bool multi = nnz > AT;
    const int rg_count = multi ? 1 : (nnz > 0 ? AT / nnz : 1);
    const int my_e = multi ? threadIdx.x : threadIdx.x % max(nnz, 1), my_rg = multi ? 0 : threadIdx.x / max(nnz, 1);
    const bool active = nnz > 0 && my_rg < rg_count;
    float acc[DA_EPT];
#pragma unroll
    for (int s = 0; s < DA_EPT; ++s) acc[s] = 0.f;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < pl.tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const int64_t next = tile + gridDim.x;
        if (threadIdx.x == 0 && next < pl.tiles) issue(next, buf ^ 1);
        const int64_t nn = (int)tile / pl.tiles_per_n;       // (tiles < 2^31: checked by the launcher; a 64-bit division per tile and thread is ~150 instructions)
        const int q0 = (int)(tile - nn * pl.tiles_per_n) * R;
        const int rows = min(R, ct - q0);
        am_mbar_wait(bar0 + 8 * buf, (uint32_t)(it >> 1) & 1u);
        const float* xs = sm + buf * stage;
        const float* gs = xs + pl.x_floats;
        if (!multi) {
            if (active) {
                const DAEntry me = ent[my_e];
                const float* xp = xs + me.xoff;
                const float* gp = gs + me.goff;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                int r = my_rg;
                for (; r + 3 * rg_count < rows; r += 4 * rg_count) {
                    a0 = fmaf(xp[r * v], gp[r * w], a0);
                    a1 = fmaf(xp[(r + rg_count) * v], gp[(r + rg_count) * w], a1);
                    a2 = fmaf(xp[(r + 2 * rg_count) * v], gp[(r + 2 * rg_count) * w], a2);
                    a3 = fmaf(xp[(r + 3 * rg_count) * v], gp[(r + 3 * rg_count) * w], a3);
                }
                for (; r < rows; r += rg_count) a0 = fmaf(xp[r * v], gp[r * w], a0);
                acc[0] += (a0 + a1) + (a2 + a3);
            }
        } else {
#pragma unroll
            for (int s = 0; s < DA_EPT; ++s) {
                const int e = s * AT + threadIdx.x;
                if (e < nnz) {
                    const DAEntry me = ent[e];
                    const float* xp = xs + me.xoff;
                    const float* gp = gs + me.goff;
                    float a0 = 0.f;
                    for (int r = 0; r < rows; ++r) a0 = fmaf(xp[r * v], gp[r * w], a0);
                    acc[s] += a0;
                }
            }
        }
        __syncthreads();
    }
    // CTA-level merge of the row groups, then one atomic per non-zero
    for (int i = threadIdx.x; i < nnz; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    if (!multi) {
        if (active) atomicAdd(red + my_e, acc[0]);
    } else {
#pragma unroll
        for (int s = 0; s < DA_EPT; ++s)
            if (s * AT + threadIdx.x < nnz) red[s * AT + threadIdx.x] = acc[s];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nnz; i += blockDim.x) atomicAdd(gA + ent[i].oidx, red[i]);
}

__global__ void mask_zero_k(float* __restrict__ gA, const float* __restrict__ mask, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && mask[i] == 0.f) gA[i] = 0.f;
}

static int check_shape(const char* what, int n, int c, int t, int v, int w, int k, size_t smem_floats) {
    KGAN_REQUIRE(n > 0 && c > 0 && t > 0 && v > 0 && w > 0 && k > 0, "%s: empty dimension", what);
    KGAN_REQUIRE((int64_t)c * t < (1ll << 31) && (int64_t)n * ceil_div64((int64_t)c * t, RB) < (1ll << 31), "%s: too many rows", what);
    KGAN_REQUIRE(smem_floats * 4 <= 200 * 1024, "%s: V=%d, W=%d, K=%d do not fit in shared memory", what, v, w, k);
    return 0;
}

}  // namespace kgan

using namespace kgan;

template <int MODE>
static int launch_rowmix(const char* what, const float* in, const float* A, float* out, int n, int c, int t, int v, int w, int k, MixEpi epi, void* stream,
                         bool dry = false) {      // dry: no launch - 0 if the selection by-product (epi.out2) has a plan, 1 if not
    KGAN_REQUIRE(n > 0 && c > 0 && t > 0 && v > 0 && w > 0 && k > 0, "%s: empty dimension", what);
    const int64_t ct64 = (int64_t)c * t;
    KGAN_REQUIRE(ct64 * (MODE == 0 ? w : v) < (1ll << 31) && ct64 * k * w < (1ll << 31), "%s: plane too large", what);
    const int ct = (int)ct64, ko = MODE == 0 ? k : 1, wo = MODE == 0 ? w : v;
    const int nlists = ko * wo;
    const size_t smem = (size_t)((nlists + 3) & ~3) * 4 + (size_t)k * v * w * sizeof(MixEntry);
    KGAN_REQUIRE(smem <= 200 * 1024, "%s: V=%d, W=%d, K=%d do not fit in shared memory", what, v, w, k);
    static SmemAttrOnce attr;
    if (!dry)
        if (int e = ensure_smem(adjmix_rowmix_k<MODE>, 200 * 1024, attr, "adjmix attribute")) return e;
    const bool want_sel = MODE == 0 && (epi.out2 != nullptr || dry);
    // bulk-copy pipelined kernel whenever every row block is 16-byte addressable
    {
        const int vi = MODE == 0 ? v : w, ki = MODE == 0 ? 1 : k;
        const int maxj = MODE == 0 ? v : k * w;
        // (planes of a few hundred bytes - the label term of the critic's first layer, 32 rows of 3 floats per sample - are latency-bound
        // in the pipelined kernel: one barrier round trip per 384-byte tile made that launch 123 us at 32 GB/s; the plain kernel serves them)
        if (k <= 4 && wo <= AT && ((int64_t)ct * vi) % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (int64_t)ct * vi * ki >= 2048) {
            Mix2Plan pl;
            // <= 24 KB per stage: 4 CTAs (2 stages each) per SM; the forward product where it writes more than it reads (K partitions out
            // of one input) runs 9-12 % faster with 12 KB stages and 6 CTAs per SM, everything else slower (same-box sweep of 12 / 16 / 24 / 32 KB,
            // tools/adjmix_bench.py)
            const int stage_bytes = (MODE == 0 && ko * wo > vi) ? 12 * 1024 : 24 * 1024;
            int R = stage_bytes / (ki * vi * 4);
            R = R / 8 * 8;
            if (R > 1024) R = 1024;
            if (R >= ct) R = (ct + 3) / 4 * 4;
            else R = (ceil_div(ct, ceil_div(ct, R)) + 7) / 8 * 8;   // equal tiles: no sliver at the end of a sample
            if (want_sel) {                                         // whole channel planes per tile
                R = R >= t ? R / t * t : t;
                if ((R & 3) || (int64_t)R * vi * 4 > 32 * 1024) R = 0;
            }
            if (R >= 8 && (int64_t)R * wo < (1 << 20) && (int64_t)R * vi * ki < (1 << 20) && (int64_t)n * ceil_div(ct, R) < (1ll << 31)) {
                pl.R = R;
                pl.tiles_per_n = ceil_div(ct, R);
                pl.tiles = (int64_t)n * pl.tiles_per_n;
                pl.in_floats = ki * R * vi;
                pl.lc_max = maxj | 1;
                pl.magic = wo == 1 ? 0u : (uint32_t)(((1ull << 32) + wo - 1) / wo);      // wo == 1: q = e (2^32 does not fit)
                const size_t smem2 = (size_t)2 * pl.in_floats * 4 + 16 + 16 + (size_t)((nlists + 3) & ~3) * 4 + (size_t)nlists * pl.lc_max * sizeof(MixEntry);
                pl.smem_bytes = (int)smem2;
                if (smem2 <= 100 * 1024) {
                    if (dry) return 0;
                    static SmemAttrOnce attr2, attr3;
                    if (int e = ensure_smem(adjmix_rowmix2_k<MODE, 0>, 100 * 1024, attr2, "adjmix attribute")) return e;
                    if (int e = ensure_smem(adjmix_rowmix2_k<MODE, 1>, 100 * 1024, attr3, "adjmix attribute")) return e;
                    const int per_sm = (int)((220 * 1024) / (smem2 + 1024));
                    const int64_t cap = (int64_t)kNumSMs * (per_sm < 1 ? 1 : per_sm > 6 ? 6 : per_sm);
                    const int64_t waves = ceil_div64(pl.tiles, cap);
                    const int64_t grid2 = ceil_div64(pl.tiles, waves);   // every CTA gets `waves` tiles (+-1), all CTAs co-resident
                    if (want_sel) {
                        static SmemAttrOnce attr5;
                        if (int e = ensure_smem(adjmix_rowmix2_k<MODE, 3>, 100 * 1024, attr5, "adjmix attribute")) return e;
                        adjmix_rowmix2_k<MODE, 3><<<(unsigned)grid2, AT, smem2, (cudaStream_t)stream>>>(in, A, out, ct, v, w, k, pl, epi);
                    } else if (epi.add && epi.inv) {
                        static SmemAttrOnce attr4;
                        if (int e = ensure_smem(adjmix_rowmix2_k<MODE, 2>, 100 * 1024, attr4, "adjmix attribute")) return e;
                        adjmix_rowmix2_k<MODE, 2><<<(unsigned)grid2, AT, smem2, (cudaStream_t)stream>>>(in, A, out, ct, v, w, k, pl, epi);
                    } else if (epi.add || epi.ysrc) adjmix_rowmix2_k<MODE, 1><<<(unsigned)grid2, AT, smem2, (cudaStream_t)stream>>>(in, A, out, ct, v, w, k, pl, epi);
                    else adjmix_rowmix2_k<MODE, 0><<<(unsigned)grid2, AT, smem2, (cudaStream_t)stream>>>(in, A, out, ct, v, w, k, pl, epi);
                    return check_launch(what);
                }
            }
        }
    }
    if (dry) return 1;
    KGAN_REQUIRE(!want_sel, "%s: no plan for the selection by-product (kgan_adjmix_fwd_sel_ok() == 0)", what);
    const int chunks = (int)ceil_div64((int64_t)ct * wo, AT * 4);
    const int64_t items = (int64_t)n * ko * chunks;
    const int64_t grid = items < 8 * kNumSMs ? items : 8 * kNumSMs;
    adjmix_rowmix_k<MODE><<<(unsigned)grid, AT, smem, (cudaStream_t)stream>>>(in, A, out, n, ct, v, w, k, chunks, items, epi);
    return check_launch(what);
}

extern "C" int kgan_adjmix_fwd(const float* x, const float* A, float* out, int n, int c, int t, int v, int w, int k, int out_tf32, void* stream) {
    KGAN_REQUIRE(x && A && out, "adjmix_fwd: null pointer");
    return launch_rowmix<0>("adjmix_fwd", x, A, out, n, c, t, v, w, k, MixEpi{nullptr, nullptr, out_tf32}, stream);
}

extern "C" int kgan_adjmix_fwd_sel_ok(const float* x, int n, int c, int t, int v, int w, int k) {
    if (!x || n <= 0 || c <= 0 || t <= 0 || v <= 0 || w <= 0 || k <= 0 || (int64_t)c * t * v >= (1ll << 31) || (int64_t)c * t * k * w >= (1ll << 31)) return 0;
    return launch_rowmix<0>("adjmix_fwd_sel", x, nullptr, nullptr, n, c, t, v, w, k, MixEpi{nullptr, nullptr, 0}, nullptr, true) == 0 ? 1 : 0;
}

extern "C" int kgan_adjmix_fwd_sel(const float* x, const float* A, const int32_t* sidx, int pc, float* out, float* out2, int n, int c, int t, int v, int w,
                                   int k, int out_tf32, void* stream) {
    KGAN_REQUIRE(x && A && out && out2 && sidx && pc > 0, "adjmix_fwd_sel: null pointer");
    MixEpi epi{nullptr, nullptr, out_tf32};
    epi.sidx = sidx;
    epi.out2 = out2;
    epi.pc = pc;
    epi.frames = t;
    return launch_rowmix<0>("adjmix_fwd_sel", x, A, out, n, c, t, v, w, k, epi, stream);
}

extern "C" int kgan_adjmix_bwd_x(const float* g, const float* A, float* gx, int n, int c, int t, int v, int w, int k, int out_tf32, void* stream) {
    KGAN_REQUIRE(g && A && gx, "adjmix_bwd_x: null pointer");
    return launch_rowmix<1>("adjmix_bwd_x", g, A, gx, n, c, t, v, w, k, MixEpi{nullptr, nullptr, out_tf32}, stream);
}

extern "C" int kgan_adjmix_bwd_x_fused(const float* g, const float* A, const float* add, const float* ysrc, float* gx, int n, int c, int t, int v,
                                       int w, int k, int out_tf32, void* stream) {
    KGAN_REQUIRE(g && A && gx, "adjmix_bwd_x_fused: null pointer");
    return launch_rowmix<1>("adjmix_bwd_x_fused", g, A, gx, n, c, t, v, w, k, MixEpi{add, ysrc, out_tf32}, stream);
}

extern "C" int kgan_adjmix_bwd_x_fused_sel(const float* g, const float* A, const float* add_c, const int32_t* inv, int pc, const float* ysrc, float* gx,
                                           int n, int c, int t, int v, int w, int k, int out_tf32, void* stream) {
    KGAN_REQUIRE(g && A && gx && add_c && inv && pc > 0, "adjmix_bwd_x_fused_sel: null pointer");
    MixEpi epi{add_c, ysrc, out_tf32};
    epi.inv = inv;
    epi.pc = pc;
    epi.frames = t;
    return launch_rowmix<1>("adjmix_bwd_x_fused_sel", g, A, gx, n, c, t, v, w, k, epi, stream);
}

extern "C" int kgan_adjmix_bwd_a(const float* x, const float* g, float* gA, int n, int c, int t, int v, int w, int k, void* stream) {
    return kgan_adjmix_bwd_a_masked(x, g, nullptr, gA, n, c, t, v, w, k, stream);
}

extern "C" int kgan_adjmix_bwd_a_masked(const float* x, const float* g, const float* mask, float* gA, int n, int c, int t, int v, int w, int k,
                                        void* stream) {
    KGAN_REQUIRE(x && g && gA, "adjmix_bwd_a: null pointer");
    KGAN_REQUIRE(n > 0 && c > 0 && t > 0 && v > 0 && w > 0 && k > 0, "adjmix_bwd_a: empty dimension");
    const int64_t ct64 = (int64_t)c * t;
    if (mask && (ct64 * v) % 4 == 0 && (ct64 * w) % 4 == 0 && ct64 < (1ll << 30) &&
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g)) & 15) == 0) {
        // sparse path: lane = non-zero of the mask (any number of non-zeros up to k*v*w <= DA_EPT * AT)
        const int per_row = (v + k * w) * 4;
        int R = (48 * 1024) / per_row;
        R = R / 8 * 8;
        if (R > 1024) R = 1024;
        if (R >= ct64) R = (int)((ct64 + 3) / 4 * 4);
        else R = (int)((ceil_div64(ct64, ceil_div64(ct64, R)) + 7) / 8 * 8);
        if (R >= 8 && k * v * w <= DA_EPT * AT && (int64_t)n * ceil_div64(ct64, R) < (1ll << 31)) {
            DA2Plan pl;
            pl.R = R;
            pl.tiles_per_n = (int)ceil_div64(ct64, R);
            pl.tiles = (int64_t)n * pl.tiles_per_n;
            pl.x_floats = R * v;
            pl.g_stride = R * w + ((12 - (R * w) % 32 + 32) % 32);
            pl.g_floats = k * pl.g_stride;
            pl.cap = k * v * w;
            const size_t smem2 = (size_t)2 * (pl.x_floats + pl.g_floats) * 4 + 16 + 16 + (sizeof(DAEntry) + 4) * (size_t)pl.cap;
            pl.smem_bytes = (int)smem2;
            static SmemAttrOnce attr2;
            if (int e = ensure_smem(adjmix_bwd_a2_k, 160 * 1024, attr2, "adjmix attribute")) return e;
            cudaStream_t s2 = (cudaStream_t)stream;
            if (cudaMemsetAsync(gA, 0, sizeof(float) * k * v * w, s2) != cudaSuccess) return check_launch("adjmix_bwd_a memset");
            const int64_t cap = 2 * kNumSMs;
            const int64_t grid2 = ceil_div64(pl.tiles, ceil_div64(pl.tiles, cap));
            adjmix_bwd_a2_k<<<(unsigned)grid2, AT, smem2, s2>>>(x, g, mask, gA, (int)ct64, v, w, k, pl);
            return check_launch("adjmix_bwd_a");
        }
    }
    KGAN_REQUIRE(k * ((v + 3) / 4) * ((w + 3) / 4) <= AT, "adjmix_bwd_a: k*v*w too large");
    const size_t fl = carve_floats(k, v, w, 0, 0);
    if (int e = check_shape("adjmix_bwd_a", n, c, t, v, w, k, fl)) return e;
    static SmemAttrOnce attr;
    if (int e = ensure_smem(adjmix_bwd_a_k, 200 * 1024, attr, "adjmix attribute")) return e;
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(gA, 0, sizeof(float) * k * v * w, s) != cudaSuccess) return check_launch("adjmix_bwd_a memset");
    const int ct = c * t, bps = ceil_div(ct, RB);
    const int64_t total = (int64_t)n * bps;
    const int64_t grid = total < 4 * kNumSMs ? total : 4 * kNumSMs;
    adjmix_bwd_a_k<<<(unsigned)grid, AT, fl * 4, s>>>(x, g, gA, ct, v, w, k, bps, (int)total);
    if (mask) mask_zero_k<<<ceil_div(k * v * w, 256), 256, 0, s>>>(gA, mask, k * v * w);      // same contract as the sparse kernel
    return check_launch("adjmix_bwd_a");
}
