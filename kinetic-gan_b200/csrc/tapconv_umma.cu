// Tap convolution on the 5th-generation tensor cores (KGAN_PREC_TF32): tcgen05.mma kind::tf32, fp32 accumulators
// in TMEM, weights staged by the bulk-copy engine (cp.async.bulk -> UBLKCP) from a pre-packed tf32 image,
// activations gathered by SIMT producer warps (position map = zero padding / joint & frame selection / tap shift),
// mbarrier full/empty ring between producers and the single MMA-issuing thread.
//
// GEMM orientation (see DESIGN.md "tcgen05 tap convolution"):
//   D[M = 128 positions][N = output channels]  +=  A[M][K] * B[N][K]^T,   K = (input-channel tile, tap)
//   A (activations): gathered from NCHW global memory, one thread per position row, written K-major into shared
//       memory in the canonical no-swizzle core-matrix layout (8 rows x 16 B), rounded to tf32 (cvt.rna).
//   B (weights):     pre-packed by tapconv_pack_k in exactly the shared-memory image, so a stage is 8 bulk copies.
//   D: TMEM lane = position, TMEM column = output channel -> the epilogue thread that owns lane l stores
//       out[n, oc, p(l)]: for a fixed oc consecutive lanes are consecutive addresses (coalesced NCHW stores).
//
// The layer is HBM/L2-bound at tensor-core rates (AI ~ 0.75*C_out flop/B unfused), so each CTA computes ALL of its
// output channels for a 128-position tile and reads the activation tile exactly once.
#include "umma.cuh"

namespace kgan {

struct UmmaPlan {
    int n_cta;        // output channels per CTA (multiple of 16; of 32 when > 256)
    int n_split;      // CTAs along output channels
    int n_rows;       // n_cta * n_split: rows of the packed weight image (zero padded)
    int n_acc;        // accumulators per CTA (1 or 2)
    int n_per_acc;    // UMMA N
    int tmem_cols;    // power of two >= n_cta, >= 32
    int stages;
    int nkt;          // input-channel tiles of UK
    int smem_bytes;
};

static bool make_plan(const kgan_tapconv_desc& d, UmmaPlan& p) {
    if (d.ck < 16 || d.co < 16 || d.w_oc_blk != 0) return false;
    const int64_t total = (int64_t)d.n * d.p_out;
    if (total < 256) return false;
    const int64_t tiles = ceil_div64(total, UM) * d.groups;
    const int n16 = round_up(d.co, 16);
    int split = ceil_div(n16, 512);
    while (tiles * split < kNumSMs && ceil_div(n16, split * 2) >= 64) split *= 2;
    int n_cta = round_up(ceil_div(n16, split), 16);
    if (n_cta > 256) n_cta = round_up(n_cta, 32);
    p.n_cta = n_cta;
    p.n_split = ceil_div(n16, n_cta);
    p.n_rows = p.n_cta * p.n_split;
    p.n_acc = n_cta > 256 ? 2 : 1;
    p.n_per_acc = n_cta / p.n_acc;
    p.tmem_cols = 32;
    while (p.tmem_cols < n_cta) p.tmem_cols *= 2;
    p.nkt = ceil_div(d.ck, UK);
    const int stage = A_STAGE_BYTES + n_cta * UK * 4;
    const int budget = (p.tmem_cols <= 256 ? 100 : 200) * 1024;      // <= 256 columns: two CTAs per SM
    p.stages = budget / stage;
    if (p.stages > 6) p.stages = 6;
    if (p.stages < 2) p.stages = 2;
    p.smem_bytes = p.stages * stage + 256;
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// weight packing: natural (strided) fp32 weights -> tf32 shared-memory image
//   wp[group][ic tile][tap][k-chunk c (8)][row r (n_rows)][4]     (zero for r >= co or ic >= ck)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tapconv_pack_k(const __grid_constant__ kgan_tapconv_desc d, const float* __restrict__ w,
                                                       float* __restrict__ wp, int n_rows, int nkt) {
    const int64_t total = (int64_t)d.groups * nkt * d.ntap * 8 * n_rows * 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(i & 3);
        int64_t r = i >> 2;
        const int row = (int)(r % n_rows);
        r /= n_rows;
        const int c = (int)(r & 7);
        r >>= 3;
        const int tap = (int)(r % d.ntap);
        r /= d.ntap;
        const int ict = (int)(r % nkt);
        const int g = (int)(r / nkt);
        const int ic = ict * UK + c * 4 + e;
        float v = 0.f;
        if (row < d.co && ic < d.ck) v = __ldg(w + (int64_t)g * d.g_w + d.tap_w_off[tap] + (int64_t)row * d.w_oc + (int64_t)ic * d.w_ic);
        wp[i] = __uint_as_float(to_tf32(v));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(UMMA_THREADS) tapconv_fwd_umma(const __grid_constant__ kgan_tapconv_desc d, const __grid_constant__ UmmaPlan pl,
                                                                 const float* __restrict__ in, const float* __restrict__ wp,
                                                                 const int32_t* __restrict__ pmap, const float* __restrict__ bias,
                                                                 const float* __restrict__ add, float* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = pl.stages;
    const int b_stage_bytes = pl.n_cta * UK * 4;
    uint8_t* a_base = smem;
    uint8_t* b_base = smem + (size_t)S * A_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + (size_t)S * b_stage_bytes);     // full[S], empty[S], accfull
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S), accfull = smem_u32(bars + 2 * S);

    const int g = blockIdx.z;
    const int oc_base = blockIdx.y * pl.n_cta;
    const int64_t m0 = (int64_t)blockIdx.x * UM;
    const int64_t total_pos = (int64_t)d.n * d.p_out;
    const int iters = pl.nkt * d.ntap;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full0 + 8 * s, UM + 1);       // 128 producer threads + the weight loader's expect_tx arrive
            mbar_init(empty0 + 8 * s, 1);           // tcgen05.commit
        }
        mbar_init(accfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(pl.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ===== activation producers: thread t owns tile row t (one output position) =====
        const int t = threadIdx.x;
        const int64_t pos = m0 + t;
        const bool valid = pos < total_pos;
        const int nn = valid ? (int)(pos / d.p_out) : 0, p = valid ? (int)(pos % d.p_out) : 0;
        const float* in_n = in + ((int64_t)nn * d.c_in_total + g * d.g_in) * d.p_in;
        // position map of this row for every tap (constant over the input-channel tiles)
        constexpr int SRC_CACHE = 4;
        int srcs[SRC_CACHE];
#pragma unroll
        for (int i = 0; i < SRC_CACHE; ++i) srcs[i] = (valid && i < d.ntap) ? __ldg(pmap + (int64_t)d.tap_row[i] * d.p_out + p) : -1;
        // all UK loads of a stage are issued back to back (volatile asm keeps them ahead of the barrier wait), and the loads
        // of stage it+1 are in flight while stage it is converted and stored: two stages of gathers outstanding per thread
        auto issue = [&](int it, float (&v)[UK]) {
            const int ict = it / d.ntap, tap = it - ict * d.ntap, ic0 = ict * UK;
            int src = -1;
            if (tap < SRC_CACHE) {
#pragma unroll
                for (int i = 0; i < SRC_CACHE; ++i) src = (i == tap) ? srcs[i] : src;
            } else if (valid) {
                src = __ldg(pmap + (int64_t)d.tap_row[tap] * d.p_out + p);
            }
            const float* xb = in_n + (int64_t)(d.tap_in_ch[tap] + ic0) * d.p_in + src;
#pragma unroll
            for (int kk = 0; kk < UK; ++kk) v[kk] = ldg_pred(xb + (int64_t)kk * d.p_in, src >= 0 && ic0 + kk < d.ck);
        };
        auto stage_out = [&](int it, float (&v)[UK]) {
            const int s = it % S;
            const uint32_t ph = (uint32_t)(it / S) & 1u;
            mbar_wait(empty0 + 8 * s, ph ^ 1u);                      // slot free (first lap passes immediately)
            const uint32_t dst = smem_u32(a_base + (size_t)s * A_STAGE_BYTES) + t * 16;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + c * A_LBO), "r"(to_tf32(v[4 * c])),
                             "r"(to_tf32(v[4 * c + 1])), "r"(to_tf32(v[4 * c + 2])), "r"(to_tf32(v[4 * c + 3]))
                             : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
            mbar_arrive(full0 + 8 * s);
        };
        float va[UK], vb[UK];
        issue(0, va);
        for (int it = 0; it < iters; it += 2) {
            if (it + 1 < iters) issue(it + 1, vb);
            stage_out(it, va);
            if (it + 1 < iters) {
                if (it + 2 < iters) issue(it + 2, va);
                stage_out(it + 1, vb);
            }
        }
        // ===== epilogue: TMEM lane t -> out[n, oc, p] =====
        mbar_wait(accfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int out_ch0 = g * d.g_out;
        for (int col0 = 0; col0 < pl.n_cta; col0 += 16) {
            if (oc_base + col0 >= d.co) break;                       // warp-uniform
            uint32_t r[16];
            tmem_ld16(taddr + col0, r);
            if (valid) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int oc = oc_base + col0 + j;
                    if (oc < d.co) {
                        const int64_t o = ((int64_t)nn * d.c_out_total + out_ch0 + oc) * d.p_out + p;
                        float val = __uint_as_float(r[j]);
                        if (bias) val += __ldg(bias + out_ch0 + oc);
                        if (add) val += __ldg(add + (d.add_period ? ((int64_t)nn * d.c_out_total + out_ch0 + oc) * d.add_period + p % d.add_period : o));
                        out[o] = apply_act(val, d.act);
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else if (warp == 4) {
        // ===== MMA issuer: one thread drives the tensor core =====
        if (lane == 0) {
            const uint32_t idesc = instr_desc_tf32(pl.n_per_acc);
            const uint32_t b_lbo = pl.n_cta * 16;
            for (int it = 0; it < iters; ++it) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                mbar_wait(full0 + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(a_base + (size_t)s * A_STAGE_BYTES);
                const uint32_t b_addr = smem_u32(b_base + (size_t)s * b_stage_bytes);
#pragma unroll
                for (int j = 0; j < UK / 8; ++j) {
                    const uint64_t adesc = smem_desc(a_addr + j * 2 * A_LBO, A_LBO, CORE_SBO);
                    for (int a = 0; a < pl.n_acc; ++a) {
                        const uint64_t bdesc = smem_desc(b_addr + a * pl.n_per_acc * 16 + j * 2 * b_lbo, b_lbo, CORE_SBO);
                        umma_tf32(tmem_base + a * pl.n_per_acc, adesc, bdesc, idesc, (it > 0 || j > 0) ? 1u : 0u);
                    }
                }
                umma_commit(empty0 + 8 * s);                          // frees the stage when these MMAs retire
            }
            umma_commit(accfull);
        }
    } else {
        // ===== weight loader: bulk copies of the packed tf32 image =====
        if (lane == 0) {
            const float* wg = wp + (int64_t)g * pl.nkt * d.ntap * pl.n_rows * UK;
            const uint32_t chunk_bytes = pl.n_cta * 16;
            for (int it = 0; it < iters; ++it) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                mbar_arrive_expect_tx(full0 + 8 * s, chunk_bytes * 8);
                const float* src = wg + (int64_t)it * pl.n_rows * UK;          // loop order == packing order (ic tile, tap)
                const uint32_t dst = smem_u32(b_base + (size_t)s * b_stage_bytes);
#pragma unroll
                for (int c = 0; c < 8; ++c) bulk_g2s(dst + c * chunk_bytes, src + ((int64_t)c * pl.n_rows + oc_base) * 4, chunk_bytes, full0 + 8 * s);
            }
        }
    }
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(pl.tmem_cols) : "memory");
    }
}

int64_t tapconv_tf32_packed_numel(const kgan_tapconv_desc& d) {
    UmmaPlan p;
    if (!make_plan(d, p)) return 0;
    return (int64_t)d.groups * p.nkt * d.ntap * p.n_rows * UK;
}

int tapconv_pack_tf32(const kgan_tapconv_desc& d, const float* w, float* wp, cudaStream_t stream) {
    UmmaPlan p;
    if (!make_plan(d, p)) {
        set_error("tapconv_pack: shape not eligible for the tf32 path");
        return 1;
    }
    const int64_t total = (int64_t)d.groups * p.nkt * d.ntap * p.n_rows * UK;
    int64_t blocks = ceil_div64(total, 256);
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    tapconv_pack_k<<<(unsigned)blocks, 256, 0, stream>>>(d, w, wp, p.n_rows, p.nkt);
    return check_launch("tapconv_pack");
}

int tapconv_fwd_tf32(const kgan_tapconv_desc& d, const float* in, const float* wp, const int32_t* pmap, const float* bias,
                     const float* add, float* out, cudaStream_t stream) {
    UmmaPlan p;
    if (!make_plan(d, p)) return -1;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(tapconv_fwd_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return check_launch("tapconv_fwd_tf32 attribute");
        attr_set = true;
    }
    const int64_t tiles = ceil_div64((int64_t)d.n * d.p_out, UM);
    dim3 grid((unsigned)tiles, p.n_split, d.groups);
    tapconv_fwd_umma<<<grid, UMMA_THREADS, p.smem_bytes, stream>>>(d, p, in, wp, pmap, bias, add, out);
    return check_launch("tapconv_fwd_tf32");
}

}  // namespace kgan
