// Tap convolution on the 5th-generation tensor cores (KGAN_PREC_TF32): tcgen05.mma kind::tf32, fp32 accumulators
// in TMEM, weights staged by the bulk-copy engine (cp.async.bulk -> UBLKCP) from a pre-packed tf32 image,
// activations gathered by SIMT producer warps (position map = zero padding / joint & frame selection / tap shift).
//
// GEMM orientation (see DESIGN.md "tcgen05 tap convolution"):
//   D[M = 128 positions][N = output channels]  +=  A[M][K] * B[N][K]^T,   K = (input-channel tile, tap)
//   A (activations): gathered from NCHW global memory, written K-major into shared memory in the canonical no-swizzle
//       core-matrix layout (8 rows x 16 B), rounded to tf32 (cvt.rna).
//   B (weights):     pre-packed by tapconv_pack_k in exactly the shared-memory image, so a stage is 8 bulk copies.
//   D: TMEM lane = position, TMEM column = output channel -> the epilogue thread that owns lane l stores
//       out[n, oc, p(l)]: for a fixed oc consecutive lanes are consecutive addresses (coalesced NCHW stores).
//
// Persistent, warp-specialised CTA (one per SM), three pipelines:
//   warps 0-7   activation producers: thread = (row, k-half); PF stages of gathers in flight per thread
//   warp  8     MMA issuer (one thread) + TMEM owner; accumulators are double-buffered in TMEM (2 x n_cta columns)
//   warp  9     weight loader (one thread, bulk copies)
//   warps 10-13 epilogue: drain accumulator b of tile i (TMEM -> bias/add/act -> global) while tile i+1 is computed
//   smem ring full[S]/empty[S] (producers+loader <-> MMA), tmem_full[2]/tmem_empty[2] (MMA <-> epilogue).
// The layer is HBM/L2-bound at tensor-core rates (AI ~ 0.75*C_out flop/B unfused): what matters is bytes in flight.
#include "umma.cuh"

namespace kgan {

constexpr int FW_PRODUCER_WARPS = 8;
constexpr int FW_MMA_WARP = 8, FW_LOAD_WARP = 9, FW_EPI_WARP0 = 10;
constexpr int FW_THREADS = 32 * 14;
constexpr int FW_KH = UK / 2;                  // channels per producer thread per stage
constexpr int FW_PF = 3;                       // stages of gathers in flight per producer thread

struct UmmaPlan {
    int n_cta;        // output channels per tile (UMMA N, multiple of 16, <= 256)
    int n_split;      // tiles along output channels
    int n_rows;       // n_cta * n_split: rows of the packed weight image (zero padded)
    int tmem_cols;    // power of two >= 2 * n_cta
    int stages;
    int nkt;          // input-channel tiles of UK
    int m_tiles;      // position tiles of 128
    int num_tiles;    // m_tiles * n_split * groups
    int smem_bytes;
};

static bool make_plan(const kgan_tapconv_desc& d, UmmaPlan& p) {
    if (d.ck < 16 || d.co < 16 || d.w_oc_blk != 0) return false;
    const int64_t total = (int64_t)d.n * d.p_out;
    if (total < 256) return false;
    const int64_t m_tiles = ceil_div64(total, UM);
    if (m_tiles * d.groups > (1 << 24)) return false;
    const int n16 = round_up(d.co, 16);
    int split = ceil_div(n16, 256);
    while (m_tiles * d.groups * split < kNumSMs && ceil_div(n16, split * 2) >= 64) split *= 2;
    p.n_cta = round_up(ceil_div(n16, split), 16);
    p.n_split = ceil_div(n16, p.n_cta);
    p.n_rows = p.n_cta * p.n_split;
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * p.n_cta) p.tmem_cols *= 2;
    p.nkt = ceil_div(d.ck, UK);
    p.m_tiles = (int)m_tiles;
    p.num_tiles = (int)m_tiles * p.n_split * d.groups;
    const int stage = A_STAGE_BYTES + p.n_cta * UK * 4;
    p.stages = (200 * 1024) / stage;
    if (p.stages > 8) p.stages = 8;
    p.smem_bytes = p.stages * stage + 512;
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

// tile id -> (group, position tile, channel split); consecutive ids share the activation tile (L2 reuse)
struct TileCoord {
    int g, mt, ns;
};
__device__ __forceinline__ TileCoord tile_coord(int tile, const UmmaPlan& pl) {
    TileCoord c;
    c.ns = tile % pl.n_split;
    const int r = tile / pl.n_split;
    c.mt = r % pl.m_tiles;
    c.g = r / pl.m_tiles;
    return c;
}

// ---------------------------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FW_THREADS, 1) tapconv_fwd_umma(const __grid_constant__ kgan_tapconv_desc d, const __grid_constant__ UmmaPlan pl,
                                                                  const float* __restrict__ in, const float* __restrict__ wp,
                                                                  const int32_t* __restrict__ pmap, const float* __restrict__ bias,
                                                                  const float* __restrict__ add, float* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = pl.stages;
    const int b_stage_bytes = pl.n_cta * UK * 4;
    uint8_t* a_base = smem;
    uint8_t* b_base = smem + (size_t)S * A_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + (size_t)S * b_stage_bytes);   // full[S], empty[S], tfull[2], tempty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);
    const uint32_t tfull0 = smem_u32(bars + 2 * S), tempty0 = smem_u32(bars + 2 * S + 2);
    const int64_t total_pos = (int64_t)d.n * d.p_out;
    const int kiters = pl.nkt * d.ntap;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full0 + 8 * s, 32 * FW_PRODUCER_WARPS + 1);    // producer threads + the weight loader's expect_tx arrive
            mbar_init(empty0 + 8 * s, 1);                            // tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull0 + 8 * b, 1);                            // tcgen05.commit after the last MMA of a tile
            mbar_init(tempty0 + 8 * b, 128);                         // epilogue threads
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == FW_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(pl.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < FW_PRODUCER_WARPS) {
        // ===== activation producers: thread = (tile row, k half): FW_KH channels of one position per stage =====
        const int row = threadIdx.x & (UM - 1), kh = threadIdx.x >> 7;
        int kit = 0;                                                  // ring position, continues across tiles
        for (int tile = blockIdx.x; tile < pl.num_tiles; tile += gridDim.x) {
            const TileCoord tc = tile_coord(tile, pl);
            const int64_t pos = (int64_t)tc.mt * UM + row;
            const bool valid = pos < total_pos;
            const int nn = valid ? (int)(pos / d.p_out) : 0, p = valid ? (int)(pos % d.p_out) : 0;
            const float* in_n = in + ((int64_t)nn * d.c_in_total + tc.g * d.g_in) * d.p_in;
            constexpr int SRC_CACHE = 4;                              // position map of this row for the first taps
            int srcs[SRC_CACHE];
#pragma unroll
            for (int i = 0; i < SRC_CACHE; ++i) srcs[i] = (valid && i < d.ntap) ? __ldg(pmap + (int64_t)d.tap_row[i] * d.p_out + p) : -1;

            auto issue = [&](int it, float (&v)[FW_KH]) {
                const int ict = it / d.ntap, tap = it - ict * d.ntap, ic0 = ict * UK + kh * FW_KH;
                int src = -1;
                if (tap < SRC_CACHE) {
#pragma unroll
                    for (int i = 0; i < SRC_CACHE; ++i) src = (i == tap) ? srcs[i] : src;
                } else if (valid) {
                    src = __ldg(pmap + (int64_t)d.tap_row[tap] * d.p_out + p);
                }
                const float* xb = in_n + (int64_t)(d.tap_in_ch[tap] + ic0) * d.p_in + src;
#pragma unroll
                for (int kk = 0; kk < FW_KH; ++kk) v[kk] = ldg_pred(xb + (int64_t)kk * d.p_in, src >= 0 && ic0 + kk < d.ck);
            };
            auto stage_out = [&](int it, float (&v)[FW_KH]) {
                const int k = kit + it, s = k % S;
                const uint32_t ph = (uint32_t)(k / S) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);                   // slot free (first lap passes immediately)
                const uint32_t dst = smem_u32(a_base + (size_t)s * A_STAGE_BYTES) + kh * (FW_KH / 4) * A_LBO + row * 16;
#pragma unroll
                for (int c = 0; c < FW_KH / 4; ++c)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + c * A_LBO), "r"(to_tf32(v[4 * c])),
                                 "r"(to_tf32(v[4 * c + 1])), "r"(to_tf32(v[4 * c + 2])), "r"(to_tf32(v[4 * c + 3]))
                                 : "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
                mbar_arrive(full0 + 8 * s);
            };
            // FW_PF stages of gathers are always in flight: issue(it + PF) follows stage_out(it) on the same registers
            float v0[FW_KH], v1[FW_KH], v2[FW_KH];
            static_assert(FW_PF == 3, "register ring below is written for PF = 3");
            if (0 < kiters) issue(0, v0);
            if (1 < kiters) issue(1, v1);
            if (2 < kiters) issue(2, v2);
            for (int it = 0; it < kiters; it += 3) {
                stage_out(it, v0);
                if (it + 3 < kiters) issue(it + 3, v0);
                if (it + 1 < kiters) {
                    stage_out(it + 1, v1);
                    if (it + 4 < kiters) issue(it + 4, v1);
                }
                if (it + 2 < kiters) {
                    stage_out(it + 2, v2);
                    if (it + 5 < kiters) issue(it + 5, v2);
                }
            }
            kit += kiters;
        }
    } else if (warp == FW_MMA_WARP) {
        // ===== MMA issuer: one thread drives the tensor core =====
        if (lane == 0) {
            const uint32_t idesc = instr_desc_tf32(pl.n_cta);
            const uint32_t b_lbo = pl.n_cta * 16;
            int kit = 0, ti = 0;
            for (int tile = blockIdx.x; tile < pl.num_tiles; tile += gridDim.x, ++ti) {
                const int buf = ti & 1;
                mbar_wait(tempty0 + 8 * buf, ((uint32_t)(ti >> 1) & 1u) ^ 1u);     // epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + buf * pl.n_cta;
                for (int it = 0; it < kiters; ++it) {
                    const int k = kit + it, s = k % S;
                    const uint32_t ph = (uint32_t)(k / S) & 1u;
                    mbar_wait(full0 + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_u32(a_base + (size_t)s * A_STAGE_BYTES);
                    const uint32_t b_addr = smem_u32(b_base + (size_t)s * b_stage_bytes);
#pragma unroll
                    for (int j = 0; j < UK / 8; ++j)
                        umma_tf32(acc, smem_desc(a_addr + j * 2 * A_LBO, A_LBO, CORE_SBO), smem_desc(b_addr + j * 2 * b_lbo, b_lbo, CORE_SBO),
                                  idesc, (it > 0 || j > 0) ? 1u : 0u);
                    umma_commit(empty0 + 8 * s);                      // frees the stage when these MMAs retire
                }
                umma_commit(tfull0 + 8 * buf);                        // accumulator complete -> epilogue
                kit += kiters;
            }
        }
    } else if (warp == FW_LOAD_WARP) {
        // ===== weight loader: bulk copies of the packed tf32 image =====
        if (lane == 0) {
            const uint32_t chunk_bytes = pl.n_cta * 16;
            int kit = 0;
            for (int tile = blockIdx.x; tile < pl.num_tiles; tile += gridDim.x) {
                const TileCoord tc = tile_coord(tile, pl);
                const float* wg = wp + (int64_t)tc.g * pl.nkt * d.ntap * pl.n_rows * UK;
                const int oc_base = tc.ns * pl.n_cta;
                for (int it = 0; it < kiters; ++it) {
                    const int k = kit + it, s = k % S;
                    const uint32_t ph = (uint32_t)(k / S) & 1u;
                    mbar_wait(empty0 + 8 * s, ph ^ 1u);
                    mbar_arrive_expect_tx(full0 + 8 * s, chunk_bytes * 8);
                    const float* src = wg + (int64_t)it * pl.n_rows * UK;          // loop order == packing order (ic tile, tap)
                    const uint32_t dst = smem_u32(b_base + (size_t)s * b_stage_bytes);
#pragma unroll
                    for (int c = 0; c < 8; ++c) bulk_g2s(dst + c * chunk_bytes, src + ((int64_t)c * pl.n_rows + oc_base) * 4, chunk_bytes, full0 + 8 * s);
                }
                kit += kiters;
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane = position row; 32 columns per step, residual loads issued before the TMEM wait =====
        const int quarter = warp & 3;                                 // warps 10..13 -> TMEM lane quarters 2,3,0,1
        const int row = quarter * 32 + lane;
        int ti = 0;
        for (int tile = blockIdx.x; tile < pl.num_tiles; tile += gridDim.x, ++ti) {
            const TileCoord tc = tile_coord(tile, pl);
            const int buf = ti & 1;
            const int64_t pos = (int64_t)tc.mt * UM + row;
            const bool valid = pos < total_pos;
            const int nn = valid ? (int)(pos / d.p_out) : 0, p = valid ? (int)(pos % d.p_out) : 0;
            const int out_ch0 = tc.g * d.g_out, oc_base = tc.ns * pl.n_cta;
            const int64_t obase = ((int64_t)nn * d.c_out_total + out_ch0) * d.p_out + p;
            const int64_t abase = d.add_period ? ((int64_t)nn * d.c_out_total + out_ch0) * d.add_period + p % d.add_period : obase;
            const int64_t astride = d.add_period ? d.add_period : d.p_out;
            mbar_wait(tfull0 + 8 * buf, (uint32_t)(ti >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * pl.n_cta;
            for (int col0 = 0; col0 < pl.n_cta; col0 += 32) {
                if (oc_base + col0 >= d.co) break;                    // warp-uniform
                float av[32];
                if (add) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) av[j] = ldg_pred(add + abase + (int64_t)(oc_base + col0 + j) * astride, valid && oc_base + col0 + j < d.co);
                }
                uint32_t r0[16], r1[16];
                tmem_ld16_nowait(taddr + col0, r0);
                if (col0 + 16 < pl.n_cta) tmem_ld16_nowait(taddr + col0 + 16, r1);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int oc = oc_base + col0 + j;
                        if (oc < d.co && col0 + j < pl.n_cta) {
                            float val = __uint_as_float(j < 16 ? r0[j & 15] : r1[j & 15]);
                            if (bias) val += __ldg(bias + out_ch0 + oc);
                            if (add) val += av[j];
                            out[obase + (int64_t)oc * d.p_out] = apply_act(val, d.act);
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(tempty0 + 8 * buf);                           // accumulator may be overwritten
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == FW_MMA_WARP) {
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
    const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;         // persistent: one CTA per SM
    tapconv_fwd_umma<<<grid, FW_THREADS, p.smem_bytes, stream>>>(d, p, in, wp, pmap, bias, add, out);
    return check_launch("tapconv_fwd_tf32");
}

}  // namespace kgan
