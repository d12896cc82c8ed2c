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
//   warps 10-17 epilogue: drain accumulator b of tile i (TMEM -> bias/add/act -> global) while tile i+1 is computed
//   smem ring full[S]/empty[S] (producers+loader <-> MMA), tmem_full[2]/tmem_empty[2] (MMA <-> epilogue).
// The layer is HBM/L2-bound at tensor-core rates (AI ~ 0.75*C_out flop/B unfused): what matters is bytes in flight.
#include "umma.cuh"

namespace kgan {

constexpr int FW_PRODUCER_WARPS = 8;
constexpr int FW_MMA_WARP = 8, FW_LOAD_WARP = 9, FW_EPI_WARP0 = 10;
constexpr int FW_EPI_WARPS = 8;                // two warps per TMEM lane quarter, alternating 16-column steps
constexpr int FW_THREADS = 32 * (10 + FW_EPI_WARPS);
constexpr int FW_KH = UK / 2;                  // channels per producer thread per stage
constexpr int FW_PF = 2;                       // stages of gathers in flight per producer thread

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

// Channel tiling shared by every tensor-core forward kernel (this file and tapconv_tma.cu): it fixes the layout of the
// packed weight image, so it depends on the descriptor only.
bool tapconv_umma_nsplit(const kgan_tapconv_desc& d, int* n_cta, int* n_split, int* n_rows, int* tmem_cols, int* nkt) {
    const int64_t total = (int64_t)d.n * d.p_out;                // any channel count: ragged K and N are zero padded
    if (total < 256 || total >= (1ll << 31) - UM) return false;
    const int64_t m_tiles = ceil_div64(total, UM);
    if (m_tiles * d.groups > (1 << 24)) return false;
    const int n16 = round_up(d.co, 16);
    // KGAN_PREC_TF32X3: tiles of at most 128 channels - a stage holds hi + lo images of both operands, and the accumulator of the
    // small terms (lo * W_hi + x * W_lo) sits beside the main one in TMEM, both double-buffered: 4 * n_cta <= 512 columns
    const bool x3 = d.precision == KGAN_PREC_TF32X3;
    int split = ceil_div(n16, x3 ? 128 : 256);
    if (m_tiles * d.groups * split < kNumSMs) {
        // fewer tiles than SMs: split the channels further - every tile streams its 128 rows of activations plus its n_cta rows of
        // weights per K step, and the launch takes ceil(tiles / SMs) waves of them: take the split that minimises waves * (128 + n_cta)
        // (the mapping network's 4096 x 632 x 632 layers: 4 splits of 160 channels = 128 tiles in one wave, not 6 x 112 = 192 in two)
        int best = split;
        int64_t best_cost = INT64_MAX;
        for (int s2 = split; s2 <= 16 * split; ++s2) {
            const int nc = round_up(ceil_div(n16, s2), 16);
            if (nc < 64 && s2 > split) break;
            const int64_t tiles = m_tiles * d.groups * ceil_div(n16, nc);
            const int64_t cost = ceil_div64(tiles, kNumSMs) * (UM + nc);
            if (cost < best_cost) {
                best_cost = cost;
                best = s2;
            }
        }
        split = best;
    }
    *n_cta = round_up(ceil_div(n16, split), 16);
    *n_split = ceil_div(n16, *n_cta);
    *n_rows = *n_cta * *n_split;
    *tmem_cols = 32;
    while (*tmem_cols < (x3 ? 4 : 2) * *n_cta) *tmem_cols *= 2;
    *nkt = ceil_div(d.ck, UK);
    return true;
}

static bool make_plan(const kgan_tapconv_desc& d, UmmaPlan& p) {
    if (!tapconv_umma_nsplit(d, &p.n_cta, &p.n_split, &p.n_rows, &p.tmem_cols, &p.nkt)) return false;
    const int64_t m_tiles = ceil_div64((int64_t)d.n * d.p_out, UM);
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
// KGAN_PREC_TF32X3: a second image of the same layout follows the first - hi = RN_tf32(w) in the first, RN_tf32(w - hi) in the second.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pack_value(float v, bool lo_half) {
    const float hi = __uint_as_float(to_tf32(v));
    return lo_half ? __uint_as_float(to_tf32(v - hi)) : hi;
}
__global__ void __launch_bounds__(256) tapconv_pack_k(const __grid_constant__ kgan_tapconv_desc d, const float* __restrict__ w,
                                                       float* __restrict__ wp, int n_rows, int nkt) {
    const int64_t total1 = (int64_t)d.groups * nkt * d.ntap * 8 * n_rows * 4;
    const int64_t total = d.precision == KGAN_PREC_TF32X3 ? 2 * total1 : total1;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += (int64_t)gridDim.x * blockDim.x) {
        const bool lo_half = i0 >= total1;
        const int64_t i = lo_half ? i0 - total1 : i0;
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
        if (row < d.co && ic < d.ck) v = __ldg(w + (int64_t)g * d.g_w + d.tap_w_off[tap] + w_oc_offset(d, row) + (int64_t)ic * d.w_ic);
        wp[i0] = pack_value(v, lo_half);
    }
}

// Batched variant: one launch re-packs every cached weight of a network after the optimizer step (blockIdx.y = item).
struct PackItem {
    kgan_tapconv_desc d;
    const float* w;
    float* wp;
    int32_t n_rows, nkt;
    int64_t total;                             // elements of ONE image (KGAN_PREC_TF32X3 writes two)
};
__global__ void __launch_bounds__(256) tapconv_pack_batched_k(const PackItem* __restrict__ items) {
    const PackItem& it = items[blockIdx.y];
    const kgan_tapconv_desc& d = it.d;
    const int n_rows = it.n_rows, nkt = it.nkt;
    const int64_t total = d.precision == KGAN_PREC_TF32X3 ? 2 * it.total : it.total;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += (int64_t)gridDim.x * blockDim.x) {
        const bool lo_half = i0 >= it.total;
        const int64_t i = lo_half ? i0 - it.total : i0;
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
        if (row < d.co && ic < d.ck) v = __ldg(it.w + (int64_t)g * d.g_w + d.tap_w_off[tap] + w_oc_offset(d, row) + (int64_t)ic * d.w_ic);
        it.wp[i0] = pack_value(v, lo_half);
    }
}

// tile id -> (position tile, group, channel split); consecutive ids share the activation tile, so the channel splits and
// the groups of one position tile (data gradient of the graph conv: 3 groups re-read the same gout) hit in L2
struct TileCoord {
    int g, mt, ns;
};
__device__ __forceinline__ TileCoord tile_coord(int tile, const UmmaPlan& pl) {
    TileCoord c;
    const int per_mt = pl.num_tiles / pl.m_tiles;      // groups * n_split
    c.mt = tile / per_mt;
    const int r = tile - c.mt * per_mt;
    c.g = r / pl.n_split;
    c.ns = r - c.g * pl.n_split;
    return c;
}

// Drain one accumulator (this thread's TMEM lane = one output position) to global memory, 16 channels per step:
// residual values are requested before the TMEM wait, the bias chunk is loaded once per step (lane j holds channel j)
// and broadcast by shuffle, the activation is a template parameter, addresses advance by one plane per channel.
template <int ACT>
__device__ __forceinline__ void epilogue_tile(uint32_t taddr, int ncols, int colpar, bool valid, float* __restrict__ op, int p_out,
                                              const float* __restrict__ ap, int64_t astride, const float* __restrict__ bp, int lane, int rnd) {
    for (int col0 = 16 * colpar; col0 < ncols; col0 += 16 * (FW_EPI_WARPS / 4)) {
        const int nc = min(16, ncols - col0);                         // warp-uniform
        float av[16];
        if (ap) {
#pragma unroll
            for (int j = 0; j < 16; ++j) av[j] = ldg_pred(ap + (int64_t)(col0 + j) * astride, valid && j < nc);
        }
        const float bl = (bp && lane < nc) ? __ldg(bp + col0 + lane) : 0.f;
        uint32_t r[16];
        tmem_ld16(taddr + col0, r);
        float* o = op + (int64_t)col0 * p_out;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float val = __uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bl, j);
            if (ap) val += av[j];
            if (ACT == KGAN_ACT_LRELU) val = val > 0.f ? val : 0.2f * val;
            if (ACT == KGAN_ACT_TANH) val = tanhf(val);
            if (valid && j < nc) *o = tf32_out(val, rnd);
            o += p_out;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FW_THREADS, 1) tapconv_fwd_umma(const __grid_constant__ kgan_tapconv_desc d, const __grid_constant__ UmmaPlan pl,
                                                                  const float* __restrict__ in, const float* __restrict__ wp,
                                                                  const int32_t* __restrict__ pmap, const float* __restrict__ bias,
                                                                  const float* __restrict__ add, float* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
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
            mbar_init(tempty0 + 8 * b, 32 * FW_EPI_WARPS);           // epilogue threads
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
        // ===== activation producers: warp = 16-byte k-chunk (4 channels), lane q = tile rows q, q+32, q+64, q+96 =====
        // 16 elements per thread per stage: loads are coalesced along positions (lanes = consecutive rows), the four
        // 16-byte shared stores of a thread go to consecutive rows across lanes (conflict-free), and addressing is one
        // 64-bit row pointer per (tap, row) plus a 32-bit channel offset.
        const int q = lane, chunk = warp;
        int kit = 0;                                                  // ring position, continues across tiles
        for (int tile = blockIdx.x; tile < pl.num_tiles; tile += gridDim.x) {
            const TileCoord tc = tile_coord(tile, pl);
            constexpr int SRC_CACHE = 3;                              // taps whose position map is kept in registers
            const float* rowp[SRC_CACHE][4];                          // in + sample/tap offset + source position (or null)
            int pp[4];
            const float* in_n[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t pos = (uint32_t)tc.mt * UM + q + 32 * i;     // make_plan guarantees total_pos < 2^31
                const bool valid = pos < (uint32_t)total_pos;
                const uint32_t nn = valid ? pos / (uint32_t)d.p_out : 0u;
                pp[i] = valid ? (int)(pos - nn * (uint32_t)d.p_out) : -1;
                in_n[i] = in + ((int64_t)nn * d.c_in_total + tc.g * d.g_in) * d.p_in;
#pragma unroll
                for (int tp = 0; tp < SRC_CACHE; ++tp) {
                    int src = -1;
                    if (tp < d.ntap && valid) src = __ldg(pmap + (int64_t)d.tap_row[tp] * d.p_out + pp[i]);
                    rowp[tp][i] = src >= 0 ? in_n[i] + (int64_t)d.tap_in_ch[tp < d.ntap ? tp : 0] * d.p_in + src : nullptr;
                }
            }
            auto issue = [&](int it, float (&v)[16]) {
                const int ict = it / d.ntap, tap = it - ict * d.ntap;
                const int ch0 = ict * UK + chunk * 4;                 // first of this thread's 4 channels
                const float* rp[4];
                if (tap < SRC_CACHE) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        rp[i] = rowp[0][i];
#pragma unroll
                        for (int tp = 1; tp < SRC_CACHE; ++tp) rp[i] = (tp == tap) ? rowp[tp][i] : rp[i];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int src = pp[i] >= 0 ? __ldg(pmap + (int64_t)d.tap_row[tap] * d.p_out + pp[i]) : -1;
                        rp[i] = src >= 0 ? in_n[i] + (int64_t)d.tap_in_ch[tap] * d.p_in + src : nullptr;
                    }
                }
                const bool all = rp[0] && rp[1] && rp[2] && rp[3] && ch0 + 3 < d.ck;
                if (all) {                                            // common case: no predication at all
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const unsigned off = (unsigned)(ch0 + j) * (unsigned)d.p_in;
#pragma unroll
                        for (int i = 0; i < 4; ++i) v[4 * i + j] = ldg_nc(rp[i] + off);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const unsigned off = (unsigned)(ch0 + j) * (unsigned)d.p_in;
#pragma unroll
                        for (int i = 0; i < 4; ++i) v[4 * i + j] = ldg_pred(rp[i] + off, rp[i] != nullptr && ch0 + j < d.ck);
                    }
                }
            };
            auto stage_out = [&](int it, float (&v)[16]) {
                const int k = kit + it, s = k % S;
                const uint32_t ph = (uint32_t)(k / S) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);                   // slot free (first lap passes immediately)
                const uint32_t dst = smem_u32(a_base + (size_t)s * A_STAGE_BYTES) + chunk * A_LBO + q * 16;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + i * 32 * 16), "r"(to_tf32_fast(v[4 * i])),
                                 "r"(to_tf32_fast(v[4 * i + 1])), "r"(to_tf32_fast(v[4 * i + 2])), "r"(to_tf32_fast(v[4 * i + 3]))
                                 : "memory");
                mbar_arrive(full0 + 8 * s);                           // release; the MMA thread runs the proxy fence (see there)
            };
            // FW_PF stages of gathers are always in flight: issue(it + PF) follows stage_out(it) on the same registers
            float v0[16], v1[16];
            static_assert(FW_PF == 2 && UK == 32 && FW_PRODUCER_WARPS == 8, "producer mapping below assumes 8 chunk-warps, PF = 2");
            if (0 < kiters) issue(0, v0);
            if (1 < kiters) issue(1, v1);
            for (int it = 0; it < kiters; it += 2) {
                stage_out(it, v0);
                if (it + 2 < kiters) issue(it + 2, v0);
                if (it + 1 < kiters) {
                    stage_out(it + 1, v1);
                    if (it + 3 < kiters) issue(it + 3, v1);
                }
            }
            kit += kiters;
        }
    } else if (warp == FW_MMA_WARP) {
        // ===== MMA issuer: the whole warp waits in uniform control flow, one elected lane drives the tensor core (umma.cuh elect_one) =====
        {
            const bool leader = elect_one();
            const uint32_t idesc = instr_desc_tf32(pl.n_cta);
            const uint32_t b_lbo = pl.n_cta * 16;
            int s = 0, ti = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < pl.num_tiles; tile += gridDim.x, ++ti) {
                const int buf = ti & 1;
                mbar_wait(tempty0 + 8 * buf, ((uint32_t)(ti >> 1) & 1u) ^ 1u);     // epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + buf * pl.n_cta;
                for (int it = 0; it < kiters; ++it) {
                    mbar_wait(full0 + 8 * s, ph);                     // acquire: the producers' st.shared are visible
                    // generic-proxy writes (observed through the barrier) -> async proxy (tcgen05.mma operand reads).  The fence
                    // sits on the consumer side: a producer-side fence would also wait for that thread's in-flight gathers of
                    // the following stages and serialise the pipeline.
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (leader) {
                        const uint32_t a_addr = smem_u32(a_base + (size_t)s * A_STAGE_BYTES);
                        const uint32_t b_addr = smem_u32(b_base + (size_t)s * b_stage_bytes);
#pragma unroll
                        for (int j = 0; j < UK / 8; ++j)
                            umma_tf32(acc, smem_desc(a_addr + j * 2 * A_LBO, A_LBO, CORE_SBO), smem_desc(b_addr + j * 2 * b_lbo, b_lbo, CORE_SBO),
                                      idesc, (it > 0 || j > 0) ? 1u : 0u);
                        umma_commit(empty0 + 8 * s);                  // frees the stage when these MMAs retire
                    }
                    __syncwarp();
                    if (++s == S) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                if (leader) umma_commit(tfull0 + 8 * buf);            // accumulator complete -> epilogue
                __syncwarp();
            }
        }
    } else if (warp == FW_LOAD_WARP) {
        // ===== weight loader: bulk copies of the packed tf32 image (one elected lane issues) =====
        {
            const bool leader = elect_one();
            const uint32_t chunk_bytes = pl.n_cta * 16;
            int s = 0;
            uint32_t ph = 1;
            for (int tile = blockIdx.x; tile < pl.num_tiles; tile += gridDim.x) {
                const TileCoord tc = tile_coord(tile, pl);
                const float* wg = wp + (int64_t)tc.g * pl.nkt * d.ntap * pl.n_rows * UK;
                const int oc_base = tc.ns * pl.n_cta;
                for (int it = 0; it < kiters; ++it) {
                    mbar_wait(empty0 + 8 * s, ph);
                    if (leader) {
                        mbar_arrive_expect_tx(full0 + 8 * s, chunk_bytes * 8);
                        const float* src = wg + (int64_t)it * pl.n_rows * UK;          // loop order == packing order (ic tile, tap)
                        const uint32_t dst = smem_u32(b_base + (size_t)s * b_stage_bytes);
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            bulk_g2s(dst + c * chunk_bytes, src + ((int64_t)c * pl.n_rows + oc_base) * 4, chunk_bytes, full0 + 8 * s);
                    }
                    __syncwarp();
                    if (++s == S) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane = position row; 32 columns per step =====
        const int quarter = warp & 3;                                 // warps 10..17 -> TMEM lane quarters 2,3,0,1,2,3,0,1
        const int colpar = (warp - FW_EPI_WARP0) >> 2;                // which of the alternating 16-column steps
        const int row = quarter * 32 + lane;
        int ti = 0;
        for (int tile = blockIdx.x; tile < pl.num_tiles; tile += gridDim.x, ++ti) {
            const TileCoord tc = tile_coord(tile, pl);
            const int buf = ti & 1;
            const uint32_t pos = (uint32_t)tc.mt * UM + row;          // make_plan guarantees total_pos < 2^31
            const bool valid = pos < (uint32_t)total_pos;
            const uint32_t nn = valid ? pos / (uint32_t)d.p_out : 0u, p = valid ? pos - nn * (uint32_t)d.p_out : 0u;
            const int out_ch0 = tc.g * d.g_out, oc_base = tc.ns * pl.n_cta;
            const int pst = out_plane(d);
            const uint32_t pg = p + (uint32_t)(tc.g * d.g_pout);      // position inside the output plane (position-block groups)
            float* op = out + ((int64_t)nn * d.c_out_total + out_ch0 + oc_base) * pst + pg;
            const int64_t astride = d.add_period ? d.add_period : pst;
            const float* ap = add ? add + ((int64_t)nn * d.c_out_total + out_ch0 + oc_base) * astride + (d.add_period ? pg % (uint32_t)d.add_period : pg)
                                  : nullptr;
            const float* bp = bias ? bias + out_ch0 + oc_base : nullptr;
            mbar_wait(tfull0 + 8 * buf, (uint32_t)(ti >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * pl.n_cta;
            const int ncols = min(pl.n_cta, d.co - oc_base);          // columns of this tile that exist
            if (d.act == KGAN_ACT_LRELU) epilogue_tile<KGAN_ACT_LRELU>(taddr, ncols, colpar, valid, op, pst, ap, astride, bp, lane, d.precision == KGAN_PREC_TF32);
            else if (d.act == KGAN_ACT_TANH) epilogue_tile<KGAN_ACT_TANH>(taddr, ncols, colpar, valid, op, pst, ap, astride, bp, lane, d.precision == KGAN_PREC_TF32);
            else epilogue_tile<KGAN_ACT_NONE>(taddr, ncols, colpar, valid, op, pst, ap, astride, bp, lane, d.precision == KGAN_PREC_TF32);
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
    return (int64_t)d.groups * p.nkt * d.ntap * p.n_rows * UK * (d.precision == KGAN_PREC_TF32X3 ? 2 : 1);
}

int tapconv_pack_tf32(const kgan_tapconv_desc& d, const float* w, float* wp, cudaStream_t stream) {
    UmmaPlan p;
    if (!make_plan(d, p)) {
        set_error("tapconv_pack: shape not eligible for the tf32 path");
        return 1;
    }
    const int64_t total = (int64_t)d.groups * p.nkt * d.ntap * p.n_rows * UK * (d.precision == KGAN_PREC_TF32X3 ? 2 : 1);
    int64_t blocks = ceil_div64(total, 256);
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    tapconv_pack_k<<<(unsigned)blocks, 256, 0, stream>>>(d, w, wp, p.n_rows, p.nkt);
    return check_launch("tapconv_pack");
}

int64_t tapconv_pack_item_bytes() { return (int64_t)sizeof(PackItem); }

int tapconv_pack_tf32_batched(int count, const kgan_tapconv_desc* descs, const float* const* w, float* const* wp, void* items_dev, int upload,
                              cudaStream_t stream) {
    if (upload) {
        PackItem* host = new PackItem[count];
        for (int i = 0; i < count; ++i) {
            UmmaPlan p;
            if (!make_plan(descs[i], p)) {
                delete[] host;
                set_error("tapconv_pack_batched: item %d not eligible for the tf32 path", i);
                return 1;
            }
            host[i].d = descs[i];
            host[i].w = w[i];
            host[i].wp = wp[i];
            host[i].n_rows = p.n_rows;
            host[i].nkt = p.nkt;
            host[i].total = (int64_t)descs[i].groups * p.nkt * descs[i].ntap * p.n_rows * UK;
        }
        // pageable source: the copy is staged by the runtime before the call returns, so `host` may be freed right away
        const cudaError_t e = cudaMemcpyAsync(items_dev, host, sizeof(PackItem) * count, cudaMemcpyHostToDevice, stream);
        delete[] host;
        if (e != cudaSuccess) return check_launch("tapconv_pack_batched upload");
    }
    tapconv_pack_batched_k<<<dim3(32, (unsigned)count), 256, 0, stream>>>(static_cast<const PackItem*>(items_dev));
    return check_launch("tapconv_pack_batched");
}

int tapconv_fwd_tf32(const kgan_tapconv_desc& d, const float* in, const float* wp, const int32_t* pmap, const float* bias,
                     const float* add, float* out, cudaStream_t stream) {
    UmmaPlan p;
    if (!make_plan(d, p)) return -1;
    static SmemAttrOnce attr;
    if (int e = ensure_smem(tapconv_fwd_umma, 227 * 1024, attr, "tapconv_fwd_tf32 attribute")) return e;
    const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;         // persistent: one CTA per SM
    tapconv_fwd_umma<<<grid, FW_THREADS, p.smem_bytes, stream>>>(d, p, in, wp, pmap, bias, add, out);
    return check_launch("tapconv_fwd_tf32");
}

}  // namespace kgan
