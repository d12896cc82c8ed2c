// Weight gradient of the tap convolution with TMA-fed operands (tcgen05.mma kind::tf32, fp32 accumulators in TMEM).
//
//     dW[tap][oc][ic] = sum_{n, p} gout[n, oc, p] * in[n, (tap, ic), p + shift_tap]          (zero outside the input plane)
//
// Same GEMM as tapconv_wgrad_umma.cu - D[M = 128 output channels][N = n_ic input channels] per tap, K = positions, split-K over
// one wave of CTAs - but for descriptors in shift form (tma_mode 1) whose output plane is a multiple of 32 positions the
// operands never pass through a register or a per-thread copy instruction: positions are contiguous per channel (NCHW), so a
// [channels] x [32 positions] box of a 3-D tensor map over (position, sample, channel) IS a K-major operand tile with
// 128-byte rows, which the TMA unit writes in the canonical SWIZZLE_128B layout (8-row atoms of 1024 bytes).  One elected
// thread issues 1 + ntap tensor loads per pipeline stage; a temporal tap is a shift of the position coordinate and the TMA
// unit zero-fills what falls outside the plane.  (The cp.async producers of tapconv_wgrad_umma.cu issue ~1000 16-byte copies
// per stage and reached 25-35 % of HBM peak.)
//
// Both operands reach the tensor core as raw fp32 words; in tf32 mode they were stored tf32-rounded by their producers
// (common.cuh tf32_out), so the tensor core's 19-bit read is exact (see tapconv_wgrad_umma.cu).  Rows of a box beyond the channel tile read neighbouring channels (or zeros beyond the tensor);
// they only feed accumulator rows / columns the epilogue never stores.
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer / TMEM owner, warps 2-9 = epilogue.
#include <cuda.h>
#include <string.h>

#include "umma.cuh"

namespace kgan {

constexpr int WT_EPI_WARPS = 8;
constexpr int WT_THREADS = 32 * (2 + WT_EPI_WARPS);
constexpr int WT_KT = 32;                        // positions per K tile: one 128-byte swizzled row per channel
// KGAN_PREC_TF32X3: four splitter warps write, for every landed stage, the lo image (x - its upper 19 bits) of BOTH operands behind the
// stage, and the MMA warp issues g_lo * x + g * x_lo + g * x per K step (see tapconv_tma.cu): the small terms into an accumulator of their
// own (2 * ntap * n_ic <= 512 TMEM columns), and split-K chunks of at most WT_X3_CHUNK K tiles - the tensor core's fp32 accumulation
// truncates, so the error of a chain grows linearly with its length (tools/x3_accuracy.py); the chunks are summed by fp32 atomics
constexpr int WT_X3_CHUNK = 32;
constexpr int WT_SPLIT_WARPS = 4;
constexpr int WT_THREADS_X3 = WT_THREADS + 32 * WT_SPLIT_WARPS;

struct WgradTmaPlan {
    int n_ic, ic_tiles, oc_tiles, tmem_cols, stages;
    int a_rows;              // rows of the gout box (multiple of 8, <= 128)
    int a_bytes, b_bytes;    // per stage: A tile (a_rows rows of 128 bytes), one tap's B tile
    int pf_dist;             // L2 prefetch distance in K tiles (0: off)
    uint32_t pf_taps;        // taps whose box is prefetched (near-duplicates - temporal shifts of the same channels - are skipped)
    int tail_pad;            // bytes after the last stage that the last A tile's 16 KB read window may touch
    int p_box, nsub;         // a stage = nsub sub-tiles of p_box positions of ONE sample each (p_box = 32 / 16 / 8, nsub = 32 / p_box)
    int row_bytes;           // 4 * p_box = swizzle span of the sub-tiles (SWIZZLE_128B / _64B / _32B)
    int sub_bytes;           // (a_rows + ntap * n_ic) * row_bytes
    uint32_t desc_hi;        // upper descriptor word: SBO = 8 rows, version, layout type of the swizzle mode
    int kt_per_plane;        // p_out / p_box
    int64_t ktiles;          // stages in total: ceil(n * kt_per_plane / nsub)
    int nchunks;
    int64_t chunk;           // K tiles per split-K chunk
    int smem_bytes;
    int x3;                  // KGAN_PREC_TF32X3: a stage is followed by its lo image (stage_stride = 2 * stage bytes)
};

int tma_encode_3d_f32(CUtensorMap* map, const float* base, const uint64_t gdim[3], const uint64_t gstr_bytes[2], const uint32_t box[3], int swizzle);

static bool make_wgrad_tma_plan(const kgan_tapconv_desc& d, WgradTmaPlan& p) {
    if (d.tma_mode != 1 || d.w_oc_blk != 0 || d.ntap > 8) return false;
    // A K tile is p_box positions of one sample, p_box = the largest of 32 / 16 / 8 that divides the plane, in the swizzle mode whose
    // span equals the box row (128 / 64 / 32 bytes); a stage holds 32 / p_box such sub-tiles.  (Boxes spanning several samples to fill
    // a 128-byte row do not work: with a swizzle span wider than the box's innermost extent the TMA unit faults on B200 - measured
    // with tools/probe_wgrad_tma.py for 16 x 2, 8 x 4 and 4 x 8.)
    if ((d.p_out & 7) || (d.p_in & 3)) return false;
    p.p_box = (d.p_out % 32) == 0 ? 32 : (d.p_out % 16) == 0 ? 16 : 8;
    p.nsub = WT_KT / p.p_box;
    p.row_bytes = 4 * p.p_box;
    p.desc_hi = (uint32_t)((8 * p.row_bytes) >> 4) | (1u << 14) | ((p.p_box == 32 ? 2u : p.p_box == 16 ? 4u : 6u) << 29);
    for (int t = 0; t < d.ntap; ++t)
        if (d.tap_shift[t] & 3) return false;                    // box origins must be 16-byte aligned
    const int64_t total = (int64_t)d.n * d.p_out;
    if (total < 1024 || total >= (1ll << 31) - WT_KT) return false;
    p.x3 = d.precision == KGAN_PREC_TF32X3 ? 1 : 0;
    int n_max = ((p.x3 ? 256 : 512) / d.ntap) / 16 * 16;
    if (n_max > 256) n_max = 256;
    if ((d.ntap >= 3 || p.x3) && n_max > 128) n_max = 128;
    if (n_max < 16) return false;
    p.n_ic = round_up(d.ck, 16) < n_max ? round_up(d.ck, 16) : n_max;
    p.ic_tiles = ceil_div(d.ck, p.n_ic);
    p.oc_tiles = ceil_div(d.co, UM);
    p.tmem_cols = 32;
    while (p.tmem_cols < (p.x3 ? 2 : 1) * d.ntap * p.n_ic) p.tmem_cols *= 2;
    if (p.tmem_cols > 512) return false;
    p.a_rows = d.co >= UM ? UM : round_up(d.co, 8);
    // The A tile only holds the a_rows rows the box writes; the M = 128 MMA reads on into the B tiles behind it (rows that feed
    // accumulator lanes nobody stores).  The kernel is bound by the bytes in flight per SM (loaded HBM latency ~3 us), so a smaller
    // stage means a deeper ring.  `tail_pad` keeps the last stage's 16 KB read window inside the allocation.
    p.a_bytes = p.a_rows * p.row_bytes;
    p.b_bytes = p.n_ic * p.row_bytes;
    p.sub_bytes = p.a_bytes + d.ntap * p.b_bytes;
    // thin layers (few channels: a 32-position K tile is only 10-16 KB): 2 or 4 K tiles per stage, up to 32 KB - the per-stage work
    // (barrier round trips, expect_tx, commit) is per stage, not per byte
    if (!p.x3)
        while (p.nsub * p.p_box < 128 && 2 * p.nsub * p.sub_bytes <= 32 * 1024) p.nsub *= 2;
    const int stage = p.nsub * p.sub_bytes * (p.x3 ? 2 : 1);
    const int tail_pad = p.sub_bytes < UM * p.row_bytes ? UM * p.row_bytes - p.sub_bytes : 0;
    p.tail_pad = tail_pad;
    p.stages = (212 * 1024 - tail_pad) / stage;
    if (p.stages > 16) p.stages = 16;
    if (p.stages < 2) return false;
    p.kt_per_plane = d.p_out / p.p_box;
    p.ktiles = ceil_div64((int64_t)d.n * p.kt_per_plane, p.nsub);
    const int tiles = p.ic_tiles * p.oc_tiles * d.groups;
    int64_t nchunks = kNumSMs / tiles;                           // one wave
    if (nchunks > p.ktiles / 4) nchunks = p.ktiles / 4;
    if (nchunks < 1) nchunks = 1;
    if (p.x3 && nchunks < ceil_div64(p.ktiles, WT_X3_CHUNK)) nchunks = ceil_div64(p.ktiles, WT_X3_CHUNK);
    p.chunk = ceil_div64(p.ktiles, nchunks);
    p.nchunks = (int)ceil_div64(p.ktiles, p.chunk);
    if ((int64_t)p.nchunks * d.groups > 65535) return false;
    p.smem_bytes = p.stages * stage + tail_pad + 1024 + 512;    // + alignment slack + barriers
    // L2 prefetch: ~256 KB of unique operand bytes ahead of the ring
    constexpr int pf_env = -1;         // compile-time switch (measured slower, see below)
    p.pf_taps = 0;
    int uniq = 0;
    for (int t = 0; t < d.ntap; ++t) {
        bool dup = false;
        for (int u = 0; u < t; ++u)
            if (((p.pf_taps >> u) & 1) && d.tap_in_ch[u] == d.tap_in_ch[t] && abs(d.tap_shift[u] - d.tap_shift[t]) < WT_KT) dup = true;
        if (!dup) {
            p.pf_taps |= 1u << t;
            ++uniq;
        }
    }
    const int per_tile = (p.a_rows + uniq * p.n_ic) * 128;
    if (p.p_box != 32) p.pf_taps = 0, uniq = 0;                 // (the prefetch path only knows whole-row tiles; it is off anyway)
    p.pf_dist = p.stages + (256 * 1024) / per_tile;
    if (p.pf_dist > 64) p.pf_dist = 64;
    // measured on B200 (profiles/r1_layer_bench_tma_prefetch_ab.txt): the prefetch makes every layer 10-50 % SLOWER - off unless asked for
    p.pf_dist = (pf_env > 0 && p.p_box == 32) ? p.stages + pf_env : 0;
    return true;
}

__device__ __forceinline__ void wt_tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}
// K-major operand in a swizzled layout whose span equals the row (128 / 64 / 32 bytes): 8-row atoms (SBO = 8 rows), LBO unused.
// `hi` = upper word from the plan: SBO >> 4 | version 1 (bit 46) | layout type (2 / 4 / 6 at bits 61-63).
__device__ __forceinline__ uint64_t smem_desc_k_sw(uint32_t addr, uint32_t hi) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)hi << 32);
}

template <bool X3>
__global__ void __launch_bounds__(X3 ? WT_THREADS_X3 : WT_THREADS, 1) tapconv_wgrad_tma_k(const __grid_constant__ kgan_tapconv_desc d, const __grid_constant__ WgradTmaPlan pl,
                                                                     const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_x,
                                                                     float* __restrict__ dw) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
    const int S = pl.stages;
    const int half_bytes = pl.nsub * pl.sub_bytes;                    // the operand boxes of a stage
    const int stage_bytes = X3 ? 2 * half_bytes : half_bytes;         // X3: + their lo images behind them
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes + pl.tail_pad);   // full[S], empty[S], accfull, (tmem slot), lofull[S]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S), accfull = smem_u32(bars + 2 * S);
    const uint32_t lofull0 = smem_u32(bars + 2 * S + 2);

    const int ic0 = blockIdx.x * pl.n_ic, oc0 = blockIdx.y * UM;
    const int g = blockIdx.z / pl.nchunks, ch = blockIdx.z % pl.nchunks;
    const int64_t kbeg = (int64_t)ch * pl.chunk, kend = min(pl.ktiles, kbeg + pl.chunk);
    const int iters = (int)(kend - kbeg);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
            if (X3) mbar_init(lofull0 + 8 * s, WT_SPLIT_WARPS);
        }
        mbar_init(accfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_g)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(pl.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // Producers: warp 0 and - until the accumulators are complete they have nothing else to do - the first two epilogue warps deal
    // the stages round-robin (a single producer warp was the bottleneck of the TMA-fed forward kernel: ~780 cycles of dependent
    // uniform-datapath instructions per stage, see tapconv_tma.cu); each stage still has exactly one producer.
    const int n_prod = S >= 3 ? 3 : 1;
    const int prod_idx = warp == 0 ? 0 : (n_prod == 3 && (warp == 2 || warp == 3)) ? warp - 1 : -1;
    if (prod_idx >= 0) {
        // ===== producer: one gout box + ntap input boxes per K tile (whole warp in uniform control flow, one lane issues) =====
        {
            const bool leader = elect_one();
            const int in_ch0 = g * d.g_in + ic0, out_ch0 = g * d.g_out + oc0;
            const uint32_t stage_tx = (uint32_t)half_bytes;
            // first K tile of this producer: stage k covers K tiles k * nsub ... + nsub - 1, K tile t = (sample t / kt_per_plane, box t % ...)
            const int64_t t0 = (kbeg + prod_idx) * pl.nsub;
            int nn = (int)(t0 / pl.kt_per_plane), pt = (int)(t0 - (int64_t)nn * pl.kt_per_plane);
            auto prefetch = [&](int64_t k) {                            // operands of stage k -> L2 (whole-row tiles only)
                const int pn = (int)(k / pl.kt_per_plane), pp = (int)(k - (int64_t)pn * pl.kt_per_plane) * pl.p_box;
                tma_prefetch_3d(&map_g, pp, pn, out_ch0);
                for (int tap = 0; tap < d.ntap; ++tap)
                    if ((pl.pf_taps >> tap) & 1) tma_prefetch_3d(&map_x, pp + d.tap_shift[tap], pn, in_ch0 + d.tap_in_ch[tap]);
            };
            if (pl.pf_dist > 0 && leader && prod_idx == 0)
                for (int64_t k = kbeg + S; k < kbeg + pl.pf_dist && k < kend; ++k) prefetch(k);
            int s = prod_idx;
            uint32_t ph = 1;                                            // parity to wait for on empty[s]
            for (int it = prod_idx; it < iters; it += n_prod) {
                mbar_wait(empty0 + 8 * s, ph);
                if (leader) {
                    if (pl.pf_dist > 0 && kbeg + it + pl.pf_dist < kend) prefetch(kbeg + it + pl.pf_dist);
                    mbar_arrive_expect_tx(full0 + 8 * s, stage_tx);
                    const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
                    int sn = nn, sp = pt;
                    for (int sub = 0; sub < pl.nsub; ++sub) {           // samples past the end (sn >= n) are zero-filled by the TMA unit
                        const uint32_t sb = st + sub * pl.sub_bytes;
                        const int p0 = sp * pl.p_box;
                        wt_tma_load_3d(sb, &map_g, p0, sn, out_ch0, full0 + 8 * s);
                        for (int tap = 0; tap < d.ntap; ++tap)
                            wt_tma_load_3d(sb + pl.a_bytes + tap * pl.b_bytes, &map_x, p0 + d.tap_shift[tap], sn, in_ch0 + d.tap_in_ch[tap], full0 + 8 * s);
                        if (++sp == pl.kt_per_plane) {
                            sp = 0;
                            ++sn;
                        }
                    }
                }
                __syncwarp();
                pt += n_prod * pl.nsub;
                while (pt >= pl.kt_per_plane) {
                    pt -= pl.kt_per_plane;
                    ++nn;
                }
                s += n_prod;
                if (s >= S) {
                    s -= S;
                    ph ^= 1u;
                }
            }
        }
    }
    if (warp == 0) {
        // producer only
    } else if (X3 && warp >= 2 + WT_EPI_WARPS) {
        // ===== splitters (X3): lo image of both operands of every landed stage (elementwise: any box layout) =====
        const uint32_t t16 = (uint32_t)(threadIdx.x - 32 * (2 + WT_EPI_WARPS)) * 16u;
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(full0 + 8 * s, ph);
            const uint32_t src = smem_u32(smem + (size_t)s * stage_bytes);
            for (uint32_t off = t16; off < (uint32_t)half_bytes; off += 4 * 2048) {
                uint32_t v[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (off + q * 2048 < (uint32_t)half_bytes)
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[q][0]), "=r"(v[q][1]), "=r"(v[q][2]), "=r"(v[q][3]) : "r"(src + off + q * 2048));
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (off + q * 2048 < (uint32_t)half_bytes) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[q][e] = __float_as_uint(__uint_as_float(v[q][e]) - __uint_as_float(v[q][e] & 0xFFFFE000u));
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(src + half_bytes + off + q * 2048), "r"(v[q][0]), "r"(v[q][1]), "r"(v[q][2]),
                                     "r"(v[q][3])
                                     : "memory");
                    }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(lofull0 + 8 * s);
            if (++s == S) {
                s = 0;
                ph ^= 1u;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp waits, one lane issues) =====
        {
            const bool leader = elect_one();
            // the taps' B tiles are adjacent in shared memory and their accumulators adjacent in TMEM: when ntap * n_ic <= 256 all taps
            // are ONE MMA of N = ntap * n_ic per K step (3x fewer instructions for the single issuing thread)
            const int nmma = d.ntap * pl.n_ic <= 256 ? 1 : d.ntap;
            const uint32_t idesc = instr_desc_tf32(nmma == 1 ? d.ntap * pl.n_ic : pl.n_ic);
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < iters; ++it) {
                mbar_wait(full0 + 8 * s, ph);
                if (X3) mbar_wait(lofull0 + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (leader) {
                    const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
                    for (int sub = 0; sub < pl.nsub; ++sub) {
                        const uint32_t a_addr = st + sub * pl.sub_bytes;
                        for (int tap = 0; tap < nmma; ++tap) {
                            const uint32_t b_addr = a_addr + pl.a_bytes + tap * pl.b_bytes;
                            for (int j = 0; j < pl.p_box / 8; ++j) {        // 32 bytes of K per MMA inside the swizzled row
                                const uint64_t ad = smem_desc_k_sw(a_addr + j * 32, pl.desc_hi), bd = smem_desc_k_sw(b_addr + j * 32, pl.desc_hi);
                                const uint32_t first = (it > 0 || sub > 0 || j > 0) ? 1u : 0u;
                                if (X3) {
                                    const uint64_t ald = smem_desc_k_sw(a_addr + half_bytes + j * 32, pl.desc_hi);
                                    const uint64_t bld = smem_desc_k_sw(b_addr + half_bytes + j * 32, pl.desc_hi);
                                    const uint32_t small = tmem_base + (d.ntap + tap) * pl.n_ic;      // the small terms' accumulator
                                    umma_tf32(small, ald, bd, idesc, first);
                                    umma_tf32(small, ad, bld, idesc, 1u);
                                    umma_tf32(tmem_base + tap * pl.n_ic, ad, bd, idesc, first);
                                } else {
                                    umma_tf32(tmem_base + tap * pl.n_ic, ad, bd, idesc, first);
                                }
                            }
                        }
                    }
                    umma_commit(empty0 + 8 * s);
                }
                __syncwarp();
                if (++s == S) {
                    s = 0;
                    ph ^= 1u;
                }
            }
            if (leader) umma_commit(accfull);
            __syncwarp();
        }
    } else {
        // ===== epilogue: TMEM lane = output channel row (warp % 4 selects the lane quarter, the other bit the column half) =====
        mbar_wait(accfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3, colhalf = (warp - 2) >> 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int oc = oc0 + quarter * 32 + lane;
        float* wb = dw + (int64_t)g * d.g_w + (int64_t)oc * d.w_oc;
        bool taps_inner = d.ntap == 3 && d.w_ic == 3;
        for (int tp = 0; tp < d.ntap; ++tp) taps_inner = taps_inner && d.tap_w_off[tp] == d.tap_w_off[0] + tp;
        auto fix = [](uint32_t bits) { return __uint_as_float(bits); };
        if (iters > 0) {
            if (taps_inner) {
                for (int col0 = colhalf * 16; col0 < pl.n_ic; col0 += 32) {
                    if (ic0 + col0 >= d.ck) break;                       // warp-uniform
                    uint32_t r[3][16];
                    tmem_ld16_nowait(taddr + col0, r[0]);
                    tmem_ld16_nowait(taddr + pl.n_ic + col0, r[1]);
                    tmem_ld16_nowait(taddr + 2 * pl.n_ic + col0, r[2]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (X3) {
#pragma unroll
                        for (int tp = 0; tp < 3; ++tp) {
                            uint32_t r2[16];
                            tmem_ld16(taddr + (3 + tp) * pl.n_ic + col0, r2);
#pragma unroll
                            for (int j = 0; j < 16; ++j) r[tp][j] = __float_as_uint(__uint_as_float(r[tp][j]) + __uint_as_float(r2[j]));
                        }
                    }
                    if (oc < d.co) {
                        float* dst = wb + d.tap_w_off[0] + (int64_t)(ic0 + col0) * 3;
                        if (ic0 + col0 + 16 <= d.ck && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                            for (int v4 = 0; v4 < 12; ++v4) {
                                float4 q;
                                q.x = fix(r[(4 * v4 + 0) % 3][(4 * v4 + 0) / 3]);
                                q.y = fix(r[(4 * v4 + 1) % 3][(4 * v4 + 1) / 3]);
                                q.z = fix(r[(4 * v4 + 2) % 3][(4 * v4 + 2) / 3]);
                                q.w = fix(r[(4 * v4 + 3) % 3][(4 * v4 + 3) / 3]);
                                atomicAdd(reinterpret_cast<float4*>(dst) + v4, q);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (ic0 + col0 + j < d.ck) {
#pragma unroll
                                    for (int tp = 0; tp < 3; ++tp) atomicAdd(dst + j * 3 + tp, fix(r[tp][j]));
                                }
                        }
                    }
                }
            } else {
                for (int tap = 0; tap < d.ntap; ++tap) {
                    for (int col0 = colhalf * 16; col0 < pl.n_ic; col0 += 32) {
                        if (ic0 + col0 >= d.ck) break;                   // warp-uniform
                        uint32_t r[16];
                        tmem_ld16(taddr + tap * pl.n_ic + col0, r);
                        if (X3) {
                            uint32_t r2[16];
                            tmem_ld16(taddr + (d.ntap + tap) * pl.n_ic + col0, r2);
#pragma unroll
                            for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
                        }
                        if (oc < d.co) {
                            float* dst = wb + d.tap_w_off[tap] + (int64_t)(ic0 + col0) * d.w_ic;
                            if (d.w_ic == 1 && ic0 + col0 + 16 <= d.ck && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                                for (int v4 = 0; v4 < 4; ++v4)
                                    atomicAdd(reinterpret_cast<float4*>(dst) + v4,
                                              make_float4(fix(r[4 * v4]), fix(r[4 * v4 + 1]), fix(r[4 * v4 + 2]), fix(r[4 * v4 + 3])));
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (ic0 + col0 + j < d.ck) atomicAdd(dst + (int64_t)j * d.w_ic, fix(r[j]));
                            }
                        }
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(pl.tmem_cols) : "memory");
    }
}

int tapconv_wgrad_tma_eligible(const kgan_tapconv_desc& d) {
    WgradTmaPlan p;
    return make_wgrad_tma_plan(d, p) ? 1 : 0;
}

// -1: not eligible (the caller falls back to the cp.async kernel)
int tapconv_wgrad_tma(const kgan_tapconv_desc& d, const float* in, const float* gout, float* dw, int64_t dw_numel, int accumulate, cudaStream_t stream) {
    WgradTmaPlan p;
    if (!make_wgrad_tma_plan(d, p)) return -1;
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(gout)) & 15) return -1;
    CUtensorMap map_g, map_x;
    {
        const uint64_t gdim[3] = {(uint64_t)d.p_out, (uint64_t)d.n, (uint64_t)d.c_out_total};
        const uint64_t gstr[2] = {(uint64_t)d.c_out_total * d.p_out * 4, (uint64_t)d.p_out * 4};
        const uint32_t box[3] = {(uint32_t)p.p_box, 1u, (uint32_t)p.a_rows};
        if (int e = tma_encode_3d_f32(&map_g, gout, gdim, gstr, box, p.p_box == 32 ? 1 : p.p_box == 16 ? 2 : 3)) return e;
    }
    {
        const uint64_t gdim[3] = {(uint64_t)d.p_in, (uint64_t)d.n, (uint64_t)d.c_in_total};
        const uint64_t gstr[2] = {(uint64_t)d.c_in_total * d.p_in * 4, (uint64_t)d.p_in * 4};
        const uint32_t box[3] = {(uint32_t)p.p_box, 1u, (uint32_t)p.n_ic};
        if (int e = tma_encode_3d_f32(&map_x, in, gdim, gstr, box, p.p_box == 32 ? 1 : p.p_box == 16 ? 2 : 3)) return e;
    }
    static SmemAttrOnce attr, attr3;
    if (int e = p.x3 ? ensure_smem(tapconv_wgrad_tma_k<true>, 227 * 1024, attr3, "tapconv_wgrad_tma (x3) attribute")
                     : ensure_smem(tapconv_wgrad_tma_k<false>, 227 * 1024, attr, "tapconv_wgrad_tma attribute"))
        return e;
    if (!accumulate && cudaMemsetAsync(dw, 0, sizeof(float) * dw_numel, stream) != cudaSuccess) return check_launch("tapconv_wgrad_tma memset");
    dim3 grid(p.ic_tiles, p.oc_tiles, (unsigned)(d.groups * p.nchunks));
    if (p.x3) tapconv_wgrad_tma_k<true><<<grid, WT_THREADS_X3, p.smem_bytes, stream>>>(d, p, map_g, map_x, dw);
    else tapconv_wgrad_tma_k<false><<<grid, WT_THREADS, p.smem_bytes, stream>>>(d, p, map_g, map_x, dw);
    return check_launch("tapconv_wgrad_tma");
}

}  // namespace kgan
