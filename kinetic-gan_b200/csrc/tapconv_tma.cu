// Tap convolution fed by the Tensor Memory Accelerator (KGAN_PREC_TF32, descriptors with tma_mode != 0).
//
// Same GEMM as tapconv_umma.cu -  D[M = 128 positions][N = output channels] += A[M][K] * B[N][K]^T,  K = (channel tile, tap) -
// but the activation operand never touches a register: activations are NCHW, i.e. for a fixed channel the positions of a
// plane are contiguous, which is exactly an "MN-major" UMMA operand.  A 3-D tensor map over (position, sample, channel)
// lets one elected thread fetch a [32 channels] x [32 positions] box (4 KB, 128-byte rows, 32-byte swizzle atoms) per M group with a single
// cp.async.bulk.tensor; four boxes make the 128 x 32 A tile of a pipeline stage.  A temporal tap is a shift of the
// position coordinate by (tap - pad) * dilation * V; positions shifted outside the plane are zero-filled by the TMA unit,
// which is the convolution's zero padding.  Planes that are not a multiple of 32 positions (P = 80, 16, 8, 4) cannot be boxed
// that way (the 32-byte-atom swizzle needs 128-byte box rows); for them four extra warps write the same shared-memory
// image with 16-byte cp.async copies (p_box positions x n_box samples per 128-byte row, swizzle applied by hand) - still
// register-free, so the whole ring of stages is in flight.
// With 4-6 stages of 16 KB in flight per SM the kernel is bandwidth- instead of latency-bound (the SIMT-gather kernel
// holds at most 32 KB of loads in registers, and measured 8-20 % of HBM peak with long-scoreboard stalls dominating).
//
// Raw fp32 activation words reach the tensor core, which reads the upper 19 bits; the weights are pre-rounded (round to
// nearest) by tapconv_pack_k.  In tf32 mode every libkgan kernel stores its activations tf32-rounded (common.cuh tf32_out -
// this kernel's own epilogue included), so the 19-bit read is exact: products are RN(x) * RN(w), unbiased, and exact for
// tf32-representable data.  (Round 1 multiplied the accumulator by 1.000353 to undo the mean truncation bias of raw
// operands: a statistical fix that put a deterministic +3.5e-4 on exactly representable inputs.)
//
// The TMA unit requires the innermost (position) coordinate of a box to be a multiple of 16 bytes - an unaligned shift
// raises an illegal-instruction fault - so only descriptors whose shifts are multiples of 4 positions are eligible.  The host
// side arranges that: joints are padded to a multiple of 4 where that is cheap (11 -> 12), and temporally strided
// convolutions read a time-unfolded copy of their input whose taps sit at block offsets (see geometry.py).
//
// Warp roles: warp 0 = producer (one thread: tensor loads + bulk copies of the packed weights), warp 1 = MMA issuer and
// TMEM owner (accumulators double-buffered), warps 2-9 = epilogue (TMEM -> bias/add/act -> coalesced NCHW stores),
// warps 10-13 = cp.async activation producers (only launched when p_box < 32).
#include <cuda.h>
#include <string.h>

#include "umma.cuh"

namespace kgan {

constexpr int TM_EPI_WARPS = 8;
constexpr int TM_EPI_WARP0 = 2;
constexpr int TM_THREADS = 32 * (TM_EPI_WARP0 + TM_EPI_WARPS);       // TMA mode
constexpr int TM_PROD_WARPS = 3;                                       // TMA mode: warp 0 + warps 10, 11 issue the loads of every 3rd stage each
constexpr int TM_THREADS_TMA = TM_THREADS + 32 * (TM_PROD_WARPS - 1);
constexpr int TM_CP_WARPS = 4;
constexpr int TM_THREADS_CP = TM_THREADS + 32 * TM_CP_WARPS;          // cp.async mode
constexpr int TM_GROUP_BYTES = 32 * 32 * 4;    // one box: 32 channel rows of 128 bytes
// KGAN_PREC_TF32X3 (fp32-accurate mode): four more warps - the splitters - follow the copy engine through the ring; for every stage that
// has landed they write  lo = x - (x & 0xFFFFE000)  (what the tensor core's 19-bit read of x drops; exact in fp32) into a second operand
// image of the same layout, and the MMA warp issues three MMAs per K step: lo * W_hi + x * W_lo + x * W_hi (W_hi / W_lo: the two halves of
// the packed weight image).  The split is elementwise, so it is the same code for every operand layout (MN-major boxes, cp.async image,
// K-major Linear plan).  The two small terms go to an accumulator of their own, added by the epilogue: the tensor core's fp32 accumulation
// truncates (measured: the error of a K-step chain grows linearly, ~2e-8 per MMA - tools/x3_accuracy.py), so the main accumulator takes
// one MMA per K step instead of three, and the small accumulator's truncation is 2^-11 smaller.
constexpr int TM_SPLIT_WARPS = 4;
constexpr int TM_SPLIT_WARP0 = TM_EPI_WARP0 + TM_EPI_WARPS + TM_CP_WARPS;      // 14 (warps 12, 13 idle in TMA mode)
constexpr int TM_THREADS_X3 = 32 * (TM_SPLIT_WARP0 + TM_SPLIT_WARPS);

// Ring size.  A deeper ring does not help: with 10 stages (216 KB) the 3-tap D1 layer measured 337 us against 327 us with 8 - the kernel is
// not bound by its own bytes in flight but by what the L2 delivers for the taps' re-reads (profiles/r2_ncu_full_d1tcn_b4096.txt)
constexpr int TM_SMEM_BUDGET = 200 * 1024;
constexpr int TM_MAX_STAGES = 8;

struct TmaPlan {
    int n_cta, n_split, n_rows, tmem_cols, nkt;   // identical to UmmaPlan (the packed weight image is shared)
    int p_box, n_box, p_shift;                    // box = p_box positions x n_box samples; p_box = 1 << p_shift
    int cps;                                      // position chunks per output plane: p_out / p_box
    int64_t groups32;                             // 32-row M groups: ceil(n / n_box) * cps
    int m_tiles, num_tiles, stages, smem_bytes;
    int per_mt;                                   // num_tiles / m_tiles = groups * n_split
    int a_lbo, a_sbo;                             // A descriptor strides (bytes): between 32-position M groups / between 8-channel K groups
    int w_res;                                    // 1: the whole packed weight image stays resident in shared memory (loaded once per CTA)
    int w_res_bytes;
    int kmajor;                                   // 1: one-position planes (Linear layers): A is a K-major [128 samples] x [32 channels] box
    int nkt2;                                     // extra K panel (fused residual 1x1 conv, see tapconv_fwd_tma_res): channel tiles of the
    int c2_total, p_in2;                          //   second input tensor (0: none), its channel count and plane size
    int img1_bytes;                               // bytes of the main packed weight image (the panel's image follows it when resident)
    int x3;                                       // KGAN_PREC_TF32X3: lo image per activation stage, hi + lo weight images
    int64_t w_lo_off;                             //   floats between the hi and the lo half of the packed weight image
    int pf_tiles;                                 // L2 prefetch distance in tiles of this CTA (0: off)
    uint32_t pf_taps;                             // taps whose boxes are prefetched (temporal shifts of the same channels are near-duplicates)
};

bool tapconv_umma_nsplit(const kgan_tapconv_desc& d, int* n_cta, int* n_split, int* n_rows, int* tmem_cols, int* nkt);

static bool make_tma_plan(const kgan_tapconv_desc& d, TmaPlan& p, int panel_ck = 0) {
    if (d.tma_mode != 1) return false;
    p.x3 = d.precision == KGAN_PREC_TF32X3 ? 1 : 0;
    if (p.x3 && panel_ck > 0) return false;                        // the fused residual panel is a tf32-mode feature
    p.nkt2 = panel_ck > 0 ? ceil_div(panel_ck, UK) : 0;
    p.c2_total = panel_ck;
    p.p_in2 = d.p_out;
    // One-position planes (nn.Linear: the mapping network, the generator's first block at T = V = 1): the activations are an
    // (N, C) row-major matrix, i.e. a plain K-major operand - one [128 samples] x [32 channels] box per stage in the standard
    // SWIZZLE_128B layout.  (As an "MN-major" operand they would need a box along the sample axis, which is strided.)
    p.kmajor = (d.p_in == 1 && d.p_out == 1 && out_plane(d) == 1) ? 1 : 0;
    if (p.kmajor) {
        if (d.c_in_total & 3) return false;                        // row stride of the (N, C) matrix must be a multiple of 16 bytes
        for (int t = 0; t < d.ntap; ++t)
            if (d.tap_shift[t] != 0) return false;
    } else {
        if ((d.p_in & 3) || (d.p_out & 3)) return false;           // global strides / box origins must be multiples of 16 bytes
        for (int t = 0; t < d.ntap; ++t)
            if (d.tap_shift[t] & 3) return false;                  // unaligned box origin: the TMA unit faults
    }
    if (!tapconv_umma_nsplit(d, &p.n_cta, &p.n_split, &p.n_rows, &p.tmem_cols, &p.nkt)) return false;
    p.p_shift = 5;
    while (p.p_shift > 2 && (d.p_out & ((1 << p.p_shift) - 1))) --p.p_shift;
    if (p.kmajor) p.p_shift = 0;                                   // M group = 32 samples x 1 position (the epilogue's row -> (sample, position) rule)
    p.p_box = 1 << p.p_shift;
    p.n_box = 32 / p.p_box;
    p.cps = d.p_out / p.p_box;
    p.groups32 = ceil_div64(d.n, p.n_box) * p.cps;
    if (p.groups32 < 8 || p.groups32 >= (1ll << 28)) return false;
    p.m_tiles = (int)ceil_div64(p.groups32, 4);
    if ((int64_t)p.m_tiles * d.groups * p.n_split > (1 << 28)) return false;
    p.num_tiles = p.m_tiles * p.n_split * d.groups;
    p.per_mt = p.n_split * d.groups;
    const int xm = p.x3 ? 2 : 1;                                   // x3: every stage holds a lo image next to each operand image
    const int stage = xm * (A_STAGE_BYTES + p.n_cta * UK * 4);
    p.stages = TM_SMEM_BUDGET / stage;
    if (p.stages > TM_MAX_STAGES) p.stages = TM_MAX_STAGES;
    if (p.stages < 2) return false;
    p.smem_bytes = p.stages * stage + 1024 + 512;                  // + alignment slack + barriers
    p.w_lo_off = (int64_t)d.groups * p.nkt * d.ntap * p.n_rows * UK;
    // Resident weights: a CTA re-fetches the same packed weight stages for every one of its tiles - a third of the bytes an SM
    // ingests at N = 64 (the marginal cost of a K stage measured ~700 cycles for 16 KB of activations + 8 KB of weights).  When
    // the whole image (all groups, all K stages) fits beside a ring of >= 5 activation stages it is loaded once per CTA.
    p.w_res = 0;
    p.w_res_bytes = 0;
    {
        const int64_t img1 = (int64_t)d.groups * p.nkt * d.ntap * p.n_cta * UK * 4;
        const int64_t img = xm * img1 + (int64_t)p.nkt2 * p.n_cta * UK * 4;      // (x3: hi image, then the lo image - contiguous in the packed buffer too)
        p.img1_bytes = (int)img1;
        const int64_t tiles_per_cta = ceil_div64(p.num_tiles, kNumSMs);
        if (p.n_split == 1 && img <= 112 * 1024 && tiles_per_cta >= 2) {
            int st = (int)((TM_SMEM_BUDGET - img) / (xm * A_STAGE_BYTES));
            if (st > TM_MAX_STAGES) st = TM_MAX_STAGES;
            if (st >= (p.x3 ? 3 : 5)) {
                p.w_res = 1;
                p.w_res_bytes = (int)img;
                p.stages = st;
                p.smem_bytes = st * xm * A_STAGE_BYTES + (int)img + 1024 + 512;
            }
        }
    }
    p.a_lbo = TM_GROUP_BYTES;
    p.a_sbo = 512;
    // L2 prefetch (TMA mode only): ~256 KB of unique activation bytes ahead of the shared-memory ring
    constexpr int pf_env = -1;         // compile-time switch (measured slower, see below)
    p.pf_taps = 0;
    int uniq = 0;
    for (int t = 0; t < d.ntap; ++t) {
        bool dup = false;
        for (int u = 0; u < t; ++u)
            if (((p.pf_taps >> u) & 1) && d.tap_in_ch[u] == d.tap_in_ch[t] && abs(d.tap_shift[u] - d.tap_shift[t]) < 32) dup = true;
        if (!dup) {
            p.pf_taps |= 1u << t;
            ++uniq;
        }
    }
    p.pf_tiles = (256 * 1024) / (p.nkt * uniq * A_STAGE_BYTES);
    if (p.pf_tiles < 1) p.pf_tiles = 1;
    if (p.pf_tiles > 16) p.pf_tiles = 16;
    // measured on B200 (profiles/r1_layer_bench_tma_prefetch_ab.txt): the prefetch makes every layer 10-50 % SLOWER - off unless asked for
    p.pf_tiles = pf_env > 0 ? (pf_env > 16 ? 16 : pf_env) : 0;
    if (p.p_box != 32) p.pf_tiles = 0;
    return true;
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}
// MN-major 32-bit operand: the only layout the tensor core accepts is SWIZZLE_128B_BASE32B (layout type 1) - 128-byte rows of
// 32 consecutive M elements, 32-byte chunks XOR-swizzled with (row & 3), atoms of 4 K rows (512 bytes).  The TMA unit writes
// exactly this with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  LBO = bytes between 32-element M groups, SBO = bytes between
// 4-row K groups.
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
           (1ull << 61);
}

// K-major 32-bit operand, SWIZZLE_128B (layout type 2): 128-byte rows (32 K elements), 8-row atoms of 1024 bytes (SBO), LBO unused;
// a K step of 8 elements advances the start address by 32 bytes inside the atom.
__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

struct TmaTile {
    int g, mt, ns;
};
__device__ __forceinline__ TmaTile tma_tile(int tile, const TmaPlan& pl) {
    // (index arithmetic per tile is on the critical path of the small-channel layers - three K stages per tile: 32-bit, and none
    // at all in the common case of one tile per position tile)
    TmaTile c;
    if (pl.per_mt == 1) {
        c.mt = tile;
        c.g = c.ns = 0;
        return c;
    }
    c.mt = tile / pl.per_mt;                           // groups * n_split consecutive tiles re-use the activation tile in L2
    const int r = tile - c.mt * pl.per_mt;
    c.g = r / pl.n_split;
    c.ns = r - c.g * pl.n_split;
    return c;
}

// The residual / label term `add` and the bias do not depend on the accumulator: their loads for the first column chunk are issued
// BEFORE the wait on the accumulator barrier, so their latency hides behind the tile's main loop instead of following it.
template <int ACT, bool SCAT = false, bool X3 = false, bool KM = false>
__device__ __forceinline__ void tma_epilogue_tile(uint32_t taddr, int ncols, int colpar, bool valid, float* __restrict__ op, int p_out,
                                                  const float* __restrict__ ap, int64_t astride, const float* __restrict__ bp, int lane,
                                                  uint32_t tfull_bar, uint32_t tfull_parity, int rnd, const float* __restrict__ bp2 = nullptr,
                                                  int d1 = 0, int dz = 0, int small_off = 0) {
    bool waited = false;
    for (int col0 = 16 * colpar; col0 < ncols; col0 += 16 * (TM_EPI_WARPS / 4)) {
        const int nc = min(16, ncols - col0);                         // warp-uniform
        float av[16];
        if (ap) {
#pragma unroll
            for (int j = 0; j < 16; ++j) av[j] = ldg_pred(ap + (int64_t)(col0 + j) * astride, valid && j < nc);
        }
        float bl = (bp && lane < nc) ? __ldg(bp + col0 + lane) : 0.f;
        if (bp2 && lane < nc) bl += __ldg(bp2 + col0 + lane);        // the fused residual conv's bias
        if (!waited) {
            mbar_wait(tfull_bar, tfull_parity);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            waited = true;
        }
        uint32_t r[16];
        tmem_ld16(taddr + col0, r);
        if (X3) {                                                  // + the accumulator of the small terms (small_off columns further)
            uint32_t r2[16];
            tmem_ld16(taddr + small_off + col0, r2);
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
        }
        float* o = op + (int64_t)col0 * p_out;
        // one-position planes (Linear layers): a thread's 16 columns are 64 contiguous bytes of its sample's row - four 16-byte stores
        // instead of 16 scalar ones that each touch their own 32-byte sector (lanes are samples, c_out_total floats apart)
        // (KM: instantiated for the K-major plan only - the hot epilogues of the small-channel layers stay as they were)
        const bool rowvec = KM && !SCAT && p_out == 1 && nc == 16 && (reinterpret_cast<uintptr_t>(o) & 15) == 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float val = __uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bl, j);
            if (ap) val += av[j];
            if (ACT == KGAN_ACT_LRELU) val = val > 0.f ? val : 0.2f * val;
            if (ACT == KGAN_ACT_TANH) val = tanhf(val);
            if (rowvec) {
                r[j] = __float_as_uint(tf32_out(val, rnd));
                if ((j & 3) == 3 && valid)
                    *reinterpret_cast<float4*>(o + j - 3) = make_float4(__uint_as_float(r[j - 3]), __uint_as_float(r[j - 2]), __uint_as_float(r[j - 1]), __uint_as_float(r[j]));
            } else if (valid && j < nc) {
                const float q = tf32_out(val, rnd);
                *o = q;
                if (SCAT) {                                        // scatter store (kgan_tapconv_fwd_tf32_scatter): second copy of this position
                    if (d1 != 0) o[d1] = q;                        // and a slot of the output plane this position keeps zero (offsets relative to *o,
                    if (dz != 0) o[dz] = 0.f;                      // 0 = none); compiled out of the ordinary launches (two predicated stores per element cost
                }                                                  // 5-25 % on the small-channel layers, whose epilogue is on the critical path)
            }
            if (!rowvec) o += p_out;
        }
    }
}

__device__ __forceinline__ void tm_cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

// KM: the instantiation launched for the K-major (Linear-layer) plan - its epilogue stores row segments (tma_epilogue_tile<.., KM>).  A
// separate kernel because the extra epilogue variants cost the main instantiation 8 % (100 -> 128 registers, measured on the whole step).
template <bool X3, bool KM>
__global__ void __launch_bounds__(X3 ? TM_THREADS_X3 : TM_THREADS_CP, 1) tapconv_fwd_tma_k(const __grid_constant__ kgan_tapconv_desc d, const __grid_constant__ TmaPlan pl,
                                                                    const __grid_constant__ CUtensorMap tmap, const float* __restrict__ wp,
                                                                    const float* __restrict__ in, const float* __restrict__ bias,
                                                                    const float* __restrict__ add, float* __restrict__ out,
                                                                    const __grid_constant__ CUtensorMap tmap2, const float* __restrict__ wp2,
                                                                    const float* __restrict__ in2, const float* __restrict__ bias2,
                                                                    const int32_t* __restrict__ omap) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // swizzle atoms: 1024-byte aligned
    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
    const int S = pl.stages;
    const int b_stage_bytes = pl.n_cta * UK * 4;
    constexpr int XM = X3 ? 2 : 1;
    uint8_t* a_base = smem;
    uint8_t* alo_base = smem + (size_t)S * A_STAGE_BYTES;           // X3: the lo images of the S activation stages
    uint8_t* b_base = smem + (size_t)XM * S * A_STAGE_BYTES;        // ring of S weight stages (X3: hi, lo per stage), or the resident image (pl.w_res)
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + (pl.w_res ? (size_t)pl.w_res_bytes : (size_t)S * XM * b_stage_bytes));
    // full[S], empty[S], tfull[2], tempty[2], wfull, (tmem slot), lofull[S]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 5);
    const uint32_t lofull0 = smem_u32(bars + 2 * S + 6);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);
    const uint32_t tfull0 = smem_u32(bars + 2 * S), tempty0 = smem_u32(bars + 2 * S + 2);
    const uint32_t wfull = smem_u32(bars + 2 * S + 4);
    const int kmain = pl.nkt * d.ntap;                               // K steps of the convolution itself ...
    const int kiters = kmain + pl.nkt2;                              // ... + the extra panel: a 1x1 conv of a second tensor into the same accumulator
    const bool kmajor = X3 ? pl.kmajor != 0 : KM;                    // (a compile-time constant in the two tf32-mode instantiations)
    const bool use_tma = pl.p_box == 32 || kmajor;                   // else: cp.async producers (warps 10-13)

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            // the producer's arrive.expect_tx (the bulk / tensor copies complete the bytes) + in cp.async mode one deferred
            // arrival per producer thread
            mbar_init(full0 + 8 * s, use_tma ? 1 : 1 + 32 * TM_CP_WARPS);
            mbar_init(empty0 + 8 * s, 1);                            // tcgen05.commit
            if (X3) mbar_init(lofull0 + 8 * s, TM_SPLIT_WARPS);      // one arrival per splitter warp
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull0 + 8 * b, 1);
            mbar_init(tempty0 + 8 * b, 32 * TM_EPI_WARPS);
        }
        mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (use_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
        if (use_tma && pl.nkt2) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap2)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(pl.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // Producer warps.  ncu (profiles/r1_ncu_d1_stalls.txt) showed ONE producer warp to be the bottleneck of the 3-tap temporal conv:
    // it never waited on an empty stage, the MMA warp waited on full ones a third of its time, and a stage cost ~780 cycles of
    // mostly dependent uniform-datapath instructions (coordinates, descriptors, 4 UTMALDG).  In TMA mode the stages are
    // therefore dealt round-robin to TM_PROD_WARPS warps (each stage still has exactly one producer and its own barrier pair).
    const int n_prod = (use_tma && S >= TM_PROD_WARPS) ? TM_PROD_WARPS : 1;      // (a producer's first stage slot is its index: needs S >= n_prod)
    const int prod_idx = warp == 0 ? 0 : (n_prod > 1 && warp >= TM_EPI_WARP0 + TM_EPI_WARPS && warp < TM_EPI_WARP0 + TM_EPI_WARPS + TM_PROD_WARPS - 1)
                                             ? warp - (TM_EPI_WARP0 + TM_EPI_WARPS) + 1
                                             : -1;
    if (prod_idx >= 0) {
        // ===== producer: tensor loads of the activation boxes + bulk copies of the packed weights =====
        // (whole warp in uniform control flow, one elected lane issues: operands stay in uniform registers, see umma.cuh elect_one)
        {
            const bool leader = elect_one();
            const uint32_t chunk_bytes = pl.n_cta * 16;
            const uint32_t stage_tx = (use_tma ? A_STAGE_BYTES : 0) + (pl.w_res ? 0u : chunk_bytes * 8 * XM);
            if (pl.w_res && leader && prod_idx == 0) {               // the whole packed image (contiguous for n_split == 1), once
                mbar_arrive_expect_tx(wfull, (uint32_t)pl.w_res_bytes);
                for (int off = 0; off < pl.img1_bytes; off += 16384) {
                    const int nb = min(16384, pl.img1_bytes - off);
                    bulk_g2s(smem_u32(b_base + off), reinterpret_cast<const uint8_t*>(wp) + off, (uint32_t)nb, wfull);
                }
                for (int off = pl.img1_bytes; off < pl.w_res_bytes; off += 16384) {      // the panel's image (X3: the lo image) behind it
                    const int nb = min(16384, pl.w_res_bytes - off);
                    const uint8_t* src2 = X3 ? reinterpret_cast<const uint8_t*>(wp + pl.w_lo_off) : reinterpret_cast<const uint8_t*>(wp2);
                    bulk_g2s(smem_u32(b_base + off), src2 + (off - pl.img1_bytes), (uint32_t)nb, wfull);
                }
            }
            // activation boxes of a later tile of this CTA -> L2 (off by default: measured slower); tiles that share an activation tile
            // (other n splits / groups reading the same channels) leave the prefetch to the first of them
            auto prefetch_tile = [&](int t2) {
                if (t2 >= pl.num_tiles) return;
                const TmaTile pc = tma_tile(t2, pl);
                if (pc.ns != 0 || (pc.g != 0 && d.g_in == 0)) return;
                for (int i = 0; i < 4; ++i) {
                    const int G = pc.mt * 4 + i;                 // (< 2^30: make_tma_plan)
                    if (G >= pl.groups32) break;
                    const int nb = G / pl.cps;
                    const int pp = (G - nb * pl.cps) << pl.p_shift;
                    for (int ict = 0; ict < pl.nkt; ++ict)
                        for (int tap = 0; tap < d.ntap; ++tap)
                            if ((pl.pf_taps >> tap) & 1)
                                tma_prefetch_3d(&tmap, pp + d.tap_shift[tap], nb * pl.n_box, pc.g * d.g_in + d.tap_in_ch[tap] + ict * UK);
                }
            };
            if (use_tma && pl.pf_tiles > 0 && leader && prod_idx == 0)
                for (int j = 1; j < pl.pf_tiles; ++j) prefetch_tile(blockIdx.x + j * gridDim.x);
            int s = prod_idx;                                         // this warp's next stage slot (n_prod <= 5 <= S)
            uint32_t ph = 1;                                          // parity to wait for on empty[s]
            int it = prod_idx;                                        // ... and its index within the current tile's K loop
            for (int tile = blockIdx.x; tile < pl.num_tiles; tile += gridDim.x) {
                if (use_tma && pl.pf_tiles > 0 && leader && prod_idx == 0) prefetch_tile(tile + pl.pf_tiles * gridDim.x);
                const TmaTile tc = tma_tile(tile, pl);
                int cp[4], cn[4];                                     // box origin (position, sample) of the 4 M groups
                {                                                     // the four groups are consecutive: one division per tile
                    int nb = (tc.mt * 4) / pl.cps, rem = tc.mt * 4 - nb * pl.cps;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        cp[i] = rem << pl.p_shift;
                        cn[i] = nb * pl.n_box;                        // >= n for the groups past the end: zero-filled
                        if (++rem == pl.cps) {
                            rem = 0;
                            ++nb;
                        }
                    }
                }
                const float* wg = wp + (int64_t)tc.g * pl.nkt * d.ntap * pl.n_rows * UK;
                const int oc_base = tc.ns * pl.n_cta;
                const int ch_g = tc.g * d.g_in;
                int ict = it / d.ntap, tap = it - ict * d.ntap;
                for (; it < kiters; it += n_prod) {
                    mbar_wait(empty0 + 8 * s, ph);
                    if (leader) {
                        mbar_arrive_expect_tx(full0 + 8 * s, stage_tx);
                        const uint32_t a_dst = smem_u32(a_base + (size_t)s * A_STAGE_BYTES);
                        const bool panel = it >= kmain;                                 // extra K panel: second tensor, no shift
                        const int ch0 = panel ? (it - kmain) * UK : ch_g + d.tap_in_ch[tap] + ict * UK, sh = panel ? 0 : d.tap_shift[tap];
                        const CUtensorMap* tm = panel ? &tmap2 : &tmap;
                        if (kmajor) {
                            tma_load_3d(a_dst, tm, ch0, tc.mt * UM, 0, full0 + 8 * s);         // (channel, sample, -): rows past n / channels past C read zero
                        } else if (use_tma) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) tma_load_3d(a_dst + i * TM_GROUP_BYTES, tm, cp[i] + sh, cn[i], ch0, full0 + 8 * s);
                        }
                        const float* src = panel ? wp2 + (int64_t)(it - kmain) * pl.n_rows * UK
                                                 : wg + (int64_t)it * pl.n_rows * UK;    // loop order == packing order (ic tile, tap)
                        const uint32_t b_dst = smem_u32(b_base + (size_t)s * XM * b_stage_bytes);
#pragma unroll
                        for (int half = 0; half < XM; ++half) {                         // X3: the lo stage behind the hi stage
                            const float* sh_src = src + half * pl.w_lo_off;
                            const uint32_t bd = b_dst + half * b_stage_bytes;
                            if (pl.w_res) {
                                // nothing: the weights are resident
                            } else if (pl.n_split == 1) {
                                bulk_g2s(bd, sh_src, chunk_bytes * 8, full0 + 8 * s);   // the whole stage is contiguous in the image
                            } else {
#pragma unroll
                                for (int c = 0; c < 8; ++c)
                                    bulk_g2s(bd + c * chunk_bytes, sh_src + ((int64_t)c * pl.n_rows + oc_base) * 4, chunk_bytes, full0 + 8 * s);
                            }
                        }
                    }
                    __syncwarp();
                    tap += n_prod;
                    while (tap >= d.ntap) {
                        tap -= d.ntap;
                        ++ict;
                    }
                    s += n_prod;
                    if (s >= S) {
                        s -= S;
                        ph ^= 1u;
                    }
                }
                it -= kiters;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp waits, one elected lane issues) =====
        {
            const bool leader = elect_one();
            const uint32_t idesc = instr_desc_tf32(pl.n_cta) | (kmajor ? 0u : (1u << 15));      // A operand MN-major unless kmajor
            const uint32_t b_lbo = pl.n_cta * 16;
            int s = 0, ti = 0;
            uint32_t ph = 0;
            if (pl.w_res) mbar_wait(wfull, 0);
            for (int tile = blockIdx.x; tile < pl.num_tiles; tile += gridDim.x, ++ti) {
                const int buf = ti & 1;
                const uint32_t b_res = smem_u32(b_base) + (uint32_t)(tma_tile(tile, pl).g * kmain) * (uint32_t)b_stage_bytes;
                mbar_wait(tempty0 + 8 * buf, ((uint32_t)(ti >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + buf * XM * pl.n_cta;      // X3: [main | small terms] per buffer
                for (int it = 0; it < kiters; ++it) {
                    mbar_wait(full0 + 8 * s, ph);
                    if (X3) mbar_wait(lofull0 + 8 * s, ph);          // the splitters' lo image of this stage (they ran the proxy fence)
                    // TMA mode: both operands were written by the async proxy.  cp.async mode: the activations went through the
                    // generic proxy and must be made visible to the tensor core's async-proxy reads
                    if (!use_tma) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (leader) {
                        const uint32_t a_addr = smem_u32(a_base + (size_t)s * A_STAGE_BYTES);
                        const uint32_t b_addr = !pl.w_res ? smem_u32(b_base + (size_t)s * XM * b_stage_bytes)
                                                : it < kmain ? b_res + (uint32_t)it * (uint32_t)b_stage_bytes
                                                             : smem_u32(b_base) + (uint32_t)pl.img1_bytes + (uint32_t)(it - kmain) * (uint32_t)b_stage_bytes;
                        if (X3) {
                            // lo * W_hi + x * W_lo -> the small accumulator, x * W_hi -> the main one (the tensor core reads x as its upper 19 bits = hi)
                            const uint32_t alo_addr = smem_u32(alo_base + (size_t)s * A_STAGE_BYTES);
                            const uint32_t blo_addr = b_addr + (pl.w_res ? (uint32_t)pl.img1_bytes : (uint32_t)b_stage_bytes);
#pragma unroll
                            for (int j = 0; j < UK / 8; ++j) {
                                const uint64_t ad = kmajor ? smem_desc_k_sw128(a_addr + j * 32) : smem_desc_mn_sw128(a_addr + j * 1024, pl.a_lbo, pl.a_sbo);
                                const uint64_t ald = kmajor ? smem_desc_k_sw128(alo_addr + j * 32) : smem_desc_mn_sw128(alo_addr + j * 1024, pl.a_lbo, pl.a_sbo);
                                const uint64_t bd = smem_desc(b_addr + j * 2 * b_lbo, b_lbo, CORE_SBO), bld = smem_desc(blo_addr + j * 2 * b_lbo, b_lbo, CORE_SBO);
                                umma_tf32(acc + pl.n_cta, ald, bd, idesc, (it > 0 || j > 0) ? 1u : 0u);
                                umma_tf32(acc + pl.n_cta, ad, bld, idesc, 1u);
                                umma_tf32(acc, ad, bd, idesc, (it > 0 || j > 0) ? 1u : 0u);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < UK / 8; ++j)
                                umma_tf32(acc, kmajor ? smem_desc_k_sw128(a_addr + j * 32) : smem_desc_mn_sw128(a_addr + j * 1024, pl.a_lbo, pl.a_sbo),
                                          smem_desc(b_addr + j * 2 * b_lbo, b_lbo, CORE_SBO), idesc, (it > 0 || j > 0) ? 1u : 0u);
                        }
                        umma_commit(empty0 + 8 * s);
                    }
                    __syncwarp();
                    if (++s == S) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                if (leader) umma_commit(tfull0 + 8 * buf);
                __syncwarp();
            }
        }
    } else if (X3 && warp >= TM_SPLIT_WARP0) {
        // ===== splitters (X3): lo image of every landed activation stage; thread = 16-byte column of the 2 KB rows of the stage =====
        const uint32_t t16 = (uint32_t)(threadIdx.x - 32 * TM_SPLIT_WARP0) * 16u;
        int s = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < pl.num_tiles; tile += gridDim.x)
            for (int it = 0; it < kiters; ++it) {
                mbar_wait(full0 + 8 * s, ph);
                const uint32_t src = smem_u32(a_base + (size_t)s * A_STAGE_BYTES) + t16, dst = smem_u32(alo_base + (size_t)s * A_STAGE_BYTES) + t16;
                uint32_t v[8][4];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[q][0]), "=r"(v[q][1]), "=r"(v[q][2]), "=r"(v[q][3]) : "r"(src + q * 2048));
#pragma unroll
                for (int q = 0; q < 8; ++q) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[q][e] = __float_as_uint(__uint_as_float(v[q][e]) - __uint_as_float(v[q][e] & 0xFFFFE000u));
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + q * 2048), "r"(v[q][0]), "r"(v[q][1]), "r"(v[q][2]), "r"(v[q][3]) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(lofull0 + 8 * s);
                if (++s == S) {
                    s = 0;
                    ph ^= 1u;
                }
            }
    } else if (warp >= TM_EPI_WARP0 + TM_EPI_WARPS) {
        // ===== cp.async activation producers (p_box < 32): thread = (16-byte chunk of the 128-byte row, channel rows rr, rr + 16) =====
        // (warps 12, 13 of an X3 launch in TMA mode have no role)
        const int t = threadIdx.x - 32 * (TM_EPI_WARP0 + TM_EPI_WARPS);
        const int chunk = t & 7, rr = t >> 3;
        const int m0 = chunk * 4;                                     // first of this chunk's 4 M elements within the group
        const int nl = m0 >> pl.p_shift, pl0 = m0 & (pl.p_box - 1);   // sample within the box, position within the chunk row
        int kit = 0;
        for (int tile = blockIdx.x; tile < (use_tma ? 0 : pl.num_tiles); tile += gridDim.x) {
            const TmaTile tc = tma_tile(tile, pl);
            const float* base[4];
            const float* base2[4];                                     // the extra K panel's tensor (same samples, plane = the output plane)
            int pq[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int G = tc.mt * 4 + i;
                const int nb = G / pl.cps;
                const int nn = nb * pl.n_box + nl;
                const bool ok = G < pl.groups32 && nn < d.n;
                pq[i] = ok ? ((G - nb * pl.cps) << pl.p_shift) + pl0 : -(1 << 30);
                base[i] = in + (int64_t)(ok ? nn : 0) * d.c_in_total * d.p_in;
                base2[i] = pl.nkt2 ? in2 + (int64_t)(ok ? nn : 0) * pl.c2_total * pl.p_in2 : in;
            }
            const int ch_g = tc.g * d.g_in;
            for (int it = 0; it < kiters; ++it) {
                const int k = kit + it, s = k % S;
                const uint32_t ph = (uint32_t)(k / S) & 1u;
                const bool panel = it >= kmain;
                const int ict = it / d.ntap, tap = panel ? 0 : it - ict * d.ntap;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                const uint32_t a_dst = smem_u32(a_base + (size_t)s * A_STAGE_BYTES);
                const int ch0 = panel ? (it - kmain) * UK : ch_g + d.tap_in_ch[tap] + ict * UK, sh = panel ? 0 : d.tap_shift[tap];
                const int c_lim = panel ? pl.c2_total : d.c_in_total, p_lim = panel ? pl.p_in2 : d.p_in;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int pos = pq[i] + sh;                       // 4-aligned: the chunk lies entirely inside or outside the plane
                    const bool pok = pos >= 0 && pos < p_lim;
                    const float* bs = panel ? base2[i] : base[i];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int c = rr + 16 * h;
                        const bool ok = pok && ch0 + c < c_lim;
                        const uint32_t dst = a_dst + i * TM_GROUP_BYTES + c * 128 + ((chunk * 16) ^ ((c & 3) << 5));
                        tm_cp_async16(dst, ok ? bs + (int64_t)(ch0 + c) * p_lim + pos : in, ok ? 16u : 0u);
                    }
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full0 + 8 * s) : "memory");
            }
            kit += kiters;
        }
    } else {
        // ===== epilogue warps: TMEM lane = row of the tile = (M group, element of the box) =====
        const int quarter = warp & 3;                                 // TMEM lane quarter this warp may read == M group of the tile
        const int rnd = d.precision == KGAN_PREC_TF32;                // activations are stored tf32-rounded in tf32 mode
        const int colpar = (warp - TM_EPI_WARP0) >> 2;
        int ti = 0;
        for (int tile = blockIdx.x; tile < pl.num_tiles; tile += gridDim.x, ++ti) {
            const TmaTile tc = tma_tile(tile, pl);
            const int buf = ti & 1;
            const int G = tc.mt * 4 + quarter;
            const int nb = G / pl.cps;
            const int nn = nb * pl.n_box + (lane >> pl.p_shift);
            const int pv = ((G - nb * pl.cps) << pl.p_shift) + (lane & (pl.p_box - 1));
            const bool valid = G < pl.groups32 && nn < d.n;
            const int po = valid ? pv : 0, nv = valid ? nn : 0;
            const int out_ch0 = tc.g * d.g_out, oc_base = tc.ns * pl.n_cta;
            int d1 = 0, dz = 0;                                       // offsets relative to the first destination; 0: none
            int pg0 = po + tc.g * d.g_pout;
            if (omap) {                                               // scatter store: destinations of source position po inside the output plane
                const int q0 = __ldg(omap + 3 * po), q1 = __ldg(omap + 3 * po + 1), q2 = __ldg(omap + 3 * po + 2);
                pg0 = q0;
                d1 = q1 >= 0 ? q1 - q0 : 0;
                dz = q2 >= 0 ? q2 - q0 : 0;
            }
            const int pst = out_plane(d), pg = pg0;                       // plane stride, position inside the output plane
            float* op = out + ((int64_t)nv * d.c_out_total + out_ch0 + oc_base) * pst + pg;
            const int64_t astride = d.add_period ? d.add_period : pst;
            const float* ap = add ? add + ((int64_t)nv * d.c_out_total + out_ch0 + oc_base) * astride + (d.add_period ? pg % d.add_period : pg)
                                  : nullptr;
            const float* bp = bias ? bias + out_ch0 + oc_base : nullptr;
            const float* bp2 = (pl.nkt2 && bias2) ? bias2 + out_ch0 + oc_base : nullptr;
            const uint32_t tbar = tfull0 + 8 * buf, tpar = (uint32_t)(ti >> 1) & 1u;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * XM * pl.n_cta;
            const int ncols = min(pl.n_cta, d.co - oc_base);
            if (16 * colpar >= ncols) {                               // this warp has no columns in the tile: it still has to observe the barrier
                mbar_wait(tbar, tpar);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            } else if (X3) {
                if (omap) tma_epilogue_tile<KGAN_ACT_NONE, true, true>(taddr, ncols, colpar, valid, op, pst, nullptr, astride, bp, lane, tbar, tpar, rnd, nullptr, d1, dz, pl.n_cta);
                else if (d.act == KGAN_ACT_LRELU) tma_epilogue_tile<KGAN_ACT_LRELU, false, true>(taddr, ncols, colpar, valid, op, pst, ap, astride, bp, lane, tbar, tpar, rnd, bp2, 0, 0, pl.n_cta);
                else if (d.act == KGAN_ACT_TANH) tma_epilogue_tile<KGAN_ACT_TANH, false, true>(taddr, ncols, colpar, valid, op, pst, ap, astride, bp, lane, tbar, tpar, rnd, bp2, 0, 0, pl.n_cta);
                else tma_epilogue_tile<KGAN_ACT_NONE, false, true>(taddr, ncols, colpar, valid, op, pst, ap, astride, bp, lane, tbar, tpar, rnd, bp2, 0, 0, pl.n_cta);
            } else if (omap) tma_epilogue_tile<KGAN_ACT_NONE, true>(taddr, ncols, colpar, valid, op, pst, nullptr, astride, bp, lane, tbar, tpar, rnd, nullptr, d1, dz);
            else if (KM) {
                if (d.act == KGAN_ACT_LRELU) tma_epilogue_tile<KGAN_ACT_LRELU, false, false, true>(taddr, ncols, colpar, valid, op, pst, ap, astride, bp, lane, tbar, tpar, rnd, bp2);
                else tma_epilogue_tile<KGAN_ACT_NONE, false, false, true>(taddr, ncols, colpar, valid, op, pst, ap, astride, bp, lane, tbar, tpar, rnd, bp2);
            } else if (d.act == KGAN_ACT_LRELU) tma_epilogue_tile<KGAN_ACT_LRELU>(taddr, ncols, colpar, valid, op, pst, ap, astride, bp, lane, tbar, tpar, rnd, bp2);
            else if (d.act == KGAN_ACT_TANH) tma_epilogue_tile<KGAN_ACT_TANH>(taddr, ncols, colpar, valid, op, pst, ap, astride, bp, lane, tbar, tpar, rnd, bp2);
            else tma_epilogue_tile<KGAN_ACT_NONE>(taddr, ncols, colpar, valid, op, pst, ap, astride, bp, lane, tbar, tpar, rnd, bp2);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(tempty0 + 8 * buf);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(pl.tmem_cols) : "memory");
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libkgan.so does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// shared with tapconv_wgrad_tma.cu: 3-D fp32 tiled tensor map (dims innermost first); swizzle 0 = 128B with 32-byte atoms, 1 = 128B,
// 2 = 64B, 3 = 32B, 4 = none
int tma_encode_3d_f32(CUtensorMap* map, const float* base, const uint64_t gdim[3], const uint64_t gstr_bytes[2], const uint32_t box[3], int swizzle) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available");
        return 1;
    }
    memset(map, 0, sizeof(*map));
    const cuuint64_t gd[3] = {gdim[0], gdim[1], gdim[2]};
    const cuuint64_t gs[2] = {gstr_bytes[0], gstr_bytes[1]};
    const cuuint32_t bx[3] = {box[0], box[1], box[2]};
    const cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_32B : swizzle == 4 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
        return 1;
    }
    return 0;
}

int tapconv_tma_eligible(const kgan_tapconv_desc& d) {
    TmaPlan p;
    return make_tma_plan(d, p) ? 1 : 0;
}

static int launch_tma(const kgan_tapconv_desc& d, TmaPlan& p, const float* in, const float* wp, const float* bias, const float* add, float* out,
                      const float* in2, const float* wp2, const float* bias2, cudaStream_t stream, const int32_t* omap = nullptr) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) {
        set_error("tapconv_fwd_tma: cuTensorMapEncodeTiled not available");
        return 1;
    }
    CUtensorMap tmap, tmap2;
    memset(&tmap, 0, sizeof(tmap));
    memset(&tmap2, 0, sizeof(tmap2));
    const cuuint64_t gdim[3] = {(cuuint64_t)d.p_in, (cuuint64_t)d.n, (cuuint64_t)d.c_in_total};
    const cuuint64_t gstr[2] = {(cuuint64_t)d.c_in_total * d.p_in * 4, (cuuint64_t)d.p_in * 4};
    const cuuint32_t box[3] = {(cuuint32_t)p.p_box, (cuuint32_t)p.n_box, 32u};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = CUDA_SUCCESS;
    if (p.kmajor) {                                                // (channel, sample, 1): 128 sample rows of 32 channels, SWIZZLE_128B
        const cuuint64_t kdim[3] = {(cuuint64_t)d.c_in_total, (cuuint64_t)d.n, 1};
        const cuuint64_t kstr[2] = {(cuuint64_t)d.c_in_total * 4, (cuuint64_t)d.c_in_total * 4 * (cuuint64_t)d.n};
        const cuuint32_t kbox[3] = {32u, (cuuint32_t)UM, 1u};
        r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(in), kdim, kstr, kbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (p.p_box == 32) {
        r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(in), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS && p.nkt2) {                          // the extra K panel's tensor: (position, sample, channel) of in2
            const cuuint64_t g2[3] = {(cuuint64_t)p.p_in2, (cuuint64_t)d.n, (cuuint64_t)p.c2_total};
            const cuuint64_t s2[2] = {(cuuint64_t)p.c2_total * p.p_in2 * 4, (cuuint64_t)p.p_in2 * 4};
            r = enc(&tmap2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(in2), g2, s2, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
    }
    if (r != CUDA_SUCCESS) {
        set_error("tapconv_fwd_tma: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return 1;
    }
    static SmemAttrOnce attr, attr3;
    const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
    if (p.x3) {
        if (int e = ensure_smem(tapconv_fwd_tma_k<true, false>, 227 * 1024, attr3, "tapconv_fwd_tma (x3) attribute")) return e;
        tapconv_fwd_tma_k<true, false><<<grid, TM_THREADS_X3, p.smem_bytes, stream>>>(d, p, tmap, wp, in, bias, add, out, tmap2, wp2, in2, bias2, omap);
        return check_launch("tapconv_fwd_tma (x3)");
    }
    if (p.kmajor) {
        static SmemAttrOnce attrk;
        if (int e = ensure_smem(tapconv_fwd_tma_k<false, true>, 227 * 1024, attrk, "tapconv_fwd_tma (K-major) attribute")) return e;
        tapconv_fwd_tma_k<false, true><<<grid, TM_THREADS_TMA, p.smem_bytes, stream>>>(d, p, tmap, wp, in, bias, add, out, tmap2, wp2, in2, bias2, omap);
        return check_launch("tapconv_fwd_tma (K-major)");
    }
    if (int e = ensure_smem(tapconv_fwd_tma_k<false, false>, 227 * 1024, attr, "tapconv_fwd_tma attribute")) return e;
    tapconv_fwd_tma_k<false, false><<<grid, p.p_box == 32 ? TM_THREADS_TMA : TM_THREADS_CP, p.smem_bytes, stream>>>(d, p, tmap, wp, in, bias, add, out,
                                                                                                                   tmap2, wp2, in2, bias2, omap);
    return check_launch("tapconv_fwd_tma");
}

// -1: not eligible (caller falls back to the gather kernel)
int tapconv_fwd_tma(const kgan_tapconv_desc& d, const float* in, const float* wp, const int32_t* pmap, const float* bias, const float* add,
                    float* out, cudaStream_t stream) {
    TmaPlan p;
    if (!make_tma_plan(d, p)) return -1;
    if (reinterpret_cast<uintptr_t>(in) & 15) return -1;
    (void)pmap;                                                    // the shift form replaces the position map
    return launch_tma(d, p, in, wp, bias, add, out, nullptr, nullptr, nullptr, stream);
}

static bool res_plans(const kgan_tapconv_desc& d, const kgan_tapconv_desc& d2, TmaPlan& p) {
    if (d.groups != 1 || d2.groups != 1 || d2.ntap != 1 || d2.tma_mode != 1 || d2.tap_shift[0] != 0 || d2.tap_in_ch[0] != 0) return false;
    if (d2.n != d.n || d2.co != d.co || d2.c_out_total != d.c_out_total || d2.p_in != d.p_out || d2.p_out != d.p_out || d2.ck != d2.c_in_total) return false;
    if (d.p_out_plane != 0 || d2.p_out_plane != 0 || d.w_oc_blk != 0 || d2.w_oc_blk != 0) return false;
    TmaPlan p2;
    if (!make_tma_plan(d, p, d2.ck) || p.kmajor) return false;
    return make_tma_plan(d2, p2) && !p2.kmajor && p2.n_cta == p.n_cta && p2.n_split == p.n_split && p2.n_rows == p.n_rows;       // same image tiling
}

int tapconv_tma_res_eligible(const kgan_tapconv_desc& d, const kgan_tapconv_desc& d2) {
    TmaPlan p;
    return res_plans(d, d2, p) ? 1 : 0;
}

// The tap convolution `d` whose result is STORED through a scatter table: source position p of the p_out computed positions goes to
// omap[3p] and (if >= 0) omap[3p + 1] of the output plane of d.p_out_plane positions, and keeps omap[3p + 2] (if >= 0) zero.  Used to
// write a graph conv's output directly in the time-unfolded layout the following strided temporal conv reads (geometry.UnfoldedTcnGeom):
// no intermediate tensor, no copy kernel.  -1: not eligible.
int tapconv_fwd_tma_scatter(const kgan_tapconv_desc& d, const float* in, const float* wp, const int32_t* omap, const float* bias, float* out,
                            cudaStream_t stream) {
    if (d.groups != 1 || d.g_pout != 0 || d.p_out_plane < d.p_out || d.add_period != 0) return -1;
    TmaPlan p;
    if (!make_tma_plan(d, p) || p.kmajor) return -1;
    if (reinterpret_cast<uintptr_t>(in) & 15) return -1;
    return launch_tma(d, p, in, wp, bias, nullptr, out, nullptr, nullptr, nullptr, stream, omap);
}

int tapconv_tma_scatter_eligible(const kgan_tapconv_desc& d) {
    TmaPlan p;
    return d.groups == 1 && d.g_pout == 0 && d.p_out_plane >= d.p_out && d.add_period == 0 && make_tma_plan(d, p) && !p.kmajor;
}

// The tap convolution `d` with an extra K panel: out = act(conv_d(in) + bias + conv_d2(in2) + bias2), d2 a 1x1 convolution (one tap, no
// shift) of a second tensor with the SAME samples, output channels and plane as d's output - the residual branch of a critic block
// (discriminator.py:115-130) accumulated into the temporal conv's TMEM accumulator instead of being written to HBM by one GEMM and read
// back as `add` by the next.  -1: not eligible (the caller runs the two convolutions separately).
int tapconv_fwd_tma_res(const kgan_tapconv_desc& d, const kgan_tapconv_desc& d2, const float* in, const float* wp, const float* in2, const float* wp2,
                        const float* bias, const float* bias2, float* out, cudaStream_t stream) {
    TmaPlan p;
    if (!res_plans(d, d2, p)) return -1;
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(in2)) & 15) return -1;
    return launch_tma(d, p, in, wp, bias, nullptr, out, in2, wp2, bias2, stream);
}

}  // namespace kgan
