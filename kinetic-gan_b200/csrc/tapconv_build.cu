// Tap convolution with the activation operand BUILT in shared memory (KGAN_PREC_TF32, any position map).
//
// Same GEMM as tapconv_tma.cu / tapconv_umma.cu -  D[M = 128 positions][N = output channels] += A[M][K] * B[N][K]^T,
// K = (channel tile, tap) - with a third way of producing A:
//
//   1. one elected thread stages the RAW activations a tile can touch with one or two bulk tensor copies per 32-channel tile
//      (cp.async.bulk.tensor, 3-D map over (position, sample, channel), no swizzle): [32 channels] x [samples] x [span positions].
//      Box origins are rounded down to 4 positions, so any map is 16-byte addressable; what lies outside the plane is zero-filled
//      by the TMA unit (the convolution's zero padding);
//   2. four builder warps (thread = tile row) gather every tap's operand from that staged tile through the position map -
//      shift, stride, joint / frame selection, anything injective - round it to tf32 (round to nearest: this kernel is exact for
//      ANY fp32 input, rounded by its producer or not) and write the K-major SWIZZLE_128B image the tensor core reads
//      (row = position, 128 bytes = 32 channels; lanes are consecutive rows: conflict-free reads and 16-byte stores);
//   3. MMA issue, TMEM double buffering, weight streaming and the epilogue are those of tapconv_tma.cu.
//
// What it buys over the other two producers:
//   * taps that read the SAME channels at different positions (the 3 x 1 temporal convolution) share ONE staged tile: the
//     activations cross L2 -> SM once instead of once per tap, and one large box replaces 4 small ones per tap (the TMA-fed kernel
//     issues 12 boxes of 4 KB per 128 x 32 x 3-tap tile; on 32/64-channel layers their issue rate, not HBM, set the pace);
//   * no alignment rule on shifts: planes of 5 or 11 joints, stride-2 frame selection and joint selection run from the tensor as it
//     is - no joint padding, no time-unfolded copy, no selected copy of the residual input (geometry.py);
//   * against the SIMT-gather kernel (tapconv_umma.cu): global memory is read by the copy engine in whole tiles, not by 16
//     scattered 4-byte loads per thread with the ring stalled behind their latency.
//
// Tiling.  Planes of more than 128 output positions: tiles of 128 rows inside one sample (the last one ragged); the staged range
// [lo, lo + span) of a tile is found by the producer warp from the position map itself (min over the tile's rows and the channel
// block's taps) and published with the tile.  Planes of at most 128 positions: a tile holds floor(128 / p_out) whole samples and
// stages their whole input planes.  `span` (a launch constant: it is the box extent of the tensor map) comes with the descriptor
// (kgan_tapconv_desc.stage_span, computed from the position map on the host side).
//
// Graph convolution with the adjacency product IN the operand builder (kgan_gcn_fwd_tf32, template MIX): the staged tile holds the
// block input x itself; for partition k the builders write  xa_k[row (t, w)][c] = sum_j coef_j * x[c][t, v_j]  - the (A (.) edge_importance)
// product of tgcn.py:66, taken over the few non-zeros of column w of A_k (compacted into shared memory once per CTA) - rounded to tf32,
// straight into the K-major operand image, and the K partitions are K taps of ONE tcgen05 GEMM: conv1x1 + einsum('nkctv,kvw->nctw') as a
// single kernel, the K*C-channel mixed tensor never exists in HBM.
//
// Warp roles: 0 = raw-tile producer, 1 = MMA issuer / TMEM owner, 2-9 = epilogue, 10-13 = operand builders, 14 = weight loader.
#include <cuda.h>
#include <string.h>

#include "umma.cuh"

namespace kgan {

constexpr int BD_EPI_WARPS = 8;
constexpr int BD_EPI_WARP0 = 2;
constexpr int BD_BUILD_WARP0 = BD_EPI_WARP0 + BD_EPI_WARPS;       // 10
constexpr int BD_BUILD_WARPS = 4;
constexpr int BD_BUILDERS = 32 * BD_BUILD_WARPS;                  // one per tile row
constexpr int BD_W_WARP = BD_BUILD_WARP0 + BD_BUILD_WARPS;        // 14
constexpr int BD_THREADS = 32 * (BD_W_WARP + 1);
constexpr int BD_MAX_CB = 4;                                       // channel blocks (taps that read the same input channels)
constexpr int BD_ROW_CACHE = 3;                                    // position-map rows whose entry a builder thread keeps in registers

struct BuildPlan {
    int n_cta, n_split, n_rows, tmem_cols, nkt;   // identical to UmmaPlan (the packed weight image is shared)
    int small;                                    // 1: tile = spt whole planes, 0: tpp tiles of 128 rows per plane
    int spt, tpp;
    int m_tiles, num_tiles;
    int per_mt;                                   // num_tiles / m_tiles = groups * n_split
    int bspan, nbox, nseg;                        // box extent (positions), boxes per staged tile (1 / 2), samples per staged tile
    int chan_stride;                              // floats between channels inside one box image: nseg * bspan
    int box_floats;                               // 32 * chan_stride
    int raw_bytes;                                // nbox * box_floats * 4
    int R, stages, smem_bytes;
    int w_res, w_res_bytes;
    int ncb;                                      // channel blocks
    int cb_in_ch[BD_MAX_CB];
    int cb_beg[BD_MAX_CB + 1];                    // taps of block c: order[cb_beg[c] .. cb_beg[c + 1])
    int order[KGAN_MAX_TAPS];
    int nmaps;                                    // distinct position-map rows, map_row[i]; tap_map[t] = index into map_row
    int map_row[KGAN_MAX_TAPS];
    int tap_map[KGAN_MAX_TAPS];
    int pm_bytes;                                 // shared-memory copy of the used position-map rows (nmaps x p_out ints), 0: read from global
    int mix;                                      // 1: adjacency product in the builders (d.mix_v joints -> d.mix_w joints, d.mix_l list entries)
};
struct MixItem {
    int v;                                        // input joint
    float coef;                                   // A[k, v, w]
};
constexpr int BD_MAX_TPP = 32;                    // tiles per plane (planes of up to 4096 positions)

bool tapconv_umma_nsplit(const kgan_tapconv_desc& d, int* n_cta, int* n_split, int* n_rows, int* tmem_cols, int* nkt);
int tma_encode_3d_f32(CUtensorMap* map, const float* base, const uint64_t gdim[3], const uint64_t gstr_bytes[2], const uint32_t box[3], int swizzle);

static bool make_build_plan(const kgan_tapconv_desc& d, BuildPlan& p) {
    if (d.precision == KGAN_PREC_TF32X3) return false;              // the fp32-accurate split lives in the TMA-fed kernels only
    p.mix = d.mix_v > 0 ? 1 : 0;
    if (p.mix) {
        if (d.mix_w <= 0 || d.mix_l <= 0 || d.mix_l > 8 || d.groups != 1 || d.ntap > 4) return false;
        if (d.p_in % d.mix_v || d.p_out % d.mix_w || d.p_in / d.mix_v != d.p_out / d.mix_w) return false;          // same frames in and out
        for (int t = 0; t < d.ntap; ++t)
            if (d.tap_in_ch[t] != 0) return false;                                                                  // every partition reads x itself
    }
    if (d.stage_span <= 0 || (d.stage_span & 3) || (d.p_in & 3)) return false;      // global strides / box extents: multiples of 16 bytes
    if ((d.p_out_plane != 0 && (d.mix_v <= 0 || d.g_pout != 0)) || d.w_oc_blk < 0) return false;      // enlarged output planes: fused gcn + scatter store only
    if (!tapconv_umma_nsplit(d, &p.n_cta, &p.n_split, &p.n_rows, &p.tmem_cols, &p.nkt)) return false;
    p.small = d.p_out <= UM ? 1 : 0;
    if (p.small) {
        p.spt = UM / d.p_out;
        // whole input planes are staged: cap the samples per tile so that a staged tile stays within 56 KB (maps that select few
        // outputs from large planes fill fewer of the 128 rows)
        const int cap = (56 * 1024) / (32 * 4 * round_up(d.stage_span, 4));
        if (cap < 1) return false;
        if (p.spt > cap) p.spt = cap;
        p.tpp = 1;
        p.nseg = p.spt;
        const int64_t mt = ceil_div64(d.n, p.spt);
        if (mt >= (1 << 24)) return false;
        p.m_tiles = (int)mt;
    } else {
        p.spt = 1;
        p.tpp = ceil_div(d.p_out, UM);
        p.nseg = 1;
        const int64_t mt = (int64_t)d.n * p.tpp;
        if (mt >= (1 << 24)) return false;
        p.m_tiles = (int)mt;
    }
    if ((int64_t)p.m_tiles * d.groups * p.n_split > (1 << 28)) return false;
    p.num_tiles = p.m_tiles * p.n_split * d.groups;
    p.per_mt = p.n_split * d.groups;
    p.nbox = d.stage_span <= 256 ? 1 : 2;
    p.bspan = p.nbox == 1 ? d.stage_span : round_up(ceil_div(d.stage_span, 2), 4);
    if (p.bspan > 256 || p.nseg > 256) return false;
    p.chan_stride = p.nseg * p.bspan;
    p.box_floats = 32 * p.chan_stride;
    p.raw_bytes = p.nbox * p.box_floats * 4;
    if (p.raw_bytes > 56 * 1024) return false;
    // channel blocks and position-map rows
    p.ncb = 0;
    int pos = 0;
    bool used[KGAN_MAX_TAPS] = {false};
    for (int t = 0; t < d.ntap; ++t) {
        if (used[t]) continue;
        if (p.ncb == BD_MAX_CB) return false;
        p.cb_in_ch[p.ncb] = d.tap_in_ch[t];
        p.cb_beg[p.ncb] = pos;
        for (int u = t; u < d.ntap; ++u)
            if (!used[u] && d.tap_in_ch[u] == d.tap_in_ch[t]) {
                used[u] = true;
                p.order[pos++] = u;
            }
        ++p.ncb;
    }
    p.cb_beg[p.ncb] = pos;
    p.nmaps = 0;
    for (int t = 0; t < d.ntap; ++t) {
        int m = -1;
        for (int i = 0; i < p.nmaps; ++i)
            if (p.map_row[i] == d.tap_row[t]) m = i;
        if (m < 0) {
            m = p.nmaps;
            p.map_row[p.nmaps++] = d.tap_row[t];
        }
        p.tap_map[t] = m;
    }
    if (!p.small && p.tpp > BD_MAX_TPP) return false;
    // Shared memory.  The kernel is bound by the bytes in flight per SM (loaded HBM latency ~2-3 us): what is in flight are the RAW
    // tiles, so they get the deep ring (up to 8); the operand stages only decouple builders and tensor core (3 are enough).
    const int b_stage = p.n_cta * UK * 4;
    p.pm_bytes = round_up(p.nmaps * d.p_out * 4, 128);
    if (p.pm_bytes > 32 * 1024) p.pm_bytes = 0;
    if (p.mix) p.pm_bytes = round_up(d.ntap * d.mix_w * d.mix_l * (int)sizeof(MixItem), 128);      // the region holds the compacted adjacency instead
    int budget = 224 * 1024 - 2048 - p.pm_bytes;
    p.w_res = 0;
    p.w_res_bytes = 0;
    const int64_t img = (int64_t)d.groups * p.nkt * d.ntap * p.n_cta * UK * 4;
    const int64_t tiles_per_cta = ceil_div64(p.num_tiles, kNumSMs);
    // operand stages: as many (up to 6) as leave room for >= 3 raw tiles - the builder -> tensor core -> builder hand-offs each cost a
    // barrier round trip of a few hundred cycles, and the ring has to cover them; resident weights when they fit beside >= 3 stages
    const bool res_ok = p.n_split == 1 && tiles_per_cta >= 2;
    p.stages = 0;
    if (res_ok)
        for (int st = 6; st >= 3 && !p.stages; --st)
            if (img + st * A_STAGE_BYTES + 3 * p.raw_bytes <= budget) {
                p.stages = st;
                p.w_res = 1;
                p.w_res_bytes = (int)img;
                budget -= (int)img + st * A_STAGE_BYTES;
            }
    if (!p.stages) {
        for (int st = 6; st >= 2 && !p.stages; --st)
            if (st * (A_STAGE_BYTES + b_stage) + 3 * p.raw_bytes <= budget) {
                p.stages = st;
                budget -= st * (A_STAGE_BYTES + b_stage);
            }
    }
    if (!p.stages) return false;
    p.R = budget / p.raw_bytes;
    if (p.R > 8) p.R = 8;
    if (p.R < 2) return false;
    p.smem_bytes = 1024 + p.R * p.raw_bytes + p.stages * A_STAGE_BYTES + (p.w_res ? p.w_res_bytes : p.stages * b_stage) + p.pm_bytes + 1024;
    return true;
}

__device__ __forceinline__ void bd_tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}
// K-major 32-bit operand, SWIZZLE_128B: 128-byte rows (32 K elements), 8-row atoms of 1024 bytes (SBO); a K step of 8 elements
// advances the start address by 32 bytes inside the atom (as the Linear-layer plan of tapconv_tma.cu)
__device__ __forceinline__ uint64_t bd_desc_k_sw128(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// 32-bit shared-space load.  The staged tile is addressed through the shared window explicitly: a pointer derived from the aligned
// dynamic-shared base is a GENERIC pointer to the compiler (the alignment arithmetic goes through an integer), and the gather loop
// then compiles to 64-bit generic LD.E with an address pair and a predicate per element - 2.6x slower than the whole TMA-fed kernel.
__device__ __forceinline__ float bd_lds(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

struct BdTile {
    int g, mt, ns;
};
__device__ __forceinline__ BdTile bd_tile(int tile, const BuildPlan& pl) {
    BdTile c;
    if (pl.per_mt == 1) {                              // (per-tile index arithmetic is on the critical path of small tiles: see tapconv_tma.cu)
        c.mt = tile;
        c.g = c.ns = 0;
        return c;
    }
    c.mt = tile / pl.per_mt;                           // groups * n_split consecutive tiles re-use the staged activations in L2
    const int r = tile - c.mt * pl.per_mt;
    c.g = r / pl.n_split;
    c.ns = r - c.g * pl.n_split;
    return c;
}

// NZ: + nw[column] * nz, nz = this row's element of a one-channel noise tensor (the generator's NoiseInjection, generator.py:12-19)
template <int ACT, bool SCAT = false, bool NZ = false>
__device__ __forceinline__ void bd_epilogue_tile(uint32_t taddr, int ncols, int colpar, bool valid, float* __restrict__ op, int p_out,
                                                 const float* __restrict__ ap, int64_t astride, const float* __restrict__ bp, int lane,
                                                 uint32_t tfull_bar, uint32_t tfull_parity, int rnd, int d1 = 0, int dz = 0, float nz = 0.f,
                                                 const float* __restrict__ nwp = nullptr) {
    bool waited = false;
    for (int col0 = 16 * colpar; col0 < ncols; col0 += 16 * (BD_EPI_WARPS / 4)) {
        const int nc = min(16, ncols - col0);                         // warp-uniform
        float av[16];
        if (ap) {
#pragma unroll
            for (int j = 0; j < 16; ++j) av[j] = ldg_pred(ap + (int64_t)(col0 + j) * astride, valid && j < nc);
        }
        const float bl = (bp && lane < nc) ? __ldg(bp + col0 + lane) : 0.f;
        const float nl = (NZ && lane < nc) ? __ldg(nwp + col0 + lane) : 0.f;
        if (!waited) {
            mbar_wait(tfull_bar, tfull_parity);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            waited = true;
        }
        uint32_t r[16];
        tmem_ld16(taddr + col0, r);
        float* o = op + (int64_t)col0 * p_out;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float val = __uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bl, j);
            if (NZ) val = fmaf(__shfl_sync(0xffffffffu, nl, j), nz, val);
            if (ap) val += av[j];
            if (ACT == KGAN_ACT_LRELU) val = val > 0.f ? val : 0.2f * val;
            if (ACT == KGAN_ACT_TANH) val = tanhf(val);
            if (valid && j < nc) {
                const float q = tf32_out(val, rnd);
                *o = q;
                if (SCAT) {                                        // scatter store (see tapconv_tma.cu): second copy / zero slot, relative to *o
                    if (d1 != 0) o[d1] = q;
                    if (dz != 0) o[dz] = 0.f;
                }
            }
            o += p_out;
        }
    }
}

// NZ: the instantiation with the noise term in the epilogue (kgan_tapconv_fwd_tf32_noise); `adj` then points to the noise tensor (n, 1, plane)
// and `nw` to the per-channel noise weights.  A separate instantiation: extra epilogue variants cost the main ones registers (tapconv_tma.cu).
template <bool MIX, bool NZ = false>
__global__ void __launch_bounds__(BD_THREADS, 1) tapconv_fwd_build_k(const __grid_constant__ kgan_tapconv_desc d, const __grid_constant__ BuildPlan pl,
                                                                     const __grid_constant__ CUtensorMap tmap, const float* __restrict__ wp,
                                                                     const int32_t* __restrict__ pmap, const float* __restrict__ bias,
                                                                     const float* __restrict__ add, float* __restrict__ out,
                                                                     const float* __restrict__ adj, const int32_t* __restrict__ omap,
                                                                     const float* __restrict__ nw = nullptr) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // swizzle atoms: 1024-byte aligned
    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
    const int S = pl.stages, R = pl.R;
    const int b_stage_bytes = pl.n_cta * UK * 4;
    uint8_t* a_base = smem;                                          // S x 16 KB, each 1024-byte aligned
    uint8_t* b_base = a_base + (size_t)S * A_STAGE_BYTES;            // ring of S weight stages, or the resident image
    uint8_t* raw_base = b_base + (pl.w_res ? (size_t)pl.w_res_bytes : (size_t)S * b_stage_bytes);
    int* pm_s = reinterpret_cast<int*>(raw_base + (size_t)R * pl.raw_bytes);          // [nmaps][p_out] copy of the used position-map rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(pm_s) + pl.pm_bytes);
    // full[S], empty[S], rawfull[R], rawempty[R], tfull[2], tempty[2], wfull
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);
    const uint32_t rfull0 = smem_u32(bars + 2 * S), rempty0 = smem_u32(bars + 2 * S + R);
    const uint32_t tfull0 = smem_u32(bars + 2 * S + 2 * R), tempty0 = smem_u32(bars + 2 * S + 2 * R + 2);
    const uint32_t wfull = smem_u32(bars + 2 * S + 2 * R + 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 2 * R + 5);
    int* raw_lo = reinterpret_cast<int*>(tmem_slot + 2);             // [R]: first staged position of the tile in raw buffer r
    int* lo_tab = raw_lo + 8;                                         // [ncb][tpp]: first staged position per (channel block, tile of the plane)
    const int kiters = pl.nkt * d.ntap;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full0 + 8 * s, BD_BUILD_WARPS + (pl.w_res ? 0 : 1));   // one arrival per builder WARP (+ the weight loader's expect_tx arrival)
            mbar_init(empty0 + 8 * s, 1);                                 // tcgen05.commit
        }
        for (int r = 0; r < R; ++r) {
            mbar_init(rfull0 + 8 * r, 1);                                 // the producer's expect_tx arrival; the tensor copies complete the bytes
            mbar_init(rempty0 + 8 * r, BD_BUILD_WARPS);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull0 + 8 * b, 1);
            mbar_init(tempty0 + 8 * b, 32 * BD_EPI_WARPS);
        }
        mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(pl.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // once per CTA: the used position-map rows -> shared memory, and from them the first staged position of every (channel block,
    // tile of the plane): min over the tile's rows and the block's taps, rounded down to a 16-byte boundary
    if (MIX) {
        // compacted adjacency: for (partition k, output joint w) the non-zeros of column w of A_k, padded with zero coefficients to mix_l
        MixItem* mt = reinterpret_cast<MixItem*>(pm_s);
        for (int e = threadIdx.x; e < d.ntap * d.mix_w; e += blockDim.x) {
            const int k = e / d.mix_w, w = e - k * d.mix_w;
            int cnt = 0;
            for (int v = 0; v < d.mix_v && cnt < d.mix_l; ++v) {
                const float a = __ldg(adj + ((int64_t)k * d.mix_v + v) * d.mix_w + w);
                if (a != 0.f) mt[e * d.mix_l + cnt++] = MixItem{v, a};
            }
            for (; cnt < d.mix_l; ++cnt) mt[e * d.mix_l + cnt] = MixItem{0, 0.f};
        }
        if (!pl.small)                                                // first staged position of tile tp: its first frame, 16-byte aligned
            for (int tp = threadIdx.x; tp < pl.tpp; tp += blockDim.x) lo_tab[tp] = ((tp * UM) / d.mix_w * d.mix_v) & ~3;
    } else if (pl.pm_bytes) {
        for (int i = threadIdx.x; i < pl.nmaps * d.p_out; i += blockDim.x) {
            const int m = i / d.p_out;
            pm_s[i] = __ldg(pmap + (int64_t)pl.map_row[m] * d.p_out + (i - m * d.p_out));
        }
    }
    if (!MIX && !pl.small) {
        for (int e = warp; e < pl.ncb * pl.tpp; e += BD_THREADS / 32) {
            const int cb = e / pl.tpp, tp = e - cb * pl.tpp;
            const int row0 = tp * UM, nrows = min(UM, d.p_out - row0);
            int m = 0x7fffffff;
            for (int k = pl.cb_beg[cb]; k < pl.cb_beg[cb + 1]; ++k) {
                const int32_t* mp = pmap + (int64_t)d.tap_row[pl.order[k]] * d.p_out + row0;
                for (int i = lane; i < nrows; i += 32) {
                    const int sidx = __ldg(mp + i);
                    if (sidx >= 0) m = min(m, sidx);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) lo_tab[e] = m == 0x7fffffff ? 0 : (m & ~3);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== raw-tile producer: per (tile, channel tile, channel block) one staged tile =====
        const bool leader = elect_one();
        int r = 0;
        uint32_t ph = 1;                                              // parity to wait for on rawempty[r]
        for (int tile = blockIdx.x; tile < pl.num_tiles; tile += gridDim.x) {
            const BdTile tc = bd_tile(tile, pl);
            int n0, row0;
            if (pl.small) {
                n0 = tc.mt * pl.spt;
                row0 = 0;                                             // whole planes: lo = 0
            } else {
                n0 = tc.mt / pl.tpp;
                row0 = (tc.mt - n0 * pl.tpp) * UM;
            }
            int lo_cb[BD_MAX_CB];
#pragma unroll
            for (int cb = 0; cb < BD_MAX_CB; ++cb) lo_cb[cb] = (!pl.small && cb < pl.ncb) ? lo_tab[cb * pl.tpp + row0 / UM] : 0;
            const int ch_g = tc.g * d.g_in;
            for (int ict = 0; ict < pl.nkt; ++ict) {
#pragma unroll
                for (int cb = 0; cb < BD_MAX_CB; ++cb) {
                    if (cb >= pl.ncb) break;
                    mbar_wait(rempty0 + 8 * r, ph);
                    if (leader) {
                        raw_lo[r] = lo_cb[cb];
                        mbar_arrive_expect_tx(rfull0 + 8 * r, (uint32_t)pl.raw_bytes);
                        const uint32_t dst = smem_u32(raw_base + (size_t)r * pl.raw_bytes);
                        const int ch0 = ch_g + pl.cb_in_ch[cb] + ict * UK;
                        for (int b = 0; b < pl.nbox; ++b)
                            bd_tma_load_3d(dst + b * pl.box_floats * 4, &tmap, lo_cb[cb] + b * pl.bspan, n0, ch0, rfull0 + 8 * r);
                    }
                    __syncwarp();
                    if (++r == R) {
                        r = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == BD_W_WARP) {
        // ===== weight loader: bulk copies of the packed tf32 image, in the builders' stage order (channel tile, block, tap) =====
        const bool leader = elect_one();
        const uint32_t chunk_bytes = pl.n_cta * 16;
        if (pl.w_res) {
            if (leader) {
                mbar_arrive_expect_tx(wfull, (uint32_t)pl.w_res_bytes);
                for (int off = 0; off < pl.w_res_bytes; off += 16384) {
                    const int nb = min(16384, pl.w_res_bytes - off);
                    bulk_g2s(smem_u32(b_base + off), reinterpret_cast<const uint8_t*>(wp) + off, (uint32_t)nb, wfull);
                }
            }
            __syncwarp();
        } else {
            int s = 0;
            uint32_t ph = 1;
            for (int tile = blockIdx.x; tile < pl.num_tiles; tile += gridDim.x) {
                const BdTile tc = bd_tile(tile, pl);
                const float* wg = wp + (int64_t)tc.g * pl.nkt * d.ntap * pl.n_rows * UK;
                const int oc_base = tc.ns * pl.n_cta;
                for (int ict = 0; ict < pl.nkt; ++ict)
                    for (int k = 0; k < d.ntap; ++k) {
                        const int tap = pl.order[k];
                        mbar_wait(empty0 + 8 * s, ph);
                        if (leader) {
                            mbar_arrive_expect_tx(full0 + 8 * s, chunk_bytes * 8);
                            const float* src = wg + (int64_t)(ict * d.ntap + tap) * pl.n_rows * UK;      // packing order: (channel tile, tap)
                            const uint32_t b_dst = smem_u32(b_base + (size_t)s * b_stage_bytes);
                            if (pl.n_split == 1) {
                                bulk_g2s(b_dst, src, chunk_bytes * 8, full0 + 8 * s);
                            } else {
#pragma unroll
                                for (int c = 0; c < 8; ++c)
                                    bulk_g2s(b_dst + c * chunk_bytes, src + ((int64_t)c * pl.n_rows + oc_base) * 4, chunk_bytes, full0 + 8 * s);
                            }
                        }
                        __syncwarp();
                        if (++s == S) {
                            s = 0;
                            ph ^= 1u;
                        }
                    }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp waits, one elected lane issues) =====
        const bool leader = elect_one();
        const uint32_t idesc = instr_desc_tf32(pl.n_cta);            // A and B K-major
        const uint32_t b_lbo = pl.n_cta * 16;
        int s = 0, ti = 0;
        uint32_t ph = 0;
        if (pl.w_res) mbar_wait(wfull, 0);
        for (int tile = blockIdx.x; tile < pl.num_tiles; tile += gridDim.x, ++ti) {
            const int buf = ti & 1;
            const uint32_t b_res = smem_u32(b_base) + (uint32_t)(bd_tile(tile, pl).g * kiters) * (uint32_t)b_stage_bytes;
            mbar_wait(tempty0 + 8 * buf, ((uint32_t)(ti >> 1) & 1u) ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem_base + buf * pl.n_cta;
            int it = 0;
            for (int ict = 0; ict < pl.nkt; ++ict)
                for (int k = 0; k < d.ntap; ++k, ++it) {
                    mbar_wait(full0 + 8 * s, ph);                    // (the builders ran the generic -> async proxy fence before arriving)
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (leader) {
                        const uint32_t a_addr = smem_u32(a_base + (size_t)s * A_STAGE_BYTES);
                        const uint32_t b_addr = pl.w_res ? b_res + (uint32_t)(ict * d.ntap + pl.order[k]) * (uint32_t)b_stage_bytes
                                                         : smem_u32(b_base + (size_t)s * b_stage_bytes);
#pragma unroll
                        for (int j = 0; j < UK / 8; ++j)
                            umma_tf32(acc, bd_desc_k_sw128(a_addr + j * 32), smem_desc(b_addr + j * 2 * b_lbo, b_lbo, CORE_SBO), idesc,
                                      (it > 0 || j > 0) ? 1u : 0u);
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
    } else if (warp >= BD_BUILD_WARP0) {
        // ===== operand builders: thread = tile row =====
        const int row = threadIdx.x - 32 * BD_BUILD_WARP0;
        int r = 0, s = 0;
        uint32_t rph = 0, sph = 1;                                   // parities: rawfull[r] to wait for, empty[s] to wait for
        const uint32_t row_off = (uint32_t)row * 128u, row_x = (uint32_t)(row & 7);
        for (int tile = blockIdx.x; tile < pl.num_tiles; tile += gridDim.x) {
            const BdTile tc = bd_tile(tile, pl);
            // this row's output position / sample slot, and its source position under every cached map row
            int p, seg;
            bool valid;
            if (pl.small) {
                seg = row / d.p_out;
                p = row - seg * d.p_out;
                valid = seg < pl.spt && tc.mt * pl.spt + seg < d.n;
            } else {
                const int n0 = tc.mt / pl.tpp;
                p = (tc.mt - n0 * pl.tpp) * UM + row;
                seg = 0;
                valid = p < d.p_out;
            }
            int srcs[BD_ROW_CACHE];
#pragma unroll
            for (int m = 0; m < BD_ROW_CACHE; ++m)
                srcs[m] = (!MIX && valid && m < pl.nmaps) ? (pl.pm_bytes ? pm_s[m * d.p_out + p] : __ldg(pmap + (int64_t)pl.map_row[m] * d.p_out + p)) : -1;
            int mix_t = 0, mix_w = 0;                                 // MIX: this row's frame and output joint
            if (MIX) {
                mix_t = p / d.mix_w;
                mix_w = p - mix_t * d.mix_w;
            }
            for (int ict = 0; ict < pl.nkt; ++ict) {
                const int kvalid = min(UK, d.ck - ict * UK);          // channels of this tile that belong to the contraction
                for (int cb = 0; cb < pl.ncb; ++cb) {
                    mbar_wait(rfull0 + 8 * r, rph);
                    const uint32_t raw = smem_u32(raw_base + (size_t)r * pl.raw_bytes);
                    const int lo = raw_lo[r];
                    for (int k = pl.cb_beg[cb]; k < pl.cb_beg[cb + 1]; ++k) {
                        const int tap = pl.order[k];
                        if (MIX) {
                            // xa_k[row][c] = sum_j coef_j * x[c][frame, v_j]: the adjacency product of partition `tap`, built into the operand image
                            const MixItem* mt = reinterpret_cast<const MixItem*>(pm_s) + (tap * d.mix_w + mix_w) * d.mix_l;
                            const uint32_t cs4 = 4u * (uint32_t)pl.chan_stride;
                            const int fbase = mix_t * d.mix_v - lo;       // offset of this row's frame inside the staged range
                            float acc[UK];
#pragma unroll
                            for (int c = 0; c < UK; ++c) acc[c] = 0.f;
                            for (int j = 0; j < d.mix_l; ++j) {
                                const MixItem it = mt[j];
                                const int off = fbase + it.v;
                                const bool ok = valid && it.coef != 0.f && off >= 0 && off < pl.nbox * pl.bspan;
                                const int bx = (pl.nbox == 2 && off >= pl.bspan) ? 1 : 0;
                                const uint32_t rp = raw + 4u * (uint32_t)(ok ? bx * pl.box_floats + seg * pl.bspan + (off - bx * pl.bspan) : 0);
                                const float cf = ok ? it.coef : 0.f;
                                float xv[UK];
#pragma unroll
                                for (int c = 0; c < UK; ++c) xv[c] = bd_lds(rp + (uint32_t)c * cs4);
#pragma unroll
                                for (int c = 0; c < UK; ++c) acc[c] = fmaf(cf, xv[c], acc[c]);
                            }
                            mbar_wait(empty0 + 8 * s, sph);
                            const uint32_t dst = smem_u32(a_base + (size_t)s * A_STAGE_BYTES) + row_off;
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                uint32_t v[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) v[e] = (4 * q + e) < kvalid ? to_tf32_fast(acc[4 * q + e]) : 0u;
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((uint32_t)q ^ row_x) << 4)), "r"(v[0]), "r"(v[1]),
                                             "r"(v[2]), "r"(v[3])
                                             : "memory");
                            }
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            __syncwarp();
                            if (lane == 0) mbar_arrive(full0 + 8 * s);
                            if (++s == S) {
                                s = 0;
                                sph ^= 1u;
                            }
                            continue;
                        }
                        const int m = pl.tap_map[tap];
                        int src = -1;
                        if (m < BD_ROW_CACHE) {
#pragma unroll
                            for (int q = 0; q < BD_ROW_CACHE; ++q) src = (q == m) ? srcs[q] : src;
                        } else if (valid) {
                            src = __ldg(pmap + (int64_t)d.tap_row[tap] * d.p_out + p);
                        }
                        // offset of (channel 0, this row's source) inside the staged tile
                        int off = src - lo;
                        const bool ok = src >= 0 && off >= 0 && off < pl.nbox * pl.bspan;
                        const int bx = (pl.nbox == 2 && off >= pl.bspan) ? 1 : 0;
                        // (an invalid row reads element 0 of the tile - a valid address - and discards it)
                        const uint32_t rp = raw + 4u * (uint32_t)(ok ? bx * pl.box_floats + seg * pl.bspan + (off - bx * pl.bspan) : 0);
                        const uint32_t cs4 = 4u * (uint32_t)pl.chan_stride;
                        const int kv = ok ? kvalid : 0;                   // channels this row really gathers
                        mbar_wait(empty0 + 8 * s, sph);
                        const uint32_t dst = smem_u32(a_base + (size_t)s * A_STAGE_BYTES) + row_off;
                        // all 32 gathers first (independent shared-memory loads in flight together), then the eight 16-byte stores: with
                        // load / store interleaved per chunk the `memory` clobber of each store serialised eight load round trips per stage
                        float xv[UK];
#pragma unroll
                        for (int c = 0; c < UK; ++c) xv[c] = bd_lds(rp + (uint32_t)c * cs4);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            uint32_t v[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) v[e] = (4 * q + e) < kv ? to_tf32_fast(xv[4 * q + e]) : 0u;
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((uint32_t)q ^ row_x) << 4)), "r"(v[0]), "r"(v[1]),
                                         "r"(v[2]), "r"(v[3])
                                         : "memory");
                        }
                        // one arrival per warp: 128 per-thread arrivals are 128 serialised atomics on one shared-memory word per stage -
                        // they, not the gather, set the pace of the first version.  __syncwarp orders the lanes' stores before lane 0's
                        // release; the MMA warp runs the proxy fence (see there)
                        // generic-proxy stores -> async-proxy reads of tcgen05.mma: fenced by the writers (no loads of theirs are in flight here;
                        // on the consumer side the fence sat on the MMA warp's critical path once per stage)
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(full0 + 8 * s);
                        if (++s == S) {
                            s = 0;
                            sph ^= 1u;
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(rempty0 + 8 * r);      // this warp's reads of raw[r] are done
                    if (++r == R) {
                        r = 0;
                        rph ^= 1u;
                    }
                }
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane = tile row =====
        const int quarter = warp & 3;                                 // TMEM lane quarter this warp may read
        const int colpar = (warp - BD_EPI_WARP0) >> 2;
        const int rnd = d.precision == KGAN_PREC_TF32;
        const int row = quarter * 32 + lane;
        int ti = 0;
        for (int tile = blockIdx.x; tile < pl.num_tiles; tile += gridDim.x, ++ti) {
            const BdTile tc = bd_tile(tile, pl);
            const int buf = ti & 1;
            int nn, pv;
            bool valid;
            if (pl.small) {
                const int seg = row / d.p_out;
                pv = row - seg * d.p_out;
                nn = tc.mt * pl.spt + seg;
                valid = seg < pl.spt && nn < d.n;
            } else {
                nn = tc.mt / pl.tpp;
                pv = (tc.mt - nn * pl.tpp) * UM + row;
                valid = pv < d.p_out;
            }
            int po = valid ? pv : 0;
            const int nv = valid ? nn : 0;
            const int out_ch0 = tc.g * d.g_out, oc_base = tc.ns * pl.n_cta;
            const int pst = omap ? d.p_out_plane : d.p_out;
            int d1 = 0, dz = 0;
            if (omap) {                                               // scatter store: destinations of source position po inside the output plane
                const int q0 = __ldg(omap + 3 * po), q1 = __ldg(omap + 3 * po + 1), q2 = __ldg(omap + 3 * po + 2);
                d1 = q1 >= 0 ? q1 - q0 : 0;
                dz = q2 >= 0 ? q2 - q0 : 0;
                po = q0;
            }
            float* op = out + ((int64_t)nv * d.c_out_total + out_ch0 + oc_base) * pst + po;
            const int64_t astride = d.add_period ? d.add_period : pst;
            const float* ap = add ? add + ((int64_t)nv * d.c_out_total + out_ch0 + oc_base) * astride + (d.add_period ? po % d.add_period : po)
                                  : nullptr;
            const float* bp = bias ? bias + out_ch0 + oc_base : nullptr;
            const uint32_t tbar = tfull0 + 8 * buf, tpar = (uint32_t)(ti >> 1) & 1u;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * pl.n_cta;
            const int ncols = min(pl.n_cta, d.co - oc_base);
            if (16 * colpar >= ncols) {                               // no columns for this warp: it still has to observe the barrier
                mbar_wait(tbar, tpar);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            } else if (NZ) {
                const float nz = valid ? __ldg(adj + (int64_t)nv * d.p_out + pv) : 0.f;       // noise[n, 0, position]
                const float* nwp = nw + out_ch0 + oc_base;
                if (d.act == KGAN_ACT_LRELU) bd_epilogue_tile<KGAN_ACT_LRELU, false, true>(taddr, ncols, colpar, valid, op, pst, ap, astride, bp, lane, tbar, tpar, rnd, 0, 0, nz, nwp);
                else if (d.act == KGAN_ACT_TANH) bd_epilogue_tile<KGAN_ACT_TANH, false, true>(taddr, ncols, colpar, valid, op, pst, ap, astride, bp, lane, tbar, tpar, rnd, 0, 0, nz, nwp);
                else bd_epilogue_tile<KGAN_ACT_NONE, false, true>(taddr, ncols, colpar, valid, op, pst, ap, astride, bp, lane, tbar, tpar, rnd, 0, 0, nz, nwp);
            } else if (omap) bd_epilogue_tile<KGAN_ACT_NONE, true>(taddr, ncols, colpar, valid, op, pst, nullptr, astride, bp, lane, tbar, tpar, rnd, d1, dz);
            else if (d.act == KGAN_ACT_LRELU) bd_epilogue_tile<KGAN_ACT_LRELU>(taddr, ncols, colpar, valid, op, pst, ap, astride, bp, lane, tbar, tpar, rnd);
            else if (d.act == KGAN_ACT_TANH) bd_epilogue_tile<KGAN_ACT_TANH>(taddr, ncols, colpar, valid, op, pst, ap, astride, bp, lane, tbar, tpar, rnd);
            else bd_epilogue_tile<KGAN_ACT_NONE>(taddr, ncols, colpar, valid, op, pst, ap, astride, bp, lane, tbar, tpar, rnd);
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

int tapconv_build_eligible(const kgan_tapconv_desc& d) {
    BuildPlan p;
    return make_build_plan(d, p) ? 1 : 0;
}

static int launch_build(const kgan_tapconv_desc& d, BuildPlan& p, const float* in, const float* wp, const int32_t* pmap, const float* bias,
                        const float* add, float* out, const float* adj, cudaStream_t stream, const int32_t* omap = nullptr, const float* nw = nullptr) {
    CUtensorMap tmap;
    const uint64_t gdim[3] = {(uint64_t)d.p_in, (uint64_t)d.n, (uint64_t)d.c_in_total};
    const uint64_t gstr[2] = {(uint64_t)d.c_in_total * d.p_in * 4, (uint64_t)d.p_in * 4};
    const uint32_t box[3] = {(uint32_t)p.bspan, (uint32_t)p.nseg, 32u};
    if (int e = tma_encode_3d_f32(&tmap, in, gdim, gstr, box, 4)) return e;
    static SmemAttrOnce attr0, attr1;
    const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
    if (nw) {                                                       // noise term in the epilogue: `adj` carries the noise tensor
        static SmemAttrOnce attr2;
        if (int e = ensure_smem(tapconv_fwd_build_k<false, true>, 227 * 1024, attr2, "tapconv_fwd_build (noise) attribute")) return e;
        tapconv_fwd_build_k<false, true><<<grid, BD_THREADS, p.smem_bytes, stream>>>(d, p, tmap, wp, pmap, bias, add, out, adj, omap, nw);
        return check_launch("tapconv_fwd_build (noise)");
    }
    if (p.mix) {
        if (int e = ensure_smem(tapconv_fwd_build_k<true>, 227 * 1024, attr1, "gcn_fwd attribute")) return e;
        tapconv_fwd_build_k<true><<<grid, BD_THREADS, p.smem_bytes, stream>>>(d, p, tmap, wp, pmap, bias, add, out, adj, omap);
    } else {
        if (int e = ensure_smem(tapconv_fwd_build_k<false>, 227 * 1024, attr0, "tapconv_fwd_build attribute")) return e;
        tapconv_fwd_build_k<false><<<grid, BD_THREADS, p.smem_bytes, stream>>>(d, p, tmap, wp, pmap, bias, add, out, adj, omap);
    }
    return check_launch("tapconv_fwd_build");
}

// Graph convolution with the adjacency product inside the GEMM's operand builder (see the head of this file); -1: not eligible
int gcn_fwd_fused(const kgan_tapconv_desc& d, const float* x, const float* wp, const float* adj, const float* bias, const float* add, float* out,
                  const int32_t* omap, cudaStream_t stream) {
    BuildPlan p;
    if (d.mix_v <= 0 || !make_build_plan(d, p)) return -1;
    if ((d.p_out_plane != 0) != (omap != nullptr) || (omap && add)) return -1;      // enlarged output planes <=> scatter store (no `add` then)
    if (reinterpret_cast<uintptr_t>(x) & 15) return -1;
    return launch_build(d, p, x, wp, nullptr, bias, add, out, adj, stream, omap);
}

// -1: not eligible (the caller goes on to the TMA-fed / gather kernels)
int tapconv_fwd_build(const kgan_tapconv_desc& d, const float* in, const float* wp, const int32_t* pmap, const float* bias, const float* add,
                      float* out, cudaStream_t stream) {
    BuildPlan p;
    if (d.mix_v > 0 || !make_build_plan(d, p)) return -1;
    if (reinterpret_cast<uintptr_t>(in) & 15) return -1;
    return launch_build(d, p, in, wp, pmap, bias, add, out, nullptr, stream);
}

// The tap convolution with the generator's noise term in its epilogue: out = act(conv(in) + bias + add + nw[oc] * noise[n, 0, p]) - the
// eval-mode generator block (BatchNorm folded into the weights) as one kernel.  -1: not eligible (no operand-building plan)
int tapconv_fwd_build_noise(const kgan_tapconv_desc& d, const float* in, const float* wp, const int32_t* pmap, const float* bias, const float* add,
                            const float* noise, const float* nw, float* out, cudaStream_t stream) {
    BuildPlan p;
    if (d.mix_v > 0 || d.p_out_plane != 0 || d.add_period != 0 || !make_build_plan(d, p)) return -1;
    if (reinterpret_cast<uintptr_t>(in) & 15) return -1;
    return launch_build(d, p, in, wp, pmap, bias, add, out, noise, stream, nullptr, nw);
}

}  // namespace kgan
