// Weight gradient of the tap convolution on the tensor cores (tcgen05.mma kind::tf32, fp32 accumulators in TMEM):
//     dW[tap][oc][ic] = sum_pos gout[oc, pos] * in[(tap, ic), pmap_tap(pos)]
//   D[M = 128 output channels][N = n_ic input channels] per tap - the ntap accumulators sit side by side in TMEM;
//   K = output positions.  Both operands are K-major in global memory already (positions are contiguous per channel),
//   so producer lanes run along positions (coalesced 128-byte loads) and scatter 4-byte stores into the K-major
//   core-matrix layout; a one-row pad per k-chunk (LBO = (rows+1)*16 B) makes those transposing stores bank-conflict
//   free.  gout (A operand) is staged once per K tile and reused by all taps.  Split-K over CTAs, fp32 atomics into dW.
// Warp roles: warps 0-7 producers (64 gathers in flight per thread) and epilogue, warp 8 MMA issuer / TMEM owner.
#include "umma.cuh"

namespace kgan {

constexpr int WG_PRODUCER_WARPS = 8;
constexpr int WG_THREADS = 32 * (WG_PRODUCER_WARPS + 1);
constexpr int WG_UNIT = 16;                      // rows a warp loads per unit (= 128 rows of a 128-row image)

struct WgradPlan {
    int n_ic;         // input channels (UMMA N) per CTA, multiple of 16, <= 256
    int ic_tiles, oc_tiles;
    int tmem_cols;
    int stages;
    int a_bytes, b_bytes;   // per stage: A image, one tap's B image
    int b_units;      // 16-row units per warp per tap: ceil(n_ic / 128)
    int nchunks;
    int64_t chunk;    // positions per split-K chunk (multiple of UK)
    int smem_bytes;
};

static bool make_wgrad_plan(const kgan_tapconv_desc& d, WgradPlan& p) {
    if (d.w_oc_blk != 0 || d.ntap > 8) return false;        // any channel count: ragged M and N are zero padded
    const int64_t total = (int64_t)d.n * d.p_out;
    if (total < 1024 || total >= (1ll << 31) - UK) return false;
    int n_max = (512 / d.ntap) / 16 * 16;
    if (n_max > 256) n_max = 256;
    if (d.ntap >= 3 && n_max > 128) n_max = 128;
    p.n_ic = round_up(d.ck, 16) < n_max ? round_up(d.ck, 16) : n_max;
    p.ic_tiles = ceil_div(d.ck, p.n_ic);
    p.oc_tiles = ceil_div(d.co, UM);
    p.tmem_cols = 32;
    while (p.tmem_cols < d.ntap * p.n_ic) p.tmem_cols *= 2;
    p.a_bytes = 8 * (UM + 1) * 16;
    p.b_bytes = 8 * (p.n_ic + 1) * 16;
    p.b_units = ceil_div(p.n_ic, WG_PRODUCER_WARPS * WG_UNIT);
    const int stage = p.a_bytes + d.ntap * p.b_bytes;
    // thin layers (<= 256 TMEM columns, small stages) run two CTAs per SM: the kernel is gather-latency bound there
    const int per_sm = (p.tmem_cols <= 256 && 2 * stage <= 100 * 1024) ? 2 : 1;
    p.stages = ((per_sm == 2 ? 100 : 200) * 1024) / stage;
    if (p.stages > 4) p.stages = 4;
    if (p.stages < 2) return false;
    const int64_t ktiles = ceil_div64(total, UK);
    const int tiles = p.ic_tiles * p.oc_tiles * d.groups;
    int64_t nchunks = ((int64_t)per_sm * kNumSMs) / tiles;        // floor: the grid must fit in ONE wave (no tail CTAs)
    if (nchunks > ktiles / 4) nchunks = ktiles / 4;
    if (nchunks < 1) nchunks = 1;
    p.chunk = ceil_div64(ktiles, nchunks) * UK;
    p.nchunks = (int)ceil_div64(total, p.chunk);
    if ((int64_t)p.nchunks * d.groups > 65535) return false;
    p.smem_bytes = p.stages * stage + 256;
    return true;
}

__global__ void __launch_bounds__(WG_THREADS) tapconv_wgrad_umma(const __grid_constant__ kgan_tapconv_desc d, const __grid_constant__ WgradPlan pl,
                                                                 const float* __restrict__ in, const float* __restrict__ gout,
                                                                 const int32_t* __restrict__ pmap, float* __restrict__ dw) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = pl.stages;
    const int stage_bytes = pl.a_bytes + d.ntap * pl.b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);           // full[S], empty[S], accfull
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S), accfull = smem_u32(bars + 2 * S);
    const uint32_t a_lbo = (UM + 1) * 16, b_lbo = (pl.n_ic + 1) * 16;

    const int ic0 = blockIdx.x * pl.n_ic, oc0 = blockIdx.y * UM;
    const int g = blockIdx.z / pl.nchunks, ch = blockIdx.z % pl.nchunks;
    const int64_t total_pos = (int64_t)d.n * d.p_out;
    const int64_t pbeg = (int64_t)ch * pl.chunk, pend = min(total_pos, pbeg + pl.chunk);
    const int iters = (int)ceil_div64(pend - pbeg, UK);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full0 + 8 * s, 32 * WG_PRODUCER_WARPS);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(accfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WG_PRODUCER_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(pl.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < WG_PRODUCER_WARPS) {
        // ===== producers: lane = position inside the K tile; warp w stages rows w, w+8, ... of every image =====
        // lean inner loops: one row pointer advanced by 8 planes per load, immediate shared-memory offsets, predicates only
        // on ragged tiles; every gather of a batch is issued before the first conversion
        const int in_ch0 = g * d.g_in + ic0, out_ch0 = g * d.g_out + oc0;
        const uint32_t kbyte = (uint32_t)(lane >> 2) * 0u + (uint32_t)(lane & 3) * 4;     // byte offset inside the 16-byte k-chunk
        const uint32_t kchunk = (uint32_t)(lane >> 2);
        const bool a_full = oc0 + UM <= d.co, b_full = ic0 + pl.n_ic <= d.ck && (pl.n_ic & 127) == 0;
        const int64_t a_step = (int64_t)WG_PRODUCER_WARPS * d.p_out, b_step = (int64_t)WG_PRODUCER_WARPS * d.p_in;
        for (int it = 0; it < iters; ++it) {
            const int s = it % S;
            const uint32_t ph = (uint32_t)(it / S) & 1u;
            const uint32_t pos = (uint32_t)(pbeg + (int64_t)it * UK) + lane;             // total_pos < 2^31 (plan)
            const bool valid = pos < (uint32_t)pend;
            const uint32_t nn = valid ? pos / (uint32_t)d.p_out : 0u, p = valid ? pos - nn * (uint32_t)d.p_out : 0u;
            const uint32_t st_a = smem_u32(smem + (size_t)s * stage_bytes);
            // batches of 16 rows per warp: batch 0 = gout (A), then (tap, half) input batches (B); the gathers of up to four
            // batches (64 per thread) are issued back to back before anything is converted or stored
            const int nbatch = 1 + d.ntap * pl.b_units;
            bool waited = false;
            for (int b0 = 0; b0 < nbatch; b0 += 4) {
                float v[4][WG_UNIT];
#pragma unroll
                for (int qb = 0; qb < 4; ++qb) {
                    const int bi = b0 + qb;
                    if (bi >= nbatch) break;
                    if (bi == 0) {
                        const float* gp = gout + ((int64_t)nn * d.c_out_total + out_ch0 + warp) * d.p_out + p;
                        if (valid && a_full) {
#pragma unroll
                            for (int j = 0; j < WG_UNIT; ++j, gp += a_step) v[qb][j] = ldg_nc(gp);
                        } else {
#pragma unroll
                            for (int j = 0; j < WG_UNIT; ++j, gp += a_step) v[qb][j] = ldg_pred(gp, valid && oc0 + warp + WG_PRODUCER_WARPS * j < d.co);
                        }
                    } else {
                        const int tap = (bi - 1) / pl.b_units, half = (bi - 1) - tap * pl.b_units;
                        const int src = valid ? __ldg(pmap + (int64_t)d.tap_row[tap] * d.p_out + p) : -1;
                        const int r0 = warp + WG_PRODUCER_WARPS * WG_UNIT * half;
                        const float* xp = in + ((int64_t)nn * d.c_in_total + in_ch0 + d.tap_in_ch[tap] + r0) * d.p_in + src;
                        if (src >= 0 && b_full) {
#pragma unroll
                            for (int j = 0; j < WG_UNIT; ++j, xp += b_step) v[qb][j] = ldg_nc(xp);
                        } else {
#pragma unroll
                            for (int j = 0; j < WG_UNIT; ++j, xp += b_step) {
                                const int r = r0 + WG_PRODUCER_WARPS * j;
                                v[qb][j] = ldg_pred(xp, src >= 0 && r < pl.n_ic && ic0 + r < d.ck);
                            }
                        }
                    }
                }
                if (!waited) {
                    mbar_wait(empty0 + 8 * s, ph ^ 1u);
                    waited = true;
                }
#pragma unroll
                for (int qb = 0; qb < 4; ++qb) {
                    const int bi = b0 + qb;
                    if (bi >= nbatch) break;
                    if (bi == 0) {
                        const uint32_t dst = st_a + kchunk * a_lbo + kbyte + warp * 16;
#pragma unroll
                        for (int j = 0; j < WG_UNIT; ++j)
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + j * (WG_PRODUCER_WARPS * 16)), "r"(to_tf32_fast(v[qb][j])) : "memory");
                    } else {
                        const int tap = (bi - 1) / pl.b_units, half = (bi - 1) - tap * pl.b_units;
                        const int r0 = warp + WG_PRODUCER_WARPS * WG_UNIT * half;
                        const uint32_t dst = st_a + pl.a_bytes + tap * pl.b_bytes + kchunk * b_lbo + kbyte + r0 * 16;
#pragma unroll
                        for (int j = 0; j < WG_UNIT; ++j)
                            if (b_full || r0 + WG_PRODUCER_WARPS * j < pl.n_ic)
                                asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + j * (WG_PRODUCER_WARPS * 16)), "r"(to_tf32_fast(v[qb][j])) : "memory");
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(full0 + 8 * s);
        }
        // ===== epilogue: TMEM lane = output channel row (warp % 4 selects the lane quarter, warp / 4 the column half) =====
        mbar_wait(accfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3, colhalf = warp >> 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int oc = oc0 + quarter * 32 + lane;
        float* wb = dw + (int64_t)g * d.g_w + (int64_t)oc * d.w_oc;
        // 16-byte vector reductions (REDG.ADD.F32x4) wherever 4 consecutive results are contiguous in dW:
        //   w_ic == 1          (graph conv / 1x1 / linear weights): 16 input channels of one tap are contiguous
        //   taps innermost     (temporal conv weights (C_out, C_in, 3, 1)): the 3 taps x 16 channels interleave to 48 floats
        bool taps_inner = d.ntap == 3 && d.w_ic == 3;
        for (int tp = 0; tp < d.ntap; ++tp) taps_inner = taps_inner && d.tap_w_off[tp] == d.tap_w_off[0] + tp;
        if (taps_inner) {
            for (int col0 = colhalf * 16; col0 < pl.n_ic; col0 += 32) {
                if (ic0 + col0 >= d.ck) break;                       // warp-uniform
                uint32_t r[3][16];
                tmem_ld16_nowait(taddr + col0, r[0]);
                tmem_ld16_nowait(taddr + pl.n_ic + col0, r[1]);
                tmem_ld16_nowait(taddr + 2 * pl.n_ic + col0, r[2]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (oc < d.co) {
                    float* dst = wb + d.tap_w_off[0] + (int64_t)(ic0 + col0) * 3;
                    if (ic0 + col0 + 16 <= d.ck && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                        for (int v4 = 0; v4 < 12; ++v4) {
                            float4 q;
                            q.x = __uint_as_float(r[(4 * v4 + 0) % 3][(4 * v4 + 0) / 3]);
                            q.y = __uint_as_float(r[(4 * v4 + 1) % 3][(4 * v4 + 1) / 3]);
                            q.z = __uint_as_float(r[(4 * v4 + 2) % 3][(4 * v4 + 2) / 3]);
                            q.w = __uint_as_float(r[(4 * v4 + 3) % 3][(4 * v4 + 3) / 3]);
                            atomicAdd(reinterpret_cast<float4*>(dst) + v4, q);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (ic0 + col0 + j < d.ck) {
#pragma unroll
                                for (int tp = 0; tp < 3; ++tp) atomicAdd(dst + j * 3 + tp, __uint_as_float(r[tp][j]));
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
                    if (oc < d.co) {
                        float* dst = wb + d.tap_w_off[tap] + (int64_t)(ic0 + col0) * d.w_ic;
                        if (d.w_ic == 1 && ic0 + col0 + 16 <= d.ck && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                            for (int v4 = 0; v4 < 4; ++v4)
                                atomicAdd(reinterpret_cast<float4*>(dst) + v4,
                                          make_float4(__uint_as_float(r[4 * v4]), __uint_as_float(r[4 * v4 + 1]), __uint_as_float(r[4 * v4 + 2]),
                                                      __uint_as_float(r[4 * v4 + 3])));
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (ic0 + col0 + j < d.ck) atomicAdd(dst + (int64_t)j * d.w_ic, __uint_as_float(r[j]));
                        }
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = instr_desc_tf32(pl.n_ic);
            for (int it = 0; it < iters; ++it) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                mbar_wait(full0 + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
                for (int tap = 0; tap < d.ntap; ++tap) {
                    const uint32_t b_addr = a_addr + pl.a_bytes + tap * pl.b_bytes;
#pragma unroll
                    for (int j = 0; j < UK / 8; ++j)
                        umma_tf32(tmem_base + tap * pl.n_ic, smem_desc(a_addr + j * 2 * a_lbo, a_lbo, CORE_SBO),
                                  smem_desc(b_addr + j * 2 * b_lbo, b_lbo, CORE_SBO), idesc, (it > 0 || j > 0) ? 1u : 0u);
                }
                umma_commit(empty0 + 8 * s);
            }
            umma_commit(accfull);
        }
    }
    __syncthreads();
    if (warp == WG_PRODUCER_WARPS) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(pl.tmem_cols) : "memory");
    }
}

int tapconv_wgrad_tf32_eligible(const kgan_tapconv_desc& d) {
    WgradPlan p;
    return make_wgrad_plan(d, p) ? 1 : 0;
}

int tapconv_wgrad_tf32(const kgan_tapconv_desc& d, const float* in, const float* gout, const int32_t* pmap, float* dw, int64_t dw_numel,
                       cudaStream_t stream) {
    WgradPlan p;
    if (!make_wgrad_plan(d, p)) return -1;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(tapconv_wgrad_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return check_launch("tapconv_wgrad_tf32 attribute");
        attr_set = true;
    }
    if (cudaMemsetAsync(dw, 0, sizeof(float) * dw_numel, stream) != cudaSuccess) return check_launch("tapconv_wgrad_tf32 memset");
    dim3 grid(p.ic_tiles, p.oc_tiles, (unsigned)(d.groups * p.nchunks));
    tapconv_wgrad_umma<<<grid, WG_THREADS, p.smem_bytes, stream>>>(d, p, in, gout, pmap, dw);
    return check_launch("tapconv_wgrad_tf32");
}

}  // namespace kgan
