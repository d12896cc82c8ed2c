// Weight gradient of the tap convolution on the tensor cores (tcgen05.mma kind::tf32, fp32 accumulators in TMEM):
//     dW[tap][oc][ic] = sum_pos gout[oc, pos] * in[(tap, ic), pmap_tap(pos)]
//   D[M = 128 output channels][N = n_ic input channels] per tap - the ntap accumulators sit side by side in TMEM;
//   K = output positions.  Both operands are K-major in global memory already (positions are contiguous per channel):
//   a 16-byte chunk of 4 consecutive positions of one channel is exactly one k-chunk of one row of the canonical
//   no-swizzle K-major core-matrix layout.  The producers therefore issue cp.async (LDGSTS) copies straight from global
//   to shared memory - no registers, no waiting: up to S stages (the whole shared memory) of loads are in flight per
//   SM, which is what this HBM/latency-bound kernel needs.  A one-row pad per k-chunk (LBO = (rows+1)*16 B) spreads
//   the scattered 16-byte writes over the banks.  Taps whose position map is not 4-contiguous (joint / frame
//   selection, odd temporal shifts) fall back to four 4-byte copies per chunk.  gout (A operand) is staged once per K
//   tile and reused by all taps.  Split-K over one wave of CTAs; the epilogue adds into dW with 16-byte vector
//   reductions (REDG.ADD.F32x4) where the weight layout allows.
//   Operands reach the tensor core as raw fp32 words: kind::tf32 reads the upper 19 bits.  In tf32 mode every kernel of the
//   library stores its activations already rounded to tf32 (round to nearest, common.cuh tf32_out), so that read is exact:
//   no bias, and tf32-representable data (masks, all-ones cotangents) give exact products.  A tensor that did not come
//   from a libkgan kernel is truncated instead (one-sided error < 2^-10 per element); callers that care pass it through
//   kgan_round_tf32 first.
// Warp roles: warps 0-7 cp.async producers, then epilogue; warp 8 MMA issuer / TMEM owner.
#include "umma.cuh"

namespace kgan {

constexpr int WG_PRODUCER_WARPS = 8;
constexpr int WG_PRODUCERS = 32 * WG_PRODUCER_WARPS;
constexpr int WG_THREADS = WG_PRODUCERS + 32;

struct WgradPlan {
    int n_ic;         // input channels (UMMA N) per CTA, multiple of 16, <= 256
    int ic_tiles, oc_tiles;
    int tmem_cols;
    int stages;
    int a_bytes, b_bytes;   // per stage: A image, one tap's B image
    int nchunks;
    int64_t chunk;    // positions per split-K chunk (multiple of UK)
    int smem_bytes;
};

static bool make_wgrad_plan(const kgan_tapconv_desc& d, WgradPlan& p) {
    if (d.w_oc_blk != 0 || d.ntap > 8) return false;         // any channel count: ragged M and N are zero filled
    if ((d.p_out & 3) || (d.p_in & 3)) return false;           // 16-byte chunks must not straddle samples
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
    const int stage = p.a_bytes + d.ntap * p.b_bytes;
    p.stages = (200 * 1024) / stage;
    if (p.stages > 8) p.stages = 8;
    if (p.stages < 2) return false;
    const int64_t ktiles = ceil_div64(total, UK);
    const int tiles = p.ic_tiles * p.oc_tiles * d.groups;
    int64_t nchunks = kNumSMs / tiles;                         // floor: the grid must fit in ONE wave (no tail CTAs)
    if (nchunks > ktiles / 4) nchunks = ktiles / 4;
    if (nchunks < 1) nchunks = 1;
    p.chunk = ceil_div64(ktiles, nchunks) * UK;
    p.nchunks = (int)ceil_div64(total, p.chunk);
    if ((int64_t)p.nchunks * d.groups > 65535) return false;
    p.smem_bytes = p.stages * stage + 256;
    return true;
}

// 16-byte (or 4-byte) asynchronous global -> shared copy; src_bytes = 0 zero-fills the destination without reading
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1) tapconv_wgrad_umma(const __grid_constant__ kgan_tapconv_desc d, const __grid_constant__ WgradPlan pl,
                                                                    const float* __restrict__ in, const float* __restrict__ gout,
                                                                    const int32_t* __restrict__ pmap, float* __restrict__ dw) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
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
            mbar_init(full0 + 8 * s, WG_PRODUCERS);            // one (asynchronous) arrival per producer thread
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
        // ===== producers: thread = (k-chunk of 4 positions, row slot); 8 consecutive lanes copy 128 contiguous bytes of a row =====
        const int chunk = threadIdx.x & 7, rslot = threadIdx.x >> 3;      // rows rslot, rslot + 32, ...
        const int in_ch0 = g * d.g_in + ic0, out_ch0 = g * d.g_out + oc0;
        const int a_rows = min(UM, d.co - oc0), b_rows = min(pl.n_ic, d.ck - ic0);   // rows that exist; the rest is zero filled
        for (int it = 0; it < iters; ++it) {
            const int s = it % S;
            const uint32_t ph = (uint32_t)(it / S) & 1u;
            const uint32_t pos = (uint32_t)(pbeg + (int64_t)it * UK) + chunk * 4;        // total_pos < 2^31 (plan); p_out % 4 == 0
            const bool valid = pos < (uint32_t)pend;
            const uint32_t nn = valid ? pos / (uint32_t)d.p_out : 0u, p = valid ? pos - nn * (uint32_t)d.p_out : 0u;
            mbar_wait(empty0 + 8 * s, ph ^ 1u);
            const uint32_t st_a = smem_u32(smem + (size_t)s * stage_bytes);
            {   // A: gout rows
                const float* gp = gout + ((int64_t)nn * d.c_out_total + out_ch0 + rslot) * d.p_out + p;
                uint32_t dst = st_a + chunk * a_lbo + rslot * 16;
#pragma unroll
                for (int i = 0; i < UM / 32; ++i, gp += (int64_t)32 * d.p_out, dst += 32 * 16) {
                    const bool ok = valid && rslot + 32 * i < a_rows;
                    cp_async16(dst, ok ? gp : gout, ok ? 16u : 0u);
                }
            }
            for (int tap = 0; tap < d.ntap; ++tap) {     // B: input rows through the tap's position map
                const int32_t* pm = pmap + (int64_t)d.tap_row[tap] * d.p_out + p;
                const float* xb = in + ((int64_t)nn * d.c_in_total + in_ch0 + d.tap_in_ch[tap] + rslot) * d.p_in;
                uint32_t dst = st_a + pl.a_bytes + tap * pl.b_bytes + chunk * b_lbo + rslot * 16;
                // shift form (tma_mode 1): the source position is arithmetic - no dependent position-map load in front of the copies
                const bool shift_form = d.tma_mode == 1;
                const int q0 = (int)p + d.tap_shift[tap];
                if ((d.pmap_vec_mask >> d.tap_row[tap]) & 1) {
                    const int src = !valid ? -1 : shift_form ? ((q0 >= 0 && q0 < d.p_in) ? q0 : -1) : __ldg(pm);
                    for (int r = rslot; r < pl.n_ic; r += 32, xb += (int64_t)32 * d.p_in, dst += 32 * 16) {
                        const bool ok = src >= 0 && r < b_rows;
                        cp_async16(dst, ok ? xb + src : in, ok ? 16u : 0u);
                    }
                } else {
                    int src[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        src[e] = !valid ? -1 : shift_form ? ((q0 + e >= 0 && q0 + e < d.p_in) ? q0 + e : -1) : __ldg(pm + e);
                    for (int r = rslot; r < pl.n_ic; r += 32, xb += (int64_t)32 * d.p_in, dst += 32 * 16) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const bool ok = src[e] >= 0 && r < b_rows;
                            cp_async4(dst + 4 * e, ok ? xb + src[e] : in, ok ? 4u : 0u);
                        }
                    }
                }
            }
            // arrives on full[s] when every copy issued by this thread so far has landed
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full0 + 8 * s) : "memory");
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
        auto fix = [](uint32_t bits) { return __uint_as_float(bits); };
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
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else {
        // ===== MMA issuer (whole warp waits in uniform control flow, one elected lane issues: umma.cuh elect_one) =====
        const bool leader = elect_one();
        const uint32_t idesc = instr_desc_tf32(pl.n_ic);
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(full0 + 8 * s, ph);                      // acquire: every producer's copies for this stage have landed
            // cp.async wrote through the generic proxy; tcgen05.mma reads operands through the async proxy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (leader) {
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
            __syncwarp();
            if (++s == S) {
                s = 0;
                ph ^= 1u;
            }
        }
        if (leader) umma_commit(accfull);
        __syncwarp();
    }
    __syncthreads();
    if (warp == WG_PRODUCER_WARPS) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(pl.tmem_cols) : "memory");
    }
}

int tapconv_wgrad_tma_eligible(const kgan_tapconv_desc& d);
int tapconv_wgrad_tma(const kgan_tapconv_desc& d, const float* in, const float* gout, float* dw, int64_t dw_numel, int accumulate, cudaStream_t stream);

int tapconv_wgrad_tf32_eligible(const kgan_tapconv_desc& d) {
    WgradPlan p;
    if (d.precision == KGAN_PREC_TF32X3) return tapconv_wgrad_tma_eligible(d);      // the fp32-accurate split lives in the TMA-fed kernel only
    return (make_wgrad_plan(d, p) || tapconv_wgrad_tma_eligible(d)) ? 1 : 0;
}

int tapconv_wgrad_tf32(const kgan_tapconv_desc& d, const float* in, const float* gout, const int32_t* pmap, float* dw, int64_t dw_numel,
                       int accumulate, cudaStream_t stream) {
    {
        const int rt = tapconv_wgrad_tma(d, in, gout, dw, dw_numel, accumulate, stream);
        if (rt != -1) return rt;
    }
    if (d.precision == KGAN_PREC_TF32X3) return -1;
    WgradPlan p;
    if (!make_wgrad_plan(d, p)) return -1;
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(gout)) & 15) {
        set_error("tapconv_wgrad_tf32: in / gout must be 16-byte aligned");
        return 1;
    }
    static SmemAttrOnce attr;
    if (int e = ensure_smem(tapconv_wgrad_umma, 227 * 1024, attr, "tapconv_wgrad_tf32 attribute")) return e;
    if (!accumulate && cudaMemsetAsync(dw, 0, sizeof(float) * dw_numel, stream) != cudaSuccess) return check_launch("tapconv_wgrad_tf32 memset");
    dim3 grid(p.ic_tiles, p.oc_tiles, (unsigned)(d.groups * p.nchunks));
    tapconv_wgrad_umma<<<grid, WG_THREADS, p.smem_bytes, stream>>>(d, p, in, gout, pmap, dw);
    return check_launch("tapconv_wgrad_tf32");
}

}  // namespace kgan
