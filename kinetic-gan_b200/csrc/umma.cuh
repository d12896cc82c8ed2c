// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers shared by the tensor-core tap-convolution kernels (sm_100a).
#pragma once
#include "common.cuh"

namespace kgan {

constexpr int UM = 128;                        // UMMA M = positions per CTA
constexpr int UK = 32;                         // contraction elements per pipeline stage (4 MMAs of K = 8)
constexpr int A_STAGE_BYTES = UM * UK * 4;     // 16 KB
constexpr int A_LBO = UM * 16;                 // bytes between the two 16-byte k-chunks of one row group
constexpr int CORE_SBO = 128;                  // bytes between 8-row groups (core matrices are contiguous)
constexpr int UMMA_THREADS = 192;              // warps 0-3: A producers + epilogue, warp 4: MMA, warp 5: weight loader

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
// One lane of a fully active warp.  The producer / MMA-issuer loops run on ALL lanes of their warp in uniform control flow and
// only the asynchronous-issue instructions sit behind this predicate: their operands (addresses, coordinates, descriptors) then
// live in uniform registers.  Written as `if (lane == 0) { loop }` the compiler has to treat every operand as per-thread data and
// wraps each UTMALDG / UTCHMMA / UBLKCP in a uniformisation loop (R2UR + ELECT + BRA.U.ANY, ~13 instructions per issue): measured
// on the weight-gradient kernel, the single issuing thread - not memory, not the tensor pipe - bounded the pipeline.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
// warp index as a value the compiler knows to be warp-uniform
__device__ __forceinline__ int uniform_warp_id() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// TMA prefetch of a 3-D box into L2 (no shared-memory destination, no barrier): widens the window of bytes in flight beyond what the
// shared-memory ring can hold - the TMA-fed kernels are bound by (bytes in flight per SM) / (loaded HBM latency, ~3 us)
__device__ __forceinline__ void tma_prefetch_3d(const void* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1),
                 "r"(c2)
                 : "memory");
}
// predicated read-only global load as a volatile asm: keeps program order relative to the barrier waits, so a whole
// stage of gathers is in flight before the thread blocks
__device__ __forceinline__ float ldg_pred(const float* ptr, bool pred) {
    float v;
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.b32 q, %2, 0;\n\t"
        "mov.f32 %0, 0f00000000;\n\t"
        "@q ld.global.nc.f32 %0, [%1];\n\t"
        "}"
        : "=f"(v)
        : "l"(ptr), "r"((int)pred));
    return v;
}
// unpredicated read-only load as volatile asm (same ordering guarantee as ldg_pred)
__device__ __forceinline__ float ldg_nc(const float* ptr) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(ptr));
    return v;
}
// fp32 -> tf32 with round-to-nearest, ties away from zero (== cvt.rna.tf32.f32 for finite inputs) in two integer ops:
// add half an ulp of the 10-bit mantissa to the magnitude bits, clear the 13 dropped bits
__device__ __forceinline__ uint32_t to_tf32_fast(float v) { return (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
// K-major, no swizzle (LayoutType::SWIZZLE_NONE = 0), descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128
__device__ __forceinline__ uint32_t instr_desc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane; the caller issues tcgen05.wait::ld before using r
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace kgan
