// Shared helpers for libkgan.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kgan.h"

namespace kgan {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return 2;
    }
    return 0;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE setting: remembered per (kernel, device ordinal), so a process
// that drives several GPUs sets it on each of them (a process-wide `static bool` left the second device at the 48 KB default).
struct SmemAttrOnce {
    std::atomic<uint64_t> done{0};
};
template <typename Kern>
inline int ensure_smem(Kern kern, int bytes, SmemAttrOnce& once, const char* what) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return check_launch(what);
    const uint64_t bit = 1ull << (dev & 63);
    if (!(once.done.load(std::memory_order_acquire) & bit)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return check_launch(what);
        once.done.fetch_or(bit, std::memory_order_release);
    }
    return 0;
}

#define KGAN_REQUIRE(cond, ...)          \
    do {                                 \
        if (!(cond)) {                   \
            kgan::set_error(__VA_ARGS__); \
            return 1;                    \
        }                                \
    } while (0)

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == KGAN_ACT_LRELU) return v > 0.f ? v : 0.2f * v;
    if (act == KGAN_ACT_TANH) return tanhf(v);
    return v;
}

// tf32 mode (KGAN_PREC_TF32 / out_tf32 = 1): every kernel that PRODUCES an activation stores it rounded to tf32 (round to nearest,
// ties away from zero == cvt.rna.tf32.f32 for finite values; fp32 container, low 13 mantissa bits zero).  The tensor-core kernels feed
// raw fp32 words to tcgen05.mma kind::tf32, which reads the upper 19 bits only: on such values that read is exact, so the product is
// RN(x) * RN(w) with no bias and tf32-representable data (0/1 masks, all-ones cotangents, small integers) go through exactly.
__device__ __forceinline__ float tf32_out(float v, int rnd) {
    return rnd ? __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u) : v;
}

__device__ __forceinline__ int64_t w_oc_offset(const kgan_tapconv_desc& d, int oc) {
    return d.w_oc_blk ? (int64_t)(oc / d.w_oc_blk) * d.w_ocblk + (int64_t)(oc % d.w_oc_blk) * d.w_oc : (int64_t)oc * d.w_oc;
}

// output plane stride / per-group plane offset (position-block groups, include/kgan.h)
__host__ __device__ inline int out_plane(const kgan_tapconv_desc& d) { return d.p_out_plane ? d.p_out_plane : d.p_out; }

constexpr int kNumSMs = 148;   // B200

bool tapconv_is_thin(const kgan_tapconv_desc& d);   // small contraction: streaming SIMT kernel in both precision modes (tapconv_simt.cu)

// tcgen05 path (tapconv_umma.cu)
int64_t tapconv_tf32_packed_numel(const kgan_tapconv_desc& d);      // 0: shape not eligible for the tensor-core path
int tapconv_pack_tf32(const kgan_tapconv_desc& d, const float* w, float* wp, cudaStream_t stream);
int64_t tapconv_pack_item_bytes();
int tapconv_pack_tf32_batched(int count, const kgan_tapconv_desc* descs, const float* const* w, float* const* wp, void* items_dev, int upload,
                              cudaStream_t stream);
int tapconv_fwd_tf32(const kgan_tapconv_desc& d, const float* in, const float* wp, const int32_t* pmap, const float* bias,
                     const float* add, float* out, cudaStream_t stream);   // -1: not eligible

// TMA-fed variant (tapconv_tma.cu) for descriptors whose taps are pure position shifts (tma_mode != 0); -1: not eligible
int tapconv_tma_eligible(const kgan_tapconv_desc& d);
int tapconv_fwd_tma(const kgan_tapconv_desc& d, const float* in, const float* wp, const int32_t* pmap, const float* bias, const float* add,
                    float* out, cudaStream_t stream);

// `d` with the 1x1 convolution `d2` of a second tensor accumulated into the same accumulator (fused residual branch); -1: not eligible
int tapconv_fwd_tma_res(const kgan_tapconv_desc& d, const kgan_tapconv_desc& d2, const float* in, const float* wp, const float* in2, const float* wp2,
                        const float* bias, const float* bias2, float* out, cudaStream_t stream);
int tapconv_tma_res_eligible(const kgan_tapconv_desc& d, const kgan_tapconv_desc& d2);

// `d` stored through a scatter table (graph conv output written in the unfolded layout of the next temporal conv); -1: not eligible
int tapconv_fwd_tma_scatter(const kgan_tapconv_desc& d, const float* in, const float* wp, const int32_t* omap, const float* bias, float* out,
                            cudaStream_t stream);
int tapconv_tma_scatter_eligible(const kgan_tapconv_desc& d);

// operand-building variant (tapconv_build.cu): raw activations staged by TMA, tap operands gathered in shared memory through the position map
int tapconv_build_eligible(const kgan_tapconv_desc& d);
int tapconv_fwd_build_noise(const kgan_tapconv_desc& d, const float* in, const float* wp, const int32_t* pmap, const float* bias, const float* add,
                            const float* noise, const float* nw, float* out, cudaStream_t stream);
int tapconv_fwd_build(const kgan_tapconv_desc& d, const float* in, const float* wp, const int32_t* pmap, const float* bias, const float* add,
                      float* out, cudaStream_t stream);

int gcn_fwd_fused(const kgan_tapconv_desc& d, const float* x, const float* wp, const float* adj, const float* bias, const float* add, float* out,
                  const int32_t* omap, cudaStream_t stream);   // adjacency product in the operand builder; -1: not eligible

int tapconv_wgrad_tf32_eligible(const kgan_tapconv_desc& d);
int tapconv_wgrad_tma_eligible(const kgan_tapconv_desc& d);   // TMA-fed variant (tapconv_wgrad_tma.cu)
int tapconv_wgrad_tf32(const kgan_tapconv_desc& d, const float* in, const float* gout, const int32_t* pmap, float* dw, int64_t dw_numel,
                       int accumulate, cudaStream_t stream);   // -1: not eligible

}  // namespace kgan
