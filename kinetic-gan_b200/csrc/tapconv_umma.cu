// tcgen05 / TMEM path of the tap convolution (KGAN_PREC_TF32).  Placeholder until the UMMA kernel lands:
// reports "not eligible" so the caller runs the exact fp32 path.
#include "common.cuh"

namespace kgan {
int tapconv_fwd_tf32(const kgan_tapconv_desc&, const float*, const float*, const int32_t*, const float*, const float*, float*, cudaStream_t) {
    return -1;
}
}  // namespace kgan
