"""ctypes binding of libkgan.so (include/kgan.h).  There is deliberately no fallback: importing the
package never needs the library, but the first operator call raises if it is missing or if the
device is not a B200-class (sm_100) GPU."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libkgan.so")
MAX_TAPS = 16

ACT_NONE, ACT_LRELU, ACT_TANH = 0, 1, 2
PREC_FP32, PREC_TF32, PREC_X3 = 0, 1, 2


class TapConvDesc(C.Structure):
    """Mirror of `kgan_tapconv_desc` (include/kgan.h)."""
    _fields_ = [
        ("n", C.c_int32), ("c_in_total", C.c_int32), ("p_in", C.c_int32), ("c_out_total", C.c_int32),
        ("p_out", C.c_int32), ("ntap", C.c_int32), ("ck", C.c_int32), ("co", C.c_int32), ("groups", C.c_int32),
        ("g_in", C.c_int32), ("g_out", C.c_int32), ("g_w", C.c_int64), ("w_oc", C.c_int64), ("w_ic", C.c_int64),
        ("w_oc_blk", C.c_int32), ("w_ocblk", C.c_int64),
        ("tap_in_ch", C.c_int32 * MAX_TAPS), ("tap_w_off", C.c_int64 * MAX_TAPS), ("tap_row", C.c_int32 * MAX_TAPS),
        ("pmap_vec_mask", C.c_int32), ("add_period", C.c_int32), ("act", C.c_int32), ("precision", C.c_int32),
        ("tma_mode", C.c_int32), ("tap_shift", C.c_int32 * MAX_TAPS),
        ("p_out_plane", C.c_int32), ("g_pout", C.c_int32), ("stage_span", C.c_int32), ("prefer_staged", C.c_int32),
        ("mix_v", C.c_int32), ("mix_w", C.c_int32), ("mix_l", C.c_int32),
    ]


_F, _I, _V = C.c_void_p, C.c_int, C.c_void_p   # device pointers travel as integers (tensor.data_ptr())
_SIGS = {
    "kgan_version": ([], C.c_int),
    "kgan_last_error": ([], C.c_char_p),
    "kgan_device_ok": ([], C.c_int),
    "kgan_tapconv_fwd": ([C.POINTER(TapConvDesc), _F, _F, _F, _F, _F, _F, _V], C.c_int),
    "kgan_tapconv_tf32_workspace": ([C.POINTER(TapConvDesc)], C.c_int64),
    "kgan_tapconv_pack_tf32": ([C.POINTER(TapConvDesc), _F, _F, _V], C.c_int),
    "kgan_tapconv_pack_item_bytes": ([], C.c_int64),
    "kgan_tapconv_pack_tf32_batched": ([_I, C.POINTER(TapConvDesc), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _F, _I, _V], C.c_int),
    "kgan_tapconv_fwd_tf32": ([C.POINTER(TapConvDesc), _F, _F, _F, _F, _F, _F, _V], C.c_int),
    "kgan_gcn_fused_ok": ([C.POINTER(TapConvDesc)], C.c_int),
    "kgan_gcn_fwd_tf32": ([C.POINTER(TapConvDesc), _F, _F, _F, _F, _F, _F, _F, _V], C.c_int),
    "kgan_tapconv_scatter_ok": ([C.POINTER(TapConvDesc)], C.c_int),
    "kgan_tapconv_fwd_tf32_scatter": ([C.POINTER(TapConvDesc), _F, _F, _F, _F, _F, _V], C.c_int),
    "kgan_tapconv_res_ok": ([C.POINTER(TapConvDesc), C.POINTER(TapConvDesc)], C.c_int),
    "kgan_tapconv_noise_ok": ([C.POINTER(TapConvDesc)], C.c_int),
    "kgan_tapconv_fwd_tf32_noise": ([C.POINTER(TapConvDesc), _F, _F, _F, _F, _F, _F, _F, _F, _V], C.c_int),
    "kgan_tapconv_fwd_tf32_res": ([C.POINTER(TapConvDesc), _F, _F, C.POINTER(TapConvDesc), _F, _F, _F, _F, _F, _V], C.c_int),
    "kgan_tapconv_tma_ok": ([C.POINTER(TapConvDesc)], C.c_int),
    "kgan_tapconv_staged_ok": ([C.POINTER(TapConvDesc)], C.c_int),
    "kgan_tapconv_wgrad_tf32_ok": ([C.POINTER(TapConvDesc)], C.c_int),
    "kgan_tapconv_wgrad_tma_ok": ([C.POINTER(TapConvDesc)], C.c_int),
    "kgan_tapconv_wgrad_tf32": ([C.POINTER(TapConvDesc), _F, _F, _F, _F, C.c_int64, _I, _V], C.c_int),
    "kgan_tapconv_wgrad": ([C.POINTER(TapConvDesc), _F, _F, _F, _F, C.c_int64, _I, _V], C.c_int),
    "kgan_adjmix_fwd": ([_F, _F, _F, _I, _I, _I, _I, _I, _I, _I, _V], C.c_int),
    "kgan_adjmix_bwd_x": ([_F, _F, _F, _I, _I, _I, _I, _I, _I, _I, _V], C.c_int),
    "kgan_adjmix_bwd_x_fused": ([_F, _F, _F, _F, _F, _I, _I, _I, _I, _I, _I, _I, _V], C.c_int),
    "kgan_adjmix_fwd_sel_ok": ([_F, _I, _I, _I, _I, _I, _I], C.c_int),
    "kgan_adjmix_fwd_sel": ([_F, _F, _F, _I, _F, _F, _I, _I, _I, _I, _I, _I, _I, _V], C.c_int),
    "kgan_adjmix_bwd_x_fused_sel": ([_F, _F, _F, _F, _I, _F, _F, _I, _I, _I, _I, _I, _I, _I, _V], C.c_int),
    "kgan_adjmix_bwd_a": ([_F, _F, _F, _I, _I, _I, _I, _I, _I, _V], C.c_int),
    "kgan_adjmix_bwd_a_masked": ([_F, _F, _F, _F, _I, _I, _I, _I, _I, _I, _V], C.c_int),
    "kgan_epilogue_fwd": ([_F, _F, _F, _F, _F, _F, _I, _I, _I, _I, _I, _V], C.c_int),
    "kgan_act_bwd": ([_F, _F, _F, C.c_int64, _I, _I, _V], C.c_int),
    "kgan_chan_reduce": ([_F, _F, _F, _I, _I, _I, _V], C.c_int),
    "kgan_plane_spmm": ([_F, _F, _F, _F, C.c_int64, _I, _I, _I, _I, _V], C.c_int),
    "kgan_plane_sum_t": ([_F, _F, C.c_int64, _I, _I, _I, _V], C.c_int),
    "kgan_label_concat": ([_F, _F, _F, _I, _I, _I, _I, _I, _V], C.c_int),
    "kgan_label_split": ([_F, _F, _F, _I, _I, _I, _I, _I, _V], C.c_int),
    "kgan_bn_workspace": ([_I, _I], C.c_int64),
    "kgan_bn_stats": ([_F, _F, _F, _F, _F, _I, _I, _I, C.c_float, C.c_float, _F, _V], C.c_int),
    "kgan_bn_apply": ([_F, _F, _F, _F, _F, _F, _I, _I, _I, _I, _V], C.c_int),
    "kgan_bn_epilogue_fwd": ([_F, _F, _F, _F, _F, _F, _F, _F, _F, _I, _I, _I, _I, _I, _V], C.c_int),
    "kgan_bn_bwd": ([_F, _F, _F, _F, _F, _F, _F, _F, _I, _I, _I, _I, _F, _V], C.c_int),
    "kgan_adam_step": ([_F, _F, _F, _F, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, _I, C.c_float, _V], C.c_int),
    "kgan_interpolate": ([_F, _F, _F, _F, _I, C.c_int64, _I, _V], C.c_int),
    "kgan_round_tf32": ([_F, _F, C.c_int64, _V], C.c_int),
}
EXPORTS = tuple(_SIGS)

_lib = None
_device_checked = False


def load():
    """Loads libkgan.so (no GPU needed) and declares every prototype."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libkgan.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C %s`. "
                "There is no CPU fallback." % os.path.dirname(LIB_PATH))
        lib = C.CDLL(LIB_PATH)
        for name, (args, res) in _SIGS.items():
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = args, res
        _lib = lib
    return _lib


def lib():
    """Library handle for compute calls: additionally requires a B200-class device."""
    global _device_checked
    l = load()
    if not _device_checked:
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("kinetic-gan_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        torch.cuda.init()
        if not l.kgan_device_ok():
            raise RuntimeError("kinetic-gan_b200 kernels are built for sm_100a only; current device is not sm_100")
        _device_checked = True
    return l


def check(code, what):
    if code != 0:
        raise RuntimeError("%s failed: %s" % (what, load().kgan_last_error().decode()))
