"""ctypes binding of libw2c.so — the C ABI declared in include/w2c.h.

There is no fallback: if the library is missing it is built with nvcc; if that fails, or a call returns an error,
an exception is raised. Nothing here computes on the CPU.
"""
import ctypes
import os
import threading

from . import build as _build

c_i32 = ctypes.c_int32
c_f32 = ctypes.c_float
c_vp = ctypes.c_void_p

ACT_BF16, ACT_BF16X2, ACT_FP16, ACT_FP16X2 = 0, 1, 2, 3
OUT_NHWC, OUT_NCHW_F32 = 0, 1
IMPL_TCGEN05, IMPL_SIMT, IMPL_TC_TAPS, IMPL_TC_PERSIST = 0, 1, 2, 4
CONV3X3_S1, CONV3X3_S2, DECONV3X3_S2, CONV1X1_S1, CONV1X1_S2 = 0, 1, 2, 3, 4
FUSE_SOFTMAX, FUSE_ACTIVATED, FUSE_ARGMAX = 0, 1, 2
GT_U8, GT_I64 = 0, 1


class ConvArgs(ctypes.Structure):
    """struct w2c_conv_args (include/w2c.h)."""
    _fields_ = [
        ("x", c_vp), ("w", c_vp), ("scale", c_vp), ("shift", c_vp), ("residual", c_vp), ("y", c_vp),
        ("n", c_i32), ("h_in", c_i32), ("w_in", c_i32),
        ("cin", c_i32), ("cout", c_i32),
        ("x_cstride", c_i32), ("x_coffset", c_i32),
        ("y_cstride", c_i32), ("y_coffset", c_i32),
        ("kind", c_i32), ("relu", c_i32), ("act", c_i32), ("out_fmt", c_i32), ("impl", c_i32), ("block_n", c_i32),
        ("labels", c_vp),
        ("passes", c_i32),
        ("bn_sums", c_vp),
    ]


class EncHeadArgs(ctypes.Structure):
    """struct w2c_enc_head_args (include/w2c.h)."""
    _fields_ = [
        ("x", c_vp), ("lut", c_vp), ("w1", c_vp), ("scale1", c_vp), ("shift1", c_vp), ("w2", c_vp), ("scale2", c_vp),
        ("shift2", c_vp), ("y", c_vp),
        ("x_u8", c_i32), ("b", c_i32), ("n_agents", c_i32), ("c_total", c_i32), ("c_first", c_i32),
        ("h", c_i32), ("w", c_i32), ("act", c_i32), ("y_cstride", c_i32), ("y_coffset", c_i32),
    ]


class AttnArgs(ctypes.Structure):
    """struct w2c_attn_args (include/w2c.h)."""
    _fields_ = [
        ("keys", c_vp), ("queries", c_vp), ("wq", c_vp), ("bq", c_vp), ("val", c_vp), ("fused", c_vp),
        ("prob_out", c_vp), ("coef_out", c_vp), ("action", c_vp), ("connect", c_vp),
        ("b_sz", c_i32), ("n_k", c_i32), ("n_q", c_i32),
        ("k_dim", c_i32), ("q_dim", c_i32),
        ("hw", c_i32), ("c", c_i32),
        ("fused_cstride", c_i32), ("fused_coffset", c_i32),
        ("act", c_i32), ("mode", c_i32), ("sparse", c_i32), ("mask_self", c_i32),
        ("temperature", c_f32), ("diag_bias", c_f32), ("thresh", c_f32),
        ("q_first", c_i32), ("q_count", c_i32), ("agents_per_rank", c_i32),
        ("keys_rank_stride", ctypes.c_int64), ("queries_rank_stride", ctypes.c_int64),
        ("val_rank_stride", ctypes.c_int64),
    ]


class MlpHead(ctypes.Structure):
    """struct w2c_mlp_head (include/w2c.h)."""
    _fields_ = [("w0", c_vp), ("b0", c_vp), ("w1", c_vp), ("b1", c_vp), ("w2", c_vp), ("b2", c_vp), ("out", c_vp),
                ("out_dim", c_i32)]


class WgradArgs(ctypes.Structure):
    """struct w2c_wgrad_args (include/w2c.h)."""
    _fields_ = [("x", c_vp), ("dy", c_vp), ("dw", c_vp), ("n", c_i32), ("h_in", c_i32), ("w_in", c_i32),
                ("cin", c_i32), ("cout", c_i32), ("x_cstride", c_i32), ("x_coffset", c_i32), ("dy_cstride", c_i32),
                ("dy_coffset", c_i32), ("kind", c_i32), ("act_x", c_i32), ("act_dy", c_i32), ("passes", c_i32)]


class BnBwdArgs(ctypes.Structure):
    """struct w2c_bn_bwd_args (include/w2c.h)."""
    _fields_ = [("dy", c_vp), ("y", c_vp), ("z", c_vp), ("dz", c_vp), ("dres", c_vp), ("n_px", ctypes.c_int64),
                ("c", c_i32), ("dy_cstride", c_i32), ("dy_coffset", c_i32), ("y_cstride", c_i32), ("y_coffset", c_i32),
                ("z_cstride", c_i32), ("z_coffset", c_i32), ("dz_cstride", c_i32), ("dz_coffset", c_i32),
                ("dres_cstride", c_i32), ("dres_coffset", c_i32), ("act_f", c_i32), ("act_g", c_i32), ("relu", c_i32),
                ("gamma", c_vp), ("stats", c_vp), ("dgamma", c_vp), ("dbeta", c_vp), ("sums_ws", c_vp),
                ("coef_ws", c_vp), ("fwd_scale", c_vp), ("fwd_shift", c_vp)]


class AttnBwdArgs(ctypes.Structure):
    """struct w2c_attn_bwd_args (include/w2c.h)."""
    _fields_ = [("keys", c_vp), ("queries", c_vp), ("wq", c_vp), ("bq", c_vp), ("val", c_vp), ("dfused", c_vp),
                ("prob", c_vp), ("dval", c_vp), ("dkeys", c_vp), ("dqueries", c_vp), ("dwq", c_vp), ("dbq", c_vp),
                ("dp_ws", c_vp), ("b_sz", c_i32), ("n_k", c_i32), ("n_q", c_i32), ("k_dim", c_i32), ("q_dim", c_i32),
                ("hw", c_i32), ("c", c_i32), ("dfused_cstride", c_i32), ("dfused_coffset", c_i32), ("act_f", c_i32),
                ("act_g", c_i32), ("sparse", c_i32), ("dval_accumulate", c_i32), ("temperature", c_f32)]


class MlpHeadGrad(ctypes.Structure):
    """struct w2c_mlp_head_grad (include/w2c.h)."""
    _fields_ = [("dout", c_vp), ("dw0", c_vp), ("db0", c_vp), ("dw1", c_vp), ("db1", c_vp), ("dw2", c_vp),
                ("db2", c_vp)]


class PackItem(ctypes.Structure):
    """struct w2c_pack_item (include/w2c.h)."""
    _fields_ = [("w", c_vp), ("packed", c_vp), ("cout", c_i32), ("cin_real", c_i32), ("cin", c_i32), ("ntaps", c_i32),
                ("transposed", c_i32), ("flip", c_i32)]


class FoldItem(ctypes.Structure):
    """struct w2c_fold_item (include/w2c.h)."""
    _fields_ = [("conv_bias", c_vp), ("gamma", c_vp), ("beta", c_vp), ("mean", c_vp), ("var", c_vp), ("scale", c_vp),
                ("shift", c_vp), ("eps", c_f32), ("cout", c_i32)]


# symbol -> (restype, argtypes); every symbol include/w2c.h declares must be listed here (tests check both ways)
_SIGNATURES = {
    "w2c_version": (ctypes.c_int, []),
    "w2c_last_error": (ctypes.c_char_p, []),
    "w2c_launch_count": (ctypes.c_uint64, []),
    "w2c_conv_bnrelu_fwd": (ctypes.c_int, [ctypes.POINTER(ConvArgs), c_vp]),
    "w2c_enc_head_fwd": (ctypes.c_int, [ctypes.POINTER(EncHeadArgs), c_vp]),
    "w2c_bn_train_fwd": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_f32,
                                        c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp]),
    "w2c_bn_train_from_sums_fwd": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp,
                                                  c_f32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp,
                                                  c_vp]),
    "w2c_conv_fuses_bn_sums": (ctypes.c_int, [ctypes.POINTER(ConvArgs)]),
    "w2c_bn_train_nchw_fwd": (ctypes.c_int, [c_vp, c_i32, c_i32, ctypes.c_int64, c_i32, c_vp, c_vp, c_f32, c_f32, c_vp,
                                             c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "w2c_conv_wgrad": (ctypes.c_int, [ctypes.POINTER(WgradArgs), c_vp]),
    "w2c_pack_conv_weight_ex": (ctypes.c_int, [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "w2c_pack_conv_weights_batch": (ctypes.c_int, [ctypes.POINTER(PackItem), c_i32, c_i32, c_vp]),
    "w2c_fold_bn_batch": (ctypes.c_int, [ctypes.POINTER(FoldItem), c_i32, c_vp]),
    "w2c_bn_train_bwd": (ctypes.c_int, [ctypes.POINTER(BnBwdArgs), c_vp]),
    "w2c_bn_train_nchw_bwd": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, ctypes.c_int64, c_i32, c_i32, c_i32,
                                             c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "w2c_attn_fuse_bwd": (ctypes.c_int, [ctypes.POINTER(AttnBwdArgs), c_vp]),
    "w2c_kq_mlp_heads_bwd": (ctypes.c_int, [c_vp, c_i32, c_i32, c_i32, ctypes.POINTER(MlpHead),
                                            ctypes.POINTER(MlpHeadGrad), c_i32, c_vp, c_vp, c_i32, c_vp, c_vp]),
    "w2c_stem_conv_wgrad": (ctypes.c_int, [c_vp, c_vp, c_vp] + [c_i32] * 11 + [c_vp]),
    "w2c_maxpool3x3s2_bwd": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "w2c_bilinear_up_bwd": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "w2c_upsample_zero2": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "w2c_cross_entropy2d": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i32, ctypes.c_int64, ctypes.c_int64, c_vp, c_vp, c_vp]),
    "w2c_grad_add": (ctypes.c_int, [c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, ctypes.c_int64, c_i32,
                                    c_i32, c_vp]),
    "w2c_stem_conv3x3_raw_fwd": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp] + [c_i32] * 9 + [c_vp]),
    "w2c_stem_conv7x7s2_raw_fwd": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp] + [c_i32] * 9 + [c_vp]),
    "w2c_cout_pad": (c_i32, [c_i32]),
    "w2c_packed_weight_bytes": (ctypes.c_size_t, [c_i32, c_i32, c_i32, c_i32]),
    "w2c_pack_conv_weight": (ctypes.c_int, [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "w2c_fold_bn": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_f32, c_i32, c_vp, c_vp, c_vp]),
    "w2c_stem_conv3x3_fwd": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp] + [c_i32] * 9 + [c_vp]),
    "w2c_stem_conv3x3_u8_fwd": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp] + [c_i32] * 9 + [c_vp]),
    "w2c_argmax_labels_fwd": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i32, ctypes.c_int64, c_vp]),
    "w2c_confusion_update": (ctypes.c_int, [c_vp, c_vp, c_i32, ctypes.c_int64, c_i32, c_vp, c_vp]),
    "w2c_confusion_update_div": (ctypes.c_int, [c_vp, c_vp, c_i32, c_vp, c_i32, ctypes.c_int64, c_i32, c_vp, c_vp, c_vp]),
    "w2c_selection_update": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "w2c_kq_mlp_fwd": (ctypes.c_int, [c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp]),
    "w2c_kq_mlp_heads_fwd": (ctypes.c_int, [c_vp, c_i32, c_i32, c_i32, ctypes.POINTER(MlpHead), c_i32, c_vp, c_vp]),
    "w2c_attn_fuse_fwd": (ctypes.c_int, [ctypes.POINTER(AttnArgs), c_vp]),
    "w2c_stem_conv7x7s2_fwd": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp] + [c_i32] * 9 + [c_vp]),
    "w2c_stem_conv7x7s2_u8_fwd": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp] + [c_i32] * 9 + [c_vp]),
    "w2c_maxpool3x3s2_fwd": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "w2c_bilinear_up_fwd": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "w2c_gather_images_fwd": (ctypes.c_int, [c_vp, c_vp, c_vp] + [c_i32] * 10 + [c_vp]),
    "w2c_bilinear_argmax_fwd": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "w2c_nhwc_to_nchw_f32": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "w2c_nchw_f32_to_nhwc": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
}

_lib = None
_lock = threading.Lock()


class W2CError(RuntimeError):
    pass


def lib_path():
    return _build.LIB_PATH


def load():
    """Load (building first if needed) libw2c.so and bind every entry point. Raises if that is impossible."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("W2C_LIB") or _build.LIB_PATH   # W2C_LIB: an alternative build, for A/B runs
        if path == _build.LIB_PATH:
            # build() is a no-op when lib/libw2c.stamp equals the fingerprint of the sources + flags, and recompiles
            # otherwise: an edited .cu / .cuh / w2c.h can never run against a stale binary. Where the sources are
            # there but nvcc is not (it always is in this image), a stale library is an error, not a warning.
            try:
                path = _build.build(force=os.environ.get("W2C_REBUILD") == "1")
            except RuntimeError as e:
                if not os.path.exists(path) or not _build.is_current():
                    raise W2CError("libw2c.so is missing or older than its sources and cannot be rebuilt: %s" % e)
        lib = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError = the .so is stale / incomplete: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc, what):
    if rc != 0:
        msg = load().w2c_last_error()
        raise W2CError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))
