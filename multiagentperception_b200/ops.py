"""Torch-tensor front end of the C ABI: each function takes CUDA tensors, passes raw device pointers + sizes and the
current CUDA stream to libw2c.so, and returns the (caller-visible) output tensor. PyTorch is used for device memory
and streams only; all arithmetic happens in the hand-written kernels.

Activation maps are NHWC bf16 tensors of shape [n, h, w, planes * c] (planes = 2 for the hi|lo "bf16x3" parity
precision, see include/w2c.h).
"""
import ctypes

import torch

from . import _lib
from ._lib import (ACT_BF16, ACT_BF16X2, ACT_FP16, ACT_FP16X2, OUT_NHWC, OUT_NCHW_F32, IMPL_TCGEN05, IMPL_SIMT, IMPL_TC_TAPS, IMPL_TC_PERSIST, CONV3X3_S1, CONV3X3_S2,
                   DECONV3X3_S2, CONV1X1_S1, CONV1X1_S2, FUSE_SOFTMAX, FUSE_ACTIVATED, FUSE_ARGMAX, GT_U8, GT_I64)

# airsim_loader.py:191 (mean_rgb['airsim'], indexed by BGR channel after the loader's flip)
LOADER_MEAN_BGR = (103.939, 116.779, 123.68)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError("libw2c operates on CUDA tensors only (got a %s tensor)" % t.device)
    if not t.is_contiguous():
        raise ValueError("libw2c needs contiguous tensors")
    return ctypes.c_void_p(t.data_ptr())


def planes_of(act):
    return 2 if act in (ACT_BF16X2, ACT_FP16X2) else 1


def launch_count():
    return int(_lib.load().w2c_launch_count())


def cout_pad(cout):
    return int(_lib.load().w2c_cout_pad(cout))


def new_act(n, h, w, c, act, device):
    return torch.empty((n, h, w, planes_of(act) * c), dtype=torch.bfloat16, device=device)


# ------------------------------------------------------------------------------------------------ setup-time ops
def fold_bn(conv_bias, bn_weight, bn_bias, bn_mean, bn_var, eps, cout, device):
    """scale/shift of the fused epilogue from an eval-mode BatchNorm2d and the conv bias (utils.py:110-114)."""
    lib = _lib.load()
    scale = torch.empty(cout, dtype=torch.float32, device=device)
    shift = torch.empty(cout, dtype=torch.float32, device=device)
    f32 = lambda t: None if t is None else t.detach().to(device=device, dtype=torch.float32).contiguous()
    cb, g, b, m, v = f32(conv_bias), f32(bn_weight), f32(bn_bias), f32(bn_mean), f32(bn_var)
    _lib.check(lib.w2c_fold_bn(_ptr(cb), _ptr(g), _ptr(b), _ptr(m), _ptr(v), float(eps), cout, _ptr(scale),
                               _ptr(shift), _stream()), "w2c_fold_bn")
    return scale, shift


def pack_conv_weight(w, cin_pad, transposed, act):
    """Conv2d.weight [co,ci,kh,kw] (or ConvTranspose2d.weight [ci,co,kh,kw]) -> packed bf16 K-major operand."""
    lib = _lib.load()
    w = w.detach().to(dtype=torch.float32).contiguous()
    if transposed:
        cin_real, cout = w.shape[0], w.shape[1]
    else:
        cout, cin_real = w.shape[0], w.shape[1]
    ntaps = w.shape[2] * w.shape[3]
    nbytes = lib.w2c_packed_weight_bytes(cout, cin_pad, ntaps, act)
    packed = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=w.device)
    _lib.check(lib.w2c_pack_conv_weight(_ptr(w), cout, cin_real, cin_pad, ntaps, int(bool(transposed)), act,
                                        _ptr(packed), _stream()), "w2c_pack_conv_weight")
    return packed


# ------------------------------------------------------------------------------------------------ hot-path ops
def conv_bnrelu(x, w_packed, scale, shift, y, *, n, h_in, w_in, cin, cout, kind, relu, act, out_fmt=OUT_NHWC,
                residual=None, x_cstride=0, x_coffset=0, y_cstride=0, y_coffset=0, impl=IMPL_TCGEN05, block_n=0,
                labels=None, passes=0, bn_sums=None):
    """labels: optional uint8 [n, h_out, w_out] tensor receiving argmax_co y (NCHW fp32 logits layout only); with
    labels given, y may be None (label map only). bn_sums: optional fp64 [2 * cout] tensor, += sum | sum of squares
    of the stored output per channel (only where w2c_conv_fuses_bn_sums says so; an error otherwise)."""
    lib = _lib.load()
    a = _lib.ConvArgs(x=_ptr(x), w=_ptr(w_packed), scale=_ptr(scale), shift=_ptr(shift), residual=_ptr(residual),
                      y=_ptr(y), labels=_ptr(labels), n=n, h_in=h_in, w_in=w_in, cin=cin, cout=cout, x_cstride=x_cstride,
                      x_coffset=x_coffset, y_cstride=y_cstride, y_coffset=y_coffset, kind=kind, relu=int(relu),
                      act=act, out_fmt=out_fmt, impl=impl, block_n=block_n, passes=passes, bn_sums=_ptr(bn_sums))
    _lib.check(lib.w2c_conv_bnrelu_fwd(ctypes.byref(a), _stream()), "w2c_conv_bnrelu_fwd")
    return y


def enc_head(x, w1, scale1, shift1, w2_packed, scale2, shift2, y, *, b, n_agents, h, w, act, c_total=0, c_first=0,
             lut=None, y_cstride=0, y_coffset=0):
    """Fused conv1 + conv2 of n_segnet_encoder (csrc/enc_head.cu). x: fp32 views (B, c_total, H, W), or - with `lut`
    given - the raw uint8 frames (B, agents_total, H, W, 3); y: NHWC (b*n_agents, H/2, W/2, planes*y_cstride)."""
    lib = _lib.load()
    u8 = lut is not None
    a = _lib.EncHeadArgs(x=_ptr(x), lut=_ptr(lut), w1=_ptr(w1), scale1=_ptr(scale1), shift1=_ptr(shift1),
                         w2=_ptr(w2_packed), scale2=_ptr(scale2), shift2=_ptr(shift2), y=_ptr(y), x_u8=int(u8), b=b,
                         n_agents=n_agents, c_total=c_total or (n_agents if u8 else 3 * n_agents), c_first=c_first,
                         h=h, w=w, act=act, y_cstride=y_cstride, y_coffset=y_coffset)
    _lib.check(lib.w2c_enc_head_fwd(ctypes.byref(a), _stream()), "w2c_enc_head_fwd")
    return y


def stem_conv3x3(x_nchw, w27, scale, shift, y, *, b, n_agents, h, w, cout, act, c_total=0, c_first=0, n_split=1):
    """n_split=2 (cout=128): y holds two dense 64-channel NHWC maps, [2, n, h, w, planes*64]."""
    lib = _lib.load()
    _lib.check(lib.w2c_stem_conv3x3_fwd(_ptr(x_nchw), _ptr(w27), _ptr(scale), _ptr(shift), _ptr(y), b, n_agents,
                                        c_total or 3 * n_agents, c_first, h, w, cout, act, n_split, _stream()),
               "w2c_stem_conv3x3_fwd")
    return y


def loader_lut(mean_bgr=LOADER_MEAN_BGR, img_norm=True, device=None):
    """fp32 [3, 256] table of the loader transform per BGR channel: float32((float64(v) - mean[c]) / 255.0)
    (airsim_loader.py:522-525 computes in float64, then torch.from_numpy(img).float(), :535)."""
    import numpy as np
    v = np.arange(256, dtype=np.float64)[None, :] - np.asarray(mean_bgr, dtype=np.float64)[:, None]
    if img_norm:
        v = v.astype(float) / 255.0
    t = torch.from_numpy(v).float().contiguous()
    return t.to(device) if device is not None else t


def stem_conv3x3_u8(frames, lut, w27, scale, shift, y, *, b, n_agents, h, w, cout, act, agents_total=0, agent_first=0,
                    n_split=1):
    """frames: uint8 RGB HWC [b, agents_total, h, w, 3] (raw loader frames); lut from loader_lut()."""
    lib = _lib.load()
    if frames.dtype != torch.uint8:
        raise ValueError("stem_conv3x3_u8 needs uint8 frames")
    _lib.check(lib.w2c_stem_conv3x3_u8_fwd(_ptr(frames), _ptr(lut), _ptr(w27), _ptr(scale), _ptr(shift), _ptr(y), b,
                                           n_agents, agents_total or n_agents, agent_first, h, w, cout, act, n_split,
                                           _stream()), "w2c_stem_conv3x3_u8_fwd")
    return y


def argmax_labels(logits, labels=None):
    """uint8 [n, h, w] = logits.max(1)[1] (first maximal index) of fp32 NCHW logits."""
    lib = _lib.load()
    n, c, h, w = logits.shape
    if labels is None:
        labels = torch.empty((n, h, w), dtype=torch.uint8, device=logits.device)
    _lib.check(lib.w2c_argmax_labels_fwd(_ptr(logits), _ptr(labels), n, c, h * w, _stream()), "w2c_argmax_labels_fwd")
    return labels


def confusion_update(pred_labels, gt, n_classes, hist):
    """hist[n_classes*gt + pred] += 1 over pixels with 0 <= gt < n_classes (runningScore._fast_hist,
    metrics.py:99-104). pred_labels uint8, gt uint8 or int64 (same number of elements), hist int64 [n, n] on device."""
    lib = _lib.load()
    if pred_labels.dtype != torch.uint8 or pred_labels.numel() != gt.numel():
        raise ValueError("confusion_update: pred must be uint8 with as many elements as gt")
    if hist.dtype != torch.int64 or hist.numel() != n_classes * n_classes:
        raise ValueError("confusion_update: hist must be int64 [%d, %d]" % (n_classes, n_classes))
    dt = {torch.uint8: GT_U8, torch.int64: GT_I64}.get(gt.dtype)
    if dt is None:
        raise ValueError("confusion_update: gt must be uint8 or int64 (got %s)" % gt.dtype)
    _lib.check(lib.w2c_confusion_update(_ptr(pred_labels), _ptr(gt), dt, gt.numel(), n_classes, _ptr(hist), _stream()),
               "w2c_confusion_update")
    return hist


def confusion_update_div(pred_labels, gt, img_flag, n_classes, hist_pos, hist_neg):
    """runningScore.update_div (metrics.py:70-97) on the device: image i of pred / gt (uint8 [n_img, H, W] / uint8 or
    int64) is histogrammed into hist_pos when img_flag[i] != 0 ("normal"), else into hist_neg ("noisy")."""
    lib = _lib.load()
    n_img = img_flag.numel()
    if pred_labels.dtype != torch.uint8 or pred_labels.numel() != gt.numel() or pred_labels.numel() % n_img:
        raise ValueError("confusion_update_div: pred must be uint8 [n_img, ...] with as many elements as gt")
    if img_flag.dtype != torch.uint8:
        raise ValueError("confusion_update_div: img_flag must be uint8 [n_img]")
    for h in (hist_pos, hist_neg):
        if h.dtype != torch.int64 or h.numel() != n_classes * n_classes:
            raise ValueError("confusion_update_div: hist must be int64 [%d, %d]" % (n_classes, n_classes))
    dt = {torch.uint8: GT_U8, torch.int64: GT_I64}.get(gt.dtype)
    if dt is None:
        raise ValueError("confusion_update_div: gt must be uint8 or int64 (got %s)" % gt.dtype)
    _lib.check(lib.w2c_confusion_update_div(_ptr(pred_labels), _ptr(gt), dt, _ptr(img_flag), n_img,
                                            pred_labels.numel() // n_img, n_classes, _ptr(hist_pos), _ptr(hist_neg),
                                            _stream()), "w2c_confusion_update_div")


def selection_update(action, commun_label, mode, counters):
    """runningScore.update_selection (metrics.py:23-68) on the device; counters int64 [3] = {total_agent,
    correct_when2com, correct_who2com}. mode 'mimo': action int64 [B, N], commun_label int64 [B, 2, N]; 'when2com':
    action int64 [B] (arg-max) or fp32 [B, N] (thresholded weights), commun_label int64 [B]."""
    lib = _lib.load()
    if counters.dtype != torch.int64 or counters.numel() != 3 or commun_label.dtype != torch.int64:
        raise ValueError("selection_update: counters int64 [3], commun_label int64")
    if mode == "mimo":
        b, n = action.shape
        if action.dtype != torch.int64 or tuple(commun_label.shape) != (b, 2, n):
            raise ValueError("selection_update('mimo'): action int64 [B, N], commun_label [B, 2, N]")
        m = 0
    elif mode == "when2com":
        if action.dim() == 1 and action.dtype == torch.int64:
            b, n, m = action.shape[0], 1, 1
        elif action.dim() == 2 and action.dtype == torch.float32:
            (b, n), m = action.shape, 2
        else:
            raise ValueError("selection_update('when2com'): action int64 [B] or fp32 [B, N]")
        if commun_label.numel() != b:
            raise ValueError("selection_update('when2com'): commun_label [B]")
    else:
        raise ValueError("selection_update: mode must be 'mimo' or 'when2com'")
    _lib.check(lib.w2c_selection_update(_ptr(action.contiguous()), _ptr(commun_label.contiguous()), b, n, m,
                                        _ptr(counters), _stream()), "w2c_selection_update")


def stem_conv7x7s2(x_nchw, w147, scale, shift, y, *, b, n_agents, h, w, act, c_total=0, c_first=0, cout=64, n_split=1):
    lib = _lib.load()
    _lib.check(lib.w2c_stem_conv7x7s2_fwd(_ptr(x_nchw), _ptr(w147), _ptr(scale), _ptr(shift), _ptr(y), b, n_agents,
                                          c_total or 3 * n_agents, c_first, h, w, cout, act, n_split, _stream()),
               "w2c_stem_conv7x7s2_fwd")
    return y


def stem_conv7x7s2_u8(frames, lut, w147, scale, shift, y, *, b, n_agents, h, w, act, agents_total=0, agent_first=0,
                      cout=64, n_split=1):
    lib = _lib.load()
    if frames.dtype != torch.uint8:
        raise ValueError("stem_conv7x7s2_u8 needs uint8 frames")
    _lib.check(lib.w2c_stem_conv7x7s2_u8_fwd(_ptr(frames), _ptr(lut), _ptr(w147), _ptr(scale), _ptr(shift), _ptr(y), b,
                                             n_agents, agents_total or n_agents, agent_first, h, w, cout, act,
                                             n_split, _stream()), "w2c_stem_conv7x7s2_u8_fwd")
    return y


def maxpool3x3s2(x, y, *, n, h, w, c, act):
    lib = _lib.load()
    _lib.check(lib.w2c_maxpool3x3s2_fwd(_ptr(x), _ptr(y), n, h, w, c, act, _stream()), "w2c_maxpool3x3s2_fwd")
    return y


def bilinear_up(x, y, *, n, c, h, w, factor):
    lib = _lib.load()
    _lib.check(lib.w2c_bilinear_up_fwd(_ptr(x), _ptr(y), n, c, h, w, factor, _stream()), "w2c_bilinear_up_fwd")
    return y


def kq_mlp(feat, act, m, n_feat, w0, b0, w1, b1, w2, b2, out_dim, out, ws):
    lib = _lib.load()
    _lib.check(lib.w2c_kq_mlp_fwd(_ptr(feat), act, m, n_feat, _ptr(w0), _ptr(b0), _ptr(w1), _ptr(b1), _ptr(w2),
                                  _ptr(b2), out_dim, _ptr(out), _ptr(ws), _stream()), "w2c_kq_mlp_fwd")
    return out


def attn_fuse(keys, queries, wq, bq, val, fused, prob_out, coef_out, action, connect, *, b_sz, n_k, n_q, k_dim,
              q_dim, hw, c, act, mode, sparse=False, mask_self=False, temperature=1.0, diag_bias=0.0, thresh=0.2,
              fused_cstride=0, fused_coffset=0):
    lib = _lib.load()
    a = _lib.AttnArgs(keys=_ptr(keys), queries=_ptr(queries), wq=_ptr(wq), bq=_ptr(bq), val=_ptr(val),
                      fused=_ptr(fused), prob_out=_ptr(prob_out), coef_out=_ptr(coef_out), action=_ptr(action),
                      connect=_ptr(connect), b_sz=b_sz, n_k=n_k, n_q=n_q, k_dim=k_dim, q_dim=q_dim, hw=hw, c=c,
                      fused_cstride=fused_cstride, fused_coffset=fused_coffset, act=act, mode=mode,
                      sparse=int(bool(sparse)), mask_self=int(bool(mask_self)), temperature=float(temperature),
                      diag_bias=float(diag_bias), thresh=float(thresh))
    _lib.check(lib.w2c_attn_fuse_fwd(ctypes.byref(a), _stream()), "w2c_attn_fuse_fwd")


# ------------------------------------------------------------------------------------------------ layout helpers
def nchw_to_act(x_nchw, act, cstride=0, coffset=0, out=None):
    lib = _lib.load()
    n, c, h, w = x_nchw.shape
    cs = cstride if cstride > 0 else c
    if out is None:
        out = torch.zeros((n, h, w, planes_of(act) * cs), dtype=torch.bfloat16, device=x_nchw.device)
    x32 = x_nchw.to(torch.float32).contiguous()
    _lib.check(lib.w2c_nchw_f32_to_nhwc(_ptr(x32), _ptr(out), n, h, w, c, cs, coffset, act, _stream()),
               "w2c_nchw_f32_to_nhwc")
    return out


def act_to_nchw(x_act, c, act, cstride=0, coffset=0):
    lib = _lib.load()
    n, h, w, _ = x_act.shape
    cs = cstride if cstride > 0 else c
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x_act.device)
    _lib.check(lib.w2c_nhwc_to_nchw_f32(_ptr(x_act), _ptr(out), n, h, w, c, cs, coffset, act, _stream()),
               "w2c_nhwc_to_nchw_f32")
    return out


# ------------------------------------------------------------------------------------------------ backward ops
def pack_conv_weight_ex(w, cout, cin_real, cin_pad, ntaps, transposed, flip, act, out=None):
    """w2c_pack_conv_weight with explicit geometry and an optional tap flip (the data-gradient operands, w2c.h)."""
    lib = _lib.load()
    nbytes = lib.w2c_packed_weight_bytes(cout, cin_pad, ntaps, act)
    if out is None:
        out = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=w.device)
    _lib.check(lib.w2c_pack_conv_weight_ex(_ptr(w), cout, cin_real, cin_pad, ntaps, int(bool(transposed)),
                                           int(bool(flip)), act, _ptr(out), _stream()), "w2c_pack_conv_weight_ex")
    return out


def conv_wgrad(x, dy, dw, *, n, h_in, w_in, cin, cout, kind, act_x, act_dy, x_cstride=0, x_coffset=0, dy_cstride=0,
               dy_coffset=0, passes=0):
    """dw (fp32, accumulated): [cout][ntaps][cin] for Conv2d kinds, [cin][ntaps][cout] for DECONV3X3_S2."""
    lib = _lib.load()
    a = _lib.WgradArgs(x=_ptr(x), dy=_ptr(dy), dw=_ptr(dw), n=n, h_in=h_in, w_in=w_in, cin=cin, cout=cout,
                       x_cstride=x_cstride, x_coffset=x_coffset, dy_cstride=dy_cstride, dy_coffset=dy_coffset, kind=kind,
                       act_x=act_x, act_dy=act_dy, passes=passes)
    _lib.check(lib.w2c_conv_wgrad(ctypes.byref(a), _stream()), "w2c_conv_wgrad")
    return dw


def bn_train_bwd(dy, y, z, dz, *, n_px, c, act_f, act_g, relu, gamma, stats, dgamma, dbeta, sums_ws, coef_ws, dres=None,
                 dy_cs=0, dy_co=0, y_cs=0, y_co=0, z_cs=0, z_co=0, dz_cs=0, dz_co=0, dres_cs=0, dres_co=0, fwd_scale=None,
                 fwd_shift=None):
    lib = _lib.load()
    a = _lib.BnBwdArgs(dy=_ptr(dy), y=_ptr(y), z=_ptr(z), dz=_ptr(dz), dres=_ptr(dres), n_px=n_px, c=c, dy_cstride=dy_cs,
                       dy_coffset=dy_co, y_cstride=y_cs, y_coffset=y_co, z_cstride=z_cs, z_coffset=z_co, dz_cstride=dz_cs,
                       dz_coffset=dz_co, dres_cstride=dres_cs, dres_coffset=dres_co, act_f=act_f, act_g=act_g,
                       relu=int(bool(relu)), gamma=_ptr(gamma), stats=_ptr(stats), dgamma=_ptr(dgamma), dbeta=_ptr(dbeta),
                       sums_ws=_ptr(sums_ws), coef_ws=_ptr(coef_ws), fwd_scale=_ptr(fwd_scale), fwd_shift=_ptr(fwd_shift))
    _lib.check(lib.w2c_bn_train_bwd(ctypes.byref(a), _stream()), "w2c_bn_train_bwd")
    return dz
