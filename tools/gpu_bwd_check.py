#!/usr/bin/env python
"""Backward-pass kernels on a real B200, each against torch autograd (float64, CPU) on the operands as the kernels see
them (quantised to the activation storage). Called by tests/test_backward_kernels_gpu.py; standalone:

    python tools/gpu_bwd_check.py            # all cases, one JSON line each
    python tools/gpu_bwd_check.py --case X
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools.gpu_kernel_check import _quant  # noqa: E402


def _imports():
    import torch
    import torch.nn.functional as F
    from multiagentperception_b200 import _lib, ops
    return torch, F, ops, _lib


def _ref_conv(F, kind, x, w):
    from multiagentperception_b200 import ops
    if kind == ops.CONV3X3_S1:
        return F.conv2d(x, w, padding=1)
    if kind == ops.CONV3X3_S2:
        return F.conv2d(x, w, padding=1, stride=2)
    if kind == ops.DECONV3X3_S2:
        return F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1)
    if kind == ops.CONV1X1_S1:
        return F.conv2d(x, w)
    return F.conv2d(x, w, stride=2)


def _verdict(got, ref, rel):
    import torch
    err = (got.double().cpu() - ref.double().cpu()).abs().max().item()
    mag = ref.abs().max().item()
    return {"max_err": err, "ref_max": mag, "tol": rel * max(mag, 1e-30),
            "ok": bool(err <= rel * max(mag, 1e-30)) and bool(torch.isfinite(got).all())}


# ------------------------------------------------------------------------------------------------ wgrad
def wgrad_case(kind, n, h, w, cin, cout, act_x, act_dy=None, passes=0, seed=0):
    torch, F, ops, _ = _imports()
    dev = torch.device("cuda:0")
    act_dy = act_x if act_dy is None else act_dy
    g = torch.Generator().manual_seed(seed)
    ks = 1 if kind in (ops.CONV1X1_S1, ops.CONV1X1_S2) else 3
    transposed = kind == ops.DECONV3X3_S2
    x = torch.randn(n, cin, h, w, generator=g)
    wshape = (cin, cout, ks, ks) if transposed else (cout, cin, ks, ks)
    w0 = torch.zeros(wshape, dtype=torch.float64, requires_grad=True)
    one = passes == 1
    xq = _quant(x, {3: 2, 1: 0}.get(act_x, act_x) if one else act_x)
    y = _ref_conv(F, kind, xq, w0)
    dy = torch.randn(y.shape, generator=g) * 0.1
    dyq = _quant(dy, {3: 2, 1: 0}.get(act_dy, act_dy) if one else act_dy)
    (y * dyq).sum().backward()
    ref = w0.grad                                                   # parameter layout [d0][d1][kh][kw]
    xa = ops.nchw_to_act(x.to(dev), act_x)
    dya = ops.nchw_to_act(dy.to(dev), act_dy)
    d0, d1 = wshape[0], wshape[1]
    dw = torch.zeros(d0, ks * ks, d1, dtype=torch.float32, device=dev)
    ops.conv_wgrad(xa, dya, dw, n=n, h_in=h, w_in=w, cin=cin, cout=cout, kind=kind, act_x=act_x, act_dy=act_dy,
                   passes=passes)
    # a second call must accumulate
    ops.conv_wgrad(xa, dya, dw, n=n, h_in=h, w_in=w, cin=cin, cout=cout, kind=kind, act_x=act_x, act_dy=act_dy,
                   passes=passes)
    torch.cuda.synchronize()
    got = dw.view(d0, ks, ks, d1).permute(0, 3, 1, 2) / 2
    planes3 = ops.planes_of(act_x) == 2 and not one
    return _verdict(got, ref, 3e-5 if planes3 else 2e-4)


WGRAD_CASES = {
    # name: (kind, n, h, w, cin, cout, act_x, kwargs)
    "wg_s1_min": (0, 1, 8, 16, 64, 128, 0, {}),
    "wg_s1_cp64": (0, 2, 16, 16, 128, 64, 0, {}),
    "wg_s1_ragged": (0, 3, 24, 40, 64, 128, 0, {}),
    "wg_s1_x2": (0, 2, 16, 16, 128, 128, 1, {}),
    "wg_s1_x2_onepass": (0, 2, 16, 16, 64, 128, 1, dict(passes=1)),
    "wg_s1_f16": (0, 2, 16, 16, 64, 128, 2, {}),
    "wg_s1_f16x2": (0, 2, 16, 16, 64, 128, 3, {}),
    "wg_s2": (1, 2, 32, 32, 64, 128, 0, {}),
    "wg_s2_x2": (1, 1, 16, 48, 128, 64, 1, {}),
    "wg_deconv": (2, 2, 8, 8, 128, 64, 0, {}),
    "wg_deconv_x2": (2, 2, 16, 16, 64, 128, 1, {}),
    "wg_1x1": (3, 2, 16, 16, 128, 128, 0, {}),
    "wg_1x1s2": (4, 2, 16, 16, 64, 128, 0, {}),
    "wg_small_4x4": (0, 5, 4, 4, 256, 256, 0, {}),
    "wg_small_s2_4x4": (1, 5, 4, 4, 256, 256, 1, {}),
    "wg_small_2x2": (0, 6, 2, 2, 256, 256, 1, {}),
    "wg_small_s2_2x2": (1, 6, 2, 2, 256, 256, 0, {}),
    "wg_big_k": (0, 4, 64, 64, 64, 64, 0, {}),
    # 256 channels on the dense operand and enough pixel blocks: two M halves per CTA
    "wg_mh2": (0, 4, 64, 64, 64, 256, 0, {}),
    "wg_mh2_x2": (0, 2, 64, 64, 128, 512, 1, {}),
    "wg_mh2_deconv": (2, 8, 32, 32, 256, 64, 0, {}),
    "wg_mh2_s2": (1, 4, 64, 64, 64, 256, 1, {}),
}


# ------------------------------------------------------------------------------------------------ dgrad
def dgrad_case(kind, n, h, w, cin, cout, act, seed=0, cout_real=None):
    """Data gradient of a forward conv of `kind` (cin -> cout on an h x w input) run on the forward kernels with the
    re-indexed weight (w2c.h, w2c_pack_conv_weight_ex)."""
    torch, F, ops, _ = _imports()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    ks = 1 if kind in (ops.CONV1X1_S1, ops.CONV1X1_S2) else 3
    transposed = kind == ops.DECONV3X3_S2
    co_r = cout_real or cout                     # real output channels (the logits layer: 11 inside a 64-channel map)
    wshape = (cin, co_r, ks, ks) if transposed else (co_r, cin, ks, ks)
    wt = torch.randn(wshape, generator=g) / (cin * ks * ks) ** 0.5
    x0 = torch.zeros(n, cin, h, w, dtype=torch.float64, requires_grad=True)
    y = _ref_conv(F, kind, x0, _quant(wt, act))
    dy = torch.randn(y.shape, generator=g)
    (y * _quant(dy, act)).sum().backward()
    ref = x0.grad
    wd = wt.to(dev)
    dy_pad = torch.zeros(n, cout, y.shape[2], y.shape[3])
    dy_pad[:, :co_r] = dy
    dya = ops.nchw_to_act(dy_pad.to(dev), act)
    one = torch.ones(cin, device=dev)
    zero = torch.zeros(cin, device=dev)
    if kind == ops.CONV3X3_S1:
        wp = ops.pack_conv_weight_ex(wd, cin, co_r, cout, 9, True, True, act)
        dk = ops.CONV3X3_S1
    elif kind == ops.CONV3X3_S2:
        wp = ops.pack_conv_weight_ex(wd, cin, co_r, cout, 9, True, False, act)
        dk = ops.DECONV3X3_S2
    elif kind == ops.DECONV3X3_S2:
        wp = ops.pack_conv_weight_ex(wd, cin, co_r, cout, 9, False, False, act)
        dk = ops.CONV3X3_S2
    else:
        wp = ops.pack_conv_weight_ex(wd, cin, co_r, cout, 1, True, False, act)
        dk = ops.CONV1X1_S1
    hy, wy = y.shape[2], y.shape[3]
    ho, wo = (h, w) if kind != ops.CONV1X1_S2 else (hy, wy)
    dx = torch.zeros((n, ho, wo, ops.planes_of(act) * cin), dtype=torch.bfloat16, device=dev)
    ops.conv_bnrelu(dya, wp, one, zero, dx, n=n, h_in=hy, w_in=wy, cin=cout, cout=cin, kind=dk, relu=False, act=act)
    if kind == ops.CONV1X1_S2:
        lib = _lib_load()
        full = torch.empty((n, h, w, ops.planes_of(act) * cin), dtype=torch.bfloat16, device=dev)
        rc = lib.w2c_upsample_zero2(_p(dx), None, _p(full), n, h, w, cin, act, ops._stream())
        assert rc == 0
        dx = full
    torch.cuda.synchronize()
    got = ops.act_to_nchw(dx, cin, act)
    return _verdict(got, ref, {0: 6e-3, 2: 8e-4}.get(act, 3e-5))


def _lib_load():
    from multiagentperception_b200 import _lib
    return _lib.load()


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


DGRAD_CASES = {
    "dg_s1": (0, 2, 16, 16, 64, 128, 1, {}),
    "dg_s1_bf16": (0, 2, 16, 16, 128, 64, 0, {}),
    "dg_s1_logits": (0, 1, 16, 16, 64, 64, 1, dict(cout_real=11)),
    "dg_s2": (1, 2, 16, 16, 64, 128, 1, {}),
    "dg_deconv": (2, 2, 8, 8, 128, 64, 1, {}),
    "dg_1x1": (3, 2, 8, 8, 64, 128, 1, {}),
    "dg_1x1s2": (4, 2, 8, 8, 64, 128, 1, {}),
}


# ------------------------------------------------------------------------------------------------ BatchNorm backward
def case_bn_bwd(act_f=1, act_g=1, relu=True, residual=False, bn=True, c=64, seed=0, mask_from_z=False, shape=(3, 6, 10)):
    """shape: (n, h, w). The default map is smaller than one trip of the kernels' raw-load loops (one-plane storages,
    >= 8 pixels per thread); the `*_big` cases run those loops and their tails."""
    torch, F, ops, _ = _imports()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    n, h, w = shape
    z = torch.randn(n, c, h, w, generator=g) * 2 + 0.3
    gamma = torch.rand(c, generator=g) + 0.5
    beta = torch.randn(c, generator=g) * 0.2
    res = torch.randn(n, c, h, w, generator=g)
    dy = torch.randn(n, c, h, w, generator=g)
    eps = 1e-5
    gq = gamma.double().requires_grad_(True)
    bq = beta.double().requires_grad_(True)
    rq = _quant(res, act_f).requires_grad_(True)

    def pre_act(zq):
        if bn:
            mean = zq.mean((0, 2, 3), keepdim=True)
            var = zq.var((0, 2, 3), unbiased=False, keepdim=True)
            u = (zq - mean) / torch.sqrt(var + eps) * gq.view(1, -1, 1, 1) + bq.view(1, -1, 1, 1)
        else:
            u = zq + bq.view(1, -1, 1, 1)
        return u + rq if residual else u
    if relu and n * h * w > 10000:
        # millions of elements: some pre-activation always lands within fp32 rounding of zero, where the device's ReLU
        # mask may legitimately differ from the float64 one - move those elements away from the kink
        for _ in range(4):
            with torch.no_grad():
                close = pre_act(_quant(z, act_f)).abs() < 2e-3
            if not bool(close.any()):
                break
            z[close] += 0.06
    zq = _quant(z, act_f).requires_grad_(True)
    u = pre_act(zq)
    yref = u.clamp_min(0) if relu else u
    (yref * _quant(dy, act_g)).sum().backward()
    # device side: run the forward kernel to get y and the statistics exactly as the train forward stores them
    lib = _lib_load()
    za = ops.nchw_to_act(z.to(dev), act_f)
    ya = torch.empty_like(za)
    ra = ops.nchw_to_act(res.to(dev), act_f) if residual else None
    gd, bd = gamma.to(dev), beta.to(dev)
    sums = torch.zeros(2 * c, dtype=torch.float64, device=dev)
    scale, shift = torch.empty(c, device=dev), torch.empty(c, device=dev)
    stats = torch.empty(2 * c, device=dev)
    n_px = n * h * w
    if bn:
        rc = lib.w2c_bn_train_fwd(_p(za), _p(ra), n_px, c, 0, 0, act_f, int(relu), _p(gd), _p(bd), eps, 0.1, None, None,
                                  None, _p(sums), _p(scale), _p(shift), _p(ya), 0, 0, _p(stats), ops._stream())
        assert rc == 0, lib.w2c_last_error()
    else:
        ya = ops.nchw_to_act(yref.detach().float().to(dev), act_f)
    dya = ops.nchw_to_act(dy.to(dev), act_g)
    dz = torch.empty((n, h, w, ops.planes_of(act_g) * c), dtype=torch.bfloat16, device=dev)
    dres = torch.empty_like(dz) if residual else None
    dgamma, dbeta = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    coef = torch.empty(3 * c, device=dev)
    ops.bn_train_bwd(dya, None if mask_from_z else ya, za, dz, n_px=n_px, c=c, act_f=act_f, act_g=act_g, relu=relu,
                     gamma=gd if bn else None, stats=stats if bn else None, dgamma=dgamma if bn else None, dbeta=dbeta,
                     sums_ws=sums, coef_ws=coef, dres=dres, fwd_scale=scale if mask_from_z else None,
                     fwd_shift=shift if mask_from_z else None)
    torch.cuda.synchronize()
    tol = {0: 8e-3, 1: 1e-4, 2: 1e-3, 3: 1e-4}[act_g]
    out = {"dz": _verdict(ops.act_to_nchw(dz, c, act_g), zq.grad, tol),
           "dbeta": _verdict(dbeta, bq.grad, 1e-4)}
    if bn:
        out["dgamma"] = _verdict(dgamma, gq.grad, 2e-4)
        out["sums_rezeroed"] = {"ok": bool((sums == 0).all())}
        # the train-mode forward that produced y (normalise + residual + ReLU pass)
        out["y_fwd"] = _verdict(ops.act_to_nchw(ya, c, act_f), yref.detach(), {0: 8e-3, 1: 1e-4, 2: 1e-3, 3: 1e-4}[act_f])
    if residual:
        out["dres"] = _verdict(ops.act_to_nchw(dres, c, act_g), rq.grad, tol)
    out["ok"] = all(v["ok"] for v in out.values())
    return out


def case_bn_bwd_nchw(bn=True, relu=True, act_g=1, seed=0):
    torch, F, ops, _ = _imports()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    n, c, h, w, c_pad = 2, 11, 12, 20, 64
    z = torch.randn(n, c, h, w, generator=g)
    gamma = torch.rand(c, generator=g) + 0.5
    beta = torch.randn(c, generator=g) * 0.2
    dy = torch.randn(n, c, h, w, generator=g)
    eps = 1e-5
    zq = z.double().requires_grad_(True)
    gq, bq = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    if bn:
        mean = zq.mean((0, 2, 3), keepdim=True)
        var = zq.var((0, 2, 3), unbiased=False, keepdim=True)
        u = (zq - mean) / torch.sqrt(var + eps) * gq.view(1, -1, 1, 1) + bq.view(1, -1, 1, 1)
    else:
        u = zq + bq.view(1, -1, 1, 1)
    yref = u.clamp_min(0) if relu else u
    (yref * dy.double()).sum().backward()
    lib = _lib_load()
    zd = z.to(dev)
    yd = torch.empty_like(zd)
    gd, bd = gamma.to(dev), beta.to(dev)
    sums = torch.zeros(2 * c, dtype=torch.float64, device=dev)
    scale, shift, stats = torch.empty(c, device=dev), torch.empty(c, device=dev), torch.empty(2 * c, device=dev)
    if bn:
        rc = lib.w2c_bn_train_nchw_fwd(_p(zd), n, c, h * w, int(relu), _p(gd), _p(bd), eps, 0.1, None, None, None, _p(sums),
                                       _p(scale), _p(shift), _p(yd), _p(stats), ops._stream())
        assert rc == 0, lib.w2c_last_error()
    else:
        yd = yref.detach().float().to(dev)
    dyd = dy.to(dev)
    dz = torch.full((n, h, w, ops.planes_of(act_g) * c_pad), float("nan"), dtype=torch.bfloat16, device=dev)
    dgamma, dbeta, coef = torch.zeros(c, device=dev), torch.zeros(c, device=dev), torch.empty(3 * c, device=dev)
    rc = lib.w2c_bn_train_nchw_bwd(_p(dyd), _p(yd), _p(zd), _p(dz), n, c, h * w, c_pad, 0, 0, act_g, int(relu),
                                   _p(gd) if bn else None, _p(stats) if bn else None, _p(dgamma) if bn else None,
                                   _p(dbeta), _p(sums), _p(coef), ops._stream())
    assert rc == 0, lib.w2c_last_error()
    torch.cuda.synchronize()
    got = ops.act_to_nchw(dz, c_pad, act_g)
    out = {"dz": _verdict(got[:, :c], zq.grad, {0: 8e-3, 1: 1e-4}[act_g]),
           "pad_zero": {"ok": bool((got[:, c:] == 0).all())},
           "dbeta": _verdict(dbeta, bq.grad, 1e-4)}
    if bn:
        out["dgamma"] = _verdict(dgamma, gq.grad, 2e-4)
    out["ok"] = all(v["ok"] for v in out.values())
    return out


# ------------------------------------------------------------------------------------------------ attention backward
def case_attn_bwd(n_k=5, n_q=5, b=2, k_dim=64, q_dim=32, proj=True, sparse=False, temperature=1.0, act_f=1, act_g=1,
                  mask_self=False, seed=0):
    torch, F, ops, _lib = _imports()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    hw, c = 12, 64
    if not proj:
        q_dim = k_dim
    keys = torch.randn(n_k * b, k_dim, generator=g) * 0.3
    queries = torch.randn(n_q * b, q_dim, generator=g) * 0.3
    wq = torch.randn(k_dim, q_dim, generator=g) * 0.3
    bq = torch.randn(k_dim, generator=g) * 0.1
    val = torch.randn(n_k * b, c, 3, 4, generator=g)
    dfu = torch.randn(n_q * b, c, 3, 4, generator=g)
    # reference in float64
    kk = keys.double().requires_grad_(True)
    qq = queries.double().requires_grad_(True)
    ww, bb = wq.double().requires_grad_(True), bq.double().requires_grad_(True)
    vv = _quant(val, act_f).requires_grad_(True)
    k3 = kk.view(n_k, b, k_dim).transpose(0, 1)           # [b][n_k][k]
    q3 = qq.view(n_q, b, q_dim).transpose(0, 1)
    qt = q3 @ ww.t() + bb if proj else q3
    S = k3 @ qt.transpose(1, 2) / temperature              # [b][n_k][n_q]
    if mask_self:
        S = S.masked_fill(torch.eye(n_k, dtype=torch.bool).unsqueeze(0), float("-inf"))
    if sparse:
        # sparsemax over dim 1
        zs, _ = torch.sort(S, dim=1, descending=True)
        rng = torch.arange(1, n_k + 1, dtype=torch.float64).view(1, -1, 1)
        css = zs.cumsum(1)
        kmax = ((1 + rng * zs) > css).to(torch.float64).sum(1, keepdim=True)
        tau = (css.gather(1, kmax.long() - 1) - 1) / kmax
        P = (S - tau).clamp_min(0)
    else:
        P = torch.softmax(S, dim=1)
    v5 = vv.view(n_k, b, c, 3, 4).transpose(0, 1)          # [b][n_k][c][h][w]
    Fu = torch.einsum("bij,bichw->bjchw", P, v5)           # [b][n_q][...]
    dfq = _quant(dfu, act_g).view(n_q, b, c, 3, 4).transpose(0, 1)
    (Fu * dfq).sum().backward()
    # device
    lib = _lib.load()
    kd, qd, wd, bd = keys.to(dev), queries.to(dev), wq.to(dev), bq.to(dev)
    va = ops.nchw_to_act(val.to(dev), act_f)
    dfa = ops.nchw_to_act(dfu.to(dev), act_g)
    Pd = P.detach().float().contiguous().to(dev)
    dval = torch.empty((n_k * b, 3, 4, ops.planes_of(act_g) * c), dtype=torch.bfloat16, device=dev)
    dkeys, dqueries = torch.empty_like(kd), torch.empty_like(qd)
    dwq, dbq = torch.zeros_like(wd), torch.zeros_like(bd)
    dp = torch.zeros(b * n_k * n_q, device=dev)
    a = _lib.AttnBwdArgs(keys=_p(kd), queries=_p(qd), wq=_p(wd) if proj else None, bq=_p(bd) if proj else None, val=_p(va),
                         dfused=_p(dfa), prob=_p(Pd), dval=_p(dval), dkeys=_p(dkeys), dqueries=_p(dqueries),
                         dwq=_p(dwq) if proj else None, dbq=_p(dbq) if proj else None, dp_ws=_p(dp), b_sz=b, n_k=n_k,
                         n_q=n_q, k_dim=k_dim, q_dim=q_dim, hw=hw, c=c, dfused_cstride=0, dfused_coffset=0, act_f=act_f,
                         act_g=act_g, sparse=int(sparse), dval_accumulate=0, temperature=temperature)
    rc = lib.w2c_attn_fuse_bwd(ctypes.byref(a), ops._stream())
    assert rc == 0, lib.w2c_last_error()
    torch.cuda.synchronize()
    tol = {0: 8e-3, 1: 1e-4}[act_g]
    out = {"dval": _verdict(ops.act_to_nchw(dval, c, act_g), vv.grad, tol),
           "dkeys": _verdict(dkeys, kk.grad, 2e-4), "dqueries": _verdict(dqueries, qq.grad, 2e-4),
           "dp_rezeroed": {"ok": bool((dp == 0).all())}}
    if proj:
        out["dwq"] = _verdict(dwq, ww.grad, 2e-4)
        out["dbq"] = _verdict(dbq, bb.grad, 2e-4)
    out["ok"] = all(v["ok"] for v in out.values())
    return out


# ------------------------------------------------------------------------------------------------ MLP heads backward
def case_mlp_bwd(m=5, side=2, n_heads=2, act_f=1, act_g=1, seed=0):
    torch, F, ops, _lib = _imports()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    n_feat = 256 * side * side
    outs = (1024, 32)[:n_heads]
    feat = torch.randn(m, 256, side, side, generator=g)
    fq = _quant(feat, act_f).requires_grad_(True)
    flat = fq.permute(0, 2, 3, 1).reshape(m, n_feat)          # NHWC flatten order
    ws, refs, douts = [], [], []
    total = 0
    for od in outs:
        w0 = (torch.randn(256, n_feat, generator=g) / n_feat ** 0.5).double().requires_grad_(True)
        b0 = (torch.randn(256, generator=g) * 0.1).double().requires_grad_(True)
        w1 = (torch.randn(128, 256, generator=g) / 16).double().requires_grad_(True)
        b1 = (torch.randn(128, generator=g) * 0.1).double().requires_grad_(True)
        w2 = (torch.randn(od, 128, generator=g) / 11).double().requires_grad_(True)
        b2 = (torch.randn(od, generator=g) * 0.1).double().requires_grad_(True)
        h0 = torch.relu(flat @ w0.t() + b0)
        h1 = torch.relu(h0 @ w1.t() + b1)
        o = h1 @ w2.t() + b2
        d = torch.randn(m, od, generator=g)
        total = total + (o * d.double()).sum()
        ws.append((w0, b0, w1, b1, w2, b2))
        douts.append(d)
    total.backward()
    lib = _lib.load()
    fa = ops.nchw_to_act(feat.to(dev), act_f)
    heads = (_lib.MlpHead * n_heads)()
    grads = (_lib.MlpHeadGrad * n_heads)()
    keep = []
    outs_d = []
    for i, od in enumerate(outs):
        wd = [t.detach().float().contiguous().to(dev) for t in ws[i]]
        od_t = torch.empty(m, od, device=dev)
        gd = [torch.zeros_like(t) for t in wd]
        dd = douts[i].to(dev)
        keep += wd + gd + [od_t, dd]
        heads[i] = _lib.MlpHead(*[t.data_ptr() for t in wd], od_t.data_ptr(), od)
        grads[i] = _lib.MlpHeadGrad(dd.data_ptr(), *[t.data_ptr() for t in gd])
        outs_d.append(gd)
    ws_fwd = torch.empty(n_heads * m * 256, device=dev)
    rc = lib.w2c_kq_mlp_heads_fwd(_p(fa), act_f, m, n_feat, heads, n_heads, _p(ws_fwd), ops._stream())
    assert rc == 0, lib.w2c_last_error()
    dfeat = torch.empty((m, side, side, ops.planes_of(act_g) * 256), dtype=torch.bfloat16, device=dev)
    ws_b = torch.empty(n_heads * m * 512 + 256, device=dev)
    rc = lib.w2c_kq_mlp_heads_bwd(_p(fa), act_f, m, n_feat, heads, grads, n_heads, _p(ws_fwd), _p(dfeat), act_g, _p(ws_b),
                                  ops._stream())
    assert rc == 0, lib.w2c_last_error()
    torch.cuda.synchronize()
    out = {"dfeat": _verdict(ops.act_to_nchw(dfeat, 256, act_g), fq.grad, {0: 8e-3, 1: 1e-4}[act_g])}
    names = ("dw0", "db0", "dw1", "db1", "dw2", "db2")
    for i in range(n_heads):
        for nm, gt, rt in zip(names, outs_d[i], ws[i]):
            out["h%d_%s" % (i, nm)] = _verdict(gt, rt.grad, 2e-4)
    out["ok"] = all(v["ok"] for v in out.values())
    return out


# ------------------------------------------------------------------------------------------------ stems, pooling, up-sampling
def case_stem_wgrad(ksize=3, act_g=1, seed=0):
    torch, F, ops, _ = _imports()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    b, n_agents, h, w, cout = 2, 3, 20, 28, 64
    x = torch.randn(b, 3 * n_agents + 3, h, w, generator=g)      # one leading view that is skipped (c_first = 3)
    stride, pad = (2, 3) if ksize == 7 else (1, 1)
    xs = torch.cat([x[:, 3 + 3 * a: 6 + 3 * a] for a in range(n_agents)], 0).double()   # agent-major
    w0 = torch.zeros(cout, 3, ksize, ksize, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(xs, w0, stride=stride, padding=pad)
    dz = torch.randn(y.shape, generator=g)
    (y * _quant(dz, act_g)).sum().backward()
    lib = _lib_load()
    xd = x.to(dev)
    dza = ops.nchw_to_act(dz.to(dev), act_g)
    dw = torch.zeros(cout, 3, ksize, ksize, device=dev)
    rc = lib.w2c_stem_conv_wgrad(_p(xd), _p(dza), _p(dw), ksize, b, n_agents, x.shape[1], 3, h, w, cout, 0, 0, act_g,
                                 ops._stream())
    assert rc == 0, lib.w2c_last_error()
    torch.cuda.synchronize()
    return _verdict(dw, w0.grad, 2e-4)


def case_maxpool_bwd(act=1, seed=0, c=64, h=12, w=20):
    torch, F, ops, _ = _imports()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    n = 2
    x = torch.relu(torch.randn(n, c, h, w, generator=g))        # ReLU'd: plenty of ties at zero
    xq = _quant(x, act).requires_grad_(True)
    y = F.max_pool2d(xq, 3, 2, 1)
    dy = torch.randn(y.shape, generator=g)
    (y * _quant(dy, act)).sum().backward()
    lib = _lib_load()
    xa, dya = ops.nchw_to_act(x.to(dev), act), ops.nchw_to_act(dy.to(dev), act)
    dx = torch.empty_like(xa)
    rc = lib.w2c_maxpool3x3s2_bwd(_p(xa), _p(dya), _p(dx), n, h, w, c, act, act, ops._stream())
    assert rc == 0, lib.w2c_last_error()
    torch.cuda.synchronize()
    return _verdict(ops.act_to_nchw(dx, c, act), xq.grad, {0: 8e-3, 1: 1e-4}[act])


def case_bilinear_bwd(factor=32, seed=0):
    torch, F, ops, _ = _imports()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    n, c, h, w = 2, 11, 4, 6
    x = torch.zeros(n, c, h, w, dtype=torch.float64, requires_grad=True)
    y = F.interpolate(x, scale_factor=factor, mode="bilinear", align_corners=False)
    dy = torch.randn(y.shape, generator=g)
    (y * dy.double()).sum().backward()
    lib = _lib_load()
    dyd = dy.to(dev)
    dx = torch.empty(n, c, h, w, device=dev)
    rc = lib.w2c_bilinear_up_bwd(_p(dyd), _p(dx), n, c, h, w, factor, ops._stream())
    assert rc == 0, lib.w2c_last_error()
    torch.cuda.synchronize()
    return _verdict(dx, x.grad, 1e-5)


def case_setup_batch(act=1, seed=0):
    """w2c_pack_conv_weights_batch / w2c_fold_bn_batch against the single calls they replace: identical bits."""
    import ctypes
    torch, F, ops, _ = _imports()
    from multiagentperception_b200 import _lib as L
    lib = _lib_load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    # (cout, cin_real, cin, ntaps, transposed, flip); 60 items: more than one launch of 48
    geoms = [(64, 64, 64, 9, 0, 0), (11, 64, 64, 9, 0, 0), (128, 40, 64, 9, 1, 1), (512, 512, 512, 9, 0, 1),
             (64, 128, 128, 1, 1, 0), (256, 192, 192, 9, 1, 0)] * 10
    ws, singles, batched, items = [], [], [], []
    for cout, cin_real, cin, ntaps, tr, flip in geoms:
        shape = (cin_real, cout, ntaps) if tr else (cout, cin_real, ntaps)
        w = torch.randn(shape, generator=g).to(dev)
        n16 = lib.w2c_packed_weight_bytes(cout, cin, ntaps, act) // 2
        a = torch.full((n16,), -1, dtype=torch.int16, device=dev)
        b = torch.full((n16,), -2, dtype=torch.int16, device=dev)
        rc = lib.w2c_pack_conv_weight_ex(_p(w), cout, cin_real, cin, ntaps, tr, flip, act, _p(a), ops._stream())
        assert rc == 0, lib.w2c_last_error()
        items.append(L.PackItem(w=_p(w), packed=_p(b), cout=cout, cin_real=cin_real, cin=cin, ntaps=ntaps, transposed=tr, flip=flip))
        ws.append(w), singles.append(a), batched.append(b)
    arr = (L.PackItem * len(items))(*items)
    rc = lib.w2c_pack_conv_weights_batch(arr, len(items), act, ops._stream())
    assert rc == 0, lib.w2c_last_error()
    torch.cuda.synchronize()
    out = {"pack": {"ok": all(bool(torch.equal(a, b)) for a, b in zip(singles, batched))}}
    # fold: with and without BatchNorm, with and without a conv bias
    fitems, pairs, keep = [], [], []
    for i, cout in enumerate([64, 11, 512, 200] * 13):
        t = [torch.randn(cout, generator=g).to(dev) for _ in range(4)] + [(torch.rand(cout, generator=g) + 0.1).to(dev)]
        bias = t[0] if i % 3 else None
        bn = i % 2 == 0
        s1, h1, s2, h2 = (torch.empty(cout, device=dev) for _ in range(4))
        rc = lib.w2c_fold_bn(_p(bias), _p(t[1]) if bn else None, _p(t[2]) if bn else None, _p(t[3]) if bn else None,
                             _p(t[4]) if bn else None, 1e-5, cout, _p(s1), _p(h1), ops._stream())
        assert rc == 0, lib.w2c_last_error()
        fitems.append(L.FoldItem(conv_bias=_p(bias), gamma=_p(t[1]) if bn else None, beta=_p(t[2]) if bn else None,
                                 mean=_p(t[3]) if bn else None, var=_p(t[4]) if bn else None, scale=_p(s2), shift=_p(h2),
                                 eps=1e-5, cout=cout))
        pairs.append((s1, h1, s2, h2)), keep.append(t)
    farr = (L.FoldItem * len(fitems))(*fitems)
    rc = lib.w2c_fold_bn_batch(farr, len(fitems), ops._stream())
    assert rc == 0, lib.w2c_last_error()
    torch.cuda.synchronize()
    out["fold"] = {"ok": all(bool(torch.equal(s1, s2)) and bool(torch.equal(h1, h2)) for s1, h1, s2, h2 in pairs)}
    out["ok"] = all(v["ok"] for v in out.values())
    return out


def case_grad_add(act=1):
    torch, F, ops, _ = _imports()
    dev = torch.device("cuda:0")
    a = torch.randn(2, 128, 4, 6)
    b = torch.randn(2, 64, 4, 6)
    aa = ops.nchw_to_act(a.to(dev), act)
    ba = ops.nchw_to_act(b.to(dev), act)
    dst = torch.zeros((2, 4, 6, ops.planes_of(act) * 64), dtype=torch.bfloat16, device=dev)
    lib = _lib_load()
    rc = lib.w2c_grad_add(_p(aa), 128, 64, _p(ba), 64, 0, _p(dst), 64, 0, 2 * 4 * 6, 64, act, ops._stream())
    assert rc == 0, lib.w2c_last_error()
    torch.cuda.synchronize()
    ref = _quant(a[:, 64:], act) + _quant(b, act)
    return _verdict(ops.act_to_nchw(dst, 64, act), ref, 1e-4)


def case_wgrad_mixed_rejected():
    """f16 forward maps against bf16 gradients: refused on the host (the mixed-type MMA faults on B200)."""
    torch, F, ops, _lib = _imports()
    dev = torch.device("cuda:0")
    x = torch.zeros((1, 8, 16, 2 * 64), dtype=torch.bfloat16, device=dev)
    dw = torch.zeros(64, 9, 64, device=dev)
    try:
        ops.conv_wgrad(x, x, dw, n=1, h_in=8, w_in=16, cin=64, cout=64, kind=0, act_x=3, act_dy=1)
    except _lib.W2CError as e:
        return {"ok": "element type" in str(e)}
    return {"ok": False}


def case_cross_entropy2d():
    torch, F, ops, _ = _imports()
    from multiagentperception_b200.loss import cross_entropy2d
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    n, c, h, w = 3, 11, 40, 56
    x = (torch.randn(n, c, h, w, generator=g) * 3).clamp_min(0)      # ReLU'd logits like the n_segnet decoder's
    t = torch.randint(0, c, (n, h, w), generator=g)
    t[0, :5] = 250
    t[2, 7, 9] = 250
    xr = x.double().requires_grad_(True)
    ref = F.cross_entropy(xr.permute(0, 2, 3, 1).reshape(-1, c), t.reshape(-1), ignore_index=250)
    (ref * 1.7).backward()
    xd = x.to(dev).requires_grad_(True)
    loss = cross_entropy2d(input=xd, target=t.to(dev))
    (loss * 1.7).backward()
    torch.cuda.synchronize()
    out = {"loss": _verdict(loss.detach().reshape(1), ref.detach().reshape(1), 2e-6),
           "grad": _verdict(xd.grad, xr.grad, 1e-5)}
    out["ok"] = all(v["ok"] for v in out.values())
    return out


OTHER_CASES = {
    "cross_entropy2d": case_cross_entropy2d,
    "wgrad_mixed_rejected": case_wgrad_mixed_rejected,
    "bn_bwd": lambda: case_bn_bwd(),
    "bn_bwd_bf16": lambda: case_bn_bwd(act_f=0, act_g=0),
    "bn_bwd_mask_from_z": lambda: case_bn_bwd(mask_from_z=True),
    "bn_bwd_mask_from_z_bf16": lambda: case_bn_bwd(act_f=0, act_g=0, mask_from_z=True, c=128),
    "bn_bwd_f16": lambda: case_bn_bwd(act_f=3, act_g=3, c=128),
    "bn_bwd_res": lambda: case_bn_bwd(residual=True),
    "bn_bwd_norelu": lambda: case_bn_bwd(relu=False),
    "bias_relu_bwd": lambda: case_bn_bwd(bn=False),
    # maps large enough for the raw-load loops of the one-plane storages (bn_bwd.cu RawPx, bn_train.cu bn_apply_span)
    "bn_bwd_big_bf16": lambda: case_bn_bwd(act_f=0, act_g=0, mask_from_z=True, shape=(4, 160, 200)),
    "bn_bwd_big_bf16_res": lambda: case_bn_bwd(act_f=0, act_g=0, residual=True, shape=(4, 150, 190)),
    "bn_bwd_big_f16_norelu": lambda: case_bn_bwd(act_f=2, act_g=2, relu=False, c=128, shape=(3, 150, 210)),
    "bias_relu_bwd_big": lambda: case_bn_bwd(act_f=0, act_g=0, bn=False, shape=(3, 180, 200)),
    "bn_bwd_big_x2": lambda: case_bn_bwd(mask_from_z=True, shape=(2, 120, 130)),
    "bn_bwd_nchw": lambda: case_bn_bwd_nchw(),
    "bias_bwd_nchw": lambda: case_bn_bwd_nchw(bn=False, relu=False),
    "attn_bwd": lambda: case_attn_bwd(),
    "attn_bwd_k1024": lambda: case_attn_bwd(n_k=5, n_q=5, b=3, k_dim=1024, q_dim=32),
    "attn_bwd_single_q": lambda: case_attn_bwd(n_k=5, n_q=1, k_dim=128, proj=True, q_dim=128),
    "attn_bwd_scaled": lambda: case_attn_bwd(n_k=4, n_q=1, k_dim=128, proj=False, temperature=128 ** 0.5),
    "attn_bwd_sparse": lambda: case_attn_bwd(n_k=5, n_q=1, k_dim=128, q_dim=128, sparse=True),
    "attn_bwd_who": lambda: case_attn_bwd(mask_self=True),
    "attn_bwd_bf16": lambda: case_attn_bwd(act_f=0, act_g=0),
    "mlp_bwd": lambda: case_mlp_bwd(),
    "mlp_bwd_big": lambda: case_mlp_bwd(m=11, side=4),
    "mlp_bwd_one": lambda: case_mlp_bwd(n_heads=1, side=1, act_f=0, act_g=0),
    "stem_wgrad3": lambda: case_stem_wgrad(3),
    "stem_wgrad7": lambda: case_stem_wgrad(7),
    "maxpool_bwd": lambda: case_maxpool_bwd(),
    "maxpool_bwd_c128_bf16": lambda: case_maxpool_bwd(act=0, c=128, h=18, w=34),   # two channel blocks, ragged tiles
    "maxpool_bwd_c72": lambda: case_maxpool_bwd(c=72, h=8, w=12),                  # per-pixel kernel (c % 64 != 0)
    "bilinear_bwd": lambda: case_bilinear_bwd(),
    "bilinear_bwd_x2": lambda: case_bilinear_bwd(2),
    "grad_add": lambda: case_grad_add(),
    "setup_batch_x2": lambda: case_setup_batch(1),
    "setup_batch_bf16": lambda: case_setup_batch(0),
}


def run_case(name):
    if name in WGRAD_CASES:
        kind, n, h, w, cin, cout, act, kw = WGRAD_CASES[name]
        return wgrad_case(kind, n, h, w, cin, cout, act, **kw)
    if name in DGRAD_CASES:
        kind, n, h, w, cin, cout, act, kw = DGRAD_CASES[name]
        return dgrad_case(kind, n, h, w, cin, cout, act, **kw)
    return OTHER_CASES[name]()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", action="append")
    args = ap.parse_args()
    names = args.case or (list(WGRAD_CASES) + list(DGRAD_CASES) + list(OTHER_CASES))
    bad = 0
    for nm in names:
        try:
            r = run_case(nm)
        except Exception as e:  # noqa: BLE001
            r = {"ok": False, "error": "%s: %s" % (type(e).__name__, e)}
        bad += not r["ok"]
        print(json.dumps({"case": nm, **r}), flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
