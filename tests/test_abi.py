"""CPU: the C-ABI library builds, loads, exports every symbol include/w2c.h declares, and validates arguments
before touching the device (no compute calls here — there is no GPU on this box)."""
import ctypes
import os
import re
import subprocess

import pytest

from multiagentperception_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "w2c.h")


def _declared_functions():
    with open(HEADER) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(w2c_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_for_sm100a():
    path = build.build()
    assert os.path.exists(path)
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_exports_match_header():
    lib = _lib.load()
    declared = _declared_functions()
    assert declared, "no functions parsed from include/w2c.h"
    for name in declared:
        assert hasattr(lib, name), "include/w2c.h declares %s but libw2c.so does not export it" % name
    assert sorted(_lib.exported_symbols()) == declared, "ctypes binding table and header disagree"


def test_struct_layouts_match_header():
    # field order / count of the two argument structs as declared in the header
    with open(HEADER) as f:
        src = f.read()

    def fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), src, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?[A-Za-z0-9_]+\s*\*?\s*", "", decl, count=1)
            names += [n.strip().lstrip("*").strip() for n in decl.split(",")]
        return names

    assert fields("w2c_conv_args") == [f[0] for f in _lib.ConvArgs._fields_]
    assert fields("w2c_attn_args") == [f[0] for f in _lib.AttnArgs._fields_]
    assert fields("w2c_mlp_head") == [f[0] for f in _lib.MlpHead._fields_]
    assert fields("w2c_pack_item") == [f[0] for f in _lib.PackItem._fields_]
    assert fields("w2c_fold_item") == [f[0] for f in _lib.FoldItem._fields_]


def test_tensor_core_and_tma_instructions_present():
    sass = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UBLKCP"):
        assert mnemonic in sass, "expected %s in the SASS of libw2c.so" % mnemonic
    assert "HMMA." not in sass.replace("UTCHMMA", ""), "legacy mma.sync path found"
    # the weight-gradient epilogue adds 16 bytes per reduction (red.global.add.v4.f32), not four scalar atomics
    assert "REDG.E.ADD.F32x4" in sass, "expected vector reductions in the wgrad epilogue"


def test_argument_validation_without_device():
    lib = _lib.load()
    assert lib.w2c_version() >= 100
    assert lib.w2c_cout_pad(11) == 16 and lib.w2c_cout_pad(64) == 64
    assert lib.w2c_packed_weight_bytes(11, 64, 9, _lib.ACT_BF16X2) == 16 * 9 * 64 * 2 * 2
    a = _lib.ConvArgs()
    assert lib.w2c_conv_bnrelu_fwd(ctypes.byref(a), None) == -1
    assert b"null pointer" in lib.w2c_last_error()
    assert lib.w2c_conv_bnrelu_fwd(None, None) == -1
    dummy = ctypes.c_void_p(256)
    a = _lib.ConvArgs(x=dummy, w=dummy, scale=dummy, shift=dummy, y=dummy, n=1, h_in=7, w_in=8, cin=64, cout=64,
                      kind=_lib.CONV3X3_S2)
    assert lib.w2c_conv_bnrelu_fwd(ctypes.byref(a), None) == -1
    assert b"even" in lib.w2c_last_error()
    a = _lib.ConvArgs(x=dummy, w=dummy, scale=dummy, shift=dummy, y=dummy, n=1, h_in=8, w_in=8, cin=48, cout=64)
    assert lib.w2c_conv_bnrelu_fwd(ctypes.byref(a), None) == -1
    assert b"multiple of 64" in lib.w2c_last_error()
    at = _lib.AttnArgs(keys=dummy, queries=dummy, val=dummy, fused=dummy, prob_out=dummy, b_sz=1, n_k=9, n_q=9,
                       k_dim=8, q_dim=8, hw=4, c=8, temperature=1.0)
    assert lib.w2c_attn_fuse_fwd(ctypes.byref(at), None) == -1
    assert b"[1, 8]" in lib.w2c_last_error()
    with pytest.raises(_lib.W2CError):
        _lib.check(-1, "unit-test")


def test_conv_fuses_bn_sums_is_host_logic():
    """w2c_conv_fuses_bn_sums mirrors the dispatch of w2c_conv_bnrelu_fwd without touching a device: the statistics come
    out of the persistent kernel's TMA-store epilogue only (NHWC, whole 64-channel groups, one storage plane, a tile
    per SM), and a launch that cannot provide them refuses bn_sums instead of silently skipping them."""
    lib = _lib.load()
    dummy = ctypes.c_void_p(256)

    def args(**kw):
        base = dict(x=dummy, w=dummy, scale=dummy, shift=dummy, y=dummy, n=10, h_in=128, w_in=128, cin=128, cout=128,
                    kind=_lib.CONV3X3_S1, act=_lib.ACT_BF16, out_fmt=_lib.OUT_NHWC, impl=_lib.IMPL_TCGEN05)
        base.update(kw)
        return _lib.ConvArgs(**base)

    assert lib.w2c_conv_fuses_bn_sums(ctypes.byref(args())) == 1
    assert lib.w2c_conv_fuses_bn_sums(ctypes.byref(args(kind=_lib.DECONV3X3_S2))) == 1
    assert lib.w2c_conv_fuses_bn_sums(ctypes.byref(args(act=_lib.ACT_FP16))) == 1
    assert lib.w2c_conv_fuses_bn_sums(ctypes.byref(args(act=_lib.ACT_BF16X2))) == 0       # two planes
    assert lib.w2c_conv_fuses_bn_sums(ctypes.byref(args(cout=72))) == 0                   # not whole 64-channel groups
    assert lib.w2c_conv_fuses_bn_sums(ctypes.byref(args(cout=11, out_fmt=_lib.OUT_NCHW_F32))) == 0
    assert lib.w2c_conv_fuses_bn_sums(ctypes.byref(args(n=1, h_in=8, w_in=8))) == 0       # sub-wave layer: one-tile kernel
    assert lib.w2c_conv_fuses_bn_sums(ctypes.byref(args(block_n=32))) == 0
    assert lib.w2c_conv_fuses_bn_sums(ctypes.byref(args(cin=48))) == -1                   # invalid arguments
    assert lib.w2c_conv_fuses_bn_sums(None) == -1
    # a launch that cannot accumulate the sums refuses them (before any device work)
    a = args(n=1, h_in=8, w_in=8, bn_sums=dummy)
    assert lib.w2c_conv_bnrelu_fwd(ctypes.byref(a), None) == -2
    assert b"bn_sums" in lib.w2c_last_error()


def test_batched_setup_calls_validate_items_on_the_host():
    lib = _lib.load()
    dummy = ctypes.c_void_p(256)
    items = (_lib.PackItem * 2)(_lib.PackItem(w=dummy, packed=dummy, cout=64, cin_real=64, cin=64, ntaps=9),
                                _lib.PackItem(w=dummy, packed=dummy, cout=64, cin_real=64, cin=48, ntaps=9))
    assert lib.w2c_pack_conv_weights_batch(items, 2, _lib.ACT_BF16, None) == -1
    assert b"item 1" in lib.w2c_last_error()
    assert lib.w2c_pack_conv_weights_batch(None, 0, _lib.ACT_BF16, None) == -1
    folds = (_lib.FoldItem * 1)(_lib.FoldItem(gamma=dummy, scale=dummy, shift=dummy, cout=8))
    assert lib.w2c_fold_bn_batch(folds, 1, None) == -1
    assert b"all present or all NULL" in lib.w2c_last_error()
