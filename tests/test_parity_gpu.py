"""GPU parity tests proper: the CUDA path, called through the model API / C ABI, against the CPU oracle and the
reference-generated golden vectors on the same seeded inputs.

Tolerances (stated here, used below):
  bf16x3 ("parity" precision): max |logit error| <= 1e-3 * max |logit|  (the north star's 1e-3 relative bound);
                               prob_action within 1e-3; action / num_connect exact; mIoU vs oracle argmax >= 0.995.
  fp16x3 / mixed:              the same 1e-3 bound. "fp16x3" = fp16 hi|lo planes, three passes; "mixed" = that storage
                               with ONE pass on the layers engine.MIXED_ONE_PASS lists (per-layer error attribution:
                               profiles/r2_precision_attribution.md) - the cheapest plan found that still meets 1e-3.
  bf16   ("fast" precision):   max |logit error| <= 5e-2 * max |logit|, mIoU >= 0.93 — bf16 activations through 27
                               stacked layers cannot meet 1e-3 (torch's own bf16 autocast of the
                               reference drifts 0.4-1.4e-2); measured drift is reported in DESIGN.md, not hidden.
"""
import os
import random

import numpy as np
import pytest
import torch

from multiagentperception_b200 import _lib, configs, ops, synth
from multiagentperception_b200.models import get_model
from oracle import when2com_oracle as orc
from tests import cases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

X3_LOGIT_TOL, X3_PROB_TOL, X3_MIOU = 1e-3, 1e-3, 0.995
BF16_LOGIT_TOL, BF16_MIOU = 5e-2, 0.93
# fp16 ("fast" speed, IEEE-half storage): eight times finer rounding than bf16 through the same 27 layers
FP16_LOGIT_TOL, FP16_MIOU = 8e-3, 0.985


def _run_case(name, dev, precision, graphs=True):
    cfg, kw, n = cases.case_config(name)
    model = get_model(cfg, 11)
    synth.randomize_(model, cases.WEIGHT_SEED)
    x = synth.synthetic_views(cases.BATCH, n, cases.IMG, cases.IMG, seed=cases.INPUT_SEED)
    random.seed(cases.RANDOM_SEED)   # the random-selection baselines draw from Python's `random`
    ref = cases.as_tuple(orc.forward(model.state_dict(), cfg, x, **kw))
    model = model.to(dev).eval().set_precision(precision).set_cuda_graphs(graphs)
    before = ops.launch_count()
    outs = None
    for _ in range(2):  # second call replays the captured graph
        random.seed(cases.RANDOM_SEED)
        outs = cases.as_tuple(model(x.to(dev), **kw))
    torch.cuda.synchronize()
    assert ops.launch_count() > before, "no libw2c launches: the CUDA path did not run"
    return ref, outs


def _rel(a, b):
    return float((a.double().cpu() - b.double()).abs().max()) / max(1e-12, float(b.double().abs().max()))


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_parity_bf16x3_against_oracle_and_golden(name, cuda_device):
    ref, outs = _run_case(name, cuda_device, "bf16x3")
    assert len(ref) == len(outs)
    assert outs[0].shape == ref[0].shape and outs[0].dtype == torch.float32
    assert _rel(outs[0], ref[0]) <= X3_LOGIT_TOL
    assert orc.miou_between(ref[0], outs[0].cpu()) >= X3_MIOU
    for o, r in zip(outs[1:], ref[1:]):
        if torch.is_tensor(r):
            assert o.shape == r.shape
            if r.dtype == torch.int64:
                assert torch.equal(o.cpu(), r)
            else:
                assert float((o.cpu() - r).abs().max()) <= X3_PROB_TOL
        else:
            assert float(o) == pytest.approx(float(r), abs=1e-9)
    # and against the vectors the unmodified reference produced
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sub = outs[0][:, :, ::4, ::4].cpu().numpy()
    assert np.abs(sub - g["out0_sub"]).max() <= X3_LOGIT_TOL * np.abs(g["out0_sub"]).max()
    assert abs(float(outs[0].double().sum()) - float(g["out0_sum"])) <= 1e-3 * float(g["out0_abs"])


@pytest.mark.parametrize("precision", ["mixed", "fp16x3"])
@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_parity_fp16_planes_against_oracle_and_golden(name, precision, cuda_device):
    """Every case again in the two fp16-plane precisions, held to the SAME 1e-3 bound as bf16x3."""
    ref, outs = _run_case(name, cuda_device, precision)
    assert _rel(outs[0], ref[0]) <= X3_LOGIT_TOL
    assert orc.miou_between(ref[0], outs[0].cpu()) >= X3_MIOU
    for o, r in zip(outs[1:], ref[1:]):
        if torch.is_tensor(r):
            if r.dtype == torch.int64:
                assert torch.equal(o.cpu(), r)
            else:
                assert float((o.cpu() - r).abs().max()) <= X3_PROB_TOL
        else:
            assert float(o) == pytest.approx(float(r), abs=1e-9)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sub = outs[0][:, :, ::4, ::4].cpu().numpy()
    assert np.abs(sub - g["out0_sub"]).max() <= X3_LOGIT_TOL * np.abs(g["out0_sub"]).max()


@pytest.mark.parametrize("name", ["single_segnet", "mimocom_segnet_activated", "mimocom_resnet_activated",
                                  "when2com_resnet_sparse"])
def test_parity_bf16_fast_precision(name, cuda_device):
    ref, outs = _run_case(name, cuda_device, "bf16")
    assert _rel(outs[0], ref[0]) <= BF16_LOGIT_TOL
    assert orc.miou_between(ref[0], outs[0].cpu()) >= BF16_MIOU


@pytest.mark.parametrize("name", ["single_segnet", "mimocom_segnet_softmax", "mimocom_resnet_activated",
                                  "when2com_resnet_sparse", "mimocomwho_segnet_argmax", "single_segnet_squeeze4"])
def test_parity_fp16_precision(name, cuda_device):
    ref, outs = _run_case(name, cuda_device, "fp16")
    assert _rel(outs[0], ref[0]) <= FP16_LOGIT_TOL
    assert orc.miou_between(ref[0], outs[0].cpu()) >= FP16_MIOU


def test_graph_replay_equals_eager(cuda_device):
    _, eager = _run_case("mimocom_segnet_activated", cuda_device, "bf16", graphs=False)
    _, graph = _run_case("mimocom_segnet_activated", cuda_device, "bf16", graphs=True)
    for a, b in zip(eager, graph):
        if torch.is_tensor(a):
            assert torch.equal(a, b)
        else:
            assert a == b


def test_full_size_properties(cuda_device):
    """BASELINE config 2 shape (5 agents, 512x512): size-independent properties instead of the slow CPU oracle."""
    dev = cuda_device
    n = 5
    cfg = configs.make_config("MIMOcom", agent_num=n, img_size=512)
    model = get_model(cfg, 11)
    synth.randomize_(model, 1337)
    model = model.to(dev).eval()
    x = synth.synthetic_views(2, n, 512, 512, seed=3).to(dev)
    kw = dict(training=False, MO_flag=True)
    pred, prob, action, nconn = model(x, inference="softmax", **kw)
    assert pred.shape == (2 * n, 11, 512, 512) and torch.isfinite(pred).all()
    assert (pred >= 0).all()  # logits pass BN+ReLU in n_segnet_decoder (backbone.py:124)
    # columns of prob_action (minus the 0.001 diagonal bias) are distributions over the supporting agents
    col = (prob - 0.001 * torch.eye(n, device=dev)).sum(1)
    assert torch.allclose(col, torch.ones_like(col), atol=1e-5)
    assert nconn == n - 1 and action.shape == (2, n)
    # scene independence: scene 0 alone gives the same prediction rows as scene 0 inside the batch
    pred1, prob1, _, _ = model(x[:1].contiguous(), inference="softmax", **kw)
    assert torch.allclose(prob1[0], prob[0], atol=1e-6)
    # bit-equal although the per-layer kernel dispatch depends on the batch: every tensor-core kernel accumulates K in
    # the same canonical order (ConvPlan::kw_major)
    assert torch.equal(pred1[:, :, :, :], pred[0::2])
    # agent permutation equivariance: swapping two agents' views swaps their predictions and permutes prob
    perm = [1, 0, 2, 3, 4]
    xp = torch.cat([x[:, 3 * p:3 * p + 3] for p in perm], 1).contiguous()
    predp, probp, _, _ = model(xp, inference="softmax", **kw)
    assert torch.allclose(probp, prob[:, perm][:, :, perm], atol=1e-5)
    base = pred.view(n, 2, 11, 512, 512)
    swapped = predp.view(n, 2, 11, 512, 512)
    assert float((swapped[0] - base[1]).abs().max()) <= 2e-2 * float(base.abs().max())
    # argmax_test == fusing exactly one supporter: its decoder input is one agent's own feature map, so with
    # action[b, j] == j the prediction equals that of a model fed only that agent (Single-agent consistency)
    pa, proba, acta, nca = model(x, inference="argmax_test", **kw)
    assert acta.shape == (2, n) and 0.0 <= nca <= 1.0
    assert int((acta != torch.arange(n, device=dev)).sum()) == round(nca * n * 2)


def test_full_size_one_scene_against_oracle(cuda_device):
    """BASELINE config 2 at full size (MIMOcom, 5 agents, 512x512, n_segnet pair), one scene: the CUDA path against
    the CPU oracle directly (the oracle needs a few seconds for one scene), both precisions."""
    dev = cuda_device
    cfg = configs.make_config("MIMOcom", agent_num=5, img_size=512)
    model = get_model(cfg, 11)
    synth.randomize_(model, 1337)
    x = synth.synthetic_views(1, 5, 512, 512, seed=1337)
    sd = model.state_dict()
    soft = dict(training=False, MO_flag=True, inference="softmax")
    act = dict(training=False, MO_flag=True, inference="activated")
    ref_soft = orc.forward(sd, cfg, x, **soft)
    ref_act = orc.forward(sd, cfg, x, **act)
    model = model.to(dev).eval()
    # logits in both precisions on the continuous (softmax-fusion) path
    for prec, tol, miou in (("bf16x3", X3_LOGIT_TOL, X3_MIOU), ("fp16x3", X3_LOGIT_TOL, X3_MIOU),
                            ("mixed", X3_LOGIT_TOL, X3_MIOU), ("bf16", BF16_LOGIT_TOL, BF16_MIOU),
                            ("fp16", FP16_LOGIT_TOL, FP16_MIOU)):
        pred, prob, action, nconn = model.set_precision(prec)(x.to(dev), **soft)
        assert _rel(pred, ref_soft[0]) <= tol
        assert orc.miou_between(ref_soft[0], pred.cpu()) >= miou
    # the thresholded path ('activated': P > 0.2 re-selection, second decoder pass) in the parity precision: the
    # communication graph must come out identical
    pred, prob, action, nconn = model.set_precision("bf16x3")(x.to(dev), **act)
    assert _rel(pred, ref_act[0]) <= X3_LOGIT_TOL
    assert float((prob.cpu() - ref_act[1]).abs().max()) <= X3_PROB_TOL
    assert torch.equal(action.cpu(), ref_act[2]) and nconn == pytest.approx(ref_act[3], abs=1e-9)


@pytest.mark.parametrize("hw", [(96, 160), (32, 32), (224, 64)])
def test_single_agent_ragged_sizes_against_oracle(hw, cuda_device):
    """Non-square and minimum sizes (any multiple of 32 is legal for Single_agent): partial 8x16 tiles at every
    scale, 1x1 feature maps at 32x32; odd batch."""
    dev = cuda_device
    h, w = hw
    cfg = configs.make_config("Single_agent", img_size=max(h, w))
    model = get_model(cfg, 11)
    synth.randomize_(model, 1337)
    x = synth.synthetic_views(3, 1, h, w, seed=9)
    ref = orc.forward(model.state_dict(), cfg, x)
    model = model.to(dev).eval()
    for prec, tol in (("bf16x3", X3_LOGIT_TOL), ("bf16", BF16_LOGIT_TOL)):
        pred = model.set_precision(prec)(x.to(dev))
        assert pred.shape == (3, 11, h, w)
        assert _rel(pred, ref) <= tol


def test_single_agent_1024_against_oracle(cuda_device):
    """BASELINE config 4 shape (Single_agent, n_segnet pair, 1024x1024; one view - the oracle takes ~10 s for it), in
    config 4's own dtype (fp16) as well as the parity and bf16 precisions."""
    dev = cuda_device
    cfg = configs.make_config("Single_agent", img_size=1024)
    model = get_model(cfg, 11)
    synth.randomize_(model, 1337)
    x = synth.synthetic_views(1, 1, 1024, 1024, seed=5)
    ref = orc.forward(model.state_dict(), cfg, x)
    model = model.to(dev).eval()
    for prec, tol, miou in (("bf16x3", X3_LOGIT_TOL, X3_MIOU), ("mixed", X3_LOGIT_TOL, X3_MIOU),
                            ("bf16", BF16_LOGIT_TOL, BF16_MIOU), ("fp16", FP16_LOGIT_TOL, FP16_MIOU)):
        pred = model.set_precision(prec)(x.to(dev))
        assert pred.shape == (1, 11, 1024, 1024)
        assert _rel(pred, ref) <= tol
        assert orc.miou_between(ref, pred.cpu()) >= miou


def test_tc_conv_matches_simt_crosscheck_at_full_size(cuda_device):
    """The tcgen05 kernel against the SIMT evaluation of the same packed operands at a BASELINE-size layer."""
    dev = cuda_device
    g = torch.Generator().manual_seed(0)
    n, h, w, cin, cout = 5, 128, 128, 128, 256
    x = torch.randn(n, cin, h, w, generator=g).to(dev)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev)
    scale = (torch.rand(cout, generator=g) + 0.5).to(dev)
    shift = (torch.randn(cout, generator=g) * 0.1).to(dev)
    for act in (ops.ACT_BF16, ops.ACT_BF16X2):
        xa = ops.nchw_to_act(x, act)
        wp = ops.pack_conv_weight(wt, cin, False, act)
        outs = []
        for impl in (ops.IMPL_TCGEN05, ops.IMPL_SIMT):
            y = ops.new_act(n, h, w, cout, act, dev)
            ops.conv_bnrelu(xa, wp, scale, shift, y, n=n, h_in=h, w_in=w, cin=cin, cout=cout, kind=ops.CONV3X3_S1,
                            relu=True, act=act, impl=impl)
            outs.append(ops.act_to_nchw(y, cout, act))
        torch.cuda.synchronize()
        err = float((outs[0] - outs[1]).abs().max())
        # same operands; the tensor core accumulates K = 1152 fp32 terms in its own order with truncating adds
        # (measured bias ~1e-4 relative), and bf16 storage may then flip one ulp (2^-8 relative)
        assert err <= (8e-3 if act == ops.ACT_BF16 else 5e-4) * float(outs[1].abs().max())


def test_library_is_the_loaded_native_code():
    path = _lib.lib_path()
    assert os.path.exists(path)
    with open("/proc/self/maps") as f:
        assert any("libw2c.so" in line for line in f), "libw2c.so is not mapped into this process"
