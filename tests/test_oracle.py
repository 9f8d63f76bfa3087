"""CPU: the oracle restatement against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py), and — where /root/reference is mounted — against the live reference."""
import os
import random

import numpy as np
import pytest
import torch

from multiagentperception_b200 import configs, synth
from multiagentperception_b200.models import get_model
from oracle import ref_harness
from oracle import when2com_oracle as orc
from tests import cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-5  # fp32 restatement vs fp32 reference: relative to the largest logit


def _oracle_outputs(name):
    cfg, kw, n = cases.case_config(name)
    model = get_model(cfg, 11)
    synth.randomize_(model, cases.WEIGHT_SEED)
    x = synth.synthetic_views(cases.BATCH, n, cases.IMG, cases.IMG, seed=cases.INPUT_SEED)
    random.seed(cases.RANDOM_SEED)   # the random-selection baselines draw from Python's `random`, like the reference
    return cases.as_tuple(orc.forward(model.state_dict(), cfg, x, **kw)), model, cfg, x, kw


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_oracle_matches_reference_golden(name):
    outs, *_ = _oracle_outputs(name)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    logits = outs[0]
    scale = float(np.abs(g["out0_sub"]).max())
    sub = logits[:, :, ::4, ::4].numpy()
    assert sub.shape == g["out0_sub"].shape
    assert np.abs(sub - g["out0_sub"]).max() <= TOL * scale
    assert abs(float(logits.double().sum()) - float(g["out0_sum"])) <= 1e-6 * float(g["out0_abs"])
    assert abs(float(logits.double().abs().sum()) - float(g["out0_abs"])) <= 1e-6 * float(g["out0_abs"])
    hist = np.bincount(logits.max(1)[1].reshape(-1).numpy(), minlength=logits.shape[1])
    # argmax ties on exact zeros (BN+ReLU logits) may fall differently only if values differ: allow a whisker
    assert np.abs(hist - g["out0_argmax_hist"]).sum() <= 1e-4 * hist.sum()
    for i, o in enumerate(outs[1:], 1):
        ref = g["out%d" % i]
        if torch.is_tensor(o):
            assert tuple(o.shape) == ref.shape
            if o.dtype in (torch.int64, torch.int32):
                assert np.array_equal(o.numpy(), ref)
            else:
                assert np.abs(o.numpy() - ref).max() <= 1e-5
        else:
            assert abs(float(o) - float(ref)) <= 1e-9


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("name", ["mimocom_segnet_activated", "when2com_resnet_sparse", "who2com_resnet_argmax",
                                  "who2com_resnet_normal_agents", "mimo_all_resnet_selection",
                                  "all_agents_resnet_selection"])
def test_oracle_matches_live_reference(name):
    outs, model, cfg, x, kw = _oracle_outputs(name)
    ref = ref_harness.build_reference_model(cfg)
    missing = ref.load_state_dict(model.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    random.seed(cases.RANDOM_SEED)
    routs = cases.as_tuple(ref_harness.reference_forward(ref, x, **kw))
    assert len(routs) == len(outs)
    for a, b in zip(routs, outs):
        if torch.is_tensor(a):
            assert a.shape == b.shape
            assert float((a.double() - b.double()).abs().max()) <= TOL * max(1.0, float(a.double().abs().max()))
        else:
            assert abs(float(a) - float(b)) <= 1e-9


def test_sparsemax_properties():
    z = torch.randn(7, 5, 1)
    p = orc.sparsemax_dim1(z)
    assert torch.all(p >= 0)
    assert torch.allclose(p.sum(1), torch.ones(7, 1), atol=1e-6)
    # a dominant logit takes all the mass
    z = torch.tensor([[[10.0], [0.0], [-1.0]]])
    assert torch.equal(orc.sparsemax_dim1(z), torch.tensor([[[1.0], [0.0], [0.0]]]))


def test_miou_metric():
    ref = torch.zeros(1, 3, 2, 2)
    ref[0, 0, 0, :] = 1
    ref[0, 1, 1, :] = 1
    got = ref.clone()
    assert orc.miou_between(ref, got, n_class=3) == 1.0
    got[0, :, 0, 0] = torch.tensor([0.0, 2.0, 0.0])  # one pixel of class 0 predicted as class 1
    m = orc.miou_between(ref, got, n_class=3)
    assert abs(m - (0.5 + 2.0 / 3.0) / 2) < 1e-12


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("arch,bb,over,kw,n", [
    ("MIMOcom", "n_segnet", dict(agent_num=3), dict(training=True, MO_flag=True), 3),
    ("MIMOcom", "resnet", dict(agent_num=2), dict(training=True, MO_flag=True), 2),
    ("Single_agent", "n_segnet", dict(feat_squeezer=2), {}, 1),
    ("LearnWhen2Com", "resnet", dict(query_size=8), dict(training=True), 5),
])
def test_train_mode_oracle_equals_the_reference_in_train_mode(arch, bb, over, kw, n):
    """model.train() forward (trainer.py:659-669): batch-statistics BatchNorm and the running-stat update of the
    oracle's train_stats path against the UNMODIFIED reference modules in train() mode, buffers included."""
    import torch
    cfg = configs.make_config(arch, img_size=128, backbones=bb, **over)
    ref = ref_harness.build_reference_model(cfg)
    synth.randomize_(ref, 1337)
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    x = synth.synthetic_views(2, n, 128, 128, seed=7)
    ref.train()
    out = ref_harness.reference_forward(ref, x, **kw)
    stats = {}
    mine = orc.forward(sd0, cfg, x, train_stats=stats, **kw)
    a = out[0] if isinstance(out, tuple) else out
    b = mine[0] if isinstance(mine, tuple) else mine
    assert float((a - b).abs().max()) <= 2e-5 * float(a.abs().max())
    after = ref.state_dict()
    assert stats, "no BatchNorm layer reported statistics"
    for k, v in stats.items():
        if torch.is_tensor(v):
            assert float((after[k] - v).abs().max()) <= 1e-6 * max(1.0, float(v.abs().max())), k
        else:
            assert int(after[k]) == v, k


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("arch,bb,over,kw,n", [
    ("MIMOcom", "n_segnet", dict(agent_num=2), dict(training=True, MO_flag=True), 2),
    ("MIMOcom", "resnet", dict(agent_num=2), dict(training=True, MO_flag=True), 2),
    ("Single_agent", "n_segnet", {}, {}, 1),
    ("LearnWhen2Com", "resnet", dict(query_size=8), dict(training=True), 5),
])
def test_oracle_gradients_equal_the_reference_autograd(arch, bb, over, kw, n):
    """The backward pass (trainer.py:668-670): the oracle's forward_with_grads against loss.backward() through the
    UNMODIFIED reference modules in train() mode with the reference's own cross_entropy2d - every parameter gradient."""
    import torch
    cfg = configs.make_config(arch, img_size=128, backbones=bb, **over)
    ref = ref_harness.build_reference_model(cfg)
    synth.randomize_(ref, 1337)
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    x = synth.synthetic_views(2, n, 128, 128, seed=7)
    n_img = 2 * (n if kw.get("MO_flag") else 1)
    labels = torch.randint(0, 11, (n_img, 128, 128), generator=torch.Generator().manual_seed(3))
    labels[0, :8] = 250     # an ignored region, like the loader's void label
    _, loss_r, g_ref = ref_harness.reference_train_step_grads(ref, x, labels, **kw)
    _, loss_o, g_orc = orc.forward_with_grads(sd0, cfg, x, labels, **kw)
    assert abs(loss_r - loss_o) <= 1e-5 * abs(loss_r)
    assert g_ref and set(g_ref) == set(g_orc)
    for k, g in g_ref.items():
        # (a conv bias in front of a train-mode BatchNorm has a mathematically zero gradient: both sides hold ~1e-7 of
        # rounding noise there, hence the absolute floor)
        scale = float(g.abs().max())
        assert float((g - g_orc[k]).abs().max()) <= 1e-2 * scale + 2e-6, k   # fp32 vs fp32 in another summation order: ~6e-3 seen
