"""Every YAML config the reference ships (configs/**/*.yml, snapshotted in tests/golden/shipped_cfgs.json) through
the drop-in boundary: get_model(cfg, n_classes) builds it with the reference's state_dict keys (CPU, against the live
reference when it is mounted), and its evaluation-time forward on the CUDA path matches the oracle (GPU)."""
import json
import os
import random

import pytest
import torch

from multiagentperception_b200 import synth
from multiagentperception_b200.models import get_model
from oracle import ref_harness
from oracle import when2com_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "shipped_cfgs.json")) as _f:
    SHIPPED = json.load(_f)

# how test.py / the trainers call each architecture at evaluation time (trainer.py:801 and siblings)
CALL = {
    "Single_agent": {}, "All_agents": {}, "MIMO_All_agents": {},
    "LearnWho2Com": dict(training=False, inference="argmax_test"),
    "LearnWhen2Com": dict(training=False, inference="activated"),
    "MIMOcom": dict(training=False, MO_flag=True, inference="activated"),
    "MIMOcomWho": dict(training=False, MO_flag=True, inference="activated"),
}


def _views_agents(cfg):
    arch = cfg["model"]["arch"]
    if arch == "Single_agent":
        return 1
    if arch in ("All_agents", "LearnWho2Com", "LearnWhen2Com"):
        return 5
    return cfg["model"]["agent_num"]


@pytest.mark.parametrize("name", sorted(SHIPPED))
def test_shipped_config_constructs_with_reference_keys(name):
    cfg = SHIPPED[name]
    model = get_model(cfg, 11)
    keys = set(model.state_dict().keys())
    assert keys, name
    if ref_harness.available():
        ref = ref_harness.build_reference_model(cfg)
        assert keys == set(ref.state_dict().keys())
        for k, v in ref.state_dict().items():
            assert tuple(v.shape) == tuple(model.state_dict()[k].shape), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SHIPPED))
def test_shipped_config_forward_matches_oracle(name, cuda_device):
    """512x512 (the size every shipped YAML sets), one scene, the parity precision: logits within 1e-3 of the logit
    range, the communication graph identical."""
    cfg = SHIPPED[name]
    arch = cfg["model"]["arch"]
    kw = CALL[arch]
    n = _views_agents(cfg)
    size = cfg["data"]["img_rows"]
    model = get_model(cfg, 11)
    synth.randomize_(model, 1337)
    x = synth.synthetic_views(1, n, size, size, seed=17)
    random.seed(5)
    ref = orc.forward(model.state_dict(), cfg, x, **kw)
    ref = ref if isinstance(ref, tuple) else (ref,)
    model = model.to(cuda_device).eval().set_precision("bf16x3")
    random.seed(5)
    out = model(x.to(cuda_device), **kw)
    out = out if isinstance(out, tuple) else (out,)
    assert len(out) == len(ref)
    scale = float(ref[0].abs().max())
    assert float((out[0].cpu() - ref[0]).abs().max()) <= 1e-3 * scale
    for o, r in zip(out[1:], ref[1:]):
        if torch.is_tensor(r):
            assert o.shape == r.shape
            if r.dtype == torch.int64:
                assert torch.equal(o.cpu(), r)
            else:
                assert float((o.cpu() - r).abs().max()) <= 1e-3
        else:
            assert float(o) == pytest.approx(float(r), abs=1e-9)
