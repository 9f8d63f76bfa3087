"""CPU: the drop-in boundary — get_model() surface, state_dict naming, error conventions, host-side plumbing."""
import pytest
import torch

from multiagentperception_b200 import configs, sharding, synth
from multiagentperception_b200.models import get_model


def _keys(arch, bb, **ov):
    return list(get_model(configs.make_config(arch, img_size=128, backbones=bb, **ov), 11).state_dict().keys())


def test_state_dict_key_counts_match_survey_appendix_d():
    # SURVEY.md appendix D [measured on the reference]: n_segnet MIMOcom 329 keys, resnet MIMOcom 551 keys
    assert len(_keys("MIMOcom", "n_segnet")) == 329
    assert len(_keys("MIMOcom", "resnet")) == 551


def test_state_dict_names():
    ks = set(_keys("MIMOcom", "n_segnet"))
    for k in ("u_encoder.feature_backbone.conv1.cbr_unit.0.weight",
              "u_encoder.feature_backbone.conv13.cbr_unit.1.running_var",
              "u_encoder.squeezer.cbr_unit.1.num_batches_tracked",
              "query_key_net.img_encoder.feature_backbone.conv7.cbr_unit.0.bias",
              "query_key_net.conv5.cbr_unit.0.weight",
              "key_net.fc.4.weight", "query_net.fc.0.bias", "attention_net.linear.weight",
              "decoder.output_decoder.deconv1.dcbr_unit.0.weight", "decoder.output_decoder.deconv12.cbr_unit.1.bias"):
        assert k in ks, k
    ks = set(_keys("Single_agent", "resnet"))
    for k in ("encoder.feature_backbone.feature_backbone.conv1.weight",
              "encoder.feature_backbone.feature_backbone.layer2.0.downsample.1.running_mean",
              "encoder.feature_backbone.feature_backbone.last_linear.bias",
              "encoder.feature_backbone.backbone_0.weight", "encoder.feature_backbone.backbone_1.3.1.bn2.weight",
              "encoder.feature_backbone.backbone_4.1.conv2.weight", "decoder.output_decoder.pred.2.bias"):
        assert k in ks, k


def test_transposed_conv_weight_layout():
    m = get_model(configs.make_config("Single_agent", backbones="n_segnet"), 11)
    sd = m.state_dict()
    assert tuple(sd["decoder.output_decoder.deconv7.dcbr_unit.0.weight"].shape) == (256, 256, 3, 3)
    assert tuple(sd["decoder.output_decoder.deconv12.cbr_unit.0.weight"].shape) == (11, 64, 3, 3)
    assert tuple(sd["decoder.output_decoder.deconv6.cbr_unit.0.weight"].shape) == (256, 512, 3, 3)


def test_unknown_arch_and_backbones_raise_value_error():
    with pytest.raises(ValueError, match="Model nope not available"):
        get_model({"model": {"arch": "nope"}, "data": {"img_rows": 128}}, 11)
    cfg = configs.make_config("Single_agent")
    cfg["model"]["enc_backbone"] = "vgg"
    with pytest.raises(ValueError, match="Encoder vgg not available"):
        get_model(cfg, 11)
    cfg = configs.make_config("Single_agent")
    cfg["model"]["dec_backbone"] = "FCN_decoder"  # broken in the reference (undefined base_4, backbone.py:179)
    with pytest.raises(ValueError, match="Decoder FCN_decoder not available"):
        get_model(cfg, 11)


def test_no_silent_fallbacks():
    m = get_model(configs.make_config("MIMOcom", agent_num=2, img_size=128), 11)
    x = synth.synthetic_views(1, 2, 128, 128)
    # train mode is a CUDA path too (batch-statistics BatchNorm, csrc/bn_train.cu): on a CPU tensor it fails loudly
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, training=True, MO_flag=True)
    m.eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, training=False, MO_flag=True, inference="softmax")
    with pytest.raises(ValueError, match="Incorrect inference mode"):
        m(x, training=False, MO_flag=True)  # default inference='argmax' is rejected like agent.py:1203-1204
    with pytest.raises(ValueError, match="MO_flag"):
        m(x, training=False, MO_flag=False, inference="softmax")
    with pytest.raises(ValueError, match="expected 6 channels"):
        m(synth.synthetic_views(1, 3, 128, 128), training=False, MO_flag=True, inference="softmax")
    with pytest.raises(RuntimeError, match="not executed eagerly"):
        m.u_encoder(x)


def test_parameter_groups_and_precision_switch():
    m = get_model(configs.make_config("MIMOcom", agent_num=3, img_size=128), 11)
    n_all = sum(p.numel() for p in m.parameters())
    assert sum(p.numel() for p in m.all_paras) == n_all
    assert len(m.policy_net_paras) > 0 and len(m.img_net_paras) > 0
    assert n_all == 55951627 - 0 or n_all > 5e7  # 55.95 M parameters for the n_segnet MIMOcom
    m.set_precision("bf16x3")
    with pytest.raises(ValueError):
        m.set_precision("fp8")
    cfg = configs.make_config("MIMOcom", precision="bf16x3")
    assert get_model(cfg, 11)._w2c["precision"] == "bf16x3"


def test_shipped_yaml_shapes_construct():
    # the ten shipped YAMLs, restated as dicts with the keys they carry (configs/**/*.yml)
    shipped = [
        ("Single_agent", dict(agent_num=5, shuffle_features="None")),
        ("Single_agent", dict(agent_num=6, shuffle_features="None")),
        ("All_agents", dict(agent_num=5, shuffle_features="selection")),
        ("MIMO_All_agents", dict(agent_num=6, shuffle_features="selection")),
        ("LearnWhen2Com", dict(agent_num=5, query_size=8, key_size=1024)),
        ("LearnWho2Com", dict(agent_num=5, query_size=8, key_size=1024)),
        ("MIMOcom", dict(agent_num=6, query_size=32, key_size=1024)),
        ("MIMOcomWho", dict(agent_num=6, query_size=32, key_size=1024, query=False)),
    ]
    for arch, ov in shipped:
        m = get_model(configs.make_config(arch, backbones="resnet", **ov), configs.N_CLASSES)
        assert isinstance(m, torch.nn.Module)


def test_synth_is_deterministic_and_order_independent():
    a = get_model(configs.make_config("Single_agent", backbones="resnet"), 11)
    b = get_model(configs.make_config("Single_agent", backbones="resnet"), 11)
    synth.randomize_(a, 5)
    synth.randomize_(b, 5)
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    x1 = synth.synthetic_views(2, 3, 32, 32, seed=9)
    x2 = synth.synthetic_views(2, 3, 32, 32, seed=9)
    assert torch.equal(x1, x2) and x1.shape == (2, 9, 32, 32)
    assert -0.5 < float(x1.min()) and float(x1.max()) < 0.61  # loader value range, SURVEY 8d


def test_shard_layout_geometry():
    lay = sharding.AgentShardLayout(agent_num=8, world=4, rank=2, batch=3, k_dim=1024, q_dim=32, h=16, w=16, c=512,
                                    planes=1)
    assert lay.apr == 2 and lay.first_agent == 4
    assert lay.keys_bytes == 2 * 3 * 1024 * 4 and lay.val_bytes == 2 * 3 * 256 * 512 * 2
    assert lay.queries_off % 256 == 0 and lay.val_slot_bytes % 256 == 0 and lay.kq_slot_bytes % 256 == 0
    assert lay.kq_region_off % 256 == 0 and lay.total_bytes == 4 * (lay.val_slot_bytes + lay.kq_slot_bytes)
    ex = lay.allocate("cpu")
    k, q, v = lay.views(ex)
    assert k.shape == (6, 1024) and q.shape == (6, 32) and v.shape == (6, 16, 16, 512)
    k.fill_(1.5)
    mine = lay.kq_region(ex)[2 * lay.kq_slot_bytes:3 * lay.kq_slot_bytes]
    assert float(mine[:lay.keys_bytes].view(torch.float32).sum()) == 1.5 * 6 * 1024
    assert float(ex.sum()) == float(mine.sum())
    with pytest.raises(ValueError):
        sharding.AgentShardLayout(5, 2, 0, 1, 8, 8, 4, 4, 8, 1)


def test_live_program_batches_its_setup_launches():
    """engine.Program._batch_setup_calls (host logic, no device): the per-layer pack / fold launches a train-mode program
    records are merged into one batched call per kind at the head of the program; everything else keeps its order."""
    import ctypes
    from multiagentperception_b200 import _lib, engine
    lib = _lib.load()
    prog = engine.Program(None, torch.device("cpu"), _lib.ACT_BF16)
    marker = object()
    prog._record(lib.w2c_pack_conv_weight, 1000, 64, 3, 64, 9, 0, _lib.ACT_BF16, 2000)
    prog._record(lib.w2c_fold_bn, None, 11, 12, 13, 14, 1e-5, 64, 3000, 3100)
    prog.calls.append((marker, ("conv-1",), 0))
    prog.calls.append((engine.Program._FORK, None, 0))
    prog._sid = 1
    prog._record(lib.w2c_pack_conv_weight_ex, 4000, 128, 64, 64, 9, 1, 1, _lib.ACT_BF16, 5000)
    prog._record(lib.w2c_fold_bn, 21, None, None, None, None, 1e-5, 128, 6000, 6100)
    prog.calls.append((marker, ("conv-2",), 1))
    prog._sid = 0
    prog.join()
    prog._batch_setup_calls()
    fns = [c[0] for c in prog.calls]
    assert fns[:2] == [lib.w2c_pack_conv_weights_batch, lib.w2c_fold_bn_batch]
    assert fns[2:] == [marker, engine.Program._FORK, marker, engine.Program._JOIN]
    assert [c[2] for c in prog.calls] == [0, 0, 0, 0, 1, 0]          # the batched calls run on the main stream, first
    items, n, act = prog.calls[0][1]
    assert n == 2 and act == _lib.ACT_BF16
    assert (items[0].w, items[0].packed, items[0].cout, items[0].cin_real, items[0].flip) == (1000, 2000, 64, 3, 0)
    assert (items[1].w, items[1].packed, items[1].transposed, items[1].flip) == (4000, 5000, 1, 1)
    folds, nf = prog.calls[1][1]
    assert nf == 2 and folds[0].gamma == 11 and folds[0].conv_bias is None and folds[1].conv_bias == 21
    assert folds[1].gamma is None and folds[1].cout == 128 and folds[1].scale == 6000
    before = list(prog.calls)
    prog._batch_setup_calls()                                          # idempotent
    assert prog.calls == before
    # an eval-mode program has no recorded setup launches: nothing to merge
    ev = engine.Program(None, torch.device("cpu"), _lib.ACT_BF16)
    ev.calls.append((marker, ("conv",), 0))
    ev._batch_setup_calls()
    assert [c[0] for c in ev.calls] == [marker]
