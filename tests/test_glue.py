"""Loader-side and evaluation-side glue of the hot path.

CPU: the oracle restatements of airsimLoader.transform and runningScore._fast_hist against vectors produced by the
UNMODIFIED reference code (tests/golden/make_golden_glue.py), and the host-built loader table against the oracle.
GPU: the fused CUDA paths (uint8 frames -> first conv; arg-max label map from the logits accumulators; device
confusion matrix) against the oracle. All of it is integer / table work: the bar is BIT-EXACT.
"""
import os

import numpy as np
import pytest
import torch

from multiagentperception_b200 import configs, ops, synth
from multiagentperception_b200.models import get_model
from oracle import when2com_oracle as orc
from tests.golden import make_golden_glue as gg

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "glue_loader_metrics.npz")


# ------------------------------------------------------------------------------------------------------ CPU
def test_oracle_loader_transform_matches_reference_vectors():
    g = np.load(GOLDEN)
    fr = gg.frames()
    for b in range(fr.shape[0]):
        for a in range(fr.shape[1]):
            got = orc.loader_transform(fr[b, a]).numpy()
            assert got.dtype == np.float32 and np.array_equal(got, g["transformed"][b, a])
    views = orc.views_from_frames(fr)
    assert tuple(views.shape) == (fr.shape[0], 3 * fr.shape[1], fr.shape[2], fr.shape[3])
    assert np.array_equal(views[1, 3:6].numpy(), g["transformed"][1, 1])  # agent a -> channels [3a, 3a+3)


def test_oracle_confusion_matches_reference_vectors():
    g = np.load(GOLDEN)
    gt, pred = gg.labels()
    hist = sum(orc.confusion_matrix(t, p, gg.N_CLASSES) for t, p in zip(gt, pred))  # runningScore.update
    assert np.array_equal(hist, g["confusion"])
    assert orc.mean_iou(hist) == pytest.approx(float(g["mean_iou"]), abs=1e-12)


def test_scores_from_confusion_match_reference_get_scores():
    """multiagentperception_b200.eval_loop.scores_from_confusion against runningScore.get_scores of the unmodified
    reference (same key strings, same values, per-class IoU)."""
    from multiagentperception_b200 import eval_loop
    g = np.load(GOLDEN)
    scores, cls_iu = eval_loop.scores_from_confusion(g["confusion"])
    assert sorted(scores) == [str(k) for k in g["score_keys"]]
    for k, v in zip(g["score_keys"], g["score_values"]):
        assert scores[str(k)] == pytest.approx(float(v), abs=1e-12)
    assert np.allclose([cls_iu[i] for i in range(gg.N_CLASSES)], g["class_iou"], atol=1e-12, equal_nan=True)


def test_loader_table_equals_the_transform_for_every_byte():
    lut = ops.loader_lut()
    assert tuple(lut.shape) == (3, 256) and lut.dtype == torch.float32
    ramp = np.repeat(np.arange(256, dtype=np.uint8)[:, None, None], 3, axis=2)  # (256, 1, 3): R = G = B = v
    ref = orc.loader_transform(ramp)                                             # (3, 256, 1), BGR channels
    assert torch.equal(lut, ref[:, :, 0])
    raw = ops.loader_lut(img_norm=False)
    assert torch.equal(raw, orc.loader_transform(ramp, img_norm=False)[:, :, 0])


def test_labels_from_logits_takes_the_first_maximum():
    x = torch.zeros(1, 4, 1, 3)
    x[0, 2, 0, 1] = 1.0
    x[0, 3, 0, 1] = 1.0
    assert orc.labels_from_logits(x).tolist() == [[[0, 2, 0]]]


# ------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("cout,act", [(64, ops.ACT_BF16), (128, ops.ACT_BF16), (128, ops.ACT_BF16X2)])
def test_stem_on_raw_frames_equals_stem_on_transformed_views(cout, act, cuda_device):
    dev = cuda_device
    g = torch.Generator().manual_seed(5)
    b, n, h, w = 2, 3, 40, 56            # ragged: 2*3*40*56 pixels is not a multiple of the 128-pixel tile
    frames = torch.randint(0, 256, (b, n, h, w, 3), dtype=torch.uint8, generator=g)
    views = orc.views_from_frames(frames.numpy())
    wt = (torch.randn(cout, 27, generator=g) * 0.2).to(dev)
    scale = (torch.rand(cout, generator=g) + 0.5).to(dev)
    shift = (torch.randn(cout, generator=g) * 0.1).to(dev)
    y_ref = ops.new_act(b * n, h, w, cout, act, dev)
    ops.stem_conv3x3(views.to(dev), wt, scale, shift, y_ref, b=b, n_agents=n, h=h, w=w, cout=cout, act=act)
    y_u8 = ops.new_act(b * n, h, w, cout, act, dev)
    ops.stem_conv3x3_u8(frames.to(dev), ops.loader_lut(device=dev), wt, scale, shift, y_u8, b=b, n_agents=n, h=h, w=w,
                        cout=cout, act=act)
    torch.cuda.synchronize()
    assert torch.equal(y_u8, y_ref)
    # a window of agents (sharded ranks convolve only their own): agents [1, 3)
    y_win = ops.new_act(b * 2, h, w, cout, act, dev)
    ops.stem_conv3x3_u8(frames.to(dev), ops.loader_lut(device=dev), wt, scale, shift, y_win, b=b, n_agents=2, h=h,
                        w=w, cout=cout, act=act, agents_total=n, agent_first=1)
    torch.cuda.synchronize()
    assert torch.equal(y_win, y_ref[b:])


@pytest.mark.gpu
@pytest.mark.parametrize("act", [ops.ACT_BF16, ops.ACT_BF16X2])
def test_7x7_stem_on_raw_frames_equals_stem_on_transformed_views(act, cuda_device):
    """resnet18's first layer (7x7 s2) on the loader's uint8 frames == on the host-transformed float views, bit for bit."""
    dev = cuda_device
    g = torch.Generator().manual_seed(6)
    b, n, h, w = 2, 2, 40, 56
    frames = torch.randint(0, 256, (b, n, h, w, 3), dtype=torch.uint8, generator=g)
    views = orc.views_from_frames(frames.numpy())
    wt = (torch.randn(64, 147, generator=g) * 0.08).to(dev)
    scale = (torch.rand(64, generator=g) + 0.5).to(dev)
    shift = (torch.randn(64, generator=g) * 0.1).to(dev)
    y_ref = ops.new_act(b * n, h // 2, w // 2, 64, act, dev)
    ops.stem_conv7x7s2(views.to(dev), wt, scale, shift, y_ref, b=b, n_agents=n, h=h, w=w, act=act)
    y_u8 = ops.new_act(b * n, h // 2, w // 2, 64, act, dev)
    ops.stem_conv7x7s2_u8(frames.to(dev), ops.loader_lut(device=dev), wt, scale, shift, y_u8, b=b, n_agents=n, h=h, w=w,
                          act=act)
    torch.cuda.synchronize()
    assert torch.equal(y_u8, y_ref)


@pytest.mark.gpu
@pytest.mark.parametrize("act", [ops.ACT_BF16, ops.ACT_BF16X2])
def test_fused_stem_pair_written_as_two_dense_maps(act, cuda_device):
    """The two encoders' fused 3 -> 128 first layer stored as two dense 64-channel maps equals the one 128-channel
    map, half by half (same accumulators, different tensor maps)."""
    dev = cuda_device
    g = torch.Generator().manual_seed(8)
    b, n, h, w = 2, 2, 24, 40
    x = torch.randn(b, 3 * n, h, w, generator=g).to(dev)
    wt = (torch.randn(128, 27, generator=g) * 0.2).to(dev)
    scale = (torch.rand(128, generator=g) + 0.5).to(dev)
    shift = (torch.randn(128, generator=g) * 0.1).to(dev)
    pl = ops.planes_of(act)
    one = ops.new_act(b * n, h, w, 128, act, dev)
    ops.stem_conv3x3(x, wt, scale, shift, one, b=b, n_agents=n, h=h, w=w, cout=128, act=act)
    two = torch.empty((2, b * n, h, w, pl * 64), dtype=torch.bfloat16, device=dev)
    ops.stem_conv3x3(x, wt, scale, shift, two, b=b, n_agents=n, h=h, w=w, cout=128, act=act, n_split=2)
    torch.cuda.synchronize()
    for half in range(2):
        for p in range(pl):
            assert torch.equal(two[half][..., p * 64:(p + 1) * 64],
                               one[..., p * 128 + half * 64:p * 128 + (half + 1) * 64])


@pytest.mark.gpu
@pytest.mark.parametrize("hw", [(16, 16), (24, 40), (128, 128)])
def test_fused_label_map_equals_argmax_of_the_logits(hw, cuda_device):
    dev = cuda_device
    g = torch.Generator().manual_seed(9)
    n, (h, w), cin, cout = 3, hw, 64, 11
    x = torch.randn(n, cin, h, w, generator=g).to(dev)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / 24.0).to(dev)
    scale = (torch.rand(cout, generator=g) + 0.5).to(dev)
    shift = (torch.randn(cout, generator=g) * 0.5 - 1.2).to(dev)   # negative shift: ReLU produces ties at 0
    xa = ops.nchw_to_act(x, ops.ACT_BF16)
    wp = ops.pack_conv_weight(wt, cin, False, ops.ACT_BF16)
    kw = dict(n=n, h_in=h, w_in=w, cin=cin, cout=cout, kind=ops.CONV3X3_S1, relu=True, act=ops.ACT_BF16,
              out_fmt=ops.OUT_NCHW_F32)
    plain = torch.empty(n, cout, h, w, device=dev)
    # same kernel as the label-writing launches (small maps otherwise dispatch to the one-tile kernel, whose tap
    # order - hence fp32 rounding - differs)
    ops.conv_bnrelu(xa, wp, scale, shift, plain, impl=ops.IMPL_TC_PERSIST, **kw)
    logits = torch.empty(n, cout, h, w, device=dev)
    labels = torch.full((n, h, w), 255, dtype=torch.uint8, device=dev)
    ops.conv_bnrelu(xa, wp, scale, shift, logits, labels=labels, **kw)
    only = torch.full((n, h, w), 255, dtype=torch.uint8, device=dev)
    ops.conv_bnrelu(xa, wp, scale, shift, None, labels=only, **kw)
    alone = ops.argmax_labels(logits)
    torch.cuda.synchronize()
    assert torch.equal(logits, plain)                              # asking for labels does not change the logits
    want = orc.labels_from_logits(logits.cpu())
    assert (logits == 0).all(1).any(), "test input should contain all-zero (tied) pixels"
    assert torch.equal(labels.cpu().long(), want)
    assert torch.equal(only, labels)
    assert torch.equal(alone, labels)


@pytest.mark.gpu
def test_evaluate_loop_equals_reference_style_evaluation(cuda_device):
    """eval_loop.evaluate (raw frames in, device label maps + device confusion matrix) against the reference-style loop:
    host transform -> forward -> max(1)[1] -> runningScore._fast_hist on the host."""
    from multiagentperception_b200 import eval_loop
    dev = cuda_device
    n = 3
    cfg = configs.make_config("MIMOcom", agent_num=n, img_size=128)
    model = get_model(cfg, 11)
    synth.randomize_(model, 1337)
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(4)
    batches = []
    for _ in range(3):
        frames = torch.randint(0, 256, (2, n, 128, 128, 3), dtype=torch.uint8, generator=g)
        labels = torch.randint(0, 12, (n * 2, 128, 128), dtype=torch.uint8, generator=g)
        labels[labels == 11] = 250           # ignore regions
        batches.append((frames, labels))
    kw = dict(training=False, MO_flag=True, inference="activated")
    hist = np.zeros((11, 11), dtype=np.int64)
    bw = []
    for frames, labels in batches:            # the reference-style loop on the same model (float views, host metrics)
        out = model(orc.views_from_frames(frames.numpy()).to(dev), **kw)
        pred = out[0].max(1)[1].cpu().numpy()
        hist += sum(orc.confusion_matrix(t, p, 11) for t, p in zip(labels.numpy(), pred))
        bw.append(out[3])
    want_scores, want_iou = eval_loop.scores_from_confusion(hist)
    scores, cls_iou, avg_bw = eval_loop.evaluate(model, batches, 11, kw)
    for k in want_scores:
        assert scores[k] == pytest.approx(want_scores[k], abs=1e-12)
    assert np.allclose([cls_iou[i] for i in range(11)], [want_iou[i] for i in range(11)], equal_nan=True)
    assert avg_bw == pytest.approx(sum(bw) / len(bw), abs=1e-12)
    assert model._w2c["io"]["u8"] is False    # the model's I/O format is restored


@pytest.mark.gpu
def test_device_confusion_matrix_equals_fast_hist(cuda_device):
    dev = cuda_device
    gref = np.load(GOLDEN)
    gt, pred = gg.labels()
    hist = torch.zeros(gg.N_CLASSES, gg.N_CLASSES, dtype=torch.int64, device=dev)
    ops.confusion_update(torch.from_numpy(pred).to(torch.uint8).to(dev), torch.from_numpy(gt).to(dev), gg.N_CLASSES,
                         hist)
    torch.cuda.synchronize()
    assert np.array_equal(hist.cpu().numpy(), gref["confusion"])
    # uint8 ground truth (ignore value 250), accumulated over two updates, at BASELINE size
    g = torch.Generator().manual_seed(3)
    gt8 = torch.randint(0, 12, (40, 512, 512), dtype=torch.uint8, generator=g)
    gt8[gt8 == 11] = 250
    pr8 = torch.randint(0, 11, (40, 512, 512), dtype=torch.uint8, generator=g)
    hist.zero_()
    for half in (slice(0, 20), slice(20, 40)):
        ops.confusion_update(pr8[half].contiguous().to(dev), gt8[half].contiguous().to(dev), 11, hist)
    torch.cuda.synchronize()
    want = orc.confusion_matrix(gt8.numpy(), pr8.numpy(), 11)
    assert np.array_equal(hist.cpu().numpy(), want)
    assert int(hist.sum()) == int((gt8 < 11).sum())


@pytest.mark.gpu
@pytest.mark.parametrize("arch,bb", [("MIMOcom", "n_segnet"), ("Single_agent", "n_segnet"), ("MIMOcom", "resnet")])
def test_model_on_raw_frames_with_label_output(arch, bb, cuda_device):
    """forward() on the loader's raw uint8 frames, returning the label map, against forward() on the host-transformed
    float views followed by max(1)[1]: bit-exact (same arithmetic, fused)."""
    dev = cuda_device
    n = 1 if arch == "Single_agent" else 3
    cfg = configs.make_config(arch, agent_num=n, img_size=128, backbones=bb)
    model = get_model(cfg, 11)
    synth.randomize_(model, 1337)
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(21)
    frames = torch.randint(0, 256, (2, n, 128, 128, 3), dtype=torch.uint8, generator=g)
    views = orc.views_from_frames(frames.numpy()).to(dev)
    kw = {} if arch == "Single_agent" else dict(training=False, MO_flag=True, inference="activated")
    ref = model(views, **kw)
    ref_pred = ref if torch.is_tensor(ref) else ref[0]

    model.set_label_output(True, logits=True)
    out = model(views, **kw)
    pred = out if torch.is_tensor(out) else out[0]
    # (the label-writing logits layer always runs in the persistent kernel; on small batches the plain forward may
    # use the one-tile kernel, whose tap order - hence fp32 rounding - differs in the last bits)
    assert float((pred - ref_pred).abs().max()) <= 1e-4 * float(ref_pred.abs().max())
    want = orc.labels_from_logits(pred.cpu())
    assert torch.equal(model.last_labels().cpu().long(), want)

    model.set_input_format("u8_hwc").set_label_output(True, logits=False)
    for _ in range(2):                                   # second call replays the CUDA graph
        out = model(frames.to(dev), **kw)
    lab = out if torch.is_tensor(out) else out[0]
    assert lab.dtype == torch.uint8 and tuple(lab.shape) == tuple(want.shape)
    assert torch.equal(lab.cpu().long(), want)
    if not torch.is_tensor(out):
        assert torch.equal(out[1], ref[1]) and torch.equal(out[2], ref[2]) and out[3] == ref[3]
    with pytest.raises(ValueError):
        model(views, **kw)                               # float views while the model expects raw frames
