"""GPU: the drop-in boundary exercised the way the reference's own callers use it.

  * the REFERENCE's Trainer_MIMOcom.load_weight + evaluate (ptsemseg/trainer.py:770-840) run unchanged against this
    repo's get_model() through the `dropin/ptsemseg` overlay package - the trainer, the metrics and the checkpoint
    key handling are the reference's code, only `ptsemseg.models` is ours;
  * DeviceScore (eval_loop.py) against the reference's runningScore on the same label maps: update, update_div,
    update_selection, get_*scores (metrics.py:19-199);
  * num_connect as a lazy device number: the reference's update_bandW / get_avg_bandW arithmetic works on it.
The reference package is the staged copy under oracle/_ref (or /root/reference in the build container).
"""
import os
import sys

import numpy as np
import pytest
import torch

from multiagentperception_b200 import configs, eval_loop, synth
from multiagentperception_b200.lazy import DeviceScalar
from multiagentperception_b200.models import get_model
from oracle import ref_harness
from oracle import when2com_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not ref_harness.available(), reason="no reference package (run oracle/make_ref.py)")

N, B, IMG, NCLS = 3, 2, 128, 11


def _loader_batches(n_batches, seed=3):
    """What airsimLoader + DataLoader hand the trainer (trainer.py:783-790): images_list (N x (B,3,H,W) float),
    labels_list (N x (B,H,W) int64), commun_label (B,2,N) int64; plus the raw frames they were made from."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n_batches):
        frames = synth.synthetic_frames(B, N, IMG, IMG, seed=seed + i)
        views = orc.views_from_frames(frames.numpy())                       # the loader transform, per frame
        images_list = [views[:, 3 * a:3 * a + 3].contiguous() for a in range(N)]
        labels_list = [torch.randint(0, NCLS + 2, (B, IMG, IMG), generator=g) for _ in range(N)]  # incl. ignore ids
        need = torch.randint(0, 2, (B, N), generator=g)
        who = torch.randint(0, N, (B, N), generator=g)
        commun = torch.stack((need, who), 1).to(torch.int64)
        out.append((frames, images_list, labels_list, commun))
    return out


def _overlay_trainer_module():
    """Import ptsemseg.trainer the way a user of the overlay would: dropin/ ahead of the reference on sys.path."""
    ref_harness._install_stubs()
    for name in [n for n in sys.modules if n == "ptsemseg" or n.startswith("ptsemseg.")]:
        del sys.modules[name]
    saved = list(sys.path)
    sys.path[:0] = [os.path.join(ROOT, "dropin"), ROOT, ref_harness.reference_root()]
    try:
        import ptsemseg
        import ptsemseg.models as models
        import ptsemseg.trainer as trainer
        import ptsemseg.visual as visual
    finally:
        sys.path[:] = saved
    assert ptsemseg.REFERENCE_PACKAGE_DIR is not None
    assert os.path.abspath(trainer.__file__).startswith(os.path.abspath(ref_harness.reference_root()))
    assert visual.draw_bounding() is None
    return models, trainer


@needs_ref
def test_reference_trainer_evaluates_the_b200_model(tmp_path, cuda_device):
    models, trainer_mod = _overlay_trainer_module()
    try:
        cfg = configs.make_config("MIMOcom", agent_num=N, img_size=IMG, backbones="n_segnet")
        cfg["data"]["commun_label"] = "mimo"
        cfg["model"]["precision"] = "bf16x3"
        model = models.get_model(cfg, NCLS)                       # <- the overlay's get_model: the B200 path
        assert type(model).__module__.startswith("multiagentperception_b200")
        # a checkpoint as train.py writes it: DataParallel 'module.' keys inside {"model_state": ...} (trainer.py:751-764)
        donor = get_model(cfg, NCLS)
        synth.randomize_(donor, 4242)
        path = str(tmp_path / "MIMOcom_airsim_best_model.pkl")
        torch.save({"epoch": 1, "model_state": {"module." + k: v for k, v in donor.state_dict().items()}}, path)
        model = model.to(cuda_device)
        tr = trainer_mod.Trainer_MIMOcom(cfg, None, None, model, None, None, None, None, None, cuda_device)
        tr.load_weight(path)                                      # convert_state_dict + load_state_dict(strict=False)
        for k, v in donor.state_dict().items():
            assert torch.equal(model.state_dict()[k].cpu(), v), k
        batches = _loader_batches(2)
        testloader = [(imgs, labs, commun) for _f, imgs, labs, commun in batches]
        score, class_iou = tr.evaluate(testloader)                # the reference's loop, 'activated' inference
        # the same evaluation through this repo's device-side loop on the raw frames: identical label maps ->
        # identical confusion matrix -> identical scores
        dev_batches = [(f, torch.cat(labs, 0), commun) for f, _imgs, labs, commun in batches]
        s2, iou2, bw2, ds = eval_loop.evaluate(model, dev_batches, NCLS, if_commun_label="mimo", return_score=True)
        for k in score:
            assert score[k] == pytest.approx(s2[k], rel=0, abs=1e-12), k
        for c in range(NCLS):
            a, b = class_iou[c], iou2[c]
            assert (np.isnan(a) and np.isnan(b)) or a == pytest.approx(b, abs=1e-12)
        # and against the fp32 oracle's predictions scored by the reference's runningScore
        metrics = ref_harness.import_reference_module("ptsemseg.metrics")
        rs = metrics.runningScore(NCLS)
        for _f, imgs, labs, _c in batches:
            out = orc.forward(donor.state_dict(), cfg, torch.cat(imgs, 1), training=False, MO_flag=True,
                              inference="activated")
            rs.update(torch.cat(labs, 0).numpy(), out[0].max(1)[1].numpy())
        ref_score, _ = rs.get_scores()
        assert score["Mean IoU : \t"] == pytest.approx(ref_score["Mean IoU : \t"], abs=2e-3)
        assert score["Overall Acc: \t"] == pytest.approx(ref_score["Overall Acc: \t"], abs=2e-3)
    finally:
        for name in [n for n in sys.modules if n == "ptsemseg" or n.startswith("ptsemseg.")]:
            del sys.modules[name]


@needs_ref
@pytest.mark.parametrize("mode", ["mimo", "when2com_argmax", "when2com_weights"])
def test_device_score_equals_reference_running_score(mode, cuda_device):
    metrics = ref_harness.import_reference_module("ptsemseg.metrics")
    g = torch.Generator().manual_seed(11)
    rs = metrics.runningScore(NCLS)
    ds = eval_loop.DeviceScore(NCLS, cuda_device)
    for _ in range(3):
        if mode == "mimo":
            n_img = N * B
            commun = torch.stack((torch.randint(0, 2, (B, N), generator=g), torch.randint(0, N, (B, N), generator=g)), 1)
            action = torch.randint(0, N, (B, N), generator=g)
            kind = "mimo"
        else:
            n_img = B * 3
            commun = torch.randint(-1, 4, (n_img,), generator=g)
            kind = "when2com"
            if mode == "when2com_argmax":
                action = torch.randint(0, 5, (n_img, 1), generator=g)
            else:
                w = torch.rand(n_img, 1, 5, generator=g)
                action = w * (w > 0.4).float()
        gt = torch.randint(0, NCLS + 2, (n_img, 64, 48), generator=g)
        pred = torch.randint(0, NCLS, (n_img, 64, 48), generator=g)
        rs.update(gt.numpy(), pred.numpy())
        rs.update_div(kind, gt.numpy(), pred.numpy(), commun.to(cuda_device))
        rs.update_selection(kind, commun.to(cuda_device), action.to(cuda_device))
        rs.update_bandW(1.5)
        gd, pd = gt.to(cuda_device), pred.to(cuda_device, torch.uint8)
        ds.update(gd, pd)
        ds.update_div(kind, gd, pd, commun.to(cuda_device))
        ds.update_selection(kind, commun.to(cuda_device), action.to(cuda_device))
        ds.update_bandW(DeviceScalar(torch.tensor(1.5, dtype=torch.float64, device=cuda_device)))
    assert np.array_equal(ds.confusion_matrix, rs.confusion_matrix.astype(np.int64))
    assert np.array_equal(ds.hist_pos.cpu().numpy(), rs.confusion_matrix_pos.astype(np.int64))
    assert np.array_equal(ds.hist_neg.cpu().numpy(), rs.confusion_matrix_neg.astype(np.int64))
    assert ds.get_selection_accuracy() == pytest.approx(rs.get_selection_accuracy(), abs=1e-12)
    assert ds.get_avg_bandW() == pytest.approx(rs.get_avg_bandW())
    for a, b in ((ds.get_only_normal_scores(), rs.get_only_normal_scores()),
                 (ds.get_only_noise_scores(), rs.get_only_noise_scores()), (ds.get_scores(), rs.get_scores())):
        for k in b[0]:
            assert a[0][k] == pytest.approx(b[0][k], abs=1e-12, nan_ok=True)


def test_num_connect_is_lazy_and_behaves_like_a_number(cuda_device):
    cfg = configs.make_config("MIMOcom", agent_num=N, img_size=IMG, backbones="n_segnet", precision="bf16x3")
    model = get_model(cfg, NCLS)
    synth.randomize_(model, 1337)
    x = synth.synthetic_views(B, N, IMG, IMG, seed=5)
    ref = orc.forward(model.state_dict(), cfg, x, training=False, MO_flag=True, inference="activated")
    model = model.to(cuda_device).eval()
    nc = [model(x.to(cuda_device), training=False, MO_flag=True, inference="activated")[3] for _ in range(3)]
    assert all(isinstance(v, DeviceScalar) for v in nc)
    total = 0
    for v in nc:                       # runningScore.update_bandW: `self.total_bandW += bandW`, metrics.py:19-21
        total += v
    assert isinstance(total, DeviceScalar) and total._v is None      # nothing has synchronised yet
    avg = total / 3.0                   # get_avg_bandW, metrics.py:110-111
    assert float(avg) == pytest.approx(ref[3], abs=1e-12)
    assert str(avg) == str(float(avg)) and avg == float(avg) and round(avg * 100, 2) == round(ref[3] * 100, 2)


class _Writer:
    """tensorboardX.SummaryWriter as the trainer uses it (add_scalar, file_writer.get_logdir)."""

    def __init__(self, logdir):
        self._dir = logdir
        self.file_writer = self
        self.scalars = []

    def get_logdir(self):
        return self._dir

    def add_scalar(self, tag, value, step):
        self.scalars.append((tag, float(value), step))


@needs_ref
def test_reference_trainer_trains_the_b200_model(tmp_path, cuda_device):
    """The REFERENCE's Trainer_MIMOcom.train() (ptsemseg/trainer.py:610-768) - scheduler.step, model.train(), forward,
    its own cross_entropy2d, loss.backward(), optimizer.step, the periodic validation pass and the best-model checkpoint
    - runs unchanged on this repo's model, and moves the weights the way the same loop moves the reference model on
    the CPU (same initial weights, same batches, plain SGD so that the parameter deltas are lr x gradient)."""
    import logging
    models, trainer_mod = _overlay_trainer_module()
    try:
        n, img, iters = 2, 256, 3
        cfg = configs.make_config("MIMOcom", agent_num=n, img_size=img, backbones="resnet")
        cfg["data"]["commun_label"] = "mimo"
        cfg["data"]["dataset"] = "airsim"
        cfg["model"]["precision"] = "bf16x3"
        cfg["training"] = {"train_iters": iters, "print_interval": 1, "val_interval": 2, "batch_size": B, "resume": None}
        loss_mod = ref_harness.import_reference_module("ptsemseg.loss.loss")
        g = torch.Generator().manual_seed(5)
        batches = []
        for i in range(iters):
            views = synth.synthetic_views(B, n, img, img, seed=20 + i)
            images_list = [views[:, 3 * a:3 * a + 3].contiguous() for a in range(n)]
            labels_list = [torch.randint(0, NCLS, (B, img, img), generator=g) for _ in range(n)]
            commun = torch.stack((torch.randint(0, 2, (B, n), generator=g), torch.randint(0, n, (B, n), generator=g)), 1)
            batches.append((images_list, labels_list, commun.to(torch.int64)))
        donor = get_model(cfg, NCLS)
        synth.randomize_(donor, 99)
        sd0 = {k: v.clone() for k, v in donor.state_dict().items()}

        def run(model, device, logdir):
            model.load_state_dict(sd0, strict=False)
            model = model.to(device)
            opt = torch.optim.SGD(model.parameters(), lr=0.02)
            sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1000)
            writer = _Writer(str(logdir))
            os.makedirs(str(logdir), exist_ok=True)
            tr = trainer_mod.Trainer_MIMOcom(cfg, writer, logging.getLogger("w2c-test"), model, loss_mod.cross_entropy2d,
                                             batches, batches[:1], opt, sched, device)
            import contextlib
            import io
            import warnings
            with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                path = tr.train()
            return model, path, writer

        ours, path, writer = run(models.get_model(cfg, NCLS), cuda_device, tmp_path / "b200")
        assert type(ours).__module__.startswith("multiagentperception_b200")
        assert os.path.isfile(path)                                # the best-model checkpoint of the validation pass
        losses = [v for tag, v, _ in writer.scalars if tag == "loss/train_loss"]
        assert len(losses) >= 2 and all(np.isfinite(losses))
        # the checkpoint holds the trained weights and loads back into a fresh model
        state = torch.load(path, weights_only=False)["model_state"]   # (the trainer stores a numpy best_iou next to the weights)
        fresh = get_model(cfg, NCLS)
        fresh.load_state_dict(state, strict=False)
        # the same loop on the UNMODIFIED reference model, on the CPU
        ref_model = ref_harness.build_reference_model(cfg)
        with ref_harness.cpu_cuda_shims():
            ref_model, _, ref_writer = run(ref_model, torch.device("cpu"), tmp_path / "ref")
        ref_losses = [v for tag, v, _ in ref_writer.scalars if tag == "loss/train_loss"]
        for a, b in zip(losses, ref_losses):
            assert a == pytest.approx(b, rel=5e-3), (losses, ref_losses)
        ours_sd, ref_sd = ours.state_dict(), ref_model.state_dict()
        num = den = 0.0
        for k, p0 in sd0.items():
            if not torch.is_floating_point(p0) or "running_" in k or k not in ref_sd:
                continue
            d_ref = (ref_sd[k].cpu() - p0).double()
            d_our = (ours_sd[k].cpu() - p0).double()
            num += float(((d_our - d_ref) ** 2).sum())
            den += float((d_ref ** 2).sum())
        assert den > 0
        assert (num / den) ** 0.5 <= 0.1, (num / den) ** 0.5   # three steps of lr x gradient: the gradient noise floor
    finally:
        for name in [n_ for n_ in sys.modules if n_ == "ptsemseg" or n_.startswith("ptsemseg.")]:
            del sys.modules[name]
