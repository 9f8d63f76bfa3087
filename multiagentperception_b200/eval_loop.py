"""Evaluation loop of the accelerated path: the device-side counterpart of Trainer_MIMOcom.evaluate
(ptsemseg/trainer.py:774-840) and of runningScore (ptsemseg/metrics.py:84-193).

The reference builds float views on the host, copies them to the device, runs the model, takes
`outputs.data.max(1)[1]`, copies the int64 label maps back and accumulates a confusion matrix in numpy per batch. Here
the raw uint8 frames go to the device as they are, the model returns uint8 label maps (loader transform fused into the
first conv, arg-max into the last one) and the confusion matrix accumulates on the device; ONE 11x11 int64 copy comes
back at the end. Label maps, confusion matrix and scores are identical to the reference's.
"""
import numpy as np
import torch

from . import ops


def scores_from_confusion(hist):
    """runningScore.get_scores (metrics.py:168-193) from a confusion matrix: ({'Overall Acc', 'Mean Acc', 'FreqW Acc',
    'Mean IoU'} with the reference's key strings, per-class IoU dict)."""
    hist = np.asarray(hist, dtype=np.float64)
    n = hist.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        acc = np.diag(hist).sum() / hist.sum()
        acc_cls = np.nanmean(np.diag(hist) / hist.sum(axis=1))
        iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist))
        mean_iu = np.nanmean(iu)
        freq = hist.sum(axis=1) / hist.sum()
        fwavacc = (freq[freq > 0] * iu[freq > 0]).sum()
    return ({"Overall Acc: \t": acc, "Mean Acc : \t": acc_cls, "FreqW Acc : \t": fwavacc, "Mean IoU : \t": mean_iu},
            dict(zip(range(n), iu)))


class DeviceScore:
    """runningScore's confusion matrix kept on the device (update = w2c_confusion_update, no host round trip)."""

    def __init__(self, n_classes, device):
        self.n_classes = n_classes
        self.hist = torch.zeros(n_classes, n_classes, dtype=torch.int64, device=device)
        self.total_bandW = 0.0
        self.count = 0

    def update(self, label_trues, label_preds):
        """label_preds: uint8 device label maps from the model; label_trues: uint8 or int64 ground truth (device)."""
        ops.confusion_update(label_preds.contiguous(), label_trues.contiguous(), self.n_classes, self.hist)

    def update_bandW(self, bandW):  # runningScore.update_bandW, metrics.py:93-95
        self.total_bandW += bandW
        self.count += 1

    def get_avg_bandW(self):
        return self.total_bandW / self.count

    @property
    def confusion_matrix(self):
        return self.hist.cpu().numpy()

    def get_scores(self):
        return scores_from_confusion(self.confusion_matrix)


def evaluate(model, batches, n_classes, forward_kwargs=None, device=None):
    """Run `model` over `batches` of (frames_u8 [B, N, H, W, 3] uint8 RGB, labels [N*B, H, W] uint8 / int64, agent-major
    like torch.cat(labels_list, 0) of trainer.py:790) and return (scores, class_iou, avg_bandwidth) exactly as
    Trainer_MIMOcom.evaluate reports them. The model is switched to raw-frame input and label-map output for the
    duration of the call."""
    kw = dict(training=False, MO_flag=True, inference="activated") if forward_kwargs is None else dict(forward_kwargs)
    device = device or next(model.parameters()).device
    io_before = dict(model._w2c["io"])
    model.eval().set_input_format("u8_hwc").set_label_output(True, logits=False)
    score = DeviceScore(n_classes, device)
    try:
        for frames, labels in batches:
            out = model(frames.to(device, non_blocking=True), **kw)
            labels_pred = out[0] if isinstance(out, tuple) else out
            score.update(labels.to(device, non_blocking=True), labels_pred)
            if isinstance(out, tuple) and len(out) >= 4:
                score.update_bandW(out[3])
    finally:
        model._w2c["io"].update(io_before)
        model._w2c["programs"].clear()
    scores, class_iou = score.get_scores()
    return scores, class_iou, (score.get_avg_bandW() if score.count else 0.0)
