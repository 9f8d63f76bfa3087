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
    """runningScore (metrics.py:7-229) with every accumulator on the device: the overall / normal / noisy confusion
    matrices (update, update_div), the selection-accuracy counters (update_selection) and the bandwidth sum
    (update_bandW). Nothing synchronises with the host until a get_* method is called."""

    def __init__(self, n_classes, device):
        self.n_classes = n_classes
        self.device = device
        self.hist = torch.zeros(n_classes, n_classes, dtype=torch.int64, device=device)
        self.hist_pos = torch.zeros(n_classes, n_classes, dtype=torch.int64, device=device)
        self.hist_neg = torch.zeros(n_classes, n_classes, dtype=torch.int64, device=device)
        self.selection = torch.zeros(3, dtype=torch.int64, device=device)  # total_agent, correct_when2com, correct_who2com
        self.total_bandW = 0
        self.count = 0

    def update(self, label_trues, label_preds):
        """label_preds: uint8 device label maps from the model; label_trues: uint8 or int64 ground truth (device)."""
        ops.confusion_update(label_preds.contiguous(), label_trues.contiguous(), self.n_classes, self.hist)

    def update_div(self, if_commun_label, label_trues, label_preds, commun_label):
        """metrics.py:70-97: images of agents that do NOT need communication ("normal") go to confusion_matrix_pos,
        the others ("noisy") to confusion_matrix_neg. label maps are agent-major (N*B, H, W) for 'mimo' (the
        reference transposes commun_label[:, 0, :] to that order), (B, H, W) for 'when2com'."""
        if if_commun_label == "mimo":
            flag = (commun_label[:, 0, :] == 0).transpose(1, 0).reshape(-1)
        elif if_commun_label == "when2com":
            flag = (commun_label == -1).reshape(-1)
        else:
            raise ValueError("if_commun_label must be 'mimo' or 'when2com'")
        ops.confusion_update_div(label_preds.contiguous(), label_trues.contiguous(),
                                 flag.to(device=self.device, dtype=torch.uint8).contiguous(), self.n_classes,
                                 self.hist_pos, self.hist_neg)

    def update_selection(self, if_commun_label, commun_label, action_argmax):
        """metrics.py:23-68 (selection accuracy against the loader's communication labels)."""
        label = commun_label.to(device=self.device, dtype=torch.int64)
        if if_commun_label == "when2com":
            action = torch.squeeze(action_argmax)
            if action.dim() == 0:
                action = action.reshape(1)
            if action.dim() == 2:
                action = action.to(torch.float32)
            ops.selection_update(action, label.reshape(-1), "when2com", self.selection)
        elif if_commun_label == "mimo":
            ops.selection_update(action_argmax.to(torch.int64), label, "mimo", self.selection)
        else:
            raise ValueError("if_commun_label must be 'mimo' or 'when2com'")

    def update_bandW(self, bandW):  # runningScore.update_bandW, metrics.py:19-21 (bandW may be a lazy.DeviceScalar)
        self.total_bandW = self.total_bandW + bandW
        self.count += 1

    def get_avg_bandW(self):
        return float(self.total_bandW / self.count)

    def get_selection_accuracy(self):
        """metrics.py:196-200: (when2com accuracy, who2com accuracy), in percent like the reference."""
        total, when_ok, who_ok = (int(v) for v in self.selection.cpu())
        return when_ok / total * 100, who_ok / total * 100

    @property
    def confusion_matrix(self):
        return self.hist.cpu().numpy()

    def get_scores(self):
        return scores_from_confusion(self.confusion_matrix)

    def get_only_normal_scores(self):   # metrics.py:113-138
        return scores_from_confusion(self.hist_pos.cpu().numpy())

    def get_only_noise_scores(self):    # metrics.py:140-165
        return scores_from_confusion(self.hist_neg.cpu().numpy())


def evaluate(model, batches, n_classes, forward_kwargs=None, device=None, if_commun_label=None, return_score=False):
    """Run `model` over `batches` of (frames_u8 [B, N, H, W, 3] uint8 RGB, labels [N*B, H, W] uint8 / int64, agent-major
    like torch.cat(labels_list, 0) of trainer.py:790[, commun_label]) and return (scores, class_iou, avg_bandwidth)
    exactly as Trainer_MIMOcom.evaluate reports them. if_commun_label ('mimo' / 'when2com', trainer.py:605) also
    accumulates the normal / noisy split and the selection accuracy from the third batch element
    (trainer.py:806-809); return_score=True appends the DeviceScore. The model is switched to raw-frame input and
    label-map output for the duration of the call; no step synchronises with the host."""
    kw = dict(training=False, MO_flag=True, inference="activated") if forward_kwargs is None else dict(forward_kwargs)
    device = device or next(model.parameters()).device
    io_before = dict(model._w2c["io"])
    model.eval().set_input_format("u8_hwc").set_label_output(True, logits=False)
    score = DeviceScore(n_classes, device)
    try:
        for batch in batches:
            frames, labels = batch[0], batch[1]
            out = model(frames.to(device, non_blocking=True), **kw)
            labels_pred = out[0] if isinstance(out, tuple) else out
            labels_dev = labels.to(device, non_blocking=True)
            score.update(labels_dev, labels_pred)
            if isinstance(out, tuple) and len(out) >= 4:
                score.update_bandW(out[3])
            if if_commun_label and len(batch) > 2:
                commun = batch[2].to(device, non_blocking=True)
                score.update_div(if_commun_label, labels_dev, labels_pred, commun)
                score.update_selection(if_commun_label, commun, out[2])
    finally:
        model._w2c["io"].update(io_before)
        model._w2c["programs"].clear()
    scores, class_iou = score.get_scores()
    res = (scores, class_iou, (score.get_avg_bandW() if score.count else 0.0))
    return res + (score,) if return_score else res
