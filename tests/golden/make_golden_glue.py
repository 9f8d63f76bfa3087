#!/usr/bin/env python
"""Generate tests/golden/glue_loader_metrics.npz from the UNMODIFIED reference loader / metrics code (run in the
build container, where /root/reference is mounted):

    python tests/golden/make_golden_glue.py

  * airsimLoader.transform (ptsemseg/loader/airsim_loader.py:515-540) is called, unbound, on seeded uint8 RGB
    frames; matplotlib (imported at the top of that file, absent here) is stubbed harness-side;
  * runningScore.update (ptsemseg/metrics.py:99-108) accumulates seeded label / prediction maps, including
    out-of-range ground-truth values (the loader's ignore regions) that _fast_hist must mask.
Inputs are regenerated from the seeds below by tests/test_glue.py; only the reference OUTPUTS are stored.
"""
import os
import sys
import types

import numpy as np

REF = os.environ.get("W2C_REFERENCE_ROOT", "/root/reference")
SEED = 4242
FRAME_SHAPE = (2, 3, 24, 40, 3)      # (B, N, H, W, RGB)
LABEL_SHAPE = (6, 24, 40)
N_CLASSES = 11


def frames(seed=SEED):
    return np.random.default_rng(seed).integers(0, 256, size=FRAME_SHAPE, dtype=np.uint8)


def labels(seed=SEED):
    rng = np.random.default_rng(seed + 1)
    gt = rng.integers(0, N_CLASSES, size=LABEL_SHAPE).astype(np.int64)
    gt[rng.random(LABEL_SHAPE) < 0.05] = 250     # ignore value outside [0, n_class)
    gt[rng.random(LABEL_SHAPE) < 0.02] = -1
    pred = rng.integers(0, N_CLASSES, size=LABEL_SHAPE).astype(np.int64)
    return gt, pred


def main():
    if "matplotlib" not in sys.modules:      # harness-side stub: the loader only calls matplotlib.use('Agg') at import
        mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        mpl.use = lambda *a, **k: None
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    sys.path.insert(0, REF)
    from ptsemseg.loader.airsim_loader import airsimLoader
    from ptsemseg.metrics import runningScore

    fake = types.SimpleNamespace(mean=np.array(airsimLoader.mean_rgb["airsim"]), img_norm=True,
                                 ignore_index=airsimLoader.ignore_index, n_classes=N_CLASSES)
    fr = frames()
    out = np.zeros(fr.shape[:2] + (3,) + fr.shape[2:4], dtype=np.float32)
    for b in range(fr.shape[0]):
        for a in range(fr.shape[1]):
            img, _ = airsimLoader.transform(fake, fr[b, a].copy(), np.zeros(fr.shape[2:4], dtype=np.uint8))
            out[b, a] = img.numpy()
    gt, pred = labels()
    rs = runningScore(N_CLASSES)
    rs.update(gt, pred)
    score, cls_iu = rs.get_scores()
    here = os.path.dirname(os.path.abspath(__file__))
    np.savez_compressed(os.path.join(here, "glue_loader_metrics.npz"), transformed=out,
                        confusion=rs.confusion_matrix.astype(np.int64),
                        mean_iou=np.float64([v for k, v in score.items() if "Mean IoU" in k][0]),
                        score_keys=np.array(sorted(score)), score_values=np.float64([score[k] for k in sorted(score)]),
                        class_iou=np.float64([cls_iu[i] for i in range(N_CLASSES)]))
    print("transformed", out.shape, float(out.min()), float(out.max()), "confusion sum", int(rs.confusion_matrix.sum()))


if __name__ == "__main__":
    main()
