#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container, where /root/reference
is mounted):

    python tests/golden/make_golden.py

For every case in tests/cases.py the reference model is built with its own get_model(), filled with the seeded
synthetic weights, run on the seeded synthetic views on CPU (fp32), and its outputs stored:
  out0_sub   logits sub-sampled [:, :, ::4, ::4]   (full logits would be ~2 MB per case)
  out0_sum / out0_abs / out0_argmax_hist   whole-tensor summaries of the logits
  out1..     prob_action / action / num_connect, complete
The inputs and weights are not stored: synth regenerates them from the seeds in tests/cases.py.
"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from multiagentperception_b200 import synth  # noqa: E402
from oracle import ref_harness  # noqa: E402
from tests import cases  # noqa: E402


def summarise(outs):
    rec = {}
    logits = outs[0].double()
    rec["out0_sub"] = outs[0][:, :, ::4, ::4].contiguous().numpy()
    rec["out0_sum"] = np.float64(logits.sum().item())
    rec["out0_abs"] = np.float64(logits.abs().sum().item())
    rec["out0_argmax_hist"] = np.bincount(outs[0].max(1)[1].reshape(-1).numpy(), minlength=outs[0].shape[1])
    for i, o in enumerate(outs[1:], 1):
        rec["out%d" % i] = o.numpy() if torch.is_tensor(o) else np.float64(o)
    return rec


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    torch.set_num_threads(os.cpu_count())
    only = sys.argv[1:]   # optional case names: regenerate just these (the others stay byte-identical in git)
    for name in cases.CASES:
        if only and name not in only:
            continue
        cfg, kw, n = cases.case_config(name)
        model = ref_harness.build_reference_model(cfg)
        synth.randomize_(model, cases.WEIGHT_SEED)
        x = synth.synthetic_views(cases.BATCH, n, cases.IMG, cases.IMG, seed=cases.INPUT_SEED)
        random.seed(cases.RANDOM_SEED)
        outs = cases.as_tuple(ref_harness.reference_forward(model, x, **kw))
        np.savez_compressed(os.path.join(here, name + ".npz"), **summarise(outs))
        print("%-32s logits %s max %.3f" % (name, tuple(outs[0].shape), outs[0].abs().max().item()))


if __name__ == "__main__":
    main()
