"""TEST-SIDE STUDY (CPU, uses the oracle): which layers of the MIMOcom / n_segnet forward can run in a cheaper number
format before the logits leave the north star's 1e-3 bound?

The oracle's conv calls are intercepted and their operands rounded the way each kernel precision rounds them
(fp32 accumulation on the CPU stands in for the tensor core's fp32 accumulators):

    x3      activations and weights as bf16 hi + bf16 lo planes (16 mantissa bits; lo*lo dropped)  - 3 MMA passes
    fp16    both operands rounded to IEEE half                                                       - 1 pass
    bf16    both operands rounded to bfloat16                                                        - 1 pass
    f16a2   activations hi + lo in fp16 (22 bits), weights fp16 hi only                              - 2 passes
    f16w2   weights hi + lo, activations fp16 hi only                                                - 2 passes

Usage:  python tests/studies/precision_study.py [--img 512] [--out profiles/r2_precision_attribution.md]
Layer numbering: 0-13 u_encoder (conv1..13, squeezer), 14-27 query_key_net.img_encoder, 28-32 query_key_net.conv1-5,
33-44 decoder.deconv1-12.
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from multiagentperception_b200 import configs, synth  # noqa: E402
from multiagentperception_b200.models import get_model  # noqa: E402
from oracle import when2com_oracle as orc  # noqa: E402

N_LAYERS = 45
ENC, POL, DEC = range(0, 14), range(14, 33), range(33, 45)


def _x3(t):
    hi = t.bfloat16().float()
    return hi + (t - hi).bfloat16().float()


def _f16x2(t):
    hi = t.half().float()
    return hi + (t - hi).half().float()


QUANT = {  # mode -> (activation rounding, weight rounding)
    "x3": (_x3, _x3),
    "fp16": (lambda t: t.half().float(), lambda t: t.half().float()),
    "bf16": (lambda t: t.bfloat16().float(), lambda t: t.bfloat16().float()),
    "f16a2": (_f16x2, lambda t: t.half().float()),
    "f16w2": (lambda t: t.half().float(), _f16x2),
    "exact": (lambda t: t, lambda t: t),
}


class _FProxy:
    """Stands in for torch.nn.functional inside the oracle module: rounds the operands of every conv per the plan."""

    def __init__(self, real, plan):
        self._real, self._plan, self.i = real, plan, 0

    def __getattr__(self, name):
        return getattr(self._real, name)

    def _q(self, x, w):
        qa, qw = QUANT[self._plan[self.i]]
        self.i += 1
        return qa(x), qw(w)

    def conv2d(self, x, w, *a, **k):
        x, w = self._q(x, w)
        return self._real.conv2d(x, w, *a, **k)

    def conv_transpose2d(self, x, w, *a, **k):
        x, w = self._q(x, w)
        return self._real.conv_transpose2d(x, w, *a, **k)


def run_plan(sd, cfg, x, plan, kw):
    real = orc.F
    proxy = _FProxy(real, plan)
    orc.F = proxy
    try:
        out = orc.forward(sd, cfg, x, **kw)
    finally:
        orc.F = real
    assert proxy.i == N_LAYERS, proxy.i
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--img", type=int, default=512)
    ap.add_argument("--agents", type=int, default=5)
    ap.add_argument("--out", default="")
    ap.add_argument("--quick", action="store_true", help="groups only, no per-layer sweep")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = configs.make_config("MIMOcom", agent_num=args.agents, img_size=args.img, backbones="n_segnet")
    model = get_model(cfg, configs.N_CLASSES)
    synth.randomize_(model, 1337)
    sd = model.state_dict()
    x = synth.synthetic_views(1, args.agents, args.img, args.img, seed=1337)
    kw = dict(training=False, MO_flag=True, inference="softmax")
    t0 = time.time()
    ref = orc.forward(sd, cfg, x, **kw)
    print("reference forward %.1f s" % (time.time() - t0), flush=True)
    scale = float(ref[0].abs().max())
    rows = []

    def measure(name, plan):
        out = run_plan(sd, cfg, x, plan, kw)
        err = float((out[0] - ref[0]).abs().max()) / scale
        perr = float((out[1] - ref[1]).abs().max())
        miou = orc.miou_between(ref[0], out[0])
        rows.append((name, err, perr, miou))
        print("%-44s logits %.3e  prob %.3e  mIoU %.5f" % (name, err, perr, miou), flush=True)
        return err

    base = ["x3"] * N_LAYERS

    def with_mode(idx, mode):
        p = list(base)
        for i in idx:
            p[i] = mode
        return p

    measure("all x3", base)
    for mode in ("fp16", "bf16", "f16a2", "f16w2"):
        measure("all " + mode, [mode] * N_LAYERS)
    for mode in ("fp16", "bf16", "f16a2", "f16w2"):
        measure("policy net %s, rest x3" % mode, with_mode(POL, mode))
    for mode in ("fp16", "f16a2", "f16w2"):
        measure("u_encoder %s, rest x3" % mode, with_mode(ENC, mode))
        measure("decoder %s, rest x3" % mode, with_mode(DEC, mode))
    if not args.quick:
        for i in list(ENC) + list(DEC):
            measure("layer %2d fp16, rest x3" % i, with_mode([i], "fp16"))
        for i in list(ENC) + list(DEC):
            measure("layer %2d f16a2, rest x3" % i, with_mode([i], "f16a2"))
    if args.out:
        with open(args.out, "w") as f:
            f.write("# Per-layer precision attribution (CPU emulation on the oracle, %d agents @%dx%d, 1 scene)\n\n"
                    % (args.agents, args.img, args.img))
            f.write("Generated by `tests/studies/precision_study.py`. Error = max|logit - fp32 logit| / max|fp32 "
                    "logit|; bound 1e-3.\n\n| plan | logits err | prob_action err | mIoU vs fp32 argmax |\n|---|---|---|---|\n")
            for r in rows:
                f.write("| %s | %.3e | %.3e | %.5f |\n" % r)


if __name__ == "__main__":
    main()
