#!/usr/bin/env python
"""Whole-model backward check on a B200: model.train(); loss = cross_entropy2d(model(x), labels); loss.backward()
through the CUDA path against the oracle's autograd gradients (pinned to the reference's, tests/test_oracle.py).
Prints one JSON line per case with per-parameter relative errors. Diagnostic twin of tests/test_backward_gpu.py."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from multiagentperception_b200 import configs, synth  # noqa: E402
from multiagentperception_b200.models import get_model  # noqa: E402
from oracle import when2com_oracle as orc  # noqa: E402

CASES = {
    "single_segnet": ("Single_agent", "n_segnet", {}, {}, 1, 128),
    "mimocom_segnet": ("MIMOcom", "n_segnet", dict(agent_num=3), dict(training=True, MO_flag=True), 3, 256),
    "single_resnet": ("Single_agent", "resnet", {}, {}, 1, 128),
    "mimocom_resnet": ("MIMOcom", "resnet", dict(agent_num=3), dict(training=True, MO_flag=True), 3, 256),
    "when2com_segnet": ("LearnWhen2Com", "n_segnet", dict(query_size=8, attention="general"), dict(training=True), 5, 128),
    "who2com_resnet": ("LearnWho2Com", "resnet", dict(query_size=8, attention="general"), dict(training=True), 5, 128),
    "mimocomwho_segnet": ("MIMOcomWho", "n_segnet", dict(agent_num=3), dict(training=True, MO_flag=True), 3, 256),
    "mimo_all_resnet": ("MIMO_All_agents", "resnet", dict(agent_num=3), {}, 3, 128),
}


def loss_fn(pred, labels):
    n, c, h, w = pred.shape
    return F.cross_entropy(pred.permute(0, 2, 3, 1).reshape(-1, c), labels.reshape(-1), ignore_index=250)


def run_case(name, precision="bf16x3", batch=2, steps=1):
    arch, bb, over, kw, n, img = CASES[name]
    dev = torch.device("cuda:0")
    cfg = configs.make_config(arch, img_size=img, backbones=bb, **over)
    model = get_model(cfg, 11)
    synth.randomize_(model, 1337)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    x = synth.synthetic_views(batch, n, img, img, seed=7)
    n_img = batch * (n if (kw.get("MO_flag") or arch == "MIMO_All_agents") else 1)
    labels = torch.randint(0, 11, (n_img, img, img), generator=torch.Generator().manual_seed(3))
    labels[0, :8] = 250
    out_o, loss_o, g_orc = orc.forward_with_grads(sd0, cfg, x, labels, **kw)
    model = model.to(dev).set_precision(precision)
    model.train()
    res = {"case": name, "precision": precision}
    for step in range(steps):
        model.zero_grad(set_to_none=True)
        out = model(x.to(dev), **kw)
        pred = out[0] if isinstance(out, tuple) else out
        loss = loss_fn(pred, labels.to(dev))
        loss.backward()
    torch.cuda.synchronize()
    res["loss"] = (float(loss.detach()), loss_o)
    named = dict(model.named_parameters())
    rows = []
    num = den = 0.0
    missing = []
    for k, g in g_orc.items():
        p = named.get(k)
        if p is None:
            continue
        if p.grad is None:
            missing.append(k)
            continue
        d = (p.grad.detach().cpu().double() - g.double())
        num += float((d ** 2).sum())
        den += float((g.double() ** 2).sum())
        rows.append((k, float(d.abs().max()), float(g.abs().max()), float(d.norm() / max(g.double().norm(), 1e-30))))
    extra = [k for k, p in named.items() if p.grad is not None and k not in g_orc]
    res["missing"] = missing
    res["extra"] = extra[:10]
    res["global_rel_l2"] = (num / max(den, 1e-60)) ** 0.5
    res["all"] = [(k, "%.1e" % m, "%.1e" % r) for k, e, m, r in rows]
    # a conv bias in front of a train-mode BatchNorm has a mathematically zero gradient: rounding noise on both sides
    # (held to an absolute bound instead: the CUDA path returns exact zeros there)
    # Likewise the last bias of key_net: it shifts every key by the same vector, which the softmax over the keys ignores.
    gmax = max(r[2] for r in rows)
    zero_grad = lambda r: r[0].endswith("cbr_unit.0.bias") or r[0] == "key_net.fc.4.bias" or r[2] < 1e-7 * gmax
    res["zero_grad_max_over_gmax"] = max([r[1] for r in rows if zero_grad(r)] or [0.0]) / gmax
    rows = [r for r in rows if not zero_grad(r)]
    rows.sort(key=lambda r: -r[3])
    res["worst"] = [(k, "%.2e" % e, "%.2e" % m, "%.2e" % r) for k, e, m, r in rows[:8]]
    res["max_rel_l2"] = rows[0][3] if rows else None
    res["n_params"] = len(rows)
    return res


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["single_segnet", "mimocom_segnet"]
    prec = "bf16x3"
    for a in sys.argv[1:]:
        if a.startswith("--precision="):
            prec = a.split("=", 1)[1]
    for nm in names:
        try:
            r = run_case(nm, prec)
        except Exception as e:  # noqa: BLE001
            import traceback
            r = {"case": nm, "error": "%s: %s" % (type(e).__name__, e), "trace": traceback.format_exc()[-2500:]}
        print(json.dumps(r), flush=True)
