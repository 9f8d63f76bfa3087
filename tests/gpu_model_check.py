#!/usr/bin/env python
"""Whole-model parity report on a B200: every case of tests/cases.py through the CUDA path (both precisions, with
and without CUDA graphs) against the CPU oracle on the same seeded inputs. Writes gpurun_out/model_check.jsonl.
Diagnostic tool; the pass/fail gates live in tests/test_parity_gpu.py.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from multiagentperception_b200 import synth  # noqa: E402
from multiagentperception_b200.models import get_model  # noqa: E402
from oracle import when2com_oracle as orc  # noqa: E402
from tests import cases  # noqa: E402


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max()) / max(1e-12, float(b.abs().max()))


def main():
    names = sys.argv[1:] or list(cases.CASES)
    out_path = os.path.join(ROOT, "gpurun_out", "model_check.jsonl")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    dev = torch.device("cuda:0")
    with open(out_path, "w") as f:
        for name in names:
            cfg, kw, n = cases.case_config(name)
            rec = {"case": name}
            try:
                model = get_model(cfg, 11)
                synth.randomize_(model, cases.WEIGHT_SEED)
                x = synth.synthetic_views(cases.BATCH, n, cases.IMG, cases.IMG, seed=cases.INPUT_SEED)
                t0 = time.time()
                ref = cases.as_tuple(orc.forward(model.state_dict(), cfg, x, **kw))
                rec["oracle_s"] = round(time.time() - t0, 2)
                model = model.to(dev).eval()
                for prec in ("bf16x3", "bf16"):
                    for graphs in (False, True):
                        model.set_precision(prec).set_cuda_graphs(graphs)
                        outs = None
                        for _ in range(3 if graphs else 1):  # graph: eager+capture, replay, replay
                            outs = cases.as_tuple(model(x.to(dev), **kw))
                        torch.cuda.synchronize()
                        tag = "%s%s" % (prec, "_graph" if graphs else "")
                        r = {"logits_rel": rel_err(outs[0], ref[0]),
                             "miou": orc.miou_between(ref[0], outs[0].cpu()),
                             "argmax_agree": float((outs[0].cpu().max(1)[1] == ref[0].max(1)[1]).float().mean())}
                        for i in range(1, len(ref)):
                            if torch.is_tensor(ref[i]):
                                if ref[i].dtype in (torch.int64, torch.int32):
                                    r["out%d_equal" % i] = bool((outs[i].cpu() == ref[i]).all())
                                else:
                                    r["out%d_abs" % i] = float((outs[i].cpu().double() - ref[i].double()).abs().max())
                            else:
                                r["out%d_delta" % i] = abs(float(outs[i]) - float(ref[i]))
                        rec[tag] = r
                rec["prob"] = ref[1].flatten()[:12].tolist() if len(ref) > 1 and torch.is_tensor(ref[1]) else None
            except Exception as e:  # keep going: this is a report
                import traceback
                rec["error"] = "%s: %s" % (type(e).__name__, e)
                rec["trace"] = traceback.format_exc()[-1500:]
            f.write(json.dumps(rec) + "\n")
            f.flush()
            print(json.dumps(rec)[:1500], flush=True)


if __name__ == "__main__":
    main()
