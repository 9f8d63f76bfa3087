#!/usr/bin/env python
"""Training-step timing on a B200: forward + cross_entropy2d + backward (+ SGD step) of MIMOcom as
Trainer_MIMOcom.train() runs it (trainer.py:659-670), through this repo's CUDA path and - the "library Blackwell path"
of SURVEY 2.1 for training - through the UNMODIFIED reference modules on the same GPU with stock torch / cuDNN autograd.
Used by bench.py (key `train_step`); standalone:  python tools/gpu_train_bench.py [--scenes 2] [--backbones n_segnet]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _loss(pred, labels):
    import torch.nn.functional as F
    n, c, h, w = pred.shape
    return F.cross_entropy(pred.permute(0, 2, 3, 1).reshape(-1, c), labels.reshape(-1), ignore_index=250)


def _timed(fn, dev, steps, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps


def train_step_rates(dev, scenes=2, agents=5, img=512, backbones="n_segnet", steps=5, precisions=("bf16", "bf16x3"),
                     library=True, split=True):
    import torch
    from multiagentperception_b200 import configs, ops, synth
    from multiagentperception_b200.models import get_model
    cfg = configs.make_config("MIMOcom", agent_num=agents, img_size=img, backbones=backbones)
    kw = dict(training=True, MO_flag=True)
    frames = scenes * agents
    x = synth.synthetic_views(scenes, agents, img, img, seed=1337).to(dev)
    labels = torch.randint(0, 11, (frames, img, img), generator=torch.Generator().manual_seed(3)).to(dev)
    out = {"what": "one training step = model.train() forward (batch-statistics BatchNorm) + cross_entropy2d + "
                   "loss.backward() + SGD step, MIMOcom %s pair, %d agents x %d scenes @%dx%d (the shipped YAMLs train "
                   "with batch_size 2)" % (backbones, agents, scenes, img, img),
           "unit": "agent-frames/s", "frames_per_step": frames}
    for prec in precisions:
        model = get_model(cfg, 11)
        synth.randomize_(model, 1337)
        model = model.to(dev).set_precision(prec)
        model.train()
        opt = torch.optim.SGD(model.parameters(), lr=1e-5)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = _loss(model(x, **kw)[0], labels)
            loss.backward()
            opt.step()
            return loss
        try:
            before = ops.launch_count()
            ms = _timed(step, dev, steps)
            rec = {"value": frames / (ms * 1e-3), "ms_per_step": ms,
                   "kernel_launches_per_step": (ops.launch_count() - before) // (steps + 3)}
            if split:
                from multiagentperception_b200.loss import cross_entropy2d

                def step_fused_loss():
                    opt.zero_grad(set_to_none=True)
                    loss = cross_entropy2d(input=model(x, **kw)[0], target=labels)
                    loss.backward()
                    opt.step()
                ms_f = _timed(step_fused_loss, dev, steps)
                rec["with_device_loss"] = {"value": frames / (ms_f * 1e-3), "ms_per_step": ms_f,
                                           "loss": "multiagentperception_b200.loss.cross_entropy2d (one pass)"}

                def fwd_only():
                    with torch.no_grad():
                        model(x, **kw)
                rec["forward_only_ms"] = _timed(fwd_only, dev, steps)
            out[prec] = rec
        except Exception as e:  # noqa: BLE001
            out[prec] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        del model, opt
        torch.cuda.empty_cache()
    if library:
        from oracle import ref_harness
        if not ref_harness.available():
            out["library"] = {"unavailable": "reference package not found"}
            return out
        import contextlib
        import io
        lib = {"what": "the UNMODIFIED reference MIMOcom on cuda: stock torch %s autograd / cuDNN %s, cudnn.benchmark=True"
                       % (torch.__version__, torch.backends.cudnn.version())}
        ref = ref_harness.build_reference_model(cfg, 11)
        synth.randomize_(ref, 1337)
        ref = ref.to(dev).train()
        opt = torch.optim.SGD(ref.parameters(), lr=1e-5)
        saved = (torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cudnn.benchmark = True

        def ref_step(autocast):
            opt.zero_grad(set_to_none=True)
            ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
            with ctx, contextlib.redirect_stdout(io.StringIO()):
                pred = ref(x, **kw)[0]
            loss = _loss(pred.float(), labels)
            loss.backward()
            opt.step()
        try:
            for name, tf32, ac in (("fp32", False, False), ("tf32", True, False), ("bf16_autocast", True, True)):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.allow_tf32 = tf32
                try:
                    ms = _timed(lambda: ref_step(ac), dev, max(2, steps // 2), warmup=2)
                    lib[name] = {"value": frames / (ms * 1e-3), "ms_per_step": ms}
                except Exception as e:  # noqa: BLE001
                    lib[name] = {"error": str(e)[:200]}
        finally:
            torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
            del ref, opt
            torch.cuda.empty_cache()
        out["library"] = lib
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=2)
    ap.add_argument("--agents", type=int, default=5)
    ap.add_argument("--img", type=int, default=512)
    ap.add_argument("--backbones", default="n_segnet")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-library", action="store_true")
    ap.add_argument("--precisions", default="bf16,bf16x3")
    a = ap.parse_args()
    import torch
    r = train_step_rates(torch.device("cuda:0"), a.scenes, a.agents, a.img, a.backbones, a.steps,
                         tuple(a.precisions.split(",")), library=not a.no_library)
    print(json.dumps(r))
