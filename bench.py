#!/usr/bin/env python
"""Headline benchmark: agent-frames/sec of the When2com forward (mrms-when2com, n_segnet 3x3-conv pair, 512x512).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one forward of the hot path over one synthetic batch of scenes. Prints ONE JSON line (rank 0).
  * N = 1: BASELINE.json configs[1] — MIMOcom, 5 agents, 512x512, bf16, one B200; 8 scenes (40 agent-frames)/step.
  * N > 1 (torchrun, one rank per GPU): configs[2]/[4] — 8 agents sharded over the ranks (8/N agents per rank),
    5*N scenes per step, so every GPU still processes 40 agent-frames per step ("weak" scaling); one NCCL
    all-gather of the packed key/query/feature slots per step.
`value` is device-timed (CUDA events, max over ranks) with inputs resident in HBM; `e2e` is the same metric through
the public API with pinned HOST inputs (H2D inside the timed region) and the predicted label map read back to the
host each step, as Trainer_MIMOcom.evaluate does (trainer.py:783-809).
--impl reference times the CPU restatement of the reference (oracle/, fp32, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "agent-frames/sec (512x512, 5 agents) at 1/2/4/8 B200; mIoU vs reference"
UNIT = "agent-frames/s"
IMG = int(os.environ.get("W2C_BENCH_IMG", "512"))  # (the contract tests shrink it; the benchmark is 512x512)
FRAMES_PER_GPU = 40
# algorithmic work per agent-frame, n_segnet MIMOcom softmax/train path (SURVEY.md 8d / BASELINE.md section 2)
GFLOP_PER_FRAME = 283.74


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3", "fp16"])
    ap.add_argument("--backbones", default="n_segnet", choices=["n_segnet", "resnet"])
    ap.add_argument("--inference", default="softmax", choices=["softmax", "activated", "argmax_test"])
    ap.add_argument("--frames-per-gpu", type=int, default=FRAMES_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fused-e2e", action="store_true")
    ap.add_argument("--no-parity-value", action="store_true")
    ap.add_argument("--layer-table", default="", help="write a per-layer timing table (markdown) to this path")
    ap.add_argument("--profile-step", action="store_true",
                    help="run one eager step inside cudaProfilerStart/Stop (for ncu) and exit without a bench line")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def traffic_per_launch(n_conv):
    """dram__bytes_read.sum + dram__bytes_write.sum per conv launch (bytes), from the committed ncu pass of this
    command (profiles/r1_traffic.json, written by tools/summarize_ncu_launches.py); None if absent or stale."""
    path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        if t.get("conv_launches") == n_conv:
            return t["conv_dram_bytes_per_step"] / n_conv
    except (OSError, ValueError, KeyError):
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_rate(backbones, inference, steps, warmup, n_agents=5):
    """agent-frames/s of the CPU restatement of the reference (oracle port), fp32, all host threads; each step is
    one scene (n_agents agent-frames) of the 512x512 workload."""
    import torch
    from multiagentperception_b200 import configs, synth
    from multiagentperception_b200.models import get_model
    from oracle import when2com_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = configs.make_config("MIMOcom", agent_num=n_agents, img_size=IMG, backbones=backbones)
    model = get_model(cfg, configs.N_CLASSES)
    synth.randomize_(model, 1337)
    sd = {k: v for k, v in model.state_dict().items()}
    x = synth.synthetic_views(1, n_agents, IMG, IMG, seed=1337)
    kw = dict(training=False, MO_flag=True, inference=inference)
    for _ in range(warmup):
        orc.forward(sd, cfg, x, **kw)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = orc.forward(sd, cfg, x, **kw)
    dt = time.perf_counter() - t0
    return {"oracle_out": out, "oracle_in": x, "value": n_agents * steps / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d step(s) x 1 scene x %d agents @%dx%d, fp32 torch-CPU restatement of the reference forward "
                      "(oracle/when2com_oracle.py), %.1f s" % (steps, n_agents, IMG, IMG, dt),
            "ms_per_step": dt / steps * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = cpu_reference_rate(args.backbones, args.inference, args.steps, max(1, min(args.warmup, 2)))
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": workload_config(args, 1, 5, 1),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit_line(line)
    return 0


def workload_config(args, world, n_agents, scenes):
    return {"workload": "mrms-when2com MIMOcom forward (inference='%s'), %s encoder/decoder pair, %d agents x %d "
                        "scenes per step @%dx%d, synthetic loader-distributed views, seeded He-init weights"
                        % (args.inference, args.backbones, n_agents, scenes, IMG, IMG),
            "agents": n_agents, "scenes_per_step": scenes, "image": IMG, "backbones": args.backbones,
            "inference": args.inference, "precision": args.precision,
            "parallelism": "single GPU" if world == 1 else "agents sharded %d/rank over %d ranks, 1 NCCL all-gather/step"
                                                           % (n_agents // world, world),
            "l2_policy": "per-step activations (>1 GB) exceed the 126 MB L2; no explicit flush",
            "cuda_graph": True}


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from multiagentperception_b200 import configs, ops, synth
    from multiagentperception_b200.models import get_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if world == 1:
        n_agents, local_agents = 5, 5
    else:
        n_agents, local_agents = 8, 8 // world
        if 8 % world:
            raise SystemExit("agent sharding needs a world size dividing 8")
    scenes = max(1, args.frames_per_gpu // local_agents)
    frames_total = scenes * n_agents

    cfg = configs.make_config("MIMOcom", agent_num=n_agents, img_size=IMG, backbones=args.backbones,
                              precision=args.precision)
    model = get_model(cfg, configs.N_CLASSES)
    synth.randomize_(model, 1337)
    model = model.to(dev).eval()
    if world > 1:
        model.shard_agents()
    kw = dict(training=False, MO_flag=True, inference=args.inference)

    views = synth.synthetic_views(scenes, n_agents, IMG, IMG, seed=1337)
    views = views[:, 3 * rank * local_agents: 3 * (rank + 1) * local_agents].contiguous()
    x_host = views.pin_memory()
    x_dev = x_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, drain=None):
        """device time of `steps` calls of fn: CUDA events, barrier + sync both sides, max over ranks (ms).
        drain(): makes the timing stream wait for side streams, so their tail is inside the timed region."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if drain is not None:
            drain()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    if args.profile_step:
        # one eager (no CUDA graph) step between cudaProfilerStart/Stop for `ncu --profile-from-start off`;
        # numbers under a profiler are never bench values, so nothing is printed
        model.set_cuda_graphs(False).set_clone_outputs(False)
        for _ in range(2):
            model(x_dev, **kw)
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
        model(x_dev, **kw)
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        return 0

    # ---- value: inputs resident in HBM
    model.set_clone_outputs(False)
    step_dev = lambda: model(x_dev, **kw)
    for _ in range(max(args.warmup, 3)):
        step_dev()
    launches_before = ops.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_dev, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches_per_step = max(model.kernel_launches_per_forward().values())
    ms_step = ms_total / args.steps
    value = frames_total / (ms_step * 1e-3)

    # ---- e2e: pinned host input -> H2D -> forward -> argmax labels -> D2H (trainer.py:783-809)
    # Every step copies ITS input from pinned host memory and ITS label map back; the copies run on two side streams,
    # double-buffered, so step i+1's H2D and step i-1's D2H overlap step i's kernels (what a serving loop does).
    main = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    x_buf = [torch.empty_like(x_dev) for _ in range(2)]
    lab_dev = [torch.empty((local_agents * scenes, IMG, IMG), dtype=torch.int64, device=dev) for _ in range(2)]
    labels_host = [torch.empty((local_agents * scenes, IMG, IMG), dtype=torch.int64).pin_memory() for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]      # H2D of buffer b finished
    ev_used = [torch.cuda.Event() for _ in range(2)]    # forward has consumed input buffer b
    ev_lab = [torch.cuda.Event() for _ in range(2)]     # labels of buffer b computed
    ev_out = [torch.cuda.Event() for _ in range(2)]     # D2H of buffer b finished
    state = {"i": 0}

    def step_e2e():
        b = state["i"] & 1
        state["i"] += 1
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_used[b])
            x_buf[b].copy_(x_host, non_blocking=True)
            ev_in[b].record(s_in)
        main.wait_event(ev_in[b])
        pred = model(x_buf[b], **kw)[0]
        ev_used[b].record(main)
        main.wait_event(ev_out[b])                     # the previous D2H from this label buffer is done
        lab_dev[b].copy_(pred.max(1)[1])               # outputs.max(1)[1], trainer.py:804
        ev_lab[b].record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_lab[b])
            labels_host[b].copy_(lab_dev[b], non_blocking=True)
            ev_out[b].record(s_out)

    def drain():
        main.wait_stream(s_in)
        main.wait_stream(s_out)

    for _ in range(4):
        step_e2e()
    drain()
    ms_e2e = timed(step_e2e, args.steps, drain) / args.steps
    e2e = {"value": frames_total / (ms_e2e * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": int(x_host.numel() * 4 * world),
           "d2h_bytes_per_step": int(labels_host[0].numel() * 8 * world), "ms_per_step": ms_e2e,
           "pipeline": "double-buffered H2D / D2H on side streams, every step's copies inside the timed region"}

    # ---- e2e through the fused evaluation path of the public API: the loader's RAW uint8 frames in
    # (transform + cat fused into the first conv), the uint8 label map out (arg-max fused into the logits layer, the
    # fp32 logits never written); same double-buffered pipeline. Bit-identical labels to the drop-in path above.
    e2e_fused = None
    if not args.no_fused_e2e:
        fmodel = get_model(cfg, configs.N_CLASSES)
        fmodel.load_state_dict(model.state_dict())
        fmodel = fmodel.to(dev).eval().set_clone_outputs(False)
        if world > 1:
            fmodel.shard_agents()
        fmodel.set_input_format("u8_hwc").set_label_output(True, logits=False)
        frames_host = synth.synthetic_frames(scenes, local_agents, IMG, IMG, seed=1337 + rank).pin_memory()
        f_buf = [torch.empty(frames_host.shape, dtype=torch.uint8, device=dev) for _ in range(2)]
        l_dev = [torch.empty((local_agents * scenes, IMG, IMG), dtype=torch.uint8, device=dev) for _ in range(2)]
        l_host = [torch.empty((local_agents * scenes, IMG, IMG), dtype=torch.uint8).pin_memory() for _ in range(2)]
        fstate = {"i": 0}

        def step_fused():
            b = fstate["i"] & 1
            fstate["i"] += 1
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_used[b])
                f_buf[b].copy_(frames_host, non_blocking=True)
                ev_in[b].record(s_in)
            main.wait_event(ev_in[b])
            labels = fmodel(f_buf[b], **kw)[0]
            ev_used[b].record(main)
            main.wait_event(ev_out[b])
            l_dev[b].copy_(labels)
            ev_lab[b].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_lab[b])
                l_host[b].copy_(l_dev[b], non_blocking=True)
                ev_out[b].record(s_out)

        for _ in range(4):
            step_fused()
        drain()
        ms_f = timed(step_fused, args.steps, drain) / args.steps
        e2e_fused = {"value": frames_total / (ms_f * 1e-3), "unit": UNIT, "ms_per_step": ms_f,
                     "h2d_bytes_per_step": int(frames_host.numel() * world),
                     "d2h_bytes_per_step": int(l_host[0].numel() * world),
                     "api": "model.set_input_format('u8_hwc').set_label_output(True, logits=False); raw uint8 RGB "
                            "frames in, uint8 label map out (what Trainer_MIMOcom.evaluate consumes)"}
        del fmodel, f_buf, l_dev

    # ---- the same step in the other two precisions: bf16x3 (hi/lo split operands, <= 1e-3 of the logit range vs fp32)
    # and fp16 (IEEE-half storage: the speed of bf16, ~5x closer to the fp32 reference)
    parity_precision = None
    fp16_precision = None
    if not args.no_parity_value and args.precision == "bf16":
        for prec in ("bf16x3", "fp16"):
            model.set_precision(prec)
            for _ in range(3):
                step_dev()
            k = max(3, args.steps // 2)
            ms_p = timed(step_dev, k) / k
            rec = {"precision": prec, "value": frames_total / (ms_p * 1e-3), "unit": UNIT, "ms_per_step": ms_p}
            if prec == "bf16x3":
                rec["logit_tolerance"] = "1e-3 of max|logit| vs the fp32 oracle (tests/test_parity_gpu.py)"
                parity_precision = rec
            else:
                rec["logit_tolerance"] = "8e-3 of max|logit| vs the fp32 oracle (tests/test_parity_gpu.py)"
                fp16_precision = rec
        model.set_precision("bf16")

    # ---- roofline of the dominant kernel (the tcgen05 conv): replay ONLY its launches, same buffers, CUDA events
    prog = max(model._w2c["programs"].values(), key=lambda c: c.prog.n_launches).prog
    conv_prog = prog.conv_only_program()
    n_conv = sum(1 for c in conv_prog.calls if c[1] is not None)
    for _ in range(3):
        conv_prog.run(True)
    ms_conv = timed(lambda: conv_prog.run(True), args.steps) / args.steps
    peaks = measured_peaks()
    # algorithmic FLOPs of the tensor-core convs per local frame = all conv layers except the two 3->64 stems
    stem_gflop = 2 * 0.906 if args.backbones == "n_segnet" else 0.0
    gflop_frame = GFLOP_PER_FRAME if args.backbones == "n_segnet" else 42.92
    conv_tflop_step = (gflop_frame - stem_gflop - 0.005) * local_agents * scenes / 1e3
    achieved = conv_tflop_step / (ms_conv * 1e-3)
    roofline = {"bound": "tensor",
                "kernel": "conv_persv1_kernel / conv_tc_kernel (tcgen05 implicit-GEMM conv + transposed conv, every "
                          "instantiation of the step: 43 launches for the n_segnet pair)",
                "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": traffic_per_launch(n_conv),
                "launches_per_step": n_conv, "avg_launch_ms": ms_conv / n_conv, "share_of_step": ms_conv / ms_step,
                "peak_source": "%s (sustained cuBLAS bf16, kernel timed inside a long step)" % peaks["source"],
                "algorithmic_tflop_per_step_per_gpu": conv_tflop_step}

    if args.layer_table and rank == 0:
        write_layer_table(args.layer_table, prog, dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu_baseline = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_rate(args.backbones, args.inference, steps=2, warmup=1)
        cpu_baseline = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        # the metric's second half, "mIoU vs reference": the oracle's fp32 output of that one scene is the checker for
        # the CUDA path on the same views, in both precisions (arg-max labels of the oracle as ground truth)
        from oracle import when2com_oracle as orc
        ref_pred = r["oracle_out"][0]
        parity = {"sample": "1 scene x 5 agents @%dx%d, same weights and views as cpu_baseline" % (IMG, IMG)}
        for prec in ("bf16", "fp16", "bf16x3"):
            model.set_precision(prec)
            pred = model(r["oracle_in"].to(dev), **kw)[0].float().cpu()
            parity[prec] = {"max_logit_err_over_max_logit": float((pred - ref_pred).abs().max() / ref_pred.abs().max()),
                            "miou_vs_reference_argmax": orc.miou_between(ref_pred, pred)}
        model.set_precision(args.precision)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"bf16": "bf16", "fp16": "fp16"}.get(args.precision, "bf16x3 (hi/lo split, fp32-grade)"),
            "data": "synthetic", "config": workload_config(args, world, n_agents, scenes),
            "e2e": e2e, "e2e_fused": e2e_fused, "parity_precision": parity_precision, "fp16_precision": fp16_precision,
            "gpu_launches": int(launches_per_step * args.steps), "launches_per_step": int(launches_per_step),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity,
            "frames_per_step": frames_total,
            "model_tflops": GFLOP_PER_FRAME * value / 1e3 if args.backbones == "n_segnet" else None}
    emit_line(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def write_layer_table(path, prog, dev):
    """Per-launch device time of every call of the step program (eager, CUDA events around each call; the tensor-core
    convs with their geometry and TFLOP/s, the other kernels by entry point)."""
    import ctypes
    import torch
    lib_conv = prog._lib.w2c_conv_bnrelu_fwd
    rows = []
    for fn, args, _sid in prog.calls:
        if fn is None or isinstance(fn, str):
            continue  # host op (collective) / fork-join marker
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            e0.record()
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if args is None:
                fn(stream)
            else:
                fn(*args, stream)
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        if fn is lib_conv:
            a = args[0]._obj
            taps = 1 if a.kind in (3, 4) else 9   # (kind 5, the dense transposed conv: useful MACs = the 9-tap count)
            if a.kind in (1, 4):
                m = a.n * (a.h_in // 2) * (a.w_in // 2)
            else:
                m = a.n * a.h_in * a.w_in
            flop = 2.0 * m * a.cin * a.cout * taps
            names = {0: "conv3x3 s1", 1: "conv3x3 s2", 2: "deconv3x3 s2", 3: "conv1x1", 4: "conv1x1 s2",
                     5: "deconv3x3 s2 (dense)"}
            rows.append((names[a.kind], a.n, a.h_in, a.w_in, a.cin, a.cout, best, "%.1f" % (flop / (best * 1e-3) / 1e12)))
        else:
            rows.append((getattr(fn, "__name__", "host-side torch op"), "", "", "", "", "", best, ""))
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        f.write("| # | launch | n | h_in | w_in | cin | cout | ms (best of 5, eager, CUDA events) | TFLOP/s |\n")
        f.write("|---|---|---|---|---|---|---|---|---|\n")
        for i, r in enumerate(rows):
            f.write("| %d | %s | %s | %s | %s | %s | %s | %.4f | %s |\n" % ((i,) + r[:6] + (r[6], r[7])))
        f.write("\ntotal %.3f ms\n" % sum(r[6] for r in rows))


class _QuietStdout:
    """Everything libraries print to fd 1 while the bench runs (e.g. NCCL's version banner) goes to stderr, so that
    stdout carries exactly ONE line: the JSON result."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.saved, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


OUT = None


def emit_line(obj):
    text = json.dumps(obj)
    if OUT is not None:
        OUT.emit(text)
    else:
        print(text, flush=True)


def main():
    global OUT
    args = parse()
    with _QuietStdout() as q:
        OUT = q
        try:
            if args.impl == "reference":
                return run_reference(args)
            return run_b200(args)
        finally:
            OUT = None


if __name__ == "__main__":
    sys.exit(main())
