#!/usr/bin/env python
"""Headline benchmark: agent-frames/sec of the When2com forward (mrms-when2com, n_segnet 3x3-conv pair, 512x512).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 2|3|4|5] [--batch B]

One "step" = one forward of the hot path over one synthetic batch of scenes. Prints ONE JSON line (rank 0).
Workloads (BASELINE.json `configs`, 0-based index in brackets):
  --config 2 (default)  [1] MIMOcom, 5 agents, 512x512, bf16, one B200: 8 scenes (40 agent-frames) per step.
                        N > 1 (torchrun, one rank per GPU): 8 agents sharded over the ranks (8/N agents per rank),
                        5*N scenes per step, so every GPU still processes 40 agent-frames per step ("weak" scaling);
                        one NCCL all-gather of the packed key/query/feature slots per step.
  --config 3            [2] MIMOcom, 8 agents, ONE scene per step (the latency point), agents sharded over N GPUs.
  --config 4            [3] Single_agent (srms-allnorm), 5 views folded into the batch, 1024x1024, fp16; N > 1 runs
                        independent replicas (no exchange step on this path).
  --config 5            [4] MIMOcom, 8 agents sharded, --batch scenes per step (default 32; sweep B = 1..32).
`value` is device-timed (CUDA events, max over ranks) with inputs resident in HBM. `e2e` is the same metric through
the public evaluation API with pinned HOST buffers: the loader's raw uint8 frames host->device and the predicted
label map device->host inside the timed region every step (what Trainer_MIMOcom.evaluate moves, trainer.py:783-809);
`e2e_dropin` is the same through the reference's own forward() types (fp32 views in, int64 labels out).
--impl reference times the UNMODIFIED reference forward on the host CPU (oracle/_ref staged copy or /root/reference;
the oracle port only if neither exists), fp32, all host threads.
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
# algorithmic work per agent-frame (SURVEY.md 8d / BASELINE.md section 2), GFLOP; the 3-channel stems are not
# tensor-core conv launches of the roofline kernel and are subtracted there
GFLOP = {"mimocom_n_segnet": 283.74, "mimocom_resnet": 42.92, "single_n_segnet": 181.80, "single_resnet": 20.78}
STEM_GFLOP = 0.906  # one 3 -> 64 first layer at 512x512


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5])
    ap.add_argument("--batch", type=int, default=0, help="scenes per step (per rank for --config 4); 0 = the config's own")
    ap.add_argument("--precision", default="", choices=["", "bf16", "bf16x3", "fp16", "fp16x3", "mixed"])
    ap.add_argument("--backbones", default="n_segnet", choices=["n_segnet", "resnet"])
    ap.add_argument("--inference", default="softmax", choices=["softmax", "activated", "argmax_test"])
    ap.add_argument("--frames-per-gpu", type=int, default=FRAMES_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--no-dropin-e2e", action="store_true")
    ap.add_argument("--no-parity-value", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true")
    ap.add_argument("--lean", action="store_true", help="value + e2e + roofline only (sweeps)")
    ap.add_argument("--layer-table", default="", help="write a per-layer timing table (markdown) to this path")
    ap.add_argument("--profile-step", action="store_true",
                    help="run one eager step inside cudaProfilerStart/Stop (for ncu) and exit without a bench line")
    a = ap.parse_args()
    if a.lean:
        a.no_cpu_baseline = a.no_library_baseline = a.no_dropin_e2e = a.no_parity_value = a.no_extra_configs = True
    return a


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
    command (profiles/r2_traffic.json, written by tools/summarize_ncu_launches.py). The file records the kernel-source
    fingerprint and the conv launch count it was measured on; if either differs from what runs now the number is
    stale and None is reported instead."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        from multiagentperception_b200 import build as _b
        with open(path) as f:
            t = json.load(f)
        if t.get("conv_launches") == n_conv and t.get("csrc_fingerprint") == _b._fingerprint(())[:16]:
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


# ------------------------------------------------------------------------------------------------ workloads
class Workload:
    """What one step is, for a (config, world size): model, inputs per rank, frames per step."""

    def __init__(self, args, world, rank=0):
        from multiagentperception_b200 import configs
        self.cfgno, self.world, self.rank = args.config, world, rank
        self.backbones, self.inference = args.backbones, args.inference
        self.img = IMG
        self.shard = False
        if args.config == 4:
            self.arch, self.img = "Single_agent", IMG * 2 if IMG == 512 else IMG
            self.n_agents = self.local_agents = 1
            views = args.batch or 5                       # 5 views folded into the batch, trainer.py:1390-1392
            self.scenes, self.frames_local = views, views
            self.frames_total = views * world             # replicas only: every rank runs its own 5 views
            self.precision = args.precision or "fp16"
            self.kw = {}
            self.gflop_frame = GFLOP["single_" + args.backbones] * (self.img / 512.0) ** 2
            self.stem_gflop = (STEM_GFLOP if args.backbones == "n_segnet" else 0.0) * (self.img / 512.0) ** 2
            self.parallelism = "single GPU" if world == 1 else "%d independent replicas (no exchange step)" % world
        else:
            self.arch = "MIMOcom"
            if world == 1 and args.config == 2:
                self.n_agents = self.local_agents = 5
            else:
                self.n_agents = 8
                if 8 % world:
                    raise SystemExit("agent sharding needs a world size dividing 8")
                self.local_agents = 8 // world
                self.shard = world > 1
            if args.config == 3:
                self.scenes = args.batch or 1
            elif args.config == 5:
                self.scenes = args.batch or 32
            else:
                self.scenes = args.batch or max(1, args.frames_per_gpu // self.local_agents)
            self.frames_local = self.scenes * self.local_agents
            self.frames_total = self.scenes * self.n_agents
            self.precision = args.precision or "bf16"
            self.kw = dict(training=False, MO_flag=True, inference=args.inference)
            self.gflop_frame = GFLOP["mimocom_" + args.backbones] * (self.img / 512.0) ** 2
            self.stem_gflop = (2 * STEM_GFLOP if args.backbones == "n_segnet" else 0.0) * (self.img / 512.0) ** 2
            self.parallelism = ("single GPU" if world == 1 else
                                "agents sharded %d/rank over %d ranks, 1 NCCL all-gather/step" % (self.local_agents, world))
        self.cfg = configs.make_config(self.arch, agent_num=self.n_agents, img_size=self.img, backbones=args.backbones)

    def build_model(self, dev, precision=None):
        from multiagentperception_b200 import configs, synth
        from multiagentperception_b200.models import get_model
        model = get_model(self.cfg, configs.N_CLASSES)
        synth.randomize_(model, 1337)
        model = model.to(dev).eval().set_precision(precision or self.precision)
        if self.shard:
            model.shard_agents()
        return model

    def views(self, scenes=None, seed=1337):
        """This rank's fp32 views (B, 3*local_agents, H, W)."""
        from multiagentperception_b200 import synth
        scenes = scenes or self.scenes
        if self.arch == "Single_agent":
            return synth.synthetic_views(scenes, 1, self.img, self.img, seed=seed + self.rank)
        v = synth.synthetic_views(scenes, self.n_agents, self.img, self.img, seed=seed)
        return v[:, 3 * self.rank * self.local_agents: 3 * (self.rank + 1) * self.local_agents].contiguous()

    def frames(self, seed=1337):
        """This rank's raw uint8 frames (B, local_agents, H, W, 3)."""
        from multiagentperception_b200 import synth
        return synth.synthetic_frames(self.scenes, self.local_agents, self.img, self.img, seed=seed + self.rank)

    def config_dict(self, impl="b200"):
        d = {"workload": "BASELINE.json configs[%d]: %s forward%s, %s encoder/decoder pair, %d agents x %d scenes per step "
                         "@%dx%d, synthetic loader-distributed views, seeded He-init weights"
                         % (self.cfgno - 1, self.arch, " (inference='%s')" % self.inference if self.kw else "",
                            self.backbones, self.n_agents, self.scenes, self.img, self.img),
             "baseline_config": self.cfgno, "arch": self.arch, "agents": self.n_agents,
             "scenes_per_step": self.scenes, "image": self.img, "backbones": self.backbones,
             "inference": self.inference if self.kw else None}
        if impl == "reference":
            d.update(precision="fp32", parallelism="host CPU, all threads", cuda_graph=False,
                     l2_policy="n/a (CPU run)")
        else:
            d.update(precision=self.precision, parallelism=self.parallelism, cuda_graph=True,
                     l2_policy="per-step activations (>1 GB) exceed the 126 MB L2; no explicit flush"
                     if self.frames_local * self.img * self.img >= 5 * 512 * 512 else
                     "small step: a 256 MB buffer is rewritten between timed steps to flush the 126 MB L2")
        return d


# ------------------------------------------------------------------------------------------------ reference arm
def _reference_forward_fn(wl, scenes, device="cpu"):
    """(callable running one reference step, kind, description). The UNMODIFIED reference when it can be found
    (oracle/ref_harness.py), else the oracle port."""
    import torch
    from multiagentperception_b200 import configs, synth
    from multiagentperception_b200.models import get_model
    from oracle import ref_harness
    if wl.arch == "Single_agent":
        x = synth.synthetic_views(scenes, 1, wl.img, wl.img, seed=1337)
    else:
        x = synth.synthetic_views(scenes, wl.n_agents, wl.img, wl.img, seed=1337)
    if ref_harness.available():
        ref = ref_harness.build_reference_model(wl.cfg, configs.N_CLASSES)
        synth.randomize_(ref, 1337)
        ref = ref.to(device).eval()
        xd = x.to(device)
        kind = "reference"
        desc = "UNMODIFIED reference ptsemseg.models (%s copy)" % ref_harness.source_kind()
        return (lambda: ref_harness.reference_forward(ref, xd, **wl.kw)), kind, desc, x
    from oracle import when2com_oracle as orc
    model = get_model(wl.cfg, configs.N_CLASSES)
    synth.randomize_(model, 1337)
    sd = dict(model.state_dict())
    return (lambda: orc.forward(sd, wl.cfg, x, **wl.kw)), "port", "oracle/when2com_oracle.py restatement", x


def cpu_reference_rate(wl, scenes, steps, warmup):
    """agent-frames/s of the reference forward on the host CPU, fp32, all host threads; `scenes` scenes per step."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fn, kind, desc, x = _reference_forward_fn(wl, scenes)
    for _ in range(warmup):
        out = fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = fn()
    dt = time.perf_counter() - t0
    frames = scenes * wl.n_agents
    return {"out": out, "in": x, "value": frames * steps / dt, "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": kind, "ms_per_step": dt / steps * 1e3,
            "sample": "%d timed step(s) (+%d warm-up) x %d scene(s) x %d agents @%dx%d, fp32, torch CPU, %s; %.1f s"
                      % (steps, warmup, scenes, wl.n_agents, wl.img, wl.img, desc, dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = Workload(args, max(1, args.gpus))
    # bounded sample: the same scenes per step as the B200 arm at N = 1; at N > 1 (where the step grows with N) the
    # CPU step stays at <= 40 agent-frames so that K + W steps still end within minutes, and says so
    scenes = wl.scenes if wl.frames_total <= 40 else max(1, 40 // wl.n_agents)
    r = cpu_reference_rate(wl, scenes, args.steps, args.warmup)
    cfg = wl.config_dict("reference")
    cfg["scenes_per_step"] = scenes
    if scenes != wl.scenes:
        cfg["bounded_sample"] = "%d of the B200 arm's %d scenes per step" % (scenes, wl.scenes)
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference", "config": cfg,
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit_line(line)
    return 0


def library_baseline(wl, dev, steps=5):
    """The "library Blackwell path" (SURVEY 2.1 / 8d): the UNMODIFIED reference modules on this GPU through stock
    PyTorch (cuDNN / cuBLAS as shipped in torch), same weights, same step, timed with CUDA events. Three settings:
    fp32 (TF32 off), TF32, and bf16 autocast + channels_last. None of this repo's kernels run on that path."""
    import torch
    from oracle import ref_harness
    if not ref_harness.available():
        return {"unavailable": "reference package not found (neither /root/reference nor oracle/_ref)"}
    from multiagentperception_b200 import configs, synth
    out = {"what": "reference ptsemseg.models.%s (%s copy) on cuda through stock torch %s / cuDNN %s, "
                   "cudnn.benchmark=True (test.py:17), no_grad, %d agent-frames per step"
                   % (wl.arch, ref_harness.source_kind(), torch.__version__, torch.backends.cudnn.version(),
                      wl.frames_total), "unit": UNIT}
    ref = ref_harness.build_reference_model(wl.cfg, configs.N_CLASSES)
    synth.randomize_(ref, 1337)
    ref = ref.to(dev).eval()
    if wl.arch == "Single_agent":
        x = synth.synthetic_views(wl.scenes, 1, wl.img, wl.img, seed=1337).to(dev)
    else:
        x = synth.synthetic_views(wl.scenes, wl.n_agents, wl.img, wl.img, seed=1337).to(dev)
    saved = (torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cudnn.benchmark = True

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        return {"value": wl.frames_total / (ms * 1e-3), "ms_per_step": ms}

    try:
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            try:
                out[name] = timed(lambda: ref_harness.reference_forward(ref, x, **wl.kw))
            except Exception as e:  # e.g. out of memory: report, keep going
                out[name] = {"error": str(e)[:200]}
        def step_bf16(m, xin):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return ref_harness.reference_forward(m, xin, **wl.kw)
        try:
            out["bf16_autocast"] = timed(lambda: step_bf16(ref, x))
        except Exception as e:
            out["bf16_autocast"] = {"error": str(e)[:200]}
        try:
            # channels_last weights and input: the reference's own `.view()` calls (agent.py:158,1114) reject
            # channels_last activations for some archs - reported as an error then, the plain autocast line stands
            ref_cl = ref.to(memory_format=torch.channels_last)
            x_cl = x.contiguous(memory_format=torch.channels_last)
            out["bf16_autocast_channels_last"] = timed(lambda: step_bf16(ref_cl, x_cl))
        except Exception as e:
            out["bf16_autocast_channels_last"] = {"error": str(e)[:200]}
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
        del ref
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ B200 arm
class Timer:
    def __init__(self, dev, world, flush_l2):
        import torch
        self.torch, self.dev, self.world = torch, dev, world
        # small steps (B = 1 latency points) fit the 126 MB L2: rewrite a 256 MB buffer between timed steps and time
        # every step with its own event pair, so the flush is outside the timed region
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush_l2 else None

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def __call__(self, fn, steps, drain=None):
        """device time of `steps` calls of fn (ms, total): CUDA events, barrier + sync both sides, max over ranks.
        drain(): makes the timing stream wait for side streams, so their tail is inside the timed region."""
        torch = self.torch
        self.barrier()
        if self.flush is None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            if drain is not None:
                drain()
            e1.record()
            self.barrier()
            ms = e0.elapsed_time(e1)
        else:
            evs = []
            for _ in range(steps):
                self.flush.add_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                if drain is not None:
                    drain()
                e1.record()
                evs.append((e0, e1))
            self.barrier()
            ms = sum(a.elapsed_time(b) for a, b in evs)
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms


def e2e_fused_leg(wl, model, dev, timed, steps):
    """End to end through the evaluation API: pinned uint8 frames -> H2D -> forward -> uint8 label map -> D2H, every
    step; copies on two side streams, double-buffered (what a serving loop does)."""
    import torch
    model.set_input_format("u8_hwc").set_label_output(True, logits=False)
    main = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    frames_host = wl.frames().pin_memory()
    n_img = wl.frames_local
    f_buf = [torch.empty(frames_host.shape, dtype=torch.uint8, device=dev) for _ in range(2)]
    l_dev = [torch.empty((n_img, wl.img, wl.img), dtype=torch.uint8, device=dev) for _ in range(2)]
    l_host = [torch.empty((n_img, wl.img, wl.img), dtype=torch.uint8).pin_memory() for _ in range(2)]
    ev = {k: [torch.cuda.Event() for _ in range(2)] for k in ("in", "used", "lab", "out")}
    state = {"i": 0}

    def step():
        b = state["i"] & 1
        state["i"] += 1
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev["used"][b])
            f_buf[b].copy_(frames_host, non_blocking=True)
            ev["in"][b].record(s_in)
        main.wait_event(ev["in"][b])
        out = model(f_buf[b], **wl.kw)
        labels = out[0] if isinstance(out, tuple) else out
        ev["used"][b].record(main)
        main.wait_event(ev["out"][b])
        l_dev[b].copy_(labels)
        ev["lab"][b].record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev["lab"][b])
            l_host[b].copy_(l_dev[b], non_blocking=True)
            ev["out"][b].record(s_out)

    def drain():
        main.wait_stream(s_in)
        main.wait_stream(s_out)

    for _ in range(4):
        step()
    drain()
    ms = timed(step, steps, drain) / steps
    model.set_input_format("f32_nchw").set_label_output(False)
    return {"value": wl.frames_total / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
            "h2d_bytes_per_step": int(frames_host.numel() * wl.world),
            "d2h_bytes_per_step": int(l_host[0].numel() * wl.world),
            "pipeline": "double-buffered H2D / D2H on side streams, every step's copies inside the timed region",
            "api": "model.set_input_format('u8_hwc').set_label_output(True, logits=False): the loader's raw uint8 RGB "
                   "frames in, the uint8 label map out (eval_loop.evaluate / Trainer_MIMOcom.evaluate's traffic)"}


def e2e_dropin_leg(wl, model, dev, timed, steps):
    """The same through the reference's own forward() types: pinned fp32 views in, int64 `max(1)[1]` labels out."""
    import torch
    main = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    x_host = wl.views().pin_memory()
    n_img = wl.frames_local
    x_buf = [torch.empty(x_host.shape, dtype=torch.float32, device=dev) for _ in range(2)]
    lab_dev = [torch.empty((n_img, wl.img, wl.img), dtype=torch.int64, device=dev) for _ in range(2)]
    lab_host = [torch.empty((n_img, wl.img, wl.img), dtype=torch.int64).pin_memory() for _ in range(2)]
    ev = {k: [torch.cuda.Event() for _ in range(2)] for k in ("in", "used", "lab", "out")}
    state = {"i": 0}

    def step():
        b = state["i"] & 1
        state["i"] += 1
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev["used"][b])
            x_buf[b].copy_(x_host, non_blocking=True)
            ev["in"][b].record(s_in)
        main.wait_event(ev["in"][b])
        out = model(x_buf[b], **wl.kw)
        pred = out[0] if isinstance(out, tuple) else out
        ev["used"][b].record(main)
        main.wait_event(ev["out"][b])
        lab_dev[b].copy_(pred.max(1)[1])               # outputs.max(1)[1], trainer.py:804
        ev["lab"][b].record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev["lab"][b])
            lab_host[b].copy_(lab_dev[b], non_blocking=True)
            ev["out"][b].record(s_out)

    def drain():
        main.wait_stream(s_in)
        main.wait_stream(s_out)

    for _ in range(4):
        step()
    drain()
    ms = timed(step, steps, drain) / steps
    return {"value": wl.frames_total / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
            "h2d_bytes_per_step": int(x_host.numel() * 4 * wl.world),
            "d2h_bytes_per_step": int(lab_host[0].numel() * 8 * wl.world),
            "api": "model(views_f32, ...) -> logits; outputs.max(1)[1] int64 labels copied back (trainer.py:783-809)"}


def roofline_leg(wl, model, timed, steps, ms_step):
    """Dominant kernel = the tcgen05 conv: a second program replaying ONLY its launches (same buffers, CUDA graph) is
    timed with CUDA events on the stream it is launched on."""
    prog = max(model._w2c["programs"].values(), key=lambda c: c.prog.n_launches).prog
    conv_prog = prog.conv_only_program()
    n_conv = sum(1 for c in conv_prog.calls if c[1] is not None)
    for _ in range(3):
        conv_prog.run(True)
    ms_conv = timed(lambda: conv_prog.run(True), steps) / steps
    peaks = measured_peaks()
    conv_tflop_step = (wl.gflop_frame - wl.stem_gflop - (0.005 if wl.arch == "MIMOcom" else 0.0)) * wl.frames_local / 1e3
    achieved = conv_tflop_step / (ms_conv * 1e-3)
    return prog, {"bound": "tensor",
                  "kernel": "conv_persv1_kernel / conv_tc_kernel / fused 64-channel end kernels (tcgen05 implicit-GEMM "
                            "conv + transposed conv: every tensor-core conv launch of the step)",
                  "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                  "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": traffic_per_launch(n_conv),
                  "launches_per_step": n_conv, "avg_launch_ms": ms_conv / n_conv, "share_of_step": ms_conv / ms_step,
                  "peak_source": "%s (sustained cuBLAS bf16, kernel timed inside a long step)" % peaks["source"],
                  "algorithmic_tflop_per_step_per_gpu": conv_tflop_step}


def quick_value(args, overrides, dev, steps):
    """value of another workload on this GPU (N = 1 extras of the default line): (record, model, workload)."""
    import copy
    a = copy.copy(args)
    for k, v in overrides.items():
        setattr(a, k, v)
    wl = Workload(a, 1)
    model = wl.build_model(dev).set_clone_outputs(False)
    x = wl.views().to(dev)
    small = wl.frames_local * wl.img * wl.img < 5 * 512 * 512
    timed = Timer(dev, 1, flush_l2=small)
    for _ in range(3):
        model(x, **wl.kw)
    ms = timed(lambda: model(x, **wl.kw), steps) / steps
    rec = {"workload": wl.config_dict()["workload"], "precision": wl.precision, "value": wl.frames_total / (ms * 1e-3),
           "unit": UNIT, "ms_per_step": ms, "frames_per_step": wl.frames_total,
           "launches_per_step": int(max(model.kernel_launches_per_forward().values())),
           "l2": "flushed between steps" if small else "step exceeds L2"}
    rec["model_tflops"] = wl.gflop_frame * rec["value"] / 1e3
    rec["frac_of_sustained_bf16_peak"] = rec["model_tflops"] / measured_peaks()["bf16_tflops_sustained"]
    return rec, model, wl


def run_b200(args):
    import torch
    import torch.distributed as dist
    from multiagentperception_b200 import engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = Workload(args, world, rank)
    model = wl.build_model(dev)
    kw = wl.kw
    x_dev = wl.views().to(dev)
    small = wl.frames_local * wl.img * wl.img < 5 * 512 * 512
    timed = Timer(dev, world, flush_l2=small)

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
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_dev, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches_per_step = max(model.kernel_launches_per_forward().values())
    ms_step = ms_total / args.steps
    value = wl.frames_total / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel
    prog, roofline = roofline_leg(wl, model, timed, args.steps, ms_step)
    if args.layer_table and rank == 0:
        write_layer_table(args.layer_table, prog, dev)

    # ---- the same step in the other precisions (all ranks: sharded programs contain a collective)
    precisions = None
    if not args.no_parity_value:
        precisions = {}
        for prec in ("bf16x3", "fp16x3", "mixed", "fp16", "bf16"):
            if prec == wl.precision or prec not in engine.PRECISIONS:
                continue
            model.set_precision(prec)
            for _ in range(3):
                step_dev()
            k = max(3, args.steps // 2)
            ms_p = timed(step_dev, k) / k
            precisions[prec] = {"precision": prec, "value": wl.frames_total / (ms_p * 1e-3), "unit": UNIT,
                                "ms_per_step": ms_p}
        model.set_precision(wl.precision)
        for prec, tol in (("bf16x3", "1e-3"), ("fp16x3", "1e-3"), ("mixed", "1e-3"), ("fp16", "8e-3")):
            if prec in precisions:
                precisions[prec]["logit_tolerance"] = ("%s of max|logit| vs the fp32 reference "
                                                       "(tests/test_parity_gpu.py)" % tol)

    # ---- parity of THIS run against the fp32 reference on one scene (rank 0 checks its own agents' logits; every
    #      rank takes part in the sharded forward)
    parity = None
    cpu_baseline = None
    if not args.no_cpu_baseline:
        ref_logits = None
        if rank == 0:
            r = cpu_reference_rate(wl, 1 if wl.arch == "MIMOcom" else min(wl.scenes, 2), steps=2, warmup=1)
            cpu_baseline = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            out = r["out"]
            ref_logits = (out[0] if isinstance(out, tuple) else out).float().cpu()
            if wl.arch == "MIMOcom":   # agent-major (N*B, C, H, W) with B = 1: this rank's agents
                ref_logits = ref_logits[rank * wl.local_agents:(rank + 1) * wl.local_agents]
        one = wl.views(scenes=1 if wl.arch == "MIMOcom" else min(wl.scenes, 2)).to(dev)
        from oracle import when2com_oracle as orc
        parity = {"sample": "1 scene x %d agents @%dx%d, same weights and views as cpu_baseline; rank 0's %d agent(s)%s"
                            % (wl.n_agents, wl.img, wl.img, wl.local_agents,
                               " of the SHARDED forward" if wl.shard else "")}
        for prec in ("bf16", "fp16", "mixed", "fp16x3", "bf16x3"):
            if prec not in engine.PRECISIONS:
                continue
            model.set_precision(prec)
            out = model(one, **kw)
            pred = (out[0] if isinstance(out, tuple) else out).float().cpu()
            if rank == 0:
                parity[prec] = {"max_logit_err_over_max_logit": float((pred - ref_logits).abs().max() / ref_logits.abs().max()),
                                "miou_vs_reference_argmax": orc.miou_between(ref_logits, pred)}
        model.set_precision(wl.precision)

    # ---- e2e (host buffers, copies inside the timed region)
    e2e_dropin = None
    if not args.no_dropin_e2e:
        e2e_dropin = e2e_dropin_leg(wl, model, dev, timed, args.steps)
    e2e = e2e_fused_leg(wl, model, dev, timed, args.steps)

    # ---- 'activated' (what the shipped trainers evaluate with, trainer.py:801): value and e2e, no host sync
    activated = None
    if wl.arch == "MIMOcom" and args.inference == "softmax" and not args.lean:
        wl_act = Workload(args, world, rank)
        wl_act.kw = dict(kw, inference="activated")
        wl_act.inference = "activated"
        for _ in range(3):
            model(x_dev, **wl_act.kw)
        ms_a = timed(lambda: model(x_dev, **wl_act.kw), args.steps) / args.steps
        ea = e2e_fused_leg(wl_act, model, dev, timed, args.steps)
        activated = {"inference": "activated", "value": wl.frames_total / (ms_a * 1e-3), "ms_per_step": ms_a,
                     "e2e": ea["value"], "e2e_ms_per_step": ea["ms_per_step"], "unit": UNIT,
                     "note": "two decoder inputs differ from 'softmax' only in the re-selected fusion weights; "
                             "num_connect stays on the device (read lazily), so the pipeline is not drained"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- N = 1 extras: library baseline and the other BASELINE configs that fit one GPU
    lib_base = None
    extra = None
    if world == 1:
        del x_dev
        model._w2c["programs"].clear()
        torch.cuda.empty_cache()
        if not args.no_library_baseline:
            try:
                lib_base = library_baseline(wl, dev)
            except Exception as e:
                lib_base = {"error": str(e)[:300]}
        if not args.no_extra_configs and args.config == 2 and args.backbones == "n_segnet":
            extra = {}
            for name, ov in (("config2_b1_latency", dict(batch=1)),
                             ("config4_single_agent_5x1024_fp16", dict(config=4)),
                             ("config2_resnet18_pair", dict(backbones="resnet"))):
                try:
                    rec, m2, _wl2 = quick_value(args, ov, dev, max(5, args.steps // 2))
                    extra[name] = rec
                    del m2
                    torch.cuda.empty_cache()
                except Exception as e:
                    extra[name] = {"error": str(e)[:300]}
            # one training step (train-mode forward + loss + backward + SGD) next to the same step of the
            # unmodified reference through stock torch autograd / cuDNN on this GPU (tools/gpu_train_bench.py)
            try:
                from tools.gpu_train_bench import train_step_rates
                extra["train_step"] = train_step_rates(dev, scenes=2, agents=5, img=IMG, steps=max(3, args.steps // 4),
                                                       library=not args.no_library_baseline)
            except Exception as e:
                extra["train_step"] = {"error": str(e)[:300]}
            torch.cuda.empty_cache()
            # the same step with the resnet18 + simple_decoder pair every shipped YAML trains
            try:
                extra["train_step_resnet18_pair"] = train_step_rates(
                    dev, scenes=2, agents=5, img=IMG, backbones="resnet", steps=max(3, args.steps // 4),
                    precisions=("bf16",), library=not args.no_library_baseline)
            except Exception as e:
                extra["train_step_resnet18_pair"] = {"error": str(e)[:300]}
            torch.cuda.empty_cache()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": {"bf16": "bf16", "fp16": "fp16", "mixed": "fp16 hi/lo planes, 1-3 MMA passes per layer"}.get(
                wl.precision, "bf16x3 (hi/lo split, fp32-grade)"),
            "data": "synthetic", "config": wl.config_dict(), "e2e": e2e, "e2e_dropin": e2e_dropin,
            "gpu_launches": int(launches_per_step * args.steps), "launches_per_step": int(launches_per_step),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "library_baseline": lib_base,
            "parity": parity, "precisions": precisions,
            "parity_precision": (precisions or {}).get("mixed") or (precisions or {}).get("bf16x3"),
            "activated": activated, "extra_configs": extra, "frames_per_step": wl.frames_total,
            "model_tflops": wl.gflop_frame * value / 1e3}
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
            taps = 1 if a.kind in (3, 4) else 9
            if a.kind in (1, 4):
                m = a.n * (a.h_in // 2) * (a.w_in // 2)
            else:
                m = a.n * a.h_in * a.w_in
            flop = 2.0 * m * a.cin * a.cout * taps
            names = {0: "conv3x3 s1", 1: "conv3x3 s2", 2: "deconv3x3 s2", 3: "conv1x1", 4: "conv1x1 s2"}
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
