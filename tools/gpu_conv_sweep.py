#!/usr/bin/env python
"""Per-layer kernel sweep on a B200: every distinct conv / deconv shape of the n_segnet MIMOcom forward at the bench
batch (40 agent-frames), timed with CUDA events for each tensor-core kernel variant (one-tile vs persistent, tuning flags, BLOCK_N).
Used to pick the dispatch heuristic; writes gpurun_out/conv_sweep.md."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from multiagentperception_b200 import ops  # noqa: E402

N = int(os.environ.get("SWEEP_FRAMES", "40"))
# (kind, h_in, cin, cout, nchw_out)
LAYERS = [
    (1, 512, 64, 64, False), (0, 256, 64, 128, False), (1, 256, 128, 128, False), (0, 128, 128, 256, False),
    (0, 128, 256, 256, False), (1, 128, 256, 256, False), (0, 64, 256, 512, False), (0, 64, 512, 512, False),
    (1, 64, 512, 512, False), (0, 32, 512, 512, False), (1, 32, 512, 512, False), (0, 16, 512, 512, False),
    (0, 16, 512, 256, False), (2, 16, 512, 512, False), (2, 32, 512, 512, False), (0, 64, 512, 256, False),
    (2, 64, 256, 256, False), (0, 128, 256, 128, False), (2, 128, 128, 128, False), (0, 256, 128, 64, False),
    (2, 256, 64, 64, False), (0, 512, 64, 11, True),
]
NAMES = {0: "conv s1", 1: "conv s2", 2: "deconv"}


def time_ms(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    dev = torch.device("cuda:0")
    act = ops.ACT_BF16X2 if os.environ.get("SWEEP_X3") == "1" else ops.ACT_BF16
    lines = ["| kind | hw | cin | cout | variant | ms | TFLOP/s |", "|---|---|---|---|---|---|---|"]
    only = os.environ.get("SWEEP_ONLY")  # "kind,h,cin,cout[;kind,h,cin,cout...]": restrict to these layers (ncu captures)
    only = [tuple(int(v) for v in o.split(",")) for o in only.split(";")] if only else None
    for kind, h, cin, cout, nchw in LAYERS:
        if only is not None and (kind, h, cin, cout) not in only:
            continue
        x = torch.randn(N, h, h, ops.planes_of(act) * cin, device=dev).to(torch.bfloat16)
        wt = torch.randn((cin, cout, 3, 3) if kind == 2 else (cout, cin, 3, 3), device=dev) * 0.05
        wp = ops.pack_conv_weight(wt, cin, kind == 2, act)
        scale = torch.ones(cout, device=dev)
        shift = torch.zeros(cout, device=dev)
        ho = h // 2 if kind == 1 else h * 2 if kind == 2 else h
        if nchw:
            y = torch.empty(N, cout, ho, ho, device=dev)
        else:
            y = torch.empty(N, ho, ho, ops.planes_of(act) * cout, device=dev, dtype=torch.bfloat16)
        m = N * (ho * ho if kind != 2 else h * h)
        flop = 2.0 * m * cin * cout * 9
        # variants: the one-tile kernel, the persistent kernel and its tuning flags (bits 8.. of impl: 1 no row-halo
        # stages, 2 one epilogue group, 4 / 8 two / three CTAs per SM, 16 one, 32 no resident weights, 64 resident
        # opt-ins, 128 class-major transposed-conv order, 256 per-thread logits stores)
        variants = [("taps", ops.IMPL_TC_TAPS, 0), ("pers", ops.IMPL_TC_PERSIST, 0),
                    ("pers 1-epi-group", ops.IMPL_TC_PERSIST | (2 << 8), 0)]
        if kind == 2:
            variants.append(("pers class-major", ops.IMPL_TC_PERSIST | (128 << 8), 0))
        if cout == 64:
            variants.append(("pers 2cta/sm", ops.IMPL_TC_PERSIST | (4 << 8), 0))
        if cout <= 16:
            variants += [("pers 1cta/sm", ops.IMPL_TC_PERSIST | (16 << 8), 0),
                         ("pers 3cta/sm", ops.IMPL_TC_PERSIST | (8 << 8), 0),
                         ("pers per-thread stores", ops.IMPL_TC_PERSIST | (256 << 8), 0)]
        cp = ops.cout_pad(cout)
        if os.environ.get("SWEEP_PROD") == "1":  # only the production dispatch (for ncu captures)
            variants = [("production dispatch", ops.IMPL_TCGEN05, 0)]
            cp = 1
        if cp % 256 == 0:
            variants += [("pers bn128", ops.IMPL_TC_PERSIST, 128), ("taps bn256", ops.IMPL_TC_TAPS, 256)]
        if cp % 128 == 0:
            variants.append(("taps bn64", ops.IMPL_TC_TAPS, 64))
        for name, impl, bn in variants:
            k2, w2 = kind, wp

            def run(k2=k2, w2=w2, impl=impl, bn=bn):
                ops.conv_bnrelu(x, w2, scale, shift, y, n=N, h_in=h, w_in=h, cin=cin, cout=cout, kind=k2, relu=True,
                                act=act, out_fmt=ops.OUT_NCHW_F32 if nchw else ops.OUT_NHWC, impl=impl, block_n=bn)
            try:
                ms = time_ms(run)
                lines.append("| %s | %d | %d | %d | %s | %.4f | %.1f |" % (NAMES[kind], h, cin, cout, name, ms,
                                                                         flop / ms / 1e9))
            except Exception as e:  # report and continue
                lines.append("| %s | %d | %d | %d | %s | ERROR %s | |" % (NAMES[kind], h, cin, cout, name, str(e)[:80]))
            print(lines[-1], flush=True)
    out = os.path.join(ROOT, "gpurun_out", "conv_sweep.md")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
