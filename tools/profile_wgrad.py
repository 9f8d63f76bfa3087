#!/usr/bin/env python
"""A few launches of the tensor-core weight-gradient kernel on one layer shape, for
  ncu --set full --import-source on --clock-control none -k regex:wgrad_kernel --launch-skip 2 -c 1 -o X python tools/profile_wgrad.py
and (without ncu) its CUDA-event time / achieved TFLOP/s.  Default: enc.conv9 / dec.deconv5 of the n_segnet pair,
512 -> 512 channels on a 64x64 map, 10 agent-frames (the training bench's step)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from multiagentperception_b200 import ops  # noqa: E402

n, h, w, cin, cout = 10, 64, 64, 512, 512
kind = ops.CONV3X3_S1
if len(sys.argv) > 5:
    n, h, w, cin, cout = (int(v) for v in sys.argv[1:6])
act = int(sys.argv[6]) if len(sys.argv) > 6 else ops.ACT_BF16
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = ops.nchw_to_act(torch.randn(n, cin, h, w, generator=g).to(dev), act)
dy = ops.nchw_to_act((torch.randn(n, cout, h, w, generator=g) * 0.1).to(dev), act)
dw = torch.zeros(cout, 9, cin, device=dev)
run = lambda: ops.conv_wgrad(x, dy, dw, n=n, h_in=h, w_in=w, cin=cin, cout=cout, kind=kind, act_x=act, act_dy=act)
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
flop = 2.0 * n * h * w * cout * 9 * cin * (3 if ops.planes_of(act) == 2 else 1)
print("wgrad %dx%dx%d %d->%d act %d: %.1f us, %.0f TFLOP/s (MMA work)" % (n, h, w, cin, cout, act, ms * 1e3, flop / ms / 1e9))
