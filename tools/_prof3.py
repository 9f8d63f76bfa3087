import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiagentperception_b200 import ops
dev = torch.device("cuda:0")
N = 40
def mk(kind, h, cin, cout, nchw):
    x = torch.randn(N, h, h, cin, device=dev).to(torch.bfloat16)
    wt = torch.randn((cin, cout, 3, 3) if kind == 2 else (cout, cin, 3, 3), device=dev) * 0.05
    wp = ops.pack_conv_weight(wt, cin, kind == 2, 0)
    scale = torch.ones(cout, device=dev); shift = torch.zeros(cout, device=dev)
    ho = h // 2 if kind == 1 else h * 2 if kind == 2 else h
    y = torch.empty(N, cout, ho, ho, device=dev) if nchw else torch.empty(N, ho, ho, cout, device=dev, dtype=torch.bfloat16)
    def run():
        ops.conv_bnrelu(x, wp, scale, shift, y, n=N, h_in=h, w_in=h, cin=cin, cout=cout, kind=kind, relu=True, act=0,
                        out_fmt=1 if nchw else 0, impl=4)
    return run
runs = [mk(0, 512, 64, 11, True), mk(2, 256, 64, 64, False), mk(1, 512, 64, 64, False), mk(0, 256, 128, 64, False)]
for r in runs:
    for _ in range(2): r()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for r in runs: r()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
