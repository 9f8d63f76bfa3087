#!/usr/bin/env python
"""One training step (forward + loss + backward) between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/profile_train_step.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from multiagentperception_b200 import configs, synth  # noqa: E402
from multiagentperception_b200.models import get_model  # noqa: E402
from tools.gpu_train_bench import _loss  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
backbones = sys.argv[2] if len(sys.argv) > 2 else "n_segnet"
scenes, agents, img = 2, 5, 512
dev = torch.device("cuda:0")
cfg = configs.make_config("MIMOcom", agent_num=agents, img_size=img, backbones=backbones)
model = get_model(cfg, 11)
synth.randomize_(model, 1337)
model = model.to(dev).set_precision(prec).set_cuda_graphs(False)
model.train()
x = synth.synthetic_views(scenes, agents, img, img, seed=1337).to(dev)
labels = torch.randint(0, 11, (scenes * agents, img, img), generator=torch.Generator().manual_seed(3)).to(dev)
for i in range(3):
    if i == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    model.zero_grad(set_to_none=True)
    loss = _loss(model(x, training=True, MO_flag=True)[0], labels)
    loss.backward()
    if i == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
print("loss", float(loss.detach()))
