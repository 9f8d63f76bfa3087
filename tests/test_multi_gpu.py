"""GPU tests that need more than one device (skipped on a 1-GPU box):

  * two models on two devices driven by two host threads of ONE process - the nn.DataParallel contract of train.py:177
    (kernel attribute opt-ins and the SM count are per-device state in libw2c, csrc/common.cuh);
  * the agent-sharded forward under torchrun (NCCL all-gather) against the unsharded forward, all precisions and
    inference modes: tools/gpu_shard_check.py as a test.
"""
import os
import subprocess
import sys
import threading

import pytest
import torch

from multiagentperception_b200 import configs, synth
from multiagentperception_b200.models import get_model

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
two_gpus = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one box")


@two_gpus
def test_two_devices_two_threads_one_process(cuda_device):
    cfg = configs.make_config("MIMOcom", agent_num=3, img_size=128, backbones="n_segnet", precision="bf16x3")
    kw = dict(training=False, MO_flag=True, inference="activated")
    x = synth.synthetic_views(2, 3, 128, 128, seed=5)
    # device 1 FIRST: with per-process (not per-device) attribute state the later device-0 launches would fail
    order = [torch.device("cuda:1"), torch.device("cuda:0")]
    models = []
    for dev in order:
        m = get_model(cfg, 11)
        synth.randomize_(m, 1337)
        models.append(m.to(dev).eval())
    first = models[0](x.to(order[0]), **kw)[0].cpu()
    results, errors = {}, []

    def work(i):
        try:
            for _ in range(3):
                out = models[i](x.to(order[i]), **kw)
            torch.cuda.synchronize(order[i])
            results[i] = [o.cpu() if torch.is_tensor(o) else float(o) for o in out]
        except Exception as e:  # surfaced below
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert torch.equal(results[0][0], first)
    for a, b in zip(results[0], results[1]):
        assert torch.equal(a, b) if torch.is_tensor(a) else a == b


@two_gpus
def test_sharded_forward_equals_unsharded_under_torchrun(cuda_device):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "gpu_shard_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-3000:])
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) >= 12, p.stdout
    assert all('"within_rounding_budget": true' in l for l in lines), p.stdout
