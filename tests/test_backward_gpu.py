"""GPU: the backward pass (second half) - model.train(); outputs = model(images, training=True,
...); loss = cross_entropy2d(outputs, labels); loss.backward(); optimizer.step() exactly as Trainer_*.train() runs it
(ptsemseg/trainer.py:659-670), through the CUDA path, against the oracle's autograd gradients (which
tests/test_oracle.py pins to loss.backward() through the UNMODIFIED reference).

What "equal" can mean here. The gradient of this network is ill-conditioned in its forward rounding: every one of the
25-45 ReLU masks flips for the few activations within rounding distance of zero, and each flip changes that element's
gradient by its full size. Measured on the fp32 oracle itself (first test below): rounding the weights and the input to
16 significant bits - nothing else - moves the gradients by 4-5 % in relative L2 on the n_segnet pair, 1-2 % on the
resnet18 pair; a 1e-6 relative perturbation still moves them by 2 %. The CUDA path in 'bf16x3' (16 significant bits,
logits within 1e-3 of the reference in train mode) lands on that floor; the gates below are the floor with a margin, and
the kernels themselves are held to 1e-4 against float64 autograd one by one (tests/test_backward_kernels_gpu.py)."""
import pytest
import torch

from multiagentperception_b200 import configs, synth
from multiagentperception_b200.models import get_model
from oracle import when2com_oracle as orc
from tools import gpu_grad_check as gc

pytestmark = pytest.mark.gpu

# name -> (global relative L2 bound, per-tensor relative L2 bound). The per-tensor bound is the loose one: the policy
# net's gradients pass through the softmax over key.query scores and sit seven orders of magnitude below the
# decoder's at initialisation (1e-8 against 1e-1); they carry 10-13 % of relative noise where the rest carries 1-6 %.
GATES = {
    "single_segnet": (0.08, 0.25), "mimocom_segnet": (0.08, 0.25), "when2com_segnet": (0.08, 0.25),
    "mimocomwho_segnet": (0.08, 0.25),
    "single_resnet": (0.04, 0.25), "mimocom_resnet": (0.04, 0.25), "who2com_resnet": (0.04, 0.25),
    "mimo_all_resnet": (0.04, 0.25),
}


def _bf16x2(t):
    hi = t.to(torch.bfloat16).float()
    return hi + (t - hi).to(torch.bfloat16).float()


def test_backward_sits_on_the_rounding_floor_of_the_reference(cuda_device):
    """Single_agent / n_segnet at 128x128: the CUDA gradients differ from the fp32 oracle's by no more than twice what
    rounding the oracle's own weights and input to the same 16 significant bits does."""
    r = gc.run_case("single_segnet")
    arch, bb, over, kw, n, img = gc.CASES["single_segnet"]
    cfg = configs.make_config(arch, img_size=img, backbones=bb, **over)
    model = get_model(cfg, 11)
    synth.randomize_(model, 1337)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    x = synth.synthetic_views(2, n, img, img, seed=7)
    labels = torch.randint(0, 11, (2, img, img), generator=torch.Generator().manual_seed(3))
    labels[0, :8] = 250
    _, _, g0 = orc.forward_with_grads(sd0, cfg, x, labels)
    sd1 = {k: (_bf16x2(v) if torch.is_floating_point(v) else v) for k, v in sd0.items()}
    _, _, g1 = orc.forward_with_grads(sd1, cfg, _bf16x2(x), labels)
    num = den = 0.0
    for k in g0:
        if float(g0[k].abs().max()) > 1e-6:
            num += float(((g1[k] - g0[k]).double() ** 2).sum())
            den += float((g0[k].double() ** 2).sum())
    floor = (num / den) ** 0.5
    assert not r["missing"] and not r["extra"], r
    assert abs(r["loss"][0] - r["loss"][1]) <= 1e-4 * abs(r["loss"][1]), r["loss"]
    assert floor > 5e-3, floor          # (the premise: the problem really is this ill-conditioned)
    assert r["global_rel_l2"] <= 2.0 * floor, (r["global_rel_l2"], floor)


@pytest.mark.parametrize("name", sorted(GATES))
def test_every_parameter_gradient_against_the_oracle(name, cuda_device):
    r = gc.run_case(name)
    assert "error" not in r, r
    assert not r["missing"], r["missing"]       # every parameter the reference differentiates gets a gradient ...
    assert not r["extra"], r["extra"]           # ... and no other
    assert abs(r["loss"][0] - r["loss"][1]) <= 2e-4 * abs(r["loss"][1]), r["loss"]
    g_all, g_one = GATES[name]
    # parameters whose gradient is mathematically zero (conv biases in front of a BatchNorm, key_net's last bias)
    assert r["zero_grad_max_over_gmax"] <= 1e-3, r["zero_grad_max_over_gmax"]   # (the fp32 reference itself holds ~1e-5 of noise there)
    assert r["global_rel_l2"] <= g_all, r
    assert r["max_rel_l2"] <= g_one, r["worst"]


def test_training_steps_track_the_reference(cuda_device):
    """Four SGD steps (forward, loss, backward, optimizer.step) from the same initial weights on the CUDA path and on
    the fp32 oracle: the losses agree step by step - the backward pass, the accumulation into .grad, and the re-packing
    of the tensor-core operands from the updated parameters inside the captured program."""
    import torch.nn.functional as F
    cfg = configs.make_config("MIMOcom", img_size=256, backbones="resnet", agent_num=2)
    kw = dict(training=True, MO_flag=True)
    model = get_model(cfg, 11)
    synth.randomize_(model, 1337)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x = synth.synthetic_views(2, 2, 256, 256, seed=7)
    labels = torch.randint(0, 11, (4, 256, 256), generator=torch.Generator().manual_seed(3))
    lr, steps = 0.05, 4
    ref_losses = []
    for _ in range(steps):
        _, loss, grads = orc.forward_with_grads(sd, cfg, x, labels, **kw)
        ref_losses.append(loss)
        for k, g in grads.items():
            sd[k] = sd[k] - lr * g
    model = model.to(cuda_device).set_precision("bf16x3")
    model.train()
    opt = torch.optim.SGD(model.parameters(), lr=lr)
    xd, ld = x.to(cuda_device), labels.to(cuda_device)
    got = []
    for _ in range(steps):
        opt.zero_grad()
        pred = model(xd, **kw)[0]
        loss = F.cross_entropy(pred.permute(0, 2, 3, 1).reshape(-1, 11), ld.reshape(-1), ignore_index=250)
        loss.backward()
        opt.step()
        got.append(float(loss.detach()))
    assert ref_losses[-1] < ref_losses[0] - 0.02, ref_losses      # the steps really train
    for a, b in zip(got, ref_losses):
        assert abs(a - b) <= 3e-3 * abs(b), (got, ref_losses)


def test_gradients_accumulate_and_stale_backward_is_refused(cuda_device):
    import torch.nn.functional as F
    cfg = configs.make_config("Single_agent", img_size=128, backbones="resnet")
    model = get_model(cfg, 11)
    synth.randomize_(model, 1337)
    model = model.to(cuda_device).set_precision("bf16x3")
    model.train()
    x = synth.synthetic_views(2, 1, 128, 128, seed=7).to(cuda_device)
    labels = torch.randint(0, 11, (2, 128, 128), generator=torch.Generator().manual_seed(3)).to(cuda_device)
    loss_of = lambda p: F.cross_entropy(p.permute(0, 2, 3, 1).reshape(-1, 11), labels.reshape(-1))
    loss_of(model(x)).backward()
    w = model.decoder.output_decoder.pred[2].weight
    g1 = w.grad.clone()
    # BatchNorm running statistics moved, the weights did not: the second backward adds the same gradient again
    loss_of(model(x)).backward()
    assert float((w.grad - 2 * g1).abs().max()) <= 1e-3 * float(g1.abs().max())
    p1 = model(x)
    model(x)                                    # a newer forward of the same shape overwrites the program's buffers
    with pytest.raises(RuntimeError, match="another forward"):
        loss_of(p1).backward()
    # under no_grad the train-mode forward carries no graph and records no backward program
    with torch.no_grad():
        assert not model(x).requires_grad
    # fp16 storages cannot run the backward pass: refused with a pointer to the precisions that can
    model.set_precision("fp16x3")
    with pytest.raises(NotImplementedError, match="bf16x3"):
        model(x)
