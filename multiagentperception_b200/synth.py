"""Deterministic synthetic inputs and weights (there is no dataset or checkpoint offline).

Inputs follow the value distribution of the reference loader (ptsemseg/loader/airsim_loader.py:515-540: uint8 image,
RGB->BGR, minus mean [103.939, 116.779, 123.68], /255), views concatenated on the channel axis (trainer.py:651).
Weights are re-drawn in place from a CPU generator so the same seed gives bit-identical state_dicts to the reference
modules, to these modules and on every machine: He-normal conv / linear weights (so activations keep O(1) scale
through the 27-layer stacks, like trained weights do), small biases, and BatchNorm affine + running statistics away
from identity so that eval-mode BN is a real per-channel affine.
"""
import zlib

import torch

LOADER_MEAN_BGR = (103.939, 116.779, 123.68)


def synthetic_views(batch, n_agents, height, width, seed=1337, device="cpu"):
    """(batch, 3*n_agents, H, W) float32, distributed like airsimLoader.transform output."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    img = torch.randint(0, 256, (batch, n_agents, 3, height, width), generator=g, dtype=torch.int32).to(torch.float32)
    mean = torch.tensor(LOADER_MEAN_BGR, dtype=torch.float32).view(1, 1, 3, 1, 1)
    x = ((img - mean) / 255.0).reshape(batch, 3 * n_agents, height, width)
    return x.to(device)


def synthetic_frames(batch, n_agents, height, width, seed=1337):
    """(batch, n_agents, H, W, 3) uint8 RGB: raw camera frames as airsimLoader.__getitem__ holds them before
    transform() (airsim_loader.py:493-496)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randint(0, 256, (batch, n_agents, height, width, 3), generator=g, dtype=torch.uint8)


def randomize_(module, seed=1337):
    """Re-initialise every parameter and buffer of `module` in place. Each tensor is drawn from its own generator
    seeded by (seed, crc32(state_dict key)), so the values do not depend on registration order."""
    g = torch.Generator(device="cpu")
    sd = module.state_dict()
    seen = {}
    with torch.no_grad():
        for name, t in sd.items():
            if t.data_ptr() in seen and t.numel() > 0:
                continue  # aliased registration (resnet_encoder registers its trunk twice, backbone.py:63-69)
            seen[t.data_ptr()] = name
            leaf = name.rsplit(".", 1)[-1]
            if leaf == "num_batches_tracked":
                continue
            g.manual_seed((int(seed) << 32) ^ zlib.crc32(name.encode()))
            if leaf == "running_mean":
                new = torch.randn(t.shape, generator=g) * 0.1
            elif leaf == "running_var":
                new = torch.rand(t.shape, generator=g) * 0.5 + 0.75
            elif t.dim() >= 2:  # conv / transposed conv / linear weight
                if t.dim() == 4:
                    # ConvTranspose2d weight is (cin, cout, kh, kw): fan-in is cin * (kh*kw / stride^2) on average
                    transposed = ".dcbr_unit." in name or "desqueezer" in name
                    fan_in = (t.shape[0] * t.shape[2] * t.shape[3] / 4.0) if transposed else t.shape[1] * t.shape[2] * t.shape[3]
                else:
                    fan_in = t.shape[1]
                new = torch.randn(t.shape, generator=g) * (2.0 / fan_in) ** 0.5
            elif leaf == "weight":  # BatchNorm gamma
                new = torch.rand(t.shape, generator=g) * 0.5 + 0.75
            else:  # conv / linear / BatchNorm bias
                new = torch.randn(t.shape, generator=g) * 0.05
            t.copy_(new.to(device=t.device, dtype=t.dtype))
    return module
