"""The trainers' loss on the device in one pass: `cross_entropy2d` with the reference's signature
(ptsemseg/loss/loss.py:5-18, selected by `training.loss.name: cross_entropy` in every shipped YAML, called at
trainer.py:668 as `self.loss_fn(input=outputs, target=labels)`).

torch's own path for this call - F.cross_entropy on the (N*H*W, C) view - reduces 2.6 M rows in a ONE-block nll_loss
kernel, forward and backward: 4.1 ms of a 37 ms training step of 10 agent-frames (ncu, profiles/r2_train_step.md).
Here one kernel reads the logits once, writes softmax - onehot for the backward pass and accumulates the loss
(csrc/bwd_misc.cu: xent2d_kernel). Opt-in: pass `multiagentperception_b200.loss.cross_entropy2d` as the trainer's
`loss_fn` (Trainer_*.__init__ takes it as an argument, trainer.py:586); the reference's own loss - or any torch loss -
keeps working on the model's outputs like on any tensor.
"""
import ctypes

import torch

from . import _lib

IGNORE_INDEX = 250


class _CrossEntropy2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        lib = _lib.load()
        n, c, h, w = logits.shape
        x = logits.detach().contiguous()
        t = target.contiguous()
        dl = torch.empty_like(x)
        totals = torch.zeros(2, dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(lib.w2c_cross_entropy2d(x.data_ptr(), t.data_ptr(), n, c, h * w, IGNORE_INDEX, dl.data_ptr(),
                                               totals.data_ptr(), stream), "w2c_cross_entropy2d")
        ctx.save_for_backward(dl, totals)
        return (totals[0] / totals[1]).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        dl, totals = ctx.saved_tensors
        return dl * (g / totals[1]).to(torch.float32), None


def cross_entropy2d(input, target, weight=None, size_average=True):
    """Mean cross-entropy over the pixels of (N, C, H, W) fp32 logits against (N, H, W) int64 labels, ignore_index 250:
    ptsemseg.loss.cross_entropy2d for equal input / target sizes, without class weights."""
    if weight is not None or not size_average:
        raise NotImplementedError("cross_entropy2d on the device: weight=None, size_average=True (the shipped configs)")
    if not (torch.is_tensor(input) and input.is_cuda and input.dtype == torch.float32 and input.dim() == 4):
        raise RuntimeError("cross_entropy2d needs (N, C, H, W) fp32 CUDA logits: there is no CPU fallback")
    if target.dtype != torch.int64 or tuple(target.shape) != (input.shape[0], input.shape[2], input.shape[3]):
        raise ValueError("cross_entropy2d: target must be int64 of shape (N, H, W) = %s, got %s %s (the reference "
                         "up-samples mismatched logits; the models here return full-size logits)"
                         % ((input.shape[0], input.shape[2], input.shape[3]), target.dtype, tuple(target.shape)))
    return _CrossEntropy2d.apply(input, target)
