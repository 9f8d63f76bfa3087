"""TEST INFRASTRUCTURE (build container only): import the UNMODIFIED reference from /root/reference on CPU.

The reference needs two harness-side shims, neither of which edits its files:
  1. `pretrainedmodels` (backbone.py:5,63; not installed): a stub module whose resnet18 is torchvision's resnet18 with
     `fc` renamed `last_linear`, which is what pretrainedmodels.resnet18 itself returns;
  2. hard `.cuda()` / `.to('cuda')` / torch.cuda.FloatTensor calls inside forward (agent.py:323,325,1040,1166,...):
     neutralised while a reference forward runs on CPU.
/root/reference does not exist on the GPU box; nothing that runs there imports this file.
"""
import contextlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("W2C_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ptsemseg", "models"))


def _install_pretrainedmodels_stub():
    if "pretrainedmodels" in sys.modules:
        return
    import torchvision

    def resnet18(num_classes=1000, pretrained=None):
        net = torchvision.models.resnet18(num_classes=num_classes)
        net.last_linear = net.fc
        del net.fc
        return net

    mod = types.ModuleType("pretrainedmodels")
    mod.resnet18 = resnet18
    sys.modules["pretrainedmodels"] = mod


def import_reference_models():
    """Returns the reference's ptsemseg.models package (get_model etc.)."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    _install_pretrainedmodels_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import ptsemseg.models as ref_models  # noqa: E402
    return ref_models


@contextlib.contextmanager
def cpu_cuda_shims():
    """Make `.cuda()`, `.to('cuda')` and `.type(torch.cuda.FloatTensor)` no-ops on CPU tensors."""
    orig_cuda, orig_to, orig_type = torch.Tensor.cuda, torch.Tensor.to, torch.Tensor.type
    orig_ft = torch.cuda.FloatTensor

    def to(self, *a, **k):
        a = tuple(x for x in a if not (isinstance(x, str) and x.startswith("cuda")))
        if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
            k.pop("device")
        return orig_to(self, *a, **k) if (a or k) else self

    def typ(self, dtype=None, *a, **k):
        if dtype is orig_ft or dtype is torch.cuda.FloatTensor:
            return self.float()
        return orig_type(self, dtype, *a, **k)

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.to = to
    torch.Tensor.type = typ
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.Tensor.to, torch.Tensor.type = orig_cuda, orig_to, orig_type


def build_reference_model(cfg, n_classes=11):
    import io
    ref_models = import_reference_models()
    with contextlib.redirect_stdout(io.StringIO()):  # the reference prints from its constructors
        model = ref_models.get_model(cfg, n_classes)
    return model.eval()


def reference_forward(model, x, **kw):
    import io
    with torch.no_grad(), cpu_cuda_shims(), contextlib.redirect_stdout(io.StringIO()):
        return model(x, **kw)
