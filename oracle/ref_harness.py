"""TEST / BENCH INFRASTRUCTURE: import and run the UNMODIFIED reference (GT-RIPL/MultiAgentPerception).

Where the reference comes from, in this order:
  1. $W2C_REFERENCE_ROOT or /root/reference (the read-only mount of the build container), else
  2. oracle/_ref/ - the byte-identical staged copy oracle/make_ref.py makes there (git-ignored; it travels to the GPU
     box with the snapshot), which is what bench.py's `--impl reference`, `cpu_baseline` and `library_baseline` legs
     and the drop-in tests run on the GPU box.
The reference needs harness-side shims, none of which edits its files:
  * import stand-ins for packages this image lacks (`pretrainedmodels`, `tensorboardX`, `matplotlib`): the stub
    modules make_ref.py writes, or in-process equivalents when running straight from /root/reference;
  * on CPU only: the hard `.cuda()` / `.to('cuda')` / torch.cuda.FloatTensor calls inside forward
    (agent.py:323,325,1040,1166,...) are neutralised while a reference forward runs. On a GPU they run as written.
Only tests/, __graft_entry__ and bench.py's baseline legs import this file; the product never does.
"""
import contextlib
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")


def reference_root():
    """Directory that contains the reference's `ptsemseg` package, or None."""
    for cand in (os.environ.get("W2C_REFERENCE_ROOT"), "/root/reference", STAGED):
        if cand and os.path.isdir(os.path.join(cand, "ptsemseg", "models")):
            return cand
    return None


REFERENCE_ROOT = reference_root() or "/root/reference"


def available():
    return reference_root() is not None


def source_kind():
    """'mounted' (/root/reference), 'staged' (oracle/_ref copy) or None."""
    root = reference_root()
    if root is None:
        return None
    return "staged" if os.path.abspath(root) == os.path.abspath(STAGED) else "mounted"


def _install_stubs():
    if "pretrainedmodels" not in sys.modules:
        import torchvision

        def resnet18(num_classes=1000, pretrained=None):
            net = torchvision.models.resnet18(num_classes=num_classes)
            net.last_linear = net.fc
            del net.fc
            return net

        mod = types.ModuleType("pretrainedmodels")
        mod.resnet18 = resnet18
        sys.modules["pretrainedmodels"] = mod
    if "tensorboardX" not in sys.modules:
        try:
            import tensorboardX  # noqa: F401
        except ImportError:
            mod = types.ModuleType("tensorboardX")

            class SummaryWriter(object):
                def __init__(self, *a, **k):
                    pass

                def add_scalar(self, *a, **k):
                    pass

                def close(self):
                    pass

            mod.SummaryWriter = SummaryWriter
            sys.modules["tensorboardX"] = mod
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            cm = types.ModuleType("matplotlib.cm")
            plt.__getattr__ = lambda name: (lambda *a, **k: None)
            cm.__getattr__ = lambda name: None
            mpl.pyplot, mpl.cm = plt, cm
            mpl.use = lambda *a, **k: None
            sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "matplotlib.cm": cm})


def _reference_on_path():
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not found (looked at $W2C_REFERENCE_ROOT, /root/reference and %s); run "
                           "`python oracle/make_ref.py` in the build container" % STAGED)
    _install_stubs()
    mod = sys.modules.get("ptsemseg")
    if mod is not None and not os.path.abspath(getattr(mod, "__file__", "") or "").startswith(os.path.abspath(root)):
        # a different `ptsemseg` (e.g. this repo's drop-in shim package) is already imported: forget it, so that
        # the names below resolve to the reference
        for name in [n for n in sys.modules if n == "ptsemseg" or n.startswith("ptsemseg.")]:
            del sys.modules[name]
    if root in sys.path:
        sys.path.remove(root)
    sys.path.insert(0, root)
    return root


def import_reference_models():
    """Returns the reference's ptsemseg.models package (get_model etc.)."""
    _reference_on_path()
    import ptsemseg.models as ref_models  # noqa: E402
    return ref_models


def import_reference_module(name):
    """Any other module of the reference package, e.g. 'ptsemseg.metrics' or 'ptsemseg.trainer'."""
    import importlib
    _reference_on_path()
    return importlib.import_module(name)


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
    """One no-grad forward of a reference model: on CPU behind the .cuda() shims, on a GPU exactly as written."""
    import io
    on_cpu = not x.is_cuda
    shims = cpu_cuda_shims() if on_cpu else contextlib.nullcontext()
    with torch.no_grad(), shims, contextlib.redirect_stdout(io.StringIO()):
        return model(x, **kw)


def reference_train_step_grads(model, x, labels, **kw):
    """model.train(); loss = cross_entropy2d(model(x, ...), labels); loss.backward() exactly as Trainer_*.train() does
    (trainer.py:659-670), with the reference's own loss (ptsemseg/loss/loss.py:5-18). Returns (outputs, loss,
    {state_dict key: gradient}). On CPU the forward runs behind the .cuda() shims."""
    import io
    loss_mod = import_reference_module("ptsemseg.loss.loss")
    on_cpu = not x.is_cuda
    shims = cpu_cuda_shims() if on_cpu else contextlib.nullcontext()
    model.train()
    model.zero_grad()
    import warnings
    with shims, contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = model(x, **kw)
        pred = out[0] if isinstance(out, tuple) else out
        loss = loss_mod.cross_entropy2d(input=pred, target=labels)
        loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    return out, float(loss.detach()), grads
