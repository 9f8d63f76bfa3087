"""TEST / BENCH INFRASTRUCTURE: stage the UNMODIFIED reference package where the GPU box can see it.

    python oracle/make_ref.py            # copies /root/reference/ptsemseg -> oracle/_ref/ptsemseg (+ import stubs)

/root/reference exists only in the build container; the GPU box receives /root/repo. oracle/_ref/ is git-ignored
(no reference source ever enters the history) but not gpurun-ignored, so the staged copy travels with the snapshot
like the built libw2c.so does. It is what `bench.py --impl reference`, the `cpu_baseline` / `library_baseline`
legs and the drop-in tests run when /root/reference itself is absent: the reference's own modules, byte for byte
(the copy is verified with sha256 below), behind the harness shims of oracle/ref_harness.py - never edits.

Also written: oracle/_ref/_stubs/ with import stand-ins for third-party packages the reference imports at module
scope and this image does not have:
  pretrainedmodels  resnet18 -> torchvision.models.resnet18 with fc renamed last_linear (backbone.py:5,63)
  tensorboardX      SummaryWriter no-op (trainer.py:28; only train() logs through it)
  matplotlib        empty pyplot / cm (airsim_loader.py:1-3 imports it for a debug plot)
__graft_entry__.build() calls stage() whenever /root/reference is present.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("W2C_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")

_STUBS = {
    "pretrainedmodels/__init__.py": '''"""Stand-in for the `pretrainedmodels` package (not installed): what pretrainedmodels.resnet18 returns is
torchvision's resnet18 with the classifier renamed `last_linear`; backbone.py:63-69 uses conv1..layer4 only."""
import torchvision


def resnet18(num_classes=1000, pretrained=None):
    net = torchvision.models.resnet18(num_classes=num_classes)
    net.last_linear = net.fc
    del net.fc
    return net
''',
    "tensorboardX/__init__.py": '''"""Stand-in for tensorboardX (not installed): trainer.py:28 imports SummaryWriter at module scope."""


class SummaryWriter(object):
    def __init__(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass

    def close(self):
        pass
''',
    "matplotlib/__init__.py": '"""Stand-in for matplotlib (not installed): airsim_loader.py imports it for a debug plot."""\n\n\ndef use(*a, **k):\n    return None\n',
    "matplotlib/pyplot.py": "def _noop(*a, **k):\n    return None\n\n\ndef __getattr__(name):\n    return _noop\n",
    "matplotlib/cm.py": "def __getattr__(name):\n    return None\n",
}


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(verbose=False):
    """Copy the reference's ptsemseg package to oracle/_ref/ and write the stubs. Returns the destination, or None
    when the reference tree is not mounted (the GPU box: use what was staged in the build container)."""
    src = os.path.join(REF_SRC, "ptsemseg")
    if not os.path.isdir(src):
        return None
    dst = os.path.join(DST, "ptsemseg")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    n = 0
    for root, _dirs, files in os.walk(src):
        for f in files:
            if not f.endswith(".py"):
                continue
            a = os.path.join(root, f)
            b = os.path.join(dst, os.path.relpath(a, src))
            if _sha(a) != _sha(b):
                raise RuntimeError("staged copy differs from the reference: %s" % b)
            n += 1
    for rel, text in _STUBS.items():
        path = os.path.join(DST, "_stubs", rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write(text)
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write("%s (%d python files, verified by sha256)\n" % (src, n))
    if verbose:
        print("staged %d reference files under %s" % (n, DST), file=sys.stderr)
    return DST


if __name__ == "__main__":
    out = stage(verbose=True)
    print(out or "reference tree not found at %s" % REF_SRC)
