"""Drop-in overlay of the reference's `ptsemseg` package (GT-RIPL/MultiAgentPerception).

Put this directory AHEAD of the reference checkout on PYTHONPATH:

    PYTHONPATH=/path/to/repo/dropin:/path/to/repo:/path/to/MultiAgentPerception  python test.py --config ...

`import ptsemseg.models` then resolves to the B200 path (ptsemseg/models/__init__.py here: the same get_model(),
ptsemseg/models/__init__.py:8-101 of the reference) and `ptsemseg.visual` to the module test.py:14 imports but the
reference never shipped, while every other submodule (trainer, loader, metrics, loss, ...) still comes from the
UNMODIFIED reference: this package extends its own __path__ with the reference's ptsemseg directory, found through
$W2C_REFERENCE_ROOT or as the next `ptsemseg` on sys.path. Nothing in the reference tree is edited.
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))


def _reference_package_dir():
    cands = [os.environ.get("W2C_REFERENCE_ROOT")] + list(sys.path)
    for root in cands:
        if not root:
            continue
        d = os.path.join(os.path.abspath(root), "ptsemseg")
        if d != _here and os.path.isfile(os.path.join(d, "trainer.py")):
            return d
    return None


_ref = _reference_package_dir()
if _ref is not None and _ref not in __path__:
    __path__.append(_ref)   # submodules this overlay does not define come from the reference
REFERENCE_PACKAGE_DIR = _ref
