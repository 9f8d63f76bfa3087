"""GPU: every C-ABI kernel entry point against a float64 torch evaluation of the same op (the bring-up cases of
tools/gpu_kernel_check.py, run in-process). Covers all conv kinds in the three tensor-core kernels (one-tile,
persistent incl. row-halo stages / dual epilogue groups / multi-CTA logits path), the activation
storages, slices, residuals, ragged edges, the stems, pooling, up-sampling, the MLP heads and the attention kernel."""
import pytest

from tools import gpu_kernel_check as kc

pytestmark = pytest.mark.gpu

_CONV = list(kc.CONV_CASES)


@pytest.mark.parametrize("name", _CONV)
def test_conv_kernel_case(name, cuda_device):
    kind, n, h, w, cin, cout, act, kw = kc.CONV_CASES[name]
    r = kc._conv_case(kind, n, h, w, cin, cout, act, **kw)
    assert r["ok"], r


@pytest.mark.parametrize("name", ["layout", "stem", "mlp", "attn", "enc_head"])
def test_other_kernels(name, cuda_device):
    r = kc.run_case(name)
    assert r["ok"], r
