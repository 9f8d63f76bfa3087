"""GPU: the backward-pass kernels (wgrad on tcgen05, data-gradient convs on the forward kernels with the
re-indexed weight, train-mode BatchNorm backward, attention / MLP-head backward, first-layer wgrad, pooling and
up-sampling adjoints), each against torch autograd in float64 on the operands as the kernels see them
(tools/gpu_bwd_check.py)."""
import pytest

from tools import gpu_bwd_check as bc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(bc.WGRAD_CASES) + list(bc.DGRAD_CASES) + list(bc.OTHER_CASES))
def test_backward_kernel_case(name, cuda_device):
    r = bc.run_case(name)
    assert r["ok"], r
