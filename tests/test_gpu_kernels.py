"""Every hand-written sm_100a kernel against a plain PyTorch fp32 reference of the same
op.  The check bodies live in utils/selfcheck.py so that
tools/gpu_check.py and __graft_entry__.smoke() run exactly the same code."""
import pytest
import torch

from comfyui_parallelanything_b200 import ops
from comfyui_parallelanything_b200.utils import selfcheck

pytestmark = pytest.mark.gpu


def test_native_library_is_loaded():
    assert ops.available(), f"native library missing: {ops.load_error()!r}"
    assert ops.native_ok("cuda:0"), "expected an sm_100 device"


@pytest.mark.parametrize("name", sorted(selfcheck.CHECKS))
def test_kernel(name):
    r = selfcheck.CHECKS[name]()
    torch.cuda.synchronize()
    assert r["ok"], r
