"""tcgen05 tile-engine self-test: one 128x128x128 GEMM through TMEM in the three operand-major / operand-format
combinations the kernels use (forward, weight-gradient, data-gradient), fp16 operands, fp32 accumulation."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_tc_engine_gemm(mode):
    from pdivgnn_b200 import _lib
    L = _lib.lib()
    torch.manual_seed(mode)
    A = torch.randn(128, 128, device="cuda")
    B = torch.randn(128, 128, device="cuda")
    D = torch.zeros(128, 128, device="cuda")
    img = torch.zeros(32768, dtype=torch.uint8, device="cuda")
    _lib.check(L.pdg_tc_selftest(mode, _lib.ptr(A), _lib.ptr(B), _lib.ptr(D), _lib.ptr(img), _lib.stream_ptr()), "selftest")
    torch.cuda.synchronize()
    a, b = A.half().double(), B.half().double()  # every operand tile is fp16 (pdg_tc.cuh)
    ref = a @ b.t() if mode == 0 else (a.t() @ b if mode == 1 else a @ b)
    err = (D.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, err
