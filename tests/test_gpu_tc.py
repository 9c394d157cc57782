"""tcgen05 / TMEM plumbing self-test: the three operand arrangements used by the tensor-core head
kernels, bf16x3 split (fp32-class accuracy) against a float64 matmul."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(pkg, mode, A, B, shape):
    C = torch.full(shape, float("nan"), device=DEV)
    stream = torch.cuda.current_stream().cuda_stream
    rc = pkg.LIB.rec_debug_tc_gemm(mode, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()),
                                   ctypes.c_void_p(C.data_ptr()), ctypes.c_void_p(stream))
    assert rc == 0
    torch.cuda.synchronize()
    return C


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_tcgen05_gemm_modes(pkg, mode):
    torch.manual_seed(mode)
    if mode == 0:
        A, B = torch.randn(128, 64, device=DEV), torch.randn(128, 64, device=DEV)
        want = A.double() @ B.double().T
        got = _run(pkg, 0, A, B, (128, 128))
    elif mode == 1:
        P, Q = torch.randn(128, 128, device=DEV), torch.randn(128, 64, device=DEV)
        want = P.double().T @ Q.double()
        got = _run(pkg, 1, P, Q, (128, 64))
    else:
        P, R = torch.randn(128, 128, device=DEV), torch.randn(128, 64, device=DEV)
        want = P.double() @ R.double()
        got = _run(pkg, 2, P, R, (128, 64))
    err = (got.double() - want).abs().max().item()
    scale = want.abs().max().item()
    assert err <= 2e-4 * scale, (mode, err, scale)  # bf16x3: ~1e-5 relative; single bf16 would be ~1e-2
