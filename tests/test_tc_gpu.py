"""GPU suite: the tcgen05 3xTF32 tile GEMM (vk_tc.cuh) against an fp64 reference, for all four
operand storage orders and ragged shapes.  Near-fp32 accuracy is the contract: the relative error
of the result matrix stays below 1e-6 + 1.2e-8 * K (measured 4e-7 at K = 8, 4e-6 at K = 512: the
tensor core's fp32 accumulator truncates, so the error grows linearly with the chain length;
plain TF32 would sit at ~5e-4 for every K)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 32), (128, 128, 64), (256, 512, 512), (4096, 512, 512), (300, 154, 154), (77, 48, 70),
          (512, 513, 1024), (154, 512, 256), (256, 32, 512), (200, 512, 32)]


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_gemm_matches_fp64(M, N, K, a_mn, b_mn):
    from vamb_b200 import _lib

    _lib.require_device()
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g)
    ref = (A.double() @ B.double().t())
    a_store = A.t().contiguous() if a_mn else A.contiguous()
    b_store = B.t().contiguous() if b_mn else B.contiguous()
    C = torch.full((M, N), float("nan"), device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib.vk_tc_gemm_test(a_store.data_ptr(), b_store.data_ptr(), C.data_ptr(), M, N, K, a_mn, b_mn, s))
    torch.cuda.synchronize()
    assert torch.isfinite(C).all()
    err = float((C.double() - ref).norm() / ref.norm())
    assert err < 1e-6 + 1.2e-8 * K, err
