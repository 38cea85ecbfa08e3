"""Diagnose the MN-major operand layout of the tcgen05 tile GEMM on a real GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vamb_b200 import _lib

_lib.require_device()
lib = _lib.lib
s = torch.cuda.current_stream().cuda_stream
for variant in (0, 1):
    lib.vk_tc_set_variant(variant)
    for (M, N, K) in [(128, 128, 8), (128, 128, 32), (128, 32, 32), (256, 96, 64)]:
        for a_mn, b_mn in [(0, 0), (0, 1), (1, 0), (1, 1)]:
            g = torch.Generator(device="cuda").manual_seed(1)
            A = torch.randn(M, K, device="cuda", generator=g)
            B = torch.randn(N, K, device="cuda", generator=g)
            ref = A.double() @ B.double().t()
            a_s = A.t().contiguous() if a_mn else A
            b_s = B.t().contiguous() if b_mn else B
            C = torch.full((M, N), float("nan"), device="cuda")
            _lib.check(lib.vk_tc_gemm_test(a_s.data_ptr(), b_s.data_ptr(), C.data_ptr(), M, N, K, a_mn, b_mn, s))
            torch.cuda.synchronize()
            err = float((C.double() - ref).norm() / ref.norm())
            print(f"variant {variant} M{M} N{N} K{K} a_mn={a_mn} b_mn={b_mn}: rel err {err:.3e}")
