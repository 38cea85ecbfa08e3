"""Launch / prologue overhead probes and the issue rate of back-to-back tcgen05.mma (vk_tc_test.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vamb_b200 import _lib
_lib.require_device()
s = torch.cuda.current_stream().cuda_stream
def timeit(fn, iters=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
x = torch.zeros(1, device="cuda")
print(f"empty torch kernel (x.add_): {timeit(lambda: x.add_(1)):.1f} us per launch")
lib = _lib.lib
import ctypes
lib.vk_tc_mma_rate_test.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for sw in (0, 2, 3):
    for n in (16, 32, 64) if sw == 3 else (16, 32, 64, 128):
        for n_mma in (1, 12, 48, 768):
            _lib.check(lib.vk_tc_mma_rate_test(n_mma, n, sw, out.data_ptr(), s)); torch.cuda.synchronize()
            _lib.check(lib.vk_tc_mma_rate_test(n_mma, n, sw, out.data_ptr(), s)); torch.cuda.synchronize()
            o = out.tolist()
            print(f"mma rate swizzle={sw} N={n} n_mma={n_mma}: issue {o[0]} cyc ({o[0]/n_mma:.1f}/mma), complete {o[1]} cyc ({o[1]/n_mma:.1f}/mma)")
if os.environ.get("MMA_ONLY"): sys.exit(0)
cnt = torch.zeros(4, dtype=torch.int32, device="cuda")
lib = _lib.lib
import ctypes
lib.vk_tc_overhead_test.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
lib.vk_tc_overhead_test.restype = ctypes.c_int
for smem in (70 * 1024, 197 * 1024):
    for mode in (0, 1, 2, 3, 4, 5):
        for grid in (32, 128):
            t = timeit(lambda: _lib.check(lib.vk_tc_overhead_test(mode, grid, smem, cnt.data_ptr(), s)))
            print(f"overhead kernel smem={smem//1024}KB mode={mode} grid={grid}: {t:.1f} us per launch")
