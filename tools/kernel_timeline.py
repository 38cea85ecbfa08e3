"""In-kernel time stamps of the tcgen05 layer kernels (block 0 of each launch).

Needs the stamped build:  python -m vamb_b200.build --timeline   (writes vamb_b200/_vk_timeline.so)
then                      VAMB_B200_SO=vamb_b200/_vk_timeline.so python tools/kernel_timeline.py
Slots: 0 entry, 1 after the dependency wait, 8 pipeline set up, 10+kt after k-tile kt was issued,
9 all issued, 2 accumulator complete, 3 tile in shared memory, 4 tile stored, 5 column sums done,
6 last block elected, 7 BatchNorm finalised.
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vamb_b200.encode as ve
from vamb_b200 import _lib
from oracle import synth

n = int(os.environ.get("N", 200_000))
ab, tnf, lens = synth.make_contigs(n, 50, seed=0)
dl = ve.make_dataloader(ab, tnf, lens, batchsize=256, destroy=True)
vae = ve.VAE(50, seed=0)
vae._net.tc_min_batch = 128
vae._bind_dataset(dl.dataset.tensors)
vae.train()
rd = _lib.lib.vk_timeline_read
rd.argtypes = [ctypes.c_void_p]
buf = np.zeros(4096, dtype=np.uint64)
for B in [int(b) for b in os.environ.get("BATCHES", "256,2048").split(",")]:
    vae._run_steps(B, 256)
    torch.cuda.synchronize()
    vae._run_steps(B, 128)
    rd(buf.ctypes.data)
    gt = buf[:2048].astype(np.int64).reshape(32, 64)
    ck = buf[2048:].astype(np.int64).reshape(32, 64)
    t_first = min(gt[k, 0] for k in range(14) if gt[k, 0] > 0)
    print(f"== B={B}: per kernel (id 0-5 forward layer, 8-13 backward ticket): start offset in step [us], then phase "
          f"durations from SM clock [us at 1.9 GHz]")
    for k in list(range(6)) + list(range(8, 14)):
        if gt[k, 0] == 0:
            continue
        c = ck[k]
        def d(a, b):
            return (c[b] - c[a]) / 1900.0 if c[a] and c[b] else float("nan")
        nk = int(np.sum(c[10:30] != 0))
        kt = [round(float(c[10 + i] - (c[10 + i - 1] if i else c[8])) / 1900.0, 2) for i in range(nk)]
        print(f"  k{k:2d} start {1e-3 * (gt[k, 0] - t_first):8.1f} | wait {d(0, 1):5.2f} setup {d(1, 8):5.2f} issue {d(8, 9):6.2f} "
              f"drain {d(9, 2):5.2f} regs->tile {d(2, 3):5.2f} store {d(3, 4):5.2f} colsum {d(4, 5):5.2f} "
              f"elect {d(5, 6):5.2f} finalize {d(6, 7):5.2f} | k-tiles {kt}")
        print(f"        k-tile 6: mma(kt-2) wait {d(13 + 2, 32):5.2f} issue copies {d(32, 33):5.2f} wait copies {d(33, 34):5.2f} split {d(34, 35):5.2f} "
              f"proxy fence {d(35, 36):5.2f} barrier {d(36, 37):5.2f} mma issue {d(37, 38):5.2f} || epilogue: first tmem load {d(2, 41):5.2f} "
              f"first float4 {d(41, 42):5.2f} next three {d(42, 43):5.2f}")
