"""In-kernel time stamps of the tcgen05 forward layer kernel (thread 0 of block 0 of each launch).

Needs the stamped build:  python -m vamb_b200.build --timeline   (writes vamb_b200/_vk_timeline.so)
then                      VK_PDL=0 VAMB_B200_SO=vamb_b200/_vk_timeline.so python tools/kernel_timeline.py
Slots: 0 entry, 1 after the dependency wait, 8 pipeline set up (barriers, tensor memory), 10+kt k-tile kt produced,
9 all produced, 2 accumulator complete, 3 tile in shared memory, 44 tensor memory released, 4 epilogue pass done
(bias / LeakyReLU / dropout, global store), 5 column sums done, 6 grid barrier passed, 45 BatchNorm constants folded,
7 staging passes done.
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
for B in [int(b) for b in os.environ.get("BATCHES", "256,1024,4096").split(",")]:
    vae._run_steps(B, 256)
    torch.cuda.synchronize()
    vae._run_steps(B, 128)
    rd(buf.ctypes.data)
    gt = buf[:2048].astype(np.int64).reshape(32, 64)
    ck = buf[2048:].astype(np.int64).reshape(32, 64)
    print(f"== B={B}: forward layers 0-5, phase durations from the SM clock [us at 1.965 GHz]")
    for k in range(6):
        if gt[k, 0] == 0:
            continue
        c = ck[k]
        def d(a, b):
            return (c[b] - c[a]) / 1965.0 if c[a] and c[b] else float("nan")
        nk = int(np.sum(c[10:30] != 0))
        kt = [round(float(c[10 + i] - (c[10 + i - 1] if i else c[8])) / 1965.0, 2) for i in range(nk)]
        print(f"  L{k}: wait {d(0, 1):5.2f} setup {d(1, 8):5.2f} produce {d(8, 9):6.2f} drain {d(9, 2):5.2f} acc->tile {d(2, 3):5.2f} "
              f"tmem free {d(3, 44):5.2f} epilogue pass {d(44, 4):5.2f} colsum {d(4, 5):5.2f} grid barrier {d(5, 6):5.2f} "
              f"bn fold {d(6, 45):5.2f} staging {d(45, 7):5.2f} | total {d(0, 7):6.2f} | k-tiles {kt}")
    print(f"   backward kernels (launch order: layer 5 .. 0), FIRST DGRAD CTA, [us]: main loop = entry -> accumulator complete, then the epilogue phases")
    for k in range(19, 26):  # kernel id = 8 + ticket_id = 19 + layer
        c = ck[k]
        if c[1] == 0:
            continue
        def d(a, b):
            return (c[b] - c[a]) / 1965.0 if c[a] and c[b] else float("nan")
        print(f"  layer {k - 19}: main loop {d(1, 2):6.2f} P tile + acc->tile {d(2, 3):5.2f} pass (store dX) {d(3, 4):5.2f} colsum {d(4, 5):5.2f} "
              f"grid barrier {d(5, 6):5.2f} fold {d(6, 45):5.2f} dY pass {d(45, 46):5.2f} staging {d(46, 7):5.2f} | total {d(1, 7):6.2f}")
