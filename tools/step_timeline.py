"""Launch-by-launch timeline of one training step inside a replayed CUDA graph (stamped build).

    python -m vamb_b200.build --timeline
    VAMB_B200_SO=vamb_b200/_vk_timeline.so python tools/step_timeline.py

Each launch records the global timer when block 0 starts and when block 0 is (nearly) done; the gaps
between one launch's end and the next one's start are dependency / launch latency.
kind: 1 batch rows, 10-13 prep (mode 0 copy, 1 BatchNorm on load, 2 dL/dY, 3 gather), 14 weight prep,
20+j forward layer j, 30 loss, 40+t backward (ticket t), 50 optimiser.
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
rd = _lib.lib.vk_timeline_kernels
rd.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
buf = np.zeros(4096 * 4, dtype=np.uint64)
cnt = np.zeros(1, dtype=np.uint32)
names = {1: "rows", 10: "prep copy", 11: "prep bn", 12: "prep dY", 13: "prep gather", 14: "prep W", 30: "loss", 50: "dadapt"}
for B in [int(b) for b in os.environ.get("BATCHES", "256,2048").split(",")]:
    vae._run_steps(B, 256)
    torch.cuda.synchronize()
    rd(buf.ctypes.data, cnt.ctypes.data, 1)
    vae._run_steps(B, 128)
    rd(buf.ctypes.data, cnt.ctypes.data, 1)
    rec = buf.reshape(4096, 4).astype(np.int64)[: int(cnt[0])]
    starts = np.flatnonzero(rec[:, 0] == 1)
    segs = [rec[a:b] for a, b in zip(starts[:-1], starts[1:])]
    lens = np.array([len(x) for x in segs])
    per = int(np.bincount(lens).argmax())
    ref = next(x for x in segs if len(x) == per)[:, 0]
    segs = [x for x in segs if len(x) == per and (x[:, 0] == ref).all()]
    print("launch kinds of one step:", ref.tolist(), f"({len(segs)} of {len(lens)} steps follow it)")
    steps = len(segs)
    k = np.stack([x[:, 0] for x in segs]); st = np.stack([x[:, 1] for x in segs]); en = np.stack([x[:, 2] for x in segs])
    sel = slice(4, steps - 1)
    dur = (en - st)[sel].mean(axis=0) / 1e3
    gap = (st[sel, 1:] - en[sel, :-1]).mean(axis=0) / 1e3
    step_len = float(np.median(np.diff(rec[starts, 1]))) / 1e3
    print(f"== B={B}: {steps} steps recorded, {step_len:.1f} us per step")
    t = 0.0
    for i in range(per):
        kind = int(k[0, i])
        nm = names.get(kind, f"fwd L{kind - 20}" if 20 <= kind < 30 else f"bwd t{kind - 40}")
        g = gap[i - 1] if i else float("nan")
        print(f"  {i:2d} {nm:12s} gap before {g:6.2f} us   block-0 busy {dur[i]:6.2f} us")
    print(f"  sum of block-0 busy {dur.sum():.1f} us, sum of gaps {gap.sum():.1f} us")
