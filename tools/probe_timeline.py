"""In-kernel time stamps of the probe kernel (stamped build: python -m vamb_b200.build --timeline).
    VAMB_B200_SO=vamb_b200/_vk_timeline.so python tools/probe_timeline.py
Block 0: entry -> prologue -> scan -> block sums merged -> ticket; last block: elected -> results in pinned memory -> flag."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vamb_b200.cluster as vc
from vamb_b200 import _lib
from oracle import synth

rd = _lib.lib.vk_cluster_timeline_read
rd.argtypes = [ctypes.c_void_p]
buf = np.zeros(4096, dtype=np.uint64)
for n in (1_000_000, 5_000_000):
    lat, ln = synth.make_latent(n, 32, seed=0, spread=0.1)
    gen = vc.ClusterGenerator(lat, ln, rng_seed=0, _driver="python")
    state = {}
    for i in range(6):
        gen._probe_mapped_once((i * 7919 + n // 2) % n, state)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gen._probe_mapped_once((7 * 7919 + n // 2) % n, state); b.record(); torch.cuda.synchronize()
    rd(buf.ctypes.data)
    g = buf[:64].astype(np.int64)
    rel = [(int(g[i]) - int(g[0])) / 1e3 if g[i] else float("nan") for i in range(8)]
    print(f"N={n}: event time {a.elapsed_time(b) * 1e3:.1f} us; global timer [us after block-0 entry]: prologue {rel[1]:.1f}, "
          f"scan done {rel[2]:.1f}, sums merged {rel[3]:.1f}, before ticket {rel[4]:.1f} | last block elected {rel[5]:.1f}, "
          f"pinned writes done {rel[6]:.1f}, flag raised {rel[7]:.1f}")
    del gen, lat
