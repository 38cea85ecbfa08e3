"""Probe kernel timing: CUDA events around the single launch the native driver issues (vk_probe_mapped);
cold = L2 flushed before every launch.  VK_PROBE_R (4 | 8 rows in flight per lane) and VK_PROBE_BPS (persistent blocks
per SM) select the launch shape, VK_PROBE_UNIT / VK_PROBE_DYNAMIC the work distribution; run once per setting (they are read once per process)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vamb_b200.cluster as vc
from oracle import synth

for n in [int(x) for x in os.environ.get("PROBE_N", "1000000,5000000").split(",")]:
    lat, ln = synth.make_latent(n, 32, seed=0, spread=0.1)
    gen = vc.ClusterGenerator(lat, ln, rng_seed=0, _driver="python")
    state = {}
    call = lambda i: gen._probe_mapped_once((i * 7919) % n, state)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    for mode in ("warm", "cold"):
        for i in range(3):
            call(i)
        ts = []
        for i in range(12):
            if mode == "cold":
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); call(i + 3); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[mode] = float(np.median(ts))
    nbytes = n * 133
    print(f"VK_PROBE_R={os.environ.get('VK_PROBE_R', '4')} VK_PROBE_BPS={os.environ.get('VK_PROBE_BPS', '4')} "
          f"UNIT={os.environ.get('VK_PROBE_UNIT', '1')} DYNAMIC={os.environ.get('VK_PROBE_DYNAMIC', '0')} N={n}: "
          f"warm {res['warm']*1e3:.1f} us ({nbytes/res['warm']/1e6:.0f} GB/s), cold {res['cold']*1e3:.1f} us "
          f"({nbytes/res['cold']/1e6:.0f} GB/s)   [one launch]")
    del gen, lat
