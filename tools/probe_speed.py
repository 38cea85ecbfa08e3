"""Probe kernel timing (CUDA events around vk_probe; cold = L2 flushed before every launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vamb_b200.cluster as vc
from vamb_b200 import _lib
from oracle import synth

for n in (1_000_000, 5_000_000):
    lat, ln = synth.make_latent(n, 32, seed=0, spread=0.1)
    gen = vc.ClusterGenerator(lat, ln, rng_seed=0, _driver="python")
    s = torch.cuda.current_stream().cuda_stream
    def call(i):
        _lib.check(_lib.lib.vk_probe(gen._m.data_ptr(), gen._len.data_ptr(), gen._kept.data_ptr(), n, gen._d,
                                     (i * 7919) % n, 0.3, gen._edges.data_ptr(), gen._hdr.data_ptr(),
                                     gen._within_over.data_ptr(), gen._nl_rows.data_ptr(), gen._nl_d.data_ptr(), s))
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
    print(f"VK_PROBE_BULK={os.environ.get('VK_PROBE_BULK', '0')} N={n}: warm {res['warm']*1e3:.1f} us ({nbytes/res['warm']/1e6:.0f} GB/s), "
          f"cold {res['cold']*1e3:.1f} us ({nbytes/res['cold']/1e6:.0f} GB/s)   [memset + rank + probe launches]")
    del gen, lat
