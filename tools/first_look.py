"""Ad-hoc GPU measurements used while developing (not the bench of record)."""
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from vamb_b200 import _lib
from oracle import synth
import vamb_b200.cluster as vc


def time_kernel(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def kernels(n, d=32):
    g = torch.Generator(device="cuda").manual_seed(0)
    m = torch.randn(n, d, device="cuda", generator=g)
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib.vk_normalize_rows(m.data_ptr(), n, d, s))
    lens = torch.randint(2000, 100000, (n,), device="cuda").float()
    kept = torch.ones(n, dtype=torch.uint8, device="cuda")
    out = torch.empty(n, device="cuda")
    hdr = torch.zeros(_lib.HDR_SIZE, dtype=torch.uint8, device="cuda")
    over = torch.empty(n, dtype=torch.int32, device="cuda")
    nl_rows = torch.empty(n, dtype=torch.int32, device="cuda")
    nl_d = torch.empty(n, dtype=torch.float32, device="cuda")
    edges = torch.from_numpy(vc._histogram_edges()).cuda()
    cnt = [0]

    def dist():
        cnt[0] += 1
        _lib.check(_lib.lib.vk_distances(m.data_ptr(), n, d, (cnt[0] * 7919) % n, out.data_ptr(), s))

    def probe():
        cnt[0] += 1
        _lib.check(_lib.lib.vk_probe(m.data_ptr(), lens.data_ptr(), kept.data_ptr(), n, d, (cnt[0] * 7919) % n, 0.3,
                                     edges.data_ptr(), hdr.data_ptr(), over.data_ptr(), nl_rows.data_ptr(),
                                     nl_d.data_ptr(), s))

    def norm():
        _lib.check(_lib.lib.vk_normalize_rows(m.data_ptr(), n, d, s))

    for name, fn, nbytes in (("distances", dist, n * (4 * d + 4)), ("probe", probe, n * (4 * d + 5)),
                             ("normalize", norm, 2 * n * 4 * d)):
        med, best = time_kernel(fn)
        print(f"n={n} d={d} {name}: median {med*1e3:.1f} us best {best*1e3:.1f} us -> {nbytes/med/1e6:.0f} GB/s (median), {nbytes/best/1e6:.0f} GB/s (best)")
    h = hdr.cpu().numpy()
    print("  last probe n_nl", h[_lib.HDR_NNL:_lib.HDR_NNL+4].view(np.int32)[0], "n_within", h[_lib.HDR_NWITHIN:_lib.HDR_NWITHIN+4].view(np.int32)[0])


def cluster_e2e(n, spread=0.1, max_clusters=None):
    lat, lens = synth.make_latent(n, 32, 0, spread)
    torch.cuda.synchronize()
    t0 = time.time()
    gen = vc.ClusterGenerator(lat, lens, rng_seed=0)
    t1 = time.time()
    k = 0
    sizes = []
    for c in gen:
        sizes.append(len(c.members))
        k += 1
        if max_clusters and k >= max_clusters:
            break
    t2 = time.time()
    print(f"cluster n={n} spread={spread}: init {t1-t0:.2f}s, {k} clusters ({sum(sizes)} contigs) in {t2-t1:.2f}s; "
          f"probes {gen._n_probes} evals {gen._n_evals}; {1e6*(t2-t1)/max(1,gen._n_probes+gen._n_evals):.0f} us per device round trip")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    kernels(1_000_000)
    kernels(5_000_000)
    kernels(200_000, 283)
    cluster_e2e(100_000)
    cluster_e2e(1_000_000, max_clusters=3000)
