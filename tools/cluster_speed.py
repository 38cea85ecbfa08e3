"""Clustering phase alone at C2 scale: the latent of a briefly trained VAE (1M x 32), clustered to exhaustion with the
native driver; prints the host-time breakdown per call kind, with scan-free moves on (default) and off."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vamb_b200.encode as ve, vamb_b200.cluster as vc
from oracle import synth

n = int(os.environ.get("N", 1_000_000))
ab, tnf, lens = synth.make_contigs(n, 50, seed=0)
dl = ve.make_dataloader(ab, tnf, lens, batchsize=256, destroy=True)
vae = ve.VAE(50, seed=0)
vae.trainmodel(dl, nepochs=int(os.environ.get("EPOCHS", 12)), batchsteps=[2, 4, 6, 8])
latent = vae.encode(dl)
for lazy in ("1", "0"):
    os.environ["VAMB_B200_CLUSTER_LAZY"] = lazy
    gen = vc.ClusterGenerator(latent.copy(), lens, windowsize=300, minsuccesses=15, rng_seed=0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    k = m = 0
    for blk in gen.iter_blocks(4096):
        k += len(blk)
        m += len(blk.members)
    dt = time.perf_counter() - t0
    t = gen._timing()
    ne = max(1, gen._n_evals)
    print(f"lazy={lazy}: {dt:.2f} s, {k} clusters / {m} contigs; probes {gen._n_probes} ({t['probe']:.2f} s, "
          f"{1e6 * t['probe'] / max(1, gen._n_probes):.1f} us each), evals {gen._n_evals} ({t['eval']:.2f} s, "
          f"{1e6 * t['eval'] / ne:.1f} us each, mean neighbour list {t['sum_nl_per_eval'] / ne:.0f} rows), select {t['select']:.2f} s, "
          f"pack {t['pack']:.2f} s, scan-free moves {t['lazy_moves']:.0f}, rebases {t['rebases']:.0f}", flush=True)
