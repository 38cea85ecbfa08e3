"""Short workload for ncu: a few training steps at three batch sizes, then clustering probes / candidate evaluations at
N = 1,000,000 (the native driver's own launches)."""
import os, sys
from itertools import islice
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vamb_b200.encode as ve, vamb_b200.cluster as vc
from oracle import synth

n = int(os.environ.get("N", 200_000))
ab, tnf, lens = synth.make_contigs(n, 50, seed=0)
dl = ve.make_dataloader(ab, tnf, lens, batchsize=256, destroy=True)
vae = ve.VAE(50, seed=0)
vae._net.tc_min_batch = int(os.environ.get("TC_MIN", 1))
vae._bind_dataset(dl.dataset.tensors)
vae.train()
for B in (4096, 1024, 256):
    for _ in range(int(os.environ.get("STEPS", 3))):
        ve._lib.check(ve._L.vk_vae_train_step(ve._ct.byref(vae._net), B, None, vae._stream()))
torch.cuda.synchronize()
lat, ln = synth.make_latent(1_000_000, 32, seed=0, spread=0.1)
gen = vc.ClusterGenerator(lat, ln, windowsize=300, minsuccesses=15, rng_seed=0)
n_clusters = sum(1 for _ in islice(gen, int(os.environ.get("CLUSTERS", 6))))
torch.cuda.synchronize()
print("done", n_clusters, gen._n_probes, gen._n_evals)
