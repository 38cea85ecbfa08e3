"""Short workload for ncu: a few training steps at two batch sizes + a few clustering probes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vamb_b200.encode as ve, vamb_b200.cluster as vc
from vamb_b200 import _lib
from oracle import synth

n = int(os.environ.get("N", 200_000))
ab, tnf, lens = synth.make_contigs(n, 50, seed=0)
dl = ve.make_dataloader(ab, tnf, lens, batchsize=256, destroy=True)
vae = ve.VAE(50, seed=0)
vae._net.tc_min_batch = int(os.environ.get("TC_MIN", 1))
vae._bind_dataset(dl.dataset.tensors)
vae.train()
for B in (4096, 256):
    for _ in range(int(os.environ.get("STEPS", 3))):
        ve._lib.check(ve._L.vk_vae_train_step(ve._ct.byref(vae._net), B, None, vae._stream()))
torch.cuda.synchronize()
lat, ln = synth.make_latent(1_000_000, 32, seed=0, spread=0.1)
gen = vc.ClusterGenerator(lat, ln, rng_seed=0, _driver="python")
for i in range(3):
    gen._probe((i * 7919 + 1) % 1_000_000)
torch.cuda.synchronize()
print("done")
