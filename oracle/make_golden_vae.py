"""TEST INFRASTRUCTURE -- VAE golden fixtures from the UNMODIFIED reference.

For each case the reference ``vamb.encode.VAE`` (with the restated DAdaptAdam shim) takes
``nsteps`` optimiser steps on explicit minibatches under ``torch.manual_seed(1000 + step)``.
The oracle (oracle/vae_oracle.py) is run on the same seeds; it draws the same dropout masks
and reparameterisation noise (asserted bit-identical results), which is how the noise that
the reference consumed gets recorded.  Stored: batch indices, noise, packed keep-masks, the
reference's five loss scalars per step, d after every step, selected final tensors and the
final eval-mode latent.
"""
import os

import numpy as np
import torch

VAE_CASES = [
    # name, nsamples, nhiddens, nlatent, dropout, N, batch, nsteps, seed
    ("v_s6_small", 6, [72, 40], 16, 0.2, 200, 48, 4, 3),
    ("v_s1_single", 1, None, 8, None, 150, 64, 3, 4),
    ("v_s50_default", 50, None, 32, 0.2, 300, 128, 3, 0),
]


def vae_inputs(nsamples, n, seed):
    rng = np.random.RandomState(100 + seed)
    tnfs = rng.random((n, 103)).astype(np.float32)
    rpkm = (rng.random((n, nsamples)).astype(np.float32) + 0.01)
    lens = rng.randint(2000, 50000, n)
    return rpkm, tnfs, lens


def make(ref, golden_dir):
    from oracle import dadapt, vae_oracle as vo

    for name, S, nh, nl, dp, n, batch, nsteps, seed in VAE_CASES:
        rpkm, tnfs, lens = vae_inputs(S, n, seed)
        dl = ref.encode.make_dataloader(rpkm.copy(), tnfs.copy(), lens, batchsize=batch)
        d, t, a, w = dl.dataset.tensors
        vae = ref.encode.VAE(S, nhiddens=nh, nlatent=nl, dropout=dp, seed=seed)
        o = vo.OracleVAE(S, nhiddens=nh, nlatent=nl, dropout=dp, seed=seed)
        assert all(torch.equal(v, o.state[k]) for k, v in vae.state_dict().items())
        opt = dadapt.DAdaptAdam(vae.parameters(), decouple=True)
        vae.train()
        g = np.random.default_rng(seed)
        idxs, epss, keeps, losses, ds = [], [], [], [], []
        for step in range(nsteps):
            idx = torch.from_numpy(g.choice(n, size=batch, replace=False).astype(np.int64))
            bd, bt, ba, bw = d[idx], t[idx], a[idx], w[idx]
            torch.manual_seed(1000 + step)
            opt.zero_grad()
            do, to, ao, mu = vae(bd, bt, ba)
            L = vae.calc_loss(bd, do, bt, to, ba, ao, mu, bw)
            L[0].backward()
            opt.step()
            torch.manual_seed(1000 + step)
            lo, eps, kp = o.train_step(bd, bt, ba, bw)
            assert [float(x.detach()) for x in L] == lo, "oracle diverged from the reference"
            idxs.append(idx.numpy())
            epss.append(eps.numpy())
            keeps.append([np.packbits((k if k is not None else torch.ones(batch, 1)).numpy().astype(np.uint8)) for k in kp])
            losses.append(lo)
            ds.append(opt.param_groups[0]["d"])
        sd = vae.state_dict()
        assert all(torch.equal(v, o.state[k]) for k, v in sd.items())
        vae.eval()
        latent = vae.encode(dl)
        out = dict(
            batch_idx=np.stack(idxs), eps=np.stack(epss), losses=np.array(losses, dtype=np.float64),
            d=np.array(ds, dtype=np.float64), latent=latent,
            mu_weight=sd["mu.weight"].numpy(), out_bias=sd["outputlayer.bias"].numpy(),
            enc0_weight_head=sd["encoderlayers.0.weight"].numpy()[:8].copy(),
            bn0_running_mean=sd["encodernorms.0.running_mean"].numpy(),
            bn0_running_var=sd["encodernorms.0.running_var"].numpy(),
            bn_last_weight=sd[f"decodernorms.{len(o.nhiddens) - 1}.weight"].numpy(),
        )
        for li in range(len(keeps[0])):
            out[f"keep{li}"] = np.stack([k[li] for k in keeps])
        np.savez_compressed(os.path.join(golden_dir, f"vae_{name}.npz"), **out)
        print(f"vae_{name}: losses {losses[-1][0]:.5f} d {ds[-1]:.3e}")
