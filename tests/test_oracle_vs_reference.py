"""CPU suite: the torch-fp32 VAE oracle is PINNED -- it reproduces the reference-made golden fixtures bit for bit
(tests/golden/vae_*.npz, written by oracle/make_golden_vae.py from the unmodified reference), and, where the
reference tree is present (build container), the live reference itself.  Also pins oracle/normalize.py."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import vae_oracle as vo
from oracle.make_golden_vae import VAE_CASES, vae_inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unpack_keep(packed, batch, n):
    return torch.from_numpy(np.unpackbits(packed)[: batch * n].reshape(batch, n).astype(np.float32))


@pytest.mark.parametrize("case", VAE_CASES, ids=[c[0] for c in VAE_CASES])
def test_vae_oracle_reproduces_reference_goldens_bitwise(case):
    """Same batches, same recorded noise -> the reference's losses, d, parameters and latent, exactly."""
    from oracle import normalize as onorm

    name, S, nh, nl, dp, n, batch, nsteps, seed = case
    g = np.load(os.path.join(GOLDEN, f"vae_{name}.npz"))
    rpkm, tnfs, lens = vae_inputs(S, n, seed)
    d, t, a, w = (torch.from_numpy(x) for x in onorm.normalize(rpkm, tnfs, lens))
    o = vo.OracleVAE(S, nhiddens=nh, nlatent=nl, dropout=dp, seed=seed)
    hidden = o.nhiddens + o.nhiddens[::-1]
    for step in range(nsteps):
        idx = torch.from_numpy(g["batch_idx"][step])
        keeps = None
        if o.dropout > 0:
            keeps = [unpack_keep(g[f"keep{li}"][step], batch, hidden[li]) for li in range(len(hidden))]
        lo, _, _ = o.train_step(d[idx], t[idx], a[idx], w[idx], eps=torch.from_numpy(g["eps"][step]), keeps=keeps)
        assert lo == list(g["losses"][step]), (step, lo, g["losses"][step])
        assert o.d == float(g["d"][step])
    L = len(o.nhiddens)
    assert np.array_equal(o.state["mu.weight"].numpy(), g["mu_weight"])
    assert np.array_equal(o.state["outputlayer.bias"].numpy(), g["out_bias"])
    assert np.array_equal(o.state["encoderlayers.0.weight"].numpy()[:8], g["enc0_weight_head"])
    assert np.array_equal(o.state["encodernorms.0.running_mean"].numpy(), g["bn0_running_mean"])
    assert np.array_equal(o.state["encodernorms.0.running_var"].numpy(), g["bn0_running_var"])
    assert np.array_equal(o.state[f"decodernorms.{L - 1}.weight"].numpy(), g["bn_last_weight"])
    latent, _ = o.encode(d, t, a, batch=batch)
    assert np.array_equal(latent, g["latent"])


def test_normalize_restatement_matches_the_product_dataloader():
    """oracle/normalize.py (used by the reference arm of bench.py) == vamb_b200.encode.make_dataloader tensors."""
    pytest.importorskip("vamb_b200._lib")  # needs the built extension to import the package (no device needed)
    import vamb_b200.encode as ve
    from oracle import normalize as onorm
    from oracle import synth

    ab, tnf, lens = synth.make_contigs(3000, 5, seed=4)
    ab[7] = 0.0  # an all-zero abundance row takes the 1/S branch (vamb/encode.py:108-113)
    want = ve.make_dataloader(ab.copy(), tnf.copy(), lens).dataset.tensors
    got = onorm.normalize(ab, tnf, lens)
    for x, y in zip(got, want):
        assert np.array_equal(x, y.numpy())


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_normalize_restatement_matches_the_live_reference():
    from oracle import normalize as onorm
    from oracle import synth

    ref = ref_loader.load()
    ab, tnf, lens = synth.make_contigs(2000, 3, seed=5)
    want = ref.encode.make_dataloader(ab.copy(), tnf.copy(), lens).dataset.tensors
    for x, y in zip(onorm.normalize(ab, tnf, lens), want):
        assert np.array_equal(x, y.numpy())


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_vae_oracle_equals_live_reference_over_training_steps():
    """Under the same ``torch.manual_seed`` stream the oracle and the unmodified reference stay bit-identical
    (losses, every parameter and buffer, encode) over 6 optimiser steps."""
    from oracle import dadapt

    ref = ref_loader.load()
    S, n, batch, seed = 5, 400, 64, 11
    rpkm, tnfs, lens = vae_inputs(S, n, seed)
    dl = ref.encode.make_dataloader(rpkm.copy(), tnfs.copy(), lens, batchsize=batch)
    d, t, a, w = dl.dataset.tensors
    vae = ref.encode.VAE(S, nhiddens=[64, 48], nlatent=12, seed=seed)
    o = vo.OracleVAE(S, nhiddens=[64, 48], nlatent=12, seed=seed)
    opt = dadapt.DAdaptAdam(vae.parameters(), decouple=True)
    vae.train()
    rng = np.random.default_rng(seed)
    for step in range(6):
        idx = torch.from_numpy(rng.choice(n, size=batch, replace=False).astype(np.int64))
        torch.manual_seed(2000 + step)
        opt.zero_grad()
        do, to, ao, mu = vae(d[idx], t[idx], a[idx])
        L = vae.calc_loss(d[idx], do, t[idx], to, a[idx], ao, mu, w[idx])
        L[0].backward()
        opt.step()
        torch.manual_seed(2000 + step)
        lo, _, _ = o.train_step(d[idx], t[idx], a[idx], w[idx])
        assert [float(x.detach()) for x in L] == lo
    for k, v in vae.state_dict().items():
        assert torch.equal(v, o.state[k]), k
    vae.eval()
    assert np.array_equal(vae.encode(dl), o.encode(d, t, a, batch=batch)[0])


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_strict_rng_emulation_follows_the_live_reference_trainmodel():
    """``vamb_b200.encode.reference_epoch_noise`` (the strict-RNG parity mode of the CUDA path) consumes torch's global
    generator exactly like the reference's DataLoader + forward pass: driving the (bit-exact) oracle with it reproduces
    the per-epoch losses that the unmodified ``VAE.trainmodel`` logs, including a batch-size doubling."""
    import re

    pytest.importorskip("vamb_b200._lib")
    from loguru import logger

    import vamb_b200.encode as ve
    from oracle import normalize as onorm
    from oracle import synth

    ref = ref_loader.load()
    S, n, seed, nepochs, bsteps = 4, 1500, 3, 4, [2]
    ab, tnf, lens = synth.make_contigs(n, S, seed=8)
    rows = []
    pat = re.compile(r"Epoch:\s*(\d+)\s+Loss:\s*(\S+)")
    logger.enable("vamb")
    sink = logger.add(lambda m: rows.append(pat.search(str(m))), level="INFO")
    dl = ref.encode.make_dataloader(ab.copy(), tnf.copy(), lens, batchsize=128)
    vae = ref.encode.VAE(S, nhiddens=[64, 32], nlatent=8, seed=seed)
    vae.trainmodel(dl, nepochs=nepochs, batchsteps=bsteps)
    logger.remove(sink)
    logger.disable("vamb")
    want = [float(m.group(2)) for m in rows if m]
    assert len(want) == nepochs

    d, t, a, w = (torch.from_numpy(x) for x in onorm.normalize(ab, tnf, lens))
    o = vo.OracleVAE(S, nhiddens=[64, 32], nlatent=8, seed=seed)  # torch.manual_seed(seed) + the reference's init draws
    batch, got = 128, []
    for epoch in range(nepochs):
        if epoch in bsteps:
            batch *= 2
        tot, k = 0.0, 0
        for idx, eps, keeps in ve.reference_epoch_noise(n, batch, n > batch, o.nhiddens, o.nlatent, o.dropout):
            lo, _, _ = o.train_step(d[idx], t[idx], a[idx], w[idx], eps=eps, keeps=keeps)
            tot += lo[0]
            k += 1
        got.append(tot / k)
    assert [f"{x:.5e}" for x in got] == [f"{x:.5e}" for x in want], (got, want)
