"""Step time of the VAE training path per batch size (graph replay) + per-launch breakdown."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vamb_b200.encode as ve
from oracle import synth

n = int(os.environ.get("N", 1_000_000))
tcmin = int(os.environ.get("TC_MIN", 512))
ab, tnf, lens = synth.make_contigs(n, 50, seed=0)
dl = ve.make_dataloader(ab, tnf, lens, batchsize=256, destroy=True)
vae = ve.VAE(50, seed=0)
vae._net.tc_min_batch = tcmin
vae._bind_dataset(dl.dataset.tensors)
vae.train()
for B in (256, 512, 1024, 2048, 4096):
    for _ in range(3):
        vae._profile_step(B)
    reps = [vae._profile_step(B) for _ in range(5)]
    fwd = np.mean([r["fwd"] for r in reps], axis=0) * 1e3
    bwd = np.mean([r["bwd"] for r in reps], axis=0) * 1e3
    oth = np.mean([[r["batch_rows"], r["loss"], r["dadapt"], r["prep"]] for r in reps], axis=0) * 1e3
    nsteps = 640
    torch.cuda.synchronize()
    vae._run_steps(B, nsteps)  # includes capture the first time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    vae._run_steps(B, nsteps)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / nsteps * 1e6
    print(f"B={B:5d} tc_min={tcmin}: {dt:7.1f} us/step (graph) | sum of launches {fwd.sum()+bwd.sum()+oth.sum():7.1f} us | "
          f"fwd {np.round(fwd,1).tolist()} bwd {np.round(bwd,1).tolist()} rows/loss/opt/prep {np.round(oth,1).tolist()} launches {reps[0]['n_launches']}")
