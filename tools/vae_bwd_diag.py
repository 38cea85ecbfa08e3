"""Pinpoint where the backward pass departs from the torch oracle: per-layer P, dH, BN statistics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
import vamb_b200.encode as ve
from oracle import vae_oracle as vo
from oracle.make_golden_vae import vae_inputs

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tc = int(sys.argv[2]) if len(sys.argv) > 2 else 0
S, n = 50, 5000
rpkm, tnfs, lens = vae_inputs(S, n, 7)
dl = ve.make_dataloader(rpkm.copy(), tnfs.copy(), lens, batchsize=B)
d, t, a, w = dl.dataset.tensors
idx = torch.from_numpy(np.random.default_rng(0).choice(n, B, replace=False))
o = vo.OracleVAE(S, seed=2)
st = {k: v.double().clone().requires_grad_(not k.endswith(("running_mean", "running_var", "num_batches_tracked")) ) if v.dtype == torch.float32 else v.clone() for k, v in o.state.items()}
torch.manual_seed(5)
# noise drawn as the fp32 oracle would
_, _, eps, keeps = vo.OracleVAE(S, seed=2).grads(d[idx], t[idx], a[idx], w[idx])
x = torch.cat((d[idx], t[idx], a[idx]), 1).double()
Ps, Hs = [], []
names = [("encoderlayers.0", "encodernorms.0"), ("encoderlayers.1", "encodernorms.1"), None,
         ("decoderlayers.0", "decodernorms.0"), ("decoderlayers.1", "decodernorms.1")]
h = x
ki = 0
for item in names:
    if item is None:
        mu = F.linear(h, st["mu.weight"], st["mu.bias"]); mu.retain_grad()
        h = mu + eps.double()
        continue
    lin, bn = item
    y = F.linear(h, st[lin + ".weight"], st[lin + ".bias"])
    p = F.leaky_relu(y, 0.01) * (keeps[ki].double() / 0.8); ki += 1
    p.retain_grad(); Ps.append(p)
    h = F.batch_norm(p, None, None, st[bn + ".weight"], st[bn + ".bias"], True, 0.1, 1e-5)
    h.retain_grad(); Hs.append(h)
rec = F.linear(h, st["outputlayer.weight"], st["outputlayer.bias"])
do = F.softmax(rec[:, :S], dim=1)
loss = vo.calc_loss(d[idx].double(), do, t[idx].double(), rec[:, S:S + 103], a[idx].double(), rec[:, S + 103:], mu, w[idx].double().reshape(-1, 1), S, 32, o.alpha, o.beta)[0]
loss.backward()

vae = ve.VAE(S, seed=2)
vae._net.tc_min_batch = 1 if tc else 0
vae._step_injected(dl.dataset.tensors, idx.numpy(), eps.numpy(), [k.numpy() for k in keeps], optimize=False)
K = vae._keep
print(f"B={B} tc={tc}")
for li, j in enumerate([0, 1, 3, 4]):
    P = K[f"act{j}"][:B].cpu().numpy(); dH = K[f"dact{j}"][:B].cpu().numpy()
    Pr = Ps[li].detach().numpy(); dHr = Hs[li].grad.numpy(); dPr = Ps[li].grad.numpy()
    mean = Pr.mean(0); var = Pr.var(0); rstd = 1 / np.sqrt(var + 1e-5); Ph = (Pr - mean) * rstd
    m1r = dHr.mean(0); m2r = (dHr * Ph).mean(0)
    m1 = K[f"bn_m1{j}"].cpu().numpy(); m2 = K[f"bn_m2{j}"].cpu().numpy()
    gm = K[f"bn_mean{j}"].cpu().numpy(); gr = K[f"bn_rstd{j}"].cpu().numpy()
    print(f" layer {j}: P {rel(P, Pr):.2e} dH {rel(dH, dHr):.2e} mean {rel(gm, mean):.2e} rstd {rel(gr, rstd):.2e} m1 {rel(m1, m1r):.2e} m2 {rel(m2, m2r):.2e}"
          f" |dH| {np.abs(dHr).mean():.2e} |m1| {np.abs(m1r).mean():.2e} |m2| {np.abs(m2r).mean():.2e} |dP| {np.abs(dPr).mean():.2e}")
dmu = K["dact2"][:B].cpu().numpy()
print(" dMU", rel(dmu, mu.grad.numpy()))
