"""In-kernel time stamps of the probe kernel (stamped build: python -m vamb_b200.build --timeline).
    VAMB_B200_SO=vamb_b200/_vk_timeline.so python tools/probe_timeline.py
Block 0: entry -> prologue -> scan -> block sums merged -> ticket; last block: elected -> results in pinned memory -> flag."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vamb_b200.cluster as vc
from vamb_b200 import _lib
from oracle import synth

rd = _lib.lib.vk_cluster_timeline_read
rd.argtypes = [ctypes.c_void_p]
buf = np.zeros(4096, dtype=np.uint64)
for n in (1_000_000, 5_000_000):
    lat, ln = synth.make_latent(n, 32, seed=0, spread=0.1)
    gen = vc.ClusterGenerator(lat, ln, rng_seed=0, _driver="python")
    state = {}
    for i in range(6):
        gen._probe_mapped_once((i * 7919 + n // 2) % n, state)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gen._probe_mapped_once((7 * 7919 + n // 2) % n, state); b.record(); torch.cuda.synchronize()
    rd(buf.ctypes.data)
    g = buf[:64].astype(np.int64)
    rel = [(int(g[i]) - int(g[0])) / 1e3 if g[i] else float("nan") for i in range(8)]
    print(f"N={n}: event time {a.elapsed_time(b) * 1e3:.1f} us; global timer [us after block-0 entry]: prologue {rel[1]:.1f}, "
          f"scan done {rel[2]:.1f}, sums merged {rel[3]:.1f}, before ticket {rel[4]:.1f} | last block elected {rel[5]:.1f}, "
          f"pinned writes done {rel[6]:.1f}, flag raised {rel[7]:.1f}")
    # one candidate evaluation over the neighbour list the probe left behind (kernel id 1 -> slots 64..)
    hdr = state["hdr"].numpy()
    n_nl = int(hdr[_lib.HDR_NNL:_lib.HDR_NNL + 4].view(np.int32)[0])
    n_within = int(hdr[_lib.HDR_NWITHIN:_lib.HDR_NWITHIN + 4].view(np.int32)[0])
    cands = hdr[_lib.HDR_WITHIN:_lib.HDR_WITHIN + 4 * min(n_within, 40)].view(np.int32).tolist()
    C, cap = _lib.VK_LIST_CAND, 1024
    out_dev = torch.zeros(_lib.VK_EVAL_SCRATCH_U64, dtype=torch.int64, device="cuda")
    out_pin = torch.zeros(4 * C, dtype=torch.int64).pin_memory()
    wdev = torch.zeros(_lib.VK_EVAL_SUBS * C * cap, dtype=torch.int32, device="cuda")
    wpin = torch.zeros(C * cap, dtype=torch.int32).pin_memory()
    ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32).pin_memory()
    arr = (_lib.c_int32 * len(cands))(*cands)
    base = (7 * 7919 + n // 2) % n
    for seq in (1, 2, 3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(_lib.lib.vk_eval_candidates_lists(gen._m.data_ptr(), gen._len.data_ptr(), 32, gen._nl_rows.data_ptr(),
                                                     gen._nl_d.data_ptr(), n_nl, 0.2, arr, len(cands), base, 0, 0, out_dev.data_ptr(),
                                                     out_pin.data_ptr(), wdev.data_ptr(), wpin.data_ptr(), cap, ticket.data_ptr(),
                                                     flag.data_ptr(), seq, gen._stream))
        b.record(); torch.cuda.synchronize()
    rd(buf.ctypes.data)
    g = buf[64:128].astype(np.int64)
    rel = [(int(g[i]) - int(g[0])) / 1e3 if g[i] else float("nan") for i in range(8)]
    print(f"   eval of {len(cands)} candidates over a {n_nl}-row list: event time {a.elapsed_time(b) * 1e3:.1f} us; [us after block-0 entry] "
          f"prologue {rel[1]:.1f}, rows done {rel[2]:.1f}, block barrier {rel[3]:.1f}, before ticket {rel[4]:.1f} | last block elected "
          f"{rel[5]:.1f}, published {rel[6]:.1f}, flag raised {rel[7]:.1f}")
    nb = min((n_nl + 31) // 32, 592, 320)
    pt = [(buf[(512 + 320 * i if i < 4 else 2560 + 320 * (i - 4)):][:nb].astype(np.int64) - int(g[0])) / 1e3 for i in range(5)]
    slow = pt[4] > np.median(pt[4]) + 3.0
    q = lambda x: " / ".join(f"{np.percentile(x, p):.1f}" for p in (0, 50, 90, 100)) if len(x) else "-"
    names = ["entry", "scan done (last round)", "ranges reserved", "loop done", "before ticket"]
    print(f"      first {nb} blocks, {int(slow.sum())} of them slow; per point [min / median / 90 % / max us]: " +
          "; ".join(f"{nm} {q(p_)}" for nm, p_ in zip(names, pt)))
    if slow.any():
        print("      slow blocks only: " + "; ".join(f"{nm} {q(p_[slow])}" for nm, p_ in zip(names, pt)))
    del gen, lat
