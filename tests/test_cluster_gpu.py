"""GPU suite: the CUDA clusterer (through the C ABI) against the oracle and the golden
fixtures made by the reference.  Integer / index results are compared bit-exactly."""
import ctypes

import numpy as np
import pytest
import torch

from tests import _util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vk():
    from vamb_b200 import _lib

    _lib.require_device()
    return _lib


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("d", [3, 32, 40, 283])
def test_normalize_and_distances_bit_exact(vk, d):
    from oracle import cluster_oracle as co

    rng = np.random.default_rng(d)
    n = 20011
    m = rng.standard_normal((n, d)).astype(np.float32)
    m[5] = 0.0
    host = m.copy()
    co.normalize(host)
    dm = _dev(m)
    s = torch.cuda.current_stream().cuda_stream
    vk.check(vk.lib.vk_normalize_rows(dm.data_ptr(), n, d, s))
    torch.cuda.synchronize()
    assert np.array_equal(dm.cpu().numpy().view(np.uint32), host.view(np.uint32))
    for idx in (0, 5, n - 1, 777):
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        vk.check(vk.lib.vk_distances(dm.data_ptr(), n, d, idx, out.data_ptr(), s))
        torch.cuda.synchronize()
        ref = co.calc_distances(host, idx)
        assert np.array_equal(out.cpu().numpy().view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("d,n", [(32, 30000), (40, 5000), (283, 3000)])
def test_probe_header_matches_oracle(vk, d, n):
    from oracle import cluster_oracle as co
    from oracle import synth

    lat, lens = synth.make_latent(n, d, seed=d, spread=0.2)
    host = lat.copy()
    co.normalize(host)
    lens32 = lens.astype(np.float32)
    rng = np.random.default_rng(0)
    kept = (rng.random(n) > 0.3).astype(np.uint8)
    dm, dl, dk = _dev(host), _dev(lens32), _dev(kept)
    hdr = torch.zeros(vk.HDR_SIZE, dtype=torch.uint8, device="cuda")
    hdr_host = torch.zeros(vk.HDR_SIZE, dtype=torch.uint8).pin_memory()
    over = torch.empty(n, dtype=torch.int32, device="cuda")
    nl_rows = torch.empty(n, dtype=torch.int32, device="cuda")
    nl_d = torch.empty(n, dtype=torch.float32, device="cuda")
    edges = co.linspace_edges()
    de = _dev(edges)
    s = torch.cuda.current_stream().cuda_stream
    lib = co._load_lib()
    for medoid in np.flatnonzero(kept)[[0, 10, -1]]:
        vk.check(vk.lib.vk_probe_sync(dm.data_ptr(), dl.data_ptr(), dk.data_ptr(), n, d, int(medoid), 0.3,
                                      de.data_ptr(), hdr.data_ptr(), over.data_ptr(), nl_rows.data_ptr(),
                                      nl_d.data_ptr(), hdr_host.data_ptr(), s))
        h = hdr_host.numpy()
        dens = (int(h[8:16].view(np.uint64)[0]) << 12) + int(h[0:8].view(np.uint64)[0])
        hist = h[vk.HDR_HIST:vk.HDR_HIST + 480].view(np.uint64)
        n_within, n_lt, n_nl, rank = (int(x) for x in h[vk.HDR_NWITHIN:vk.HDR_NWITHIN + 16].view(np.int32))
        # oracle on the kept rows only (the reference's CPU path has physically removed the others)
        sel = np.flatnonzero(kept)
        sub = np.ascontiguousarray(host[sel])
        dist = co.calc_distances(sub, int(np.searchsorted(sel, medoid)))
        idx = np.empty(len(sel), dtype=np.int64)
        od = (ctypes.c_uint64 * 2)()
        cnt = lib.ok_sample(co._p(dist, ctypes.c_float), co._p(np.ascontiguousarray(lens32[sel]), ctypes.c_float),
                            len(sel), ctypes.c_float(0.05), co._p(idx, ctypes.c_int64), od)
        oh = np.zeros(60, dtype=np.uint64)
        o_nlt = lib.ok_hist(co._p(dist, ctypes.c_float), co._p(np.ascontiguousarray(lens32[sel]), ctypes.c_float),
                            len(sel), co._p(edges, ctypes.c_float), 60, ctypes.c_float(0.05), co._p(oh, ctypes.c_uint64))
        assert n_within == cnt and dens == (int(od[1]) << 12) + int(od[0]) and n_lt == o_nlt
        assert np.array_equal(hist, oh)
        assert rank == int(np.searchsorted(sel, medoid))
        within = np.sort(h[vk.HDR_WITHIN:vk.HDR_WITHIN + 4 * min(n_within, vk.VK_PROBE_INLINE)].view(np.int32))
        if n_within <= vk.VK_PROBE_INLINE:
            assert np.array_equal(sel[idx[:cnt]], within)
        assert n_nl == int(np.count_nonzero(dist <= np.float32(0.3)))
        got = np.sort(nl_rows[:n_nl].cpu().numpy())
        assert np.array_equal(got, sel[dist <= np.float32(0.3)])


@pytest.mark.parametrize("driver", ["native", "python"])
@pytest.mark.parametrize("name", list(_util.CLUSTER_CASES))
def test_cuda_clusterer_matches_reference_golden(name, driver):
    import vamb_b200.cluster as vc

    g, lat, lens, rng_seed = _util.load_cluster_golden(name)
    clusters = list(vc.ClusterGenerator(lat, lens, rng_seed=rng_seed, _driver=driver, **_util.CLUSTER_CASES[name]))
    _util.assert_clusters_equal_golden(clusters, g)


@pytest.mark.parametrize("n,d,spread,seed", [(20000, 32, 0.1, 21), (50000, 32, 0.3, 22), (8000, 283, 0.1, 23)])
def test_cuda_clusterer_matches_oracle_larger(n, d, spread, seed):
    import vamb_b200.cluster as vc
    from oracle import cluster_oracle as co
    from oracle import synth

    lat, lens = synth.make_latent(n, d, seed, spread)
    oc = list(co.OracleClusterGenerator(lat, lens, rng_seed=seed))
    gc = list(vc.ClusterGenerator(lat, lens, rng_seed=seed))
    _util.assert_clusters_equal(gc, oc)
    gp = list(vc.ClusterGenerator(lat, lens, rng_seed=seed, _driver="python"))
    _util.assert_clusters_equal(gp, oc)


def test_forced_packing_and_unpruned_paths_agree():
    """pack() at every cluster and the no-pruning path give the same clusters."""
    import vamb_b200.cluster as vc
    from oracle import synth

    lat, lens = synth.make_latent(6000, 32, 31, 0.3)
    base = list(vc.ClusterGenerator(lat, lens, rng_seed=1))
    for driver in ("native", "python"):
        # pack after every emitted cluster, like the reference's CPU path
        gen = vc.ClusterGenerator(lat, lens, rng_seed=1, _driver=driver, _pack_fraction=2.0)
        _util.assert_clusters_equal(list(gen), base)
    gen = vc.ClusterGenerator(lat, lens, rng_seed=1, _driver="python")
    gen._prune_radius = float("inf")
    gen._nl_radius = float("inf")
    _util.assert_clusters_equal(list(gen), base)


# ---- the reference's own acceptance tests for this boundary (test/test_cluster.py) ----
class TestReferenceClusterSuite:
    rng = np.random.RandomState(5)
    data = rng.random((1024, 40)).astype(np.float32)
    lens = rng.randint(500, 1000, size=1024)

    def test_basics(self):
        import vamb_b200.cluster as vc

        clstr = vc.ClusterGenerator(self.data, self.lens)
        assert clstr is iter(clstr)
        x = next(clstr)
        assert isinstance(x, vc.Cluster)
        clusters = list(clstr)
        clusters.append(x)
        assert sum(len(c.members) for c in clusters) == len(self.data)
        mems = set()
        for c in clusters:
            mems.update(c.members)
        assert mems == set(range(len(self.data)))

    def test_matches_oracle_with_ties(self):
        import vamb_b200.cluster as vc
        from oracle import cluster_oracle as co

        order = np.argsort(self.lens)[::-1]
        oc = list(co.OracleClusterGenerator(self.data, self.lens, order=order))
        gc = list(vc.ClusterGenerator(self.data, self.lens))
        _util.assert_clusters_equal(gc, oc)

    def test_destruction(self):
        import vamb_b200.cluster as vc

        copy = self.data.copy()
        clstr = vc.ClusterGenerator(self.data, self.lens)
        assert np.any(np.abs(self.data - clstr.matrix.numpy()) > 0.001)
        clstr = vc.ClusterGenerator(copy, self.lens, destroy=True)
        assert np.all(np.abs(copy - clstr.matrix.numpy()) < 1e-6)
        assert np.any(np.abs(self.data - clstr.matrix.numpy()) > 0.001)

    def test_normalization(self):
        import vamb_b200.cluster as vc
        from hashlib import md5

        def xor_rows_hash(matrix):
            m = np.frombuffer(matrix.copy().data, dtype=np.uint32)
            m.shape = matrix.shape
            v = m[0].copy()
            for i in range(1, len(m)):
                v ^= m[i]
            return md5(v).digest().hex()

        before = md5(self.data.data.tobytes()).digest().hex()
        vc.ClusterGenerator(self.data, self.lens)
        assert before == md5(self.data.data.tobytes()).digest().hex()
        cp = self.data.copy()
        vc.ClusterGenerator(cp, self.lens, destroy=True)
        assert before != md5(cp.data.tobytes()).digest().hex()
        bx = xor_rows_hash(cp)
        vc.ClusterGenerator(cp, self.lens, destroy=True, normalized=True)
        assert bx == xor_rows_hash(cp)

    def test_cluster(self):
        import vamb_b200.cluster as vc

        x = next(vc.ClusterGenerator(self.data, self.lens))
        assert isinstance(x.members, np.ndarray)


def test_eval_candidates_lists_matches_oracle(vk):
    """vk_eval_candidates_lists: per-candidate density / count / within-ids / distance to the base == the oracle's
    sample_medoid on the kept rows, for candidates INSIDE the coverage radius 0.12 of the base."""
    from oracle import cluster_oracle as co
    from oracle import synth

    n, d = 40000, 32
    lat, lens = synth.make_latent(n, d, seed=5, spread=0.25)
    host = lat.copy()
    co.normalize(host)
    lens32 = lens.astype(np.float32)
    kept = (np.random.default_rng(1).random(n) > 0.2).astype(np.uint8)
    dm, dl, dk = _dev(host), _dev(lens32), _dev(kept)
    hdr = torch.zeros(vk.HDR_SIZE, dtype=torch.uint8, device="cuda")
    hdr_host = torch.zeros(vk.HDR_SIZE, dtype=torch.uint8).pin_memory()
    over = torch.empty(n, dtype=torch.int32, device="cuda")
    nl_rows = torch.empty(n, dtype=torch.int32, device="cuda")
    nl_d = torch.empty(n, dtype=torch.float32, device="cuda")
    de = _dev(co.linspace_edges())
    s = torch.cuda.current_stream().cuda_stream
    base = int(np.flatnonzero(kept)[17])
    vk.check(vk.lib.vk_probe_sync(dm.data_ptr(), dl.data_ptr(), dk.data_ptr(), n, d, base, 0.3, de.data_ptr(),
                                  hdr.data_ptr(), over.data_ptr(), nl_rows.data_ptr(), nl_d.data_ptr(),
                                  hdr_host.data_ptr(), s))
    n_nl = int(hdr_host.numpy()[vk.HDR_NNL:vk.HDR_NNL + 4].view(np.int32)[0])
    dist_base = co.calc_distances(host, base)
    near = np.flatnonzero((dist_base <= np.float32(0.1199)) & (kept != 0))
    far = np.flatnonzero((dist_base > np.float32(0.125)) & (dist_base <= np.float32(0.3)) & (kept != 0))
    cands = list(near[:: max(1, len(near) // 50)][:50]) + list(far[:3])  # more than VK_MAX_CAND: the lists kernel takes 64
    C, cap = vk.VK_LIST_CAND, 512
    out_dev = torch.zeros(vk.VK_EVAL_SCRATCH_U64, dtype=torch.int64, device="cuda")
    out_pin = torch.zeros(4 * C, dtype=torch.int64).pin_memory()
    within_pin = torch.zeros(C * cap, dtype=torch.int32).pin_memory()
    within_dev = torch.zeros(vk.VK_EVAL_SUBS * C * cap, dtype=torch.int32, device="cuda")
    ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32).pin_memory()
    arr = (vk.c_int32 * len(cands))(*[int(c) for c in cands])
    vk.check(vk.lib.vk_eval_candidates_lists(dm.data_ptr(), dl.data_ptr(), d, nl_rows.data_ptr(), nl_d.data_ptr(), n_nl,
                                             0.3, arr, len(cands), base, 0, 0, out_dev.data_ptr(), out_pin.data_ptr(),
                                             within_dev.data_ptr(), within_pin.data_ptr(), cap, ticket.data_ptr(),
                                             flag.data_ptr(), 7, s))
    assert int(flag[0]) == 7 and int(out_dev.abs().sum()) == 0  # accumulators left zeroed for the next call
    res = out_pin.numpy().view(np.uint64)
    sel = np.flatnonzero(kept)
    sub = np.ascontiguousarray(host[sel])
    lib = co._load_lib()
    for k, c in enumerate(cands):
        dbase = np.array([res[3 * C + k]], dtype=np.uint64).astype(np.uint32).view(np.float32)[0]
        assert dbase == dist_base[c]
        if dist_base[c] > np.float32(0.1199):
            continue  # outside the coverage radius the list may miss neighbours: the driver re-probes instead
        dist = co.calc_distances(sub, int(np.searchsorted(sel, c)))
        idx = np.empty(len(sel), dtype=np.int64)
        od = (ctypes.c_uint64 * 2)()
        cnt = lib.ok_sample(co._p(dist, ctypes.c_float), co._p(np.ascontiguousarray(lens32[sel]), ctypes.c_float),
                            len(sel), ctypes.c_float(0.05), co._p(idx, ctypes.c_int64), od)
        assert int(res[2 * C + k]) == cnt
        assert (int(res[C + k]) << 12) + int(res[k]) == (int(od[1]) << 12) + int(od[0])
        got = np.sort(within_pin.numpy()[k * cap:k * cap + cnt])
        assert np.array_equal(got, sel[idx[:cnt]])
    # with a density threshold the sums and counts are the same and only denser candidates get their ids published
    dens = [(int(res[C + k]) << 12) + int(res[k]) for k in range(len(cands))]
    thr = sorted(dens)[len(dens) // 2]
    first = out_pin.clone()
    within_pin.fill_(-7)
    vk.check(vk.lib.vk_eval_candidates_lists(dm.data_ptr(), dl.data_ptr(), d, nl_rows.data_ptr(), nl_d.data_ptr(), n_nl,
                                             0.3, arr, len(cands), base, thr >> 12, thr & 4095, out_dev.data_ptr(),
                                             out_pin.data_ptr(), within_dev.data_ptr(), within_pin.data_ptr(), cap,
                                             ticket.data_ptr(), flag.data_ptr(), 8, s))
    assert int(flag[0]) == 8 and torch.equal(out_pin, first)
    for k, c in enumerate(cands):
        ids = within_pin.numpy()[k * cap:k * cap + int(res[2 * C + k])]
        if dens[k] > thr:
            assert (ids >= 0).all() and len(np.unique(ids)) == len(ids)
        else:
            assert (ids == -7).all()


@pytest.mark.parametrize("n,d,spread,seed", [(30000, 32, 0.3, 41), (30000, 32, 0.08, 42), (6000, 40, 0.2, 43)])
def test_lazy_medoid_moves_emit_the_same_clusters(n, d, spread, seed, monkeypatch):
    """Native driver with scan-free medoid moves (default) == a full scan per move (round-1 behaviour) == oracle;
    and the scan-free path is actually taken."""
    import vamb_b200.cluster as vc
    from oracle import cluster_oracle as co
    from oracle import synth

    lat, lens = synth.make_latent(n, d, seed, spread)
    gen = vc.ClusterGenerator(lat, lens, rng_seed=seed)
    lazy = list(gen)
    t = gen._timing()
    monkeypatch.setenv("VAMB_B200_CLUSTER_LAZY", "0")
    gen0 = vc.ClusterGenerator(lat, lens, rng_seed=seed)
    eager = list(gen0)
    t0 = gen0._timing()
    monkeypatch.delenv("VAMB_B200_CLUSTER_LAZY")
    _util.assert_clusters_equal(lazy, eager)
    assert t0["lazy_moves"] == 0 and t["lazy_moves"] > 0
    assert gen._n_probes < gen0._n_probes
    _util.assert_clusters_equal(lazy, list(co.OracleClusterGenerator(lat, lens, rng_seed=seed)))


def test_next_block_equals_iteration_and_writes_reference_text(tmp_path):
    """Block emission (one foreign call per block): same clusters as iterating, mixing ``next()`` and
    ``next_block`` is allowed, and the block writer's files equal the per-cluster restatement of the reference's loop."""
    import os

    import vamb_b200.cluster as vc
    from oracle import synth
    from tests.test_cluster_writer_cpu import reference_style

    lat, lens = synth.make_latent(12000, 32, 51, 0.25)
    want = list(vc.ClusterGenerator(lat, lens, rng_seed=9))
    gen = vc.ClusterGenerator(lat, lens, rng_seed=9)
    got = [next(gen), next(gen)]
    for blk in gen.iter_blocks(97):
        assert len(blk) <= 97 and blk.offsets[-1] == len(blk.members)
        got.extend(blk.clusters())
    _util.assert_clusters_equal(got, want)
    assert gen.n_remaining_points == 0 and gen.n_emitted_clusters == len(want)
    with pytest.raises(StopIteration):
        next(gen)
    names = [f"S{i % 3}C{i}" for i in range(len(lat))]
    base = os.path.join(tmp_path, "vae")
    n_cl, n_ct = vc.write_clusters_tsv(vc.ClusterGenerator(lat, lens, rng_seed=9), names, lens, base, bin_prefix="bin_")
    u, m = reference_style(want, names, lens, "bin_")
    assert (n_cl, n_ct) == (len(want), len(lat))
    assert open(base + "_unsplit.tsv").read() == u and open(base + "_metadata.tsv").read() == m
