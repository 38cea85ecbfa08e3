"""CPU suite: the block-wise TSV writer prints exactly what the reference's per-cluster loop prints
(vamb/__main__.py:1310-1377 with the binsplitter disabled), checked against a line-by-line restatement of that loop."""
import io
import os

import numpy as np
import pytest

pytest.importorskip("vamb_b200._lib")


def reference_style(clusters, sequence_names, sequence_lens, bin_prefix):
    """The reference's loop, restated: returns (unsplit text, metadata text)."""
    unsplit, meta = io.StringIO(), io.StringIO()
    print("name\tradius\tpeak valley ratio\tkind\tbp\tncontigs\tmedoid", file=meta)
    print("clustername\tcontigname", file=unsplit)
    for cluster_index, cluster in enumerate(clusters):
        cluster_members = [sequence_names[int(i)] for i in cluster.members]
        cluster_name = str(cluster_index + 1)
        if bin_prefix is not None:
            cluster_name = bin_prefix + cluster_name
        for member in cluster_members:
            print(cluster_name, member, sep="\t", file=unsplit)
        print(cluster_name, None if cluster.radius is None else round(cluster.radius, 3),
              None if cluster.observed_pvr is None else round(cluster.observed_pvr, 2), cluster.kind_str,
              sum(sequence_lens[i] for i in cluster.members), len(cluster_members), sequence_names[cluster.medoid],
              file=meta, sep="\t")
    return unsplit.getvalue(), meta.getvalue()


class FakeGenerator:
    """Stands in for ClusterGenerator: serves a fixed list of clusters through ``next_block`` (Python-driver path)."""

    def __init__(self, clusters):
        import vamb_b200.cluster as vc

        self._clusters = list(clusters)
        self._native = None
        self._vc = vc
        self._pos = 0

    def __next__(self):
        if self._pos >= len(self._clusters):
            raise StopIteration
        self._pos += 1
        return self._clusters[self._pos - 1]

    def next_block(self, max_clusters=1024):
        return self._vc.ClusterGenerator.next_block(self, max_clusters)


@pytest.mark.parametrize("prefix,max_clusters,block", [(None, None, 7), ("S1C", None, 1000), ("b", 13, 5)])
def test_block_writer_equals_reference_loop(tmp_path, prefix, max_clusters, block):
    import vamb_b200.cluster as vc

    rng = np.random.default_rng(3)
    n = 500
    names = [f"contig_{i}_len" for i in range(n)]
    lens = rng.integers(2000, 90000, n)
    perm = rng.permutation(n)
    clusters, off = [], 0
    while off < n:
        sz = int(min(n - off, rng.integers(1, 30)))
        mem = np.sort(perm[off:off + sz]).astype(np.int64)
        off += sz
        kind = rng.integers(0, 3) if sz > 1 else 0
        radius = None if kind == 0 else float(rng.choice([0.06, 0.015, 0.1234567, 0.2]))
        pvr = float(rng.random()) if kind == 2 else None
        clusters.append(vc.Cluster(int(mem[rng.integers(0, sz)]), 0, mem, 0.1, pvr, radius, 0, 0))
    base = os.path.join(tmp_path, "vae")
    got = vc.write_clusters_tsv(FakeGenerator(clusters), names, lens, base, bin_prefix=prefix, max_clusters=max_clusters,
                                block=block)
    want_clusters = clusters if max_clusters is None else clusters[:max_clusters]
    u, m = reference_style(want_clusters, names, lens, prefix)
    assert open(base + "_unsplit.tsv").read() == u
    assert open(base + "_metadata.tsv").read() == m
    assert got == (len(want_clusters), sum(len(c.members) for c in want_clusters))
