"""Helpers shared by the CPU and GPU test-suites."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# must stay in sync with oracle/make_golden.py:CLUSTER_CASES (name -> constructor kwargs)
CLUSTER_CASES = {
    "c_2k_d32": {},
    "c_5k_d32": {},
    "c_10k_d32_mixed": {},
    "c_3k_d40": {},
    "c_1500_d283": {},
    "c_2k_d3": {},
    "c_4k_d32_small_window": {"windowsize": 20, "minsuccesses": 5, "maxsteps": 10},
}


def load_cluster_golden(name):
    g = np.load(os.path.join(GOLDEN, f"cluster_{name}.npz"))
    n, nlatent, data_seed, rng_seed = (int(x) for x in g["params"])
    spread = float(g["spread"][0])
    from oracle import synth

    lat, lens = synth.make_latent(n, nlatent, data_seed, spread, unique_lengths=True)
    return g, lat, lens, rng_seed


def none_or(x):
    return None if (x is None or (isinstance(x, float) and np.isnan(x))) else float(x)


def assert_clusters_equal_golden(clusters, g):
    assert len(clusters) == len(g["medoid"]), (len(clusters), len(g["medoid"]))
    off = 0
    for k, c in enumerate(clusters):
        sz = int(g["sizes"][k])
        mem = g["members"][off:off + sz]
        off += sz
        assert int(c.medoid) == int(g["medoid"][k]), f"cluster {k}: medoid"
        assert np.array_equal(np.asarray(c.members, dtype=np.int64), mem), f"cluster {k}: members"
        assert int(c.seed) == int(g["seed"][k]), f"cluster {k}: seed"
        assert none_or(c.radius) == none_or(float(g["radius"][k])), f"cluster {k}: radius"
        assert none_or(c.observed_pvr) == none_or(float(g["observed_pvr"][k])), f"cluster {k}: pvr"
        assert float(c.maximal_pvr) == float(g["maximal_pvr"][k]), f"cluster {k}: maximal_pvr"
        assert int(c.successes) == int(g["successes"][k]) and int(c.attempts) == int(g["attempts"][k]), f"cluster {k}: window"


def assert_clusters_equal(a, b):
    assert len(a) == len(b), (len(a), len(b))
    for k, (x, y) in enumerate(zip(a, b)):
        assert int(x.medoid) == int(y.medoid), f"cluster {k}: medoid {x.medoid} != {y.medoid}"
        assert np.array_equal(np.asarray(x.members, dtype=np.int64), np.asarray(y.members, dtype=np.int64)), f"cluster {k}: members"
        assert int(x.seed) == int(y.seed), f"cluster {k}: seed"
        assert x.radius == y.radius and x.observed_pvr == y.observed_pvr, f"cluster {k}: radius/pvr"
        assert x.maximal_pvr == y.maximal_pvr and x.successes == y.successes and x.attempts == y.attempts, f"cluster {k}: window"
