"""The data formats around the path and the TNF projection.  CPU: npz round trips in the reference's
formats (vamb/parsecontigs.py:110-129, vamb/parsebam.py:55-86, vamb/vambtools.py:738-762).  GPU: ``project_tnf`` against
a float64 restatement of ``Composition._project`` (vamb/parsecontigs.py:141-150)."""
import io
import os

import numpy as np
import pytest

pytest.importorskip("vamb_b200._lib")


def test_npz_formats_round_trip(tmp_path):
    from vamb_b200 import inputs

    rng = np.random.default_rng(0)
    n = 50
    comp = dict(matrix=rng.random((n, 103), dtype=np.float32), identifiers=np.array([f"c{i}" for i in range(n)], dtype=object),
                lengths=rng.integers(2000, 9000, n), mask=np.ones(n + 7, dtype=bool), minlength=2000)
    p = os.path.join(tmp_path, "composition.npz")
    np.savez_compressed(p, **comp)  # what Composition.save writes
    got = inputs.load_composition(p)
    assert np.array_equal(got["matrix"], comp["matrix"]) and got["minlength"] == 2000
    assert list(got["identifiers"]) == list(comp["identifiers"]) and np.array_equal(got["lengths"], comp["lengths"])
    ab = dict(matrix=rng.random((n, 4), dtype=np.float32), samplenames=np.array(["a", "b", "c", "d"], dtype=object),
              minid=0.9, refhash=b"\x01\x02")
    p = os.path.join(tmp_path, "abundance.npz")
    np.savez_compressed(p, **ab)
    got = inputs.load_abundance(p)
    assert np.array_equal(got["matrix"], ab["matrix"]) and got["minid"] == 0.9 and got["refhash"] == b"\x01\x02"
    p = os.path.join(tmp_path, "old.npz")
    np.savez_compressed(p, ab["matrix"])
    assert np.array_equal(inputs.load_abundance(p)["matrix"], ab["matrix"])
    buf = io.BytesIO()
    inputs.write_npz(buf, comp["matrix"])
    buf.seek(0)
    assert np.array_equal(inputs.read_npz(buf), comp["matrix"])
    with pytest.raises(ValueError):
        np.savez_compressed(p, matrix=comp["matrix"].astype(np.float64), identifiers=comp["identifiers"],
                            lengths=comp["lengths"], mask=comp["mask"], minlength=2000)
        inputs.load_composition(p)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 31, 1000, 100_003])
def test_project_tnf_matches_float64_restatement(n):
    from vamb_b200 import inputs

    rng = np.random.default_rng(n)
    counts = rng.integers(0, 400, size=(n, 256)).astype(np.float32)
    if n > 5:
        counts[3] = 0.0  # an all-zero row keeps its sum at 1 (parsecontigs.py:144-145)
    kernel = rng.standard_normal((256, 103)).astype(np.float32)
    x = counts.astype(np.float64)
    s = x.sum(axis=1).reshape(-1, 1)
    s[s == 0] = 1.0
    want = (x / s - 1.0 / 256.0) @ kernel.astype(np.float64)
    keep = counts.copy()
    got = inputs.project_tnf(counts, kernel)
    assert np.array_equal(counts, keep)  # the input is not modified
    assert got.shape == (n, 103) and got.dtype == np.float32
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 2e-6 * scale + 1e-7
