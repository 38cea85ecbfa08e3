"""ctypes mirror of the native clusterer driver (vk_cluster_* in include/vamb_b200.h).

``vamb_b200.cluster.ClusterGenerator`` uses it by default: one foreign call per emitted cluster,
the reference's decision logic (incl. CPython's ``random.Random.sample``) restated in C++
(vamb_b200/csrc/vk_cluster_host.cu).  ``VAMB_B200_CLUSTER_DRIVER=python`` selects the Python
rendition of the same logic instead; both emit identical clusters (tests/test_cluster_gpu.py).
"""
import ctypes as _ct

from . import _lib

_p = _ct.c_void_p


class VkClusterConfig(_ct.Structure):
    _fields_ = [
        ("n", _ct.c_int64), ("d", _ct.c_int32),
        ("maxsteps", _ct.c_int32), ("windowsize", _ct.c_int32), ("minsuccesses", _ct.c_int32),
        ("nl_radius", _ct.c_float), ("prune_radius", _ct.c_float),
        ("pack_fraction", _ct.c_double),
        ("matrix", _p), ("matrix2", _p), ("lengths", _p), ("lengths2", _p),
        ("kept", _p), ("kept2", _p), ("orig", _p), ("orig2", _p),
        ("nl_rows", _p), ("nl_dists", _p), ("hdr", _p), ("within_overflow", _p), ("edges", _p),
        ("cand_out", _p), ("members", _p), ("tile_scratch", _p),
        ("hdr_host", _p), ("cand_out_host", _p), ("members_host", _p),
        ("members_host_cap", _ct.c_int32), ("seed_key_len", _ct.c_int32),
        ("seed_key", _p), ("order_host", _p), ("normalpdf_host", _p),
        ("stream", _p),
    ]


class VkClusterResult(_ct.Structure):
    _fields_ = [
        ("medoid", _ct.c_int64), ("seed", _ct.c_int64), ("n_members", _ct.c_int64), ("n_remaining", _ct.c_int64),
        ("members_host", _ct.POINTER(_ct.c_int64)),
        ("maximal_pvr", _ct.c_double), ("observed_pvr", _ct.c_double), ("radius", _ct.c_double),
        ("peak_valley_ratio", _ct.c_double),
        ("kind", _ct.c_int32), ("successes", _ct.c_int32), ("attempts", _ct.c_int32),
    ]


class VkClusterBlock(_ct.Structure):
    _fields_ = [
        ("max_clusters", _ct.c_int64), ("n_clusters", _ct.c_int64), ("n_members_total", _ct.c_int64),
        ("n_remaining", _ct.c_int64),
        ("medoid", _p), ("seed", _p), ("n_members", _p), ("members", _p),
        ("maximal_pvr", _p), ("observed_pvr", _p), ("radius", _p),
        ("kind", _p), ("successes", _p), ("attempts", _p),
        ("peak_valley_ratio", _ct.c_double),
    ]


_L = _lib.lib
_L.vk_cluster_next_block.argtypes = [_ct.c_void_p, _ct.POINTER(VkClusterBlock)]
_L.vk_cluster_next_block.restype = _ct.c_int
_L.vk_cluster_create.argtypes = [_ct.POINTER(_ct.c_void_p), _ct.POINTER(VkClusterConfig)]
_L.vk_cluster_create.restype = _ct.c_int
_L.vk_cluster_next.argtypes = [_ct.c_void_p, _ct.POINTER(VkClusterResult)]
_L.vk_cluster_next.restype = _ct.c_int
_L.vk_cluster_stats.argtypes = [_ct.c_void_p, _ct.POINTER(_ct.c_int64)]
_L.vk_cluster_stats.restype = _ct.c_int
_L.vk_cluster_timing.argtypes = [_ct.c_void_p, _ct.POINTER(_ct.c_double)]
_L.vk_cluster_timing.restype = _ct.c_int
_L.vk_cluster_destroy.argtypes = [_ct.c_void_p]
_L.vk_cluster_destroy.restype = None
_L.vk_cluster_rng_selftest.argtypes = [_ct.POINTER(_ct.c_uint32), _ct.c_int, _ct.POINTER(_ct.c_int32), _ct.c_int,
                                       _ct.c_int, _ct.POINTER(_ct.c_int32)]
_L.vk_cluster_rng_selftest.restype = _ct.c_int
_L.vk_cluster_sizeof.argtypes = [_ct.c_int]
_L.vk_cluster_sizeof.restype = _ct.c_int64
if (_L.vk_cluster_sizeof(0) != _ct.sizeof(VkClusterConfig) or _L.vk_cluster_sizeof(1) != _ct.sizeof(VkClusterResult)
        or _L.vk_cluster_sizeof(2) != _ct.sizeof(VkClusterBlock)):
    raise ImportError("vamb_b200: cluster driver structs are out of sync with include/vamb_b200.h")


def seed_key(seed: int):
    """The 32-bit little-endian words of |seed| -- what CPython's random.seed(int) feeds init_by_array."""
    a = abs(int(seed))
    words = []
    while True:
        words.append(a & 0xFFFFFFFF)
        a >>= 32
        if a == 0:
            break
    return (_ct.c_uint32 * len(words))(*words)


def rng_selftest(seed: int, ns, k: int):
    """Positions chosen by the C++ restatement of random.Random(seed).sample(range(n), min(n, k))."""
    key = seed_key(seed)
    arr = (_ct.c_int32 * len(ns))(*ns)
    out = (_ct.c_int32 * (len(ns) * k))()
    _lib.check(_L.vk_cluster_rng_selftest(key, len(key), arr, len(ns), k, out))
    res = []
    for i, n in enumerate(ns):
        res.append([out[i * k + j] for j in range(min(n, k))])
    return res
