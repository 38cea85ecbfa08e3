__doc__ = """The data formats either side of the hot path and the one O(N) numeric step before it (SURVEY 8f-4).

  * ``load_composition`` / ``load_abundance`` / ``read_npz`` / ``write_npz``: the ``.npz`` files `vamb bin default` reads
    and writes around the path -- ``composition.npz`` (vamb/parsecontigs.py:110-129: matrix [N, 103] fp32, identifiers,
    lengths, mask, minlength), ``abundance.npz`` (vamb/parsebam.py:55-86: matrix [N, S] fp32, samplenames, minid,
    refhash; or the old single-array format) and ``latent.npz`` (vamb/vambtools.py:738-762: key ``arr_0``).
  * ``to_device``: npz array -> pinned host staging -> HBM in one asynchronous copy.
  * ``project_tnf``: vamb/parsecontigs.py:141-150 (``Composition._project``): four-mer counts [N, 256] -> TNF [N, 103]
    on the GPU (vk_tnf_project).  The projection kernel matrix is an argument (the reference ships it as package data).

FASTA parsing / k-mer counting (``vambcore.kmercounts``) and BAM parsing are I/O-bound and stay with the reference.
"""
from typing import Optional

import numpy as _np
import torch as _torch

from . import _lib


def _validate(array: _np.ndarray) -> _np.ndarray:
    "vamb/vambtools.py validate_input_array: plain ndarray, C-contiguous, not object dtype for numeric payloads."
    if not isinstance(array, _np.ndarray):
        raise ValueError("Array must be of type numpy.ndarray")
    return _np.ascontiguousarray(array)


def read_npz(file) -> _np.ndarray:
    "Array stored under ``arr_0`` (``latent.npz``, vamb/vambtools.py:738-750)."
    with _np.load(file) as npz:
        return _validate(npz["arr_0"])


def write_npz(file, array: _np.ndarray) -> None:
    "vamb/vambtools.py:753-762."
    _np.savez_compressed(file, array)


def load_composition(file) -> dict:
    "Fields of ``Composition.save`` (vamb/parsecontigs.py:110-129)."
    arrs = _np.load(file, allow_pickle=True)
    out = {
        "matrix": _validate(arrs["matrix"]),
        "identifiers": _validate(arrs["identifiers"]),
        "lengths": _validate(arrs["lengths"]),
        "mask": _validate(arrs["mask"]),
        "minlength": arrs["minlength"].item(),
    }
    if out["matrix"].dtype != _np.float32 or out["matrix"].ndim != 2:
        raise ValueError("composition matrix must be a 2-dimensional float32 array")
    if len(out["matrix"]) != len(out["identifiers"]) or len(out["lengths"]) != len(out["matrix"]):
        raise ValueError("composition arrays disagree in length")
    return out


def load_abundance(file) -> dict:
    "Fields of ``Abundance.save`` (vamb/parsebam.py:55-86), or the old single-array format (key ``arr_0``)."
    arrs = _np.load(file, allow_pickle=True)
    if "arr_0" in arrs.keys():
        return {"matrix": _validate(arrs["arr_0"]), "samplenames": None, "minid": None, "refhash": None}
    return {"matrix": _validate(arrs["matrix"]), "samplenames": arrs["samplenames"], "minid": arrs["minid"].item(),
            "refhash": arrs["refhash"].item()}


def to_device(array: _np.ndarray, device: Optional[_torch.device] = None) -> _torch.Tensor:
    "Host array -> pinned staging buffer -> HBM (one asynchronous H2D copy on the current stream)."
    _lib.require_device()
    dev = device if device is not None else _torch.device("cuda", _torch.cuda.current_device())
    staged = _torch.from_numpy(_np.ascontiguousarray(array)).pin_memory()
    return staged.to(dev, non_blocking=True)


def project_tnf(fourmers, kernel: _np.ndarray) -> _np.ndarray:
    """``Composition._project(fourmers, kernel)`` (vamb/parsecontigs.py:141-150) on the GPU.  ``fourmers``: [N, 256]
    fp32 counts (NumPy or a CUDA tensor; not modified), ``kernel``: [256, 103] fp32.  Returns [N, 103] fp32 (NumPy for
    NumPy input, a CUDA tensor for tensor input)."""
    _lib.require_device()
    kernel = _np.ascontiguousarray(kernel, dtype=_np.float32)
    if kernel.ndim != 2 or kernel.shape[0] != 256:
        raise ValueError("kernel must be a [256, n_out] matrix")
    was_numpy = isinstance(fourmers, _np.ndarray)
    x = to_device(fourmers.astype(_np.float32, copy=False)) if was_numpy else fourmers.contiguous()
    if x.ndim != 2 or x.shape[1] != 256 or x.dtype != _torch.float32:
        raise ValueError("fourmers must be an [N, 256] float32 matrix")
    k = _torch.from_numpy(kernel).to(x.device)
    out = _torch.empty((x.shape[0], kernel.shape[1]), dtype=_torch.float32, device=x.device)
    _lib.check(_lib.lib.vk_tnf_project(x.data_ptr(), k.data_ptr(), out.data_ptr(), x.shape[0], kernel.shape[1],
                                       _torch.cuda.current_stream().cuda_stream))
    if was_numpy:
        return out.cpu().numpy()
    return out
