"""Host-side helpers of the hot path -- counterparts of vamb/vambtools.py:250-330.

Only the four functions the encoder / clusterer touch are provided (``zscore``,
``numpy_inplace_maskarray``, ``torch_inplace_maskarray``, ``mask_lower_bits``); they are
O(N) one-off NumPy passes, not kernels.  ``vambcore.overwrite_matrix`` (an external Rust
wheel in the reference) is restated as a stable in-place row compaction.
"""
from typing import Optional

import numpy as _np


def zscore(array: _np.ndarray, axis: Optional[int] = None, inplace: bool = False) -> _np.ndarray:
    """z-score along ``axis`` (whole array when None); constant slices map to 0.
    Same contract as vamb/vambtools.py:250-288 (pinned by test_vambtools.py:212-269)."""
    if axis is not None and (axis >= array.ndim or axis < 0):
        raise _np.exceptions.AxisError(str(axis))
    if inplace and not _np.issubdtype(array.dtype, _np.floating):
        raise TypeError("Cannot convert a non-float array to zscores")
    mean = array.mean(axis=axis)
    std = array.std(axis=axis)
    if axis is None:
        if std == 0:
            std = 1
    else:
        std[std == 0.0] = 1
        keep = tuple(1 if ax == axis else dim for ax, dim in enumerate(array.shape))
        mean = mean.reshape(keep)
        std = std.reshape(keep)
    if inplace:
        array -= mean
        array /= std
        return array
    return (array - mean) / std


def overwrite_matrix(matrix: _np.ndarray, mask: _np.ndarray) -> int:
    """Move the rows where ``mask`` is true to the front (order kept); return how many."""
    idx = _np.flatnonzero(_np.asarray(mask, dtype=bool))
    n = len(idx)
    if n and idx[-1] != n - 1:
        matrix[:n] = matrix[idx]  # forward copy is safe: rows only move towards the front
    return int(n)


def numpy_inplace_maskarray(array: _np.ndarray, mask: _np.ndarray) -> _np.ndarray:
    """``array[mask]`` without allocating (vamb/vambtools.py:291-304)."""
    if len(mask) != len(array):
        raise ValueError("Lengths of array and mask must match")
    elif len(array.shape) != 2:
        raise ValueError("Can only take a 2 dimensional-array.")
    index = overwrite_matrix(array, mask)
    array.resize((index, array.shape[1]), refcheck=False)
    return array


def torch_inplace_maskarray(array, mask):
    """``array[mask]`` for a CPU tensor without allocating (vamb/vambtools.py:307-321)."""
    if len(mask) != len(array):
        raise ValueError("Lengths of array and mask must match")
    elif array.dim() != 2:
        raise ValueError("Can only take a 2 dimensional-array.")
    np_array = array.numpy()
    index = overwrite_matrix(np_array, _np.frombuffer(mask.numpy(), dtype=bool))
    array.resize_((index, array.shape[1]))
    return array


def mask_lower_bits(floats: _np.ndarray, bits: int) -> None:
    """Zero the lowest ``bits`` mantissa bits in place (vamb/vambtools.py:324-330)."""
    if bits < 0 or bits > 23:
        raise ValueError("Must mask between 0 and 23 bits")
    mask = ~_np.uint32(2 ** bits - 1)
    u = floats.view(_np.uint32)
    u &= mask
