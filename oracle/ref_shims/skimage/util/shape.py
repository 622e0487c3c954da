"""Shim for skimage.util.shape.view_as_windows (only used by non-SIMPLE sampling grids)."""
import numpy as np


def view_as_windows(arr, window_shape, step=1):
    if isinstance(step, int):
        step = (step,) * arr.ndim
    v = np.lib.stride_tricks.sliding_window_view(arr, window_shape)
    return v[tuple(slice(None, None, s) for s in step)]
