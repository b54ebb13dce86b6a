"""TEST INFRASTRUCTURE ONLY - numpy stand-in for the one TFP call in aslrest_disp.py:63."""
import types as _types

import numpy as _np


def _batch_interp_regular_1d_grid(x, x_ref_min, x_ref_max, y_ref, axis=-1, **_kw):
    """Linear interpolation of y_ref (regular grid on [min,max] along the last axis) at x,
    constant extension outside the grid (TFP default fill_value='constant_extension')."""
    assert axis == -1
    x, y_ref = _np.asarray(x), _np.asarray(y_ref)
    n = y_ref.shape[-1]
    pos = (x - x_ref_min) / (x_ref_max - x_ref_min) * (n - 1)
    pos = _np.clip(pos, 0, n - 1)
    lo = _np.clip(_np.floor(pos).astype(_np.int64), 0, n - 2)
    frac = pos - lo
    lo_b = _np.broadcast_to(lo, _np.broadcast_shapes(lo.shape, y_ref.shape[:-1] + (1,)))
    y_b = _np.broadcast_to(y_ref, lo_b.shape[:-1] + (n,))
    y0 = _np.take_along_axis(y_b, lo_b, axis=-1)
    y1 = _np.take_along_axis(y_b, lo_b + 1, axis=-1)
    return y0 + frac * (y1 - y0)


math = _types.SimpleNamespace(batch_interp_regular_1d_grid=_batch_interp_regular_1d_grid)
