"""
TEST INFRASTRUCTURE ONLY - numpy stand-in for the TensorFlow ops the reference
plugins call, so that /root/reference/svb_models_asl/*.py
can be executed eagerly on numpy arrays to produce golden vectors.
Each function follows the TF op's documented semantics; dtype follows the inputs
(feed float64 for high-precision goldens, float32 for the reference's own precision).
"""
import types as _types

import numpy as _np
from scipy import special as _sp

float32 = _np.float32
float64 = _np.float64
int32 = _np.int32
Tensor = _np.ndarray


def _a(x):
    return _np.asarray(x)


def constant(value, dtype=None, **_kw):
    return _np.asarray(value, dtype=dtype)


def identity(x, **_kw):
    return _a(x)


def shape(x, **_kw):
    return _np.array(_a(x).shape, dtype=_np.int64)


class _Immutable(_np.ndarray):
    """TF tensors are immutable: `x += y` rebinds to a new (broadcast) tensor (aslrest.py:331,338)."""
    def __iadd__(self, other):
        return _np.add(_np.asarray(self), other)


def zeros(shp, dtype=float32, **_kw):
    return _np.zeros(tuple(int(s) for s in _np.atleast_1d(shp)), dtype=dtype).view(_Immutable)


def ones(shp, dtype=float32, **_kw):
    return _np.ones(tuple(int(s) for s in _np.atleast_1d(shp)), dtype=dtype)


def exp(x, **_kw):
    return _np.exp(_a(x))


def add(a, b, **_kw):
    return _np.add(a, b)


def multiply(a, b, **_kw):
    return _np.multiply(a, b)


def greater(a, b, **_kw):
    return _np.greater(a, b)


def less(a, b, **_kw):
    return _np.less(a, b)


def logical_and(a, b, **_kw):
    return _np.logical_and(a, b)


def logical_not(a, **_kw):
    return _np.logical_not(a)


def where(cond, x, y, **_kw):
    return _np.where(cond, x, y)


def minimum(a, b, **_kw):
    return _np.minimum(a, b)


def maximum(a, b, **_kw):
    return _np.maximum(a, b)


def clip_by_value(x, lo, hi, **_kw):
    return _np.clip(x, lo, hi)


def reduce_max(x, axis=None, **_kw):
    return _np.max(_a(x), axis=axis)


def reduce_mean(x, axis=None, **_kw):
    return _np.mean(_a(x), axis=axis)


def reduce_sum(x, axis=None, **_kw):
    return _np.sum(_a(x), axis=axis)


def expand_dims(x, axis, **_kw):
    return _np.expand_dims(_a(x), axis)


def squeeze(x, axis=None, **_kw):
    return _np.squeeze(_a(x), axis=axis)


def reshape(x, shp, **_kw):
    return _np.reshape(_a(x), tuple(int(s) for s in _np.atleast_1d(shp)))


def tile(x, multiples, **_kw):
    return _np.tile(_a(x), tuple(int(m) for m in multiples))


def stack(values, axis=0, **_kw):
    return _np.stack(values, axis=axis)


def repeat(x, repeats, axis=None, **_kw):
    return _np.repeat(_a(x), int(repeats), axis=axis)


def pad(x, paddings, **_kw):
    return _np.pad(_a(x), [(int(a), int(b)) for a, b in paddings])


def matmul(a, b, **_kw):
    return _np.matmul(a, b)


def gather(params, indices, axis=0, batch_dims=0, **_kw):
    if batch_dims == 1 and axis == 1:
        return _np.take_along_axis(_a(params), _a(indices), axis=1)
    if batch_dims == 0:
        return _np.take(_a(params), indices, axis=axis)
    raise NotImplementedError("gather(batch_dims=%i, axis=%i)" % (batch_dims, axis))


def _conv1d(value, filters, stride, padding, **_kw):
    """tf.nn.conv1d: value [N, W, Cin], filters [K, Cin, Cout] - cross-correlation."""
    value, filters = _a(value), _a(filters)
    assert stride == 1 and padding == "SAME" and filters.shape[1] == 1 and filters.shape[2] == 1
    k = filters.shape[0]
    left = (k - 1) // 2
    right = k - 1 - left
    padded = _np.pad(value[..., 0], [(0, 0), (left, right)])
    w = value.shape[1]
    out = _np.zeros(value.shape[:2], dtype=_np.result_type(value, filters))
    for j in range(k):
        out += padded[:, j:j + w] * filters[j, 0, 0]
    return out[..., None]


math = _types.SimpleNamespace(
    erf=lambda x, **_kw: _sp.erf(_a(x)),
    exp=exp,
    maximum=maximum,
    minimum=minimum,
    argmax=lambda x, axis=None, **_kw: _np.argmax(_a(x), axis=axis),
    igammac=lambda a, x, **_kw: _sp.gammaincc(a, x),
    igamma=lambda a, x, **_kw: _sp.gammainc(a, x),
)
nn = _types.SimpleNamespace(tanh=lambda x, **_kw: _np.tanh(_a(x)), conv1d=_conv1d)


class Session:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def run(self, x, **_kw):
        return x
