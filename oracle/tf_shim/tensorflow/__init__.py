"""Minimal stand-in for the `tensorflow` module, backed by torch CPU tensors.

TEST INFRASTRUCTURE ONLY.  TensorFlow is not installed in the authoring container, so the
reference's own DCNv3 files (/root/reference/layers/dcn_v3/{op,utils,dcn_v3}.py) -- and the two sibling
gather layers, layers/deformable_multihead_self_attention.py and layers/dcn_v2.py -- cannot be imported
as they are.  This package implements exactly the TF primitives those files touch, with the
semantics TF documents for them, so that `oracle/ref_runner.py` can execute the UNMODIFIED reference
source here and freeze its outputs (and, through torch autograd, the gradients TF autodiff would
produce) as golden fixtures under tests/golden/.

What this is not: it is not TensorFlow and not XLA.  Primitive semantics that are TF/XLA
implementation details (fused multiply-add contraction, XLA turning `x / const` into
`x * (1/const)`, reduction order inside `reduce_sum`) are not reproduced; they are <= few-ulp
effects (see DESIGN.md "oracle pinning").

Nothing in the product path (iseg_b200/) imports this.
"""
import builtins as _bi
import math as _math
import types as _types

import torch as _torch

float16 = _torch.float16
bfloat16 = _torch.bfloat16
float32 = _torch.float32
float64 = _torch.float64
int32 = _torch.int32
int64 = _torch.int64
bool = _torch.bool  # noqa: A001  (mirrors tf.bool)



class Tensor(_torch.Tensor):
    """TF tensors are immutable: `a += b` rebinds the name to a new tensor.  torch would update in
    place (and fail to broadcast / corrupt autograd), so the augmented assignments the reference
    uses are made out-of-place here."""

    def __iadd__(self, other):
        return self + other

    def __isub__(self, other):
        return self - other

    def __imul__(self, other):
        return self * other

    def __itruediv__(self, other):
        return self / other


version = _types.SimpleNamespace(VERSION="2.15.0-torchshim")
__version__ = version.VERSION


def _t(x, dtype=None):
    """tf.convert_to_tensor: python ints -> int32, python floats -> float32."""
    if isinstance(x, _torch.Tensor):
        x = x if dtype is None else x.to(dtype)
        return x if isinstance(x, Tensor) else x.as_subclass(Tensor)
    if dtype is None:
        if isinstance(x, _bi.bool):
            dtype = _torch.bool
        elif isinstance(x, int):
            dtype = _torch.int32
        elif isinstance(x, float):
            dtype = _torch.float32
    return _torch.as_tensor(x, dtype=dtype).as_subclass(Tensor)


def function(func=None, **_kwargs):
    """tf.function(jit_compile=..., autograph=..., reduce_retracing=...) -> run eagerly."""
    if func is not None and callable(func):
        return func

    def deco(f):
        return f

    return deco


def identity(x, name=None):
    return _t(x)


def convert_to_tensor(x, dtype=None, name=None):
    return _t(x, dtype)


def cast(x, dtype, name=None):
    x = _t(x)
    if dtype in (int32, int64) and x.is_floating_point():
        return _torch.trunc(x).to(dtype)  # TF float->int cast truncates toward zero
    return x.to(dtype)


def shape(x, name=None):
    return tuple(x.shape)


def reshape(x, shape, name=None):  # noqa: A002
    return _torch.reshape(_t(x), tuple(int(s) for s in shape))


def transpose(x, perm=None, name=None):
    x = _t(x)
    if perm is None:
        perm = list(range(x.dim()))[::-1]
    return x.permute(*perm).contiguous()


def expand_dims(x, axis, name=None):
    return _torch.unsqueeze(_t(x), axis)


def pad(x, paddings, mode="CONSTANT", constant_values=0, name=None):
    assert mode == "CONSTANT"
    flat = []
    for lo, hi in reversed([tuple(p) for p in paddings]):
        flat += [int(lo), int(hi)]
    return _torch.nn.functional.pad(_t(x), flat, mode="constant", value=constant_values)


def stack(values, axis=0, name=None):
    vals = [_t(v) for v in values]
    return _torch.stack(vals, dim=axis)


def tile(x, multiples, name=None):
    return _t(x).repeat(*[int(m) for m in multiples])


def zeros(shape, dtype=float32, name=None):  # noqa: A002
    return _t(_torch.zeros(tuple(int(s) for s in shape), dtype=dtype))


def ones(shape, dtype=float32, name=None):  # noqa: A002
    return _t(_torch.ones(tuple(int(s) for s in shape), dtype=dtype))


def range(start, limit=None, delta=1, dtype=None, name=None):  # noqa: A001
    if limit is None:
        start, limit = 0, start
    return _t(_torch.arange(int(start), int(limit), int(delta), dtype=dtype or _torch.int32))


def floor(x, name=None):
    return _torch.floor(_t(x))


def clip_by_value(x, clip_value_min, clip_value_max, name=None):
    lo = _t(clip_value_min).to(x.dtype)
    hi = _t(clip_value_max).to(x.dtype)
    return _torch.minimum(_torch.maximum(x, lo), hi)


def reduce_sum(x, axis=None, keepdims=False, name=None):
    if axis is None:
        return _torch.sum(x)
    return _torch.sum(x, dim=axis, keepdim=keepdims)


def linspace(start, stop, num, name=None, axis=0):
    """tf.linspace (math_ops.linspace_nd): start + delta * [1..num-2] between exact end points.

    Integer start/stop are true-divided (float64 result), as TF's `/` does on int32 tensors."""
    start = _t(start)
    stop = _t(stop, dtype=start.dtype)
    num = int(num)
    n_steps = max(num - 1, 1)
    if not start.is_floating_point():
        start = start.to(_torch.float64)
        stop = stop.to(_torch.float64)
    delta = (stop - start) / _torch.tensor(n_steps, dtype=start.dtype)
    inner = _torch.arange(1, n_steps, dtype=_torch.int64).to(start.dtype)
    res = start + delta * inner
    out = _torch.cat([start.reshape(1), res, stop.reshape(1)])
    return out[:num]


def meshgrid(*args, indexing="xy", name=None):
    return list(_torch.meshgrid(*[_t(a) for a in args], indexing=indexing))


def _pack(values, axis=0, name=None):
    return stack(values, axis=axis)


def _gather_nd(params, indices, name=None):
    idx = indices.long()
    k = idx.shape[-1]
    return params[tuple(idx[..., i] for i in _bi.range(k))]


def _mul(x, y, name=None):
    return x * y


raw_ops = _types.SimpleNamespace(Pack=_pack, GatherNd=_gather_nd, Mul=_mul)


# ---- primitives of the sibling gather ops (layers/deformable_multihead_self_attention.py, layers/dcn_v2.py) ----
gather_nd = _gather_nd


def broadcast_to(x, shape, name=None):  # noqa: A002
    return _t(_torch.broadcast_to(_t(x), tuple(int(v) for v in shape)))


_DTYPE_NAMES = {"int32": int32, "int64": int64, "float32": float32, "float64": float64, "bfloat16": bfloat16, "float16": float16}


def constant(value, dtype=None, shape=None, name=None):  # noqa: A002
    return _t(value, _DTYPE_NAMES.get(dtype, dtype))


def add(x, y, name=None):
    return _t(x) + _t(y)


def multiply(x, y, name=None):
    return x * y


def matmul(a, b, name=None):
    return _torch.matmul(a, b)


def squeeze(x, axis=None, name=None):
    return _torch.squeeze(x) if axis is None else _torch.squeeze(x, dim=axis)


def concat(values, axis, name=None):
    return _torch.cat([_t(v) for v in values], dim=axis)


def split(value, num_or_size_splits, axis=0, name=None):
    if isinstance(num_or_size_splits, int):
        return list(_torch.chunk(value, num_or_size_splits, dim=axis))
    return list(_torch.split(value, [int(v) for v in num_or_size_splits], dim=axis))


def unstack(value, num=None, axis=0, name=None):
    return list(_torch.unbind(value, dim=axis))


def ones_like(x, dtype=None, name=None):
    return _t(_torch.ones_like(x, dtype=dtype))


def tanh(x, name=None):
    return _torch.tanh(x)


class _InitScope:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def init_scope():
    return _InitScope()


def _conv2d(input, filters, strides, padding, dilations=None, name=None):  # noqa: A002
    """tf.nn.conv2d, NHWC x HWIO, stride 1, 'SAME' with an odd kernel (what layers/dcn_v2.py:128-134 calls)."""
    st = [int(v) for v in (strides if isinstance(strides, (list, tuple)) else [strides] * 4)]
    assert padding == "SAME" and all(v == 1 for v in st), (padding, st)
    dil = [int(v) for v in (dilations or (1, 1))][-2:]
    kh, kw = int(filters.shape[0]), int(filters.shape[1])
    assert kh % 2 == 1 and kw % 2 == 1
    y = _torch.nn.functional.conv2d(input.permute(0, 3, 1, 2), filters.permute(3, 2, 0, 1), None, 1,
                                    (dil[0] * (kh - 1) // 2, dil[1] * (kw - 1) // 2), tuple(dil))
    return _t(y.permute(0, 2, 3, 1).contiguous())


def _gelu(x, approximate=False, name=None):
    return _torch.nn.functional.gelu(x, approximate="tanh" if approximate else "none")


def _softmax(x, axis=-1, name=None):
    return _torch.softmax(x, dim=axis)


nn = _types.SimpleNamespace(gelu=_gelu, softmax=_softmax, sigmoid=lambda x, name=None: _torch.sigmoid(x), conv2d=_conv2d)


def zeros_initializer():
    return "zeros"


# --------------------------------------------------------------------------------------------------
# tf.keras.layers used by layers/dcn_v3/dcn_v3.py.  Weights are plain torch tensors created on first
# call (Keras build-on-call), Glorot-uniform kernels / zero biases like the Keras defaults, drawn
# from torch's global RNG so that a torch.manual_seed() before the first call pins them.
# --------------------------------------------------------------------------------------------------
class _Layer:
    def __init__(self, name=None):
        self.name = name
        self.built = False

    def __call__(self, x, *args, **kwargs):
        if not self.built:
            self.build(tuple(x.shape))
            self.built = True
        return self.call(x, *args, **kwargs)


def _glorot(shape, fan_in, fan_out, dtype):
    limit = _math.sqrt(6.0 / (fan_in + fan_out))
    return (_torch.rand(shape, dtype=_torch.float64) * 2 - 1).mul_(limit).to(dtype)


class Dense(_Layer):
    def __init__(self, units, kernel_initializer=None, bias_initializer=None, name=None, **_):
        super().__init__(name)
        self.units = int(units)
        self.kernel_initializer = kernel_initializer

    def build(self, input_shape):
        cin = int(input_shape[-1])
        if self.kernel_initializer == "zeros":
            self.kernel = _torch.zeros(cin, self.units)
        else:
            self.kernel = _glorot((cin, self.units), cin, self.units, _torch.float32)
        self.bias = _torch.zeros(self.units)

    def call(self, x):
        return _torch.matmul(x, self.kernel.to(x.dtype)) + self.bias.to(x.dtype)


class DepthwiseConv2D(_Layer):
    def __init__(self, kernel_size, strides=1, padding="valid", name=None, **_):
        super().__init__(name)
        self.k = int(kernel_size)
        assert int(strides) == 1
        self.padding = padding

    def build(self, input_shape):
        c = int(input_shape[-1])
        k = self.k
        self.depthwise_kernel = _glorot((k, k, c, 1), k * k * c, k * k, _torch.float32)
        self.bias = _torch.zeros(c)

    def call(self, x):
        c = x.shape[-1]
        w = self.depthwise_kernel.to(x.dtype).permute(2, 3, 0, 1).contiguous()  # [C,1,k,k]
        xin = x.permute(0, 3, 1, 2)
        p = self.k // 2 if self.padding == "same" else 0
        y = _torch.nn.functional.conv2d(xin, w, self.bias.to(x.dtype), padding=p, groups=c)
        return y.permute(0, 2, 3, 1).contiguous()


class LayerNormalization(_Layer):
    def __init__(self, epsilon=1e-3, name=None, **_):
        super().__init__(name)
        self.epsilon = epsilon

    def build(self, input_shape):
        c = int(input_shape[-1])
        self.gamma = _torch.ones(c)
        self.beta = _torch.zeros(c)

    def call(self, x):
        return _torch.nn.functional.layer_norm(
            x, (x.shape[-1],), self.gamma.to(x.dtype), self.beta.to(x.dtype), self.epsilon
        )


class _Model(_Layer):
    def __init__(self, *args, name=None, **kwargs):
        super().__init__(name)

    def build(self, input_shape):
        pass


keras = _types.SimpleNamespace(
    layers=_types.SimpleNamespace(
        Dense=Dense,
        DepthwiseConv2D=DepthwiseConv2D,
        LayerNormalization=LayerNormalization,
        Layer=_Layer,
    ),
    Model=_Model,
    __version__="2.15.0-torchshim",
)
