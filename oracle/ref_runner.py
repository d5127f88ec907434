"""Run the UNMODIFIED reference DCNv3 source over the torch-backed TF shim (authoring container only).

TEST INFRASTRUCTURE ONLY -- used by tests/golden/make_golden.py to freeze golden vectors and by
tests marked `needs_reference`.  It executes, by file path, the reference's own

    /root/reference/layers/dcn_v3/utils.py   (get_reference_points :14, generate_dilation_grids :65,
                                              dcnv3_bilinear_sampler :110)
    /root/reference/layers/dcn_v3/op.py      (dcnv3_op :16)
    /root/reference/layers/dcn_v3/dcn_v3.py  (DeformableConvolutionV3 :16)

with `import tensorflow` resolving to oracle/tf_shim/tensorflow (torch CPU tensors) and with the two
`iseg.utils` helpers those files import replaced by local stand-ins (the real `iseg/utils/__init__`
pulls in keras, distutils and the whole library).  No reference source is copied: the files are
loaded from where they lie.  /root/reference does not exist on the GPU box; callers must check
`available()`.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ISEG_REFERENCE_ROOT", "/root/reference")
_SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tf_shim")

_loaded = None


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "layers", "dcn_v3", "op.py"))


def _load_file(modname, path):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Returns a namespace with the reference's dcnv3_op, sampler helpers and layer class."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    if "tensorflow" in sys.modules and not getattr(
        sys.modules["tensorflow"], "__version__", ""
    ).endswith("torchshim"):
        raise RuntimeError("a real tensorflow is already imported; use it directly instead")
    if _SHIM_DIR not in sys.path:
        sys.path.insert(0, _SHIM_DIR)
    import tensorflow as tf  # the shim

    # Stand-ins for iseg.utils.get_tensor_shape (utils/common.py:67: static shape, dynamic fallback;
    # shapes are always static here) and iseg.utils.keras3_utils.Keras3_Model_Wrapper
    # (utils/keras3_utils.py:23: keras.Model that rewrites '/' in the name).
    def get_tensor_shape(x, return_list=False):
        shp = [int(s) for s in x.shape]
        return shp if return_list else tuple(shp)

    class Keras3_Model_Wrapper(tf.keras.Model):
        def __init__(self, *args, name=None, **kwargs):
            super().__init__(*args, name=None if name is None else name.replace("/", "."), **kwargs)

    def pkg(name, path=None):
        m = types.ModuleType(name)
        m.__path__ = [path] if path else []
        sys.modules[name] = m
        return m

    for name in list(sys.modules):
        if name == "iseg" or name.startswith("iseg."):
            del sys.modules[name]
    pkg("iseg")
    u = pkg("iseg.utils")
    u.get_tensor_shape = get_tensor_shape
    k3 = types.ModuleType("iseg.utils.keras3_utils")
    k3.Keras3_Model_Wrapper = Keras3_Model_Wrapper
    sys.modules["iseg.utils.keras3_utils"] = k3
    pkg("iseg.layers")
    d = os.path.join(REFERENCE_ROOT, "layers", "dcn_v3")
    pkg("iseg.layers.dcn_v3", d)
    utils = _load_file("iseg.layers.dcn_v3.utils", os.path.join(d, "utils.py"))
    op = _load_file("iseg.layers.dcn_v3.op", os.path.join(d, "op.py"))
    layer = _load_file("iseg.layers.dcn_v3.dcn_v3", os.path.join(d, "dcn_v3.py"))
    _loaded = types.SimpleNamespace(
        tf=tf,
        dcnv3_op=op.dcnv3_op,
        get_reference_points=utils.get_reference_points,
        generate_dilation_grids=utils.generate_dilation_grids,
        dcnv3_bilinear_sampler=utils.dcnv3_bilinear_sampler,
        DeformableConvolutionV3=layer.DeformableConvolutionV3,
    )
    return _loaded


def run_op(x, offset, mask, kernel_size=(3, 3), strides=(1, 1), padding="SAME",
           dilation_rate=(1, 1), groups=4, group_channels=16, offset_scale=1.0, grad_out=None):
    """Runs the reference dcnv3_op on numpy inputs.  Returns out, or (out, gx, goff, gmask) when
    grad_out is given (gradients by torch autograd through the reference's own forward graph,
    standing in for TF autodiff: floor / int casts / clip cut the graph in both)."""
    import numpy as np
    import torch

    ref = load()
    tx, to, tm = (ref.tf.convert_to_tensor(torch.from_numpy(np.ascontiguousarray(a)))
                  for a in (x, offset, mask))
    if grad_out is not None:
        tx.requires_grad_(True), to.requires_grad_(True), tm.requires_grad_(True)
    out = ref.dcnv3_op(tx, to, tm, list(kernel_size), list(strides), padding, list(dilation_rate),
                       groups, group_channels, offset_scale)
    if grad_out is None:
        return out.detach().numpy()
    out.backward(torch.from_numpy(np.ascontiguousarray(grad_out)))
    return (out.detach().numpy(), tx.grad.numpy(), to.grad.numpy(), tm.grad.numpy())


# ---- sibling op: the sampler of the reference's deformable multi-head self-attention --------
_loaded_dmsa = None


def load_deform_attn():
    """The reference's DeformableMultiHeadSelfAttentionLayer class, loaded from
    /root/reference/layers/deformable_multihead_self_attention.py over the shim.  The helpers that file imports from
    the rest of the library (NaN scrubbing, check_numerics, the Keras-3 switch) are stand-ins: `_bilinear_sample`
    (:102-175), the only method used, touches none of them."""
    global _loaded_dmsa
    if _loaded_dmsa is not None:
        return _loaded_dmsa
    ref = load()
    tf = ref.tf
    import torch

    # primitives only this file needs
    if not hasattr(tf, "broadcast_to"):
        tf.broadcast_to = lambda x, shape, name=None: tf.convert_to_tensor(torch.broadcast_to(x, tuple(int(v) for v in shape)))
    if not hasattr(tf, "gather_nd"):
        tf.gather_nd = tf.raw_ops.GatherNd
    keras = types.ModuleType("keras")
    keras.backend = types.SimpleNamespace(epsilon=lambda: 1e-7)
    keras.layers = tf.keras.layers
    sys.modules.setdefault("keras", keras)
    iseg = sys.modules["iseg"]
    iseg.check_numerics = lambda x, *a, **k: x
    ou = types.ModuleType("iseg.utils.op_utils")
    ou.replace_nan_or_inf = lambda x, *a, **k: x
    ou.safed_softmax = lambda x, *a, **k: tf.nn.softmax(x)
    sys.modules["iseg.utils.op_utils"] = ou
    vu = types.ModuleType("iseg.utils.version_utils")
    vu.is_keras3 = lambda: True
    sys.modules["iseg.utils.version_utils"] = vu
    mod = _load_file("iseg.layers.deformable_multihead_self_attention",
                     os.path.join(REFERENCE_ROOT, "layers", "deformable_multihead_self_attention.py"))
    _loaded_dmsa = mod.DeformableMultiHeadSelfAttentionLayer
    return _loaded_dmsa


def run_deform_attn(value, y, x, attn, grad_out=None):
    """Reference `_bilinear_sample(value, y, x)` (:102-175) followed by the two statements that aggregate it
    (:233-235: `tf.reduce_sum(sampled * tf.expand_dims(attn, -1), axis=-2)`).  Returns out, or (out, grad_value, grad_y,
    grad_x, grad_attn) when grad_out is given (torch autograd through that graph, standing in for TF autodiff)."""
    import numpy as np
    import torch

    cls = load_deform_attn()
    tf = load().tf
    tv, ty, tx, ta = (tf.convert_to_tensor(torch.from_numpy(np.ascontiguousarray(a))) for a in (value, y, x, attn))
    if grad_out is not None:
        for t in (tv, ty, tx, ta):
            t.requires_grad_(True)
    sampled = cls._bilinear_sample(None, tv, ty, tx)                       # [N,H,W,heads,P,C]
    out = tf.reduce_sum(sampled * tf.expand_dims(ta, axis=-1), axis=-2)    # :233-235
    if grad_out is None:
        return out.detach().numpy()
    out.backward(torch.from_numpy(np.ascontiguousarray(grad_out)))
    return (out.detach().numpy(), tv.grad.numpy(), ty.grad.numpy(), tx.grad.numpy(), ta.grad.numpy())


# ---- sibling op: DCNv2 ------------------------------------------------------------------------
_loaded_dcnv2 = None


def load_dcn_v2():
    """The reference's DCNv2 layer class, loaded from /root/reference/layers/dcn_v2.py over the shim.  Stand-ins: a
    minimal keras Layer with add_weight (zeros / glorot initialisers), `keras.activations.get`, and the three helpers
    the file imports from iseg.utils."""
    global _loaded_dcnv2
    if _loaded_dcnv2 is not None:
        return _loaded_dcnv2
    ref = load()
    tf = ref.tf
    import torch

    class Layer:
        def __init__(self, *args, name=None, **kwargs):
            self.name = name

        def add_weight(self, name=None, shape=None, initializer="zeros", regularizer=None, trainable=True, dtype="float32"):
            return tf.convert_to_tensor(torch.zeros(tuple(int(v) for v in shape), dtype=torch.float32))

        def build(self, input_shape):
            self.built = True

    keras = sys.modules.get("keras") or types.ModuleType("keras")
    keras.layers = types.SimpleNamespace(Layer=Layer, **{k: getattr(tf.keras.layers, k) for k in ("Dense",)})
    keras.activations = types.SimpleNamespace(get=lambda a: (lambda t: t) if a is None else a)
    keras.backend = types.SimpleNamespace(epsilon=lambda: 1e-7)
    sys.modules["keras"] = keras
    vu = types.ModuleType("iseg.utils.version_utils")
    vu.is_keras3 = lambda: True
    sys.modules["iseg.utils.version_utils"] = vu
    val = types.ModuleType("iseg.utils.value_utils")
    val.values_to_tuple_2d = lambda v: tuple(v) if isinstance(v, (list, tuple)) else (v, v)   # utils/value_utils.py:22-31
    sys.modules["iseg.utils.value_utils"] = val
    k3 = sys.modules["iseg.utils.keras3_utils"]

    class Keras3_Layer_Wrapper(Layer):   # utils/keras3_utils.py:32: keras Layer that rewrites '/' in the name
        def __init__(self, trainable=True, name=None, dtype=None, dynamic=False, **kwargs):
            super().__init__(name=None if name is None else name.replace("/", "."))

    k3.Keras3_Layer_Wrapper = Keras3_Layer_Wrapper
    mod = _load_file("iseg.layers.dcn_v2", os.path.join(REFERENCE_ROOT, "layers", "dcn_v2.py"))
    _loaded_dcnv2 = mod.DCNv2
    return _loaded_dcnv2


def run_dcn_v2(x, kernel, bias, offset_kernel, offset_bias, dilation_rate=1, grads_for=None):
    """Reference DCNv2 (`layers/dcn_v2.py`): the class's own build() (:61-113: weights, patch offsets) and _forward()
    (:121-265) on numpy inputs, with the given weights written over the freshly built ones.  Returns the output
    [N,H,W,filters], or (out, grad_x, grad_kernel, grad_bias, grad_offset_kernel, grad_offset_bias) when `grads_for`
    (a grad_out array) is given."""
    import numpy as np
    import torch

    cls = load_dcn_v2()
    tf = load().tf
    kh, kw, ic, oc = kernel.shape
    layer = cls(oc, (kh, kw), dilation_rate=dilation_rate, use_bias=bias is not None)
    layer.build((None, None, None, ic))
    t = lambda a: tf.convert_to_tensor(torch.from_numpy(np.ascontiguousarray(a)))  # noqa: E731
    tx = t(x)
    layer.kernel, layer.offset_kernel, layer.offset_bias = t(kernel), t(offset_kernel), t(offset_bias)
    if bias is not None:
        layer.bias = t(bias)
    leaves = [tx, layer.kernel] + ([layer.bias] if bias is not None else []) + [layer.offset_kernel, layer.offset_bias]
    if grads_for is not None:
        for v in leaves:
            v.requires_grad_(True)
    out = layer.call(tx)
    if grads_for is None:
        return out.detach().numpy()
    out.backward(torch.from_numpy(np.ascontiguousarray(grads_for)))
    return (out.detach().numpy(),) + tuple(v.grad.numpy() for v in leaves)
