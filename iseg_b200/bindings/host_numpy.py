"""ctypes binding of the HOST-BUFFER entry points of include/dcnv3_b200.h for numpy arrays -- the piece a
maintainer of the reference wraps in `tf.numpy_function` + `tf.custom_gradient` to replace the body of
`iseg/layers/dcn_v3/op.py:16 dcnv3_op` (INTEGRATION.md shows those few TF lines; everything below them is this
module and is exercised by tests/test_gpu_parity.py::test_host_numpy_binding).

Unlike a DLPack hand-over it never writes into framework-owned tensors (TF tensors are immutable and small
constants may be shared): inputs are read from the caller's buffers, results are fresh numpy arrays.  The library
copies host -> device, runs the sm_100a kernels, copies back and synchronises (dcnv3_forward_host /
dcnv3_forward_backward_host).  There is no CPU fallback: without a CUDA device the calls raise.
"""
import ctypes

import numpy as np

from .. import _cabi

_DTYPES = {np.dtype(np.float32): _cabi.F32}


def _resolve_padding(kernel_size, padding):
    # reference op.py:29-39
    if not isinstance(padding, str):
        raise TypeError("padding must be a string in 'SAME' or 'VALID'")
    padding = padding.upper()
    if padding == "SAME":
        return (kernel_size[0] // 2, kernel_size[1] // 2)
    if padding == "VALID":
        return (0, 0)
    raise ValueError("padding must be 'SAME' or 'VALID'")


def _prepare(x, offset, mask, kernel_size, strides, padding, dilation_rate, groups, group_channels, offset_scale,
             mask_is_logits):
    x, offset, mask = (np.ascontiguousarray(a) for a in (x, offset, mask))
    if x.dtype not in _DTYPES or offset.dtype != x.dtype or mask.dtype != x.dtype:
        raise TypeError("x, offset and mask must be float32 arrays (bfloat16 has no numpy dtype: use the torch host)")
    pad = _resolve_padding(kernel_size, padding)
    p = _cabi.make_params(x.shape, offset.shape[1:3], tuple(kernel_size), tuple(strides), pad, tuple(dilation_rate),
                          int(groups), int(group_channels), float(offset_scale), _DTYPES[x.dtype],
                          _cabi.FLAG_MASK_LOGITS if mask_is_logits else 0)
    _cabi.check(_cabi.lib.dcnv3_check_params(ctypes.byref(p)))
    c, gp = groups * group_channels, groups * kernel_size[0] * kernel_size[1]
    if x.shape[3] != c or offset.shape != (x.shape[0], p.ho, p.wo, 2 * gp) or mask.shape != (x.shape[0], p.ho, p.wo, gp):
        raise ValueError("tensor shapes do not match groups / group_channels / kernel_size")
    return x, offset, mask, p


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def dcnv3_op_numpy(x, offset, mask, kernel_size, strides, padding, dilation_rate, groups, group_channels,
                   offset_scale, mask_is_logits=False, device=0):
    """The reference's `dcnv3_op` signature on numpy arrays: returns out [N, Ho, Wo, G*gc]."""
    x, offset, mask, p = _prepare(x, offset, mask, kernel_size, strides, padding, dilation_rate, groups,
                                  group_channels, offset_scale, mask_is_logits)
    out = np.empty((x.shape[0], p.ho, p.wo, groups * group_channels), x.dtype)
    _cabi.check(_cabi.lib.dcnv3_forward_host(_ptr(x), _ptr(offset), _ptr(mask), _ptr(out), ctypes.byref(p), device))
    return out


def dcnv3_op_with_grads_numpy(x, offset, mask, grad_out, kernel_size, strides, padding, dilation_rate, groups,
                              group_channels, offset_scale, mask_is_logits=False, device=0):
    """Forward and the gradient TF autodiff derives from the reference function, in one device round trip:
    returns (out, grad_x, grad_offset, grad_mask)."""
    x, offset, mask, p = _prepare(x, offset, mask, kernel_size, strides, padding, dilation_rate, groups,
                                  group_channels, offset_scale, mask_is_logits)
    grad_out = np.ascontiguousarray(grad_out, dtype=x.dtype)
    out = np.empty((x.shape[0], p.ho, p.wo, groups * group_channels), x.dtype)
    if grad_out.shape != out.shape:
        raise ValueError("grad_out must have the shape of the output")
    gx, goff, gm = np.empty_like(x), np.empty_like(offset), np.empty_like(mask)
    _cabi.check(_cabi.lib.dcnv3_forward_backward_host(_ptr(x), _ptr(offset), _ptr(mask), _ptr(grad_out), _ptr(out),
                                                      _ptr(gx), _ptr(goff), _ptr(gm), ctypes.byref(p), device))
    return out, gx, goff, gm
