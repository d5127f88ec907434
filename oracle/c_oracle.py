"""ctypes front end of oracle/dcnv3_oracle.c (TEST INFRASTRUCTURE ONLY -- see that file's header).

`build()` compiles the C restatement with gcc into oracle/_build/libdcnv3_oracle.so (git-ignored;
it travels to the GPU box with the snapshot, and is rebuilt there if missing since gcc is present).
"""
import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_DIR, "dcnv3_oracle.c")
_SO = os.path.join(_DIR, "_build", "libdcnv3_oracle.so")
_lib = None


class Params(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int) for k in
                ("n", "h", "w", "ho", "wo", "groups", "gc", "kh", "kw", "sh", "sw", "ph", "pw", "dh",
                 "dw")] + [("scale", ctypes.c_float)]


def build(force=False, native=False):
    """native=True: -march=native build for timing on THIS host (bench.py's CPU baseline); it gets
    its own file name and is always rebuilt, because the snapshot may come from another CPU."""
    so = _SO.replace(".so", "_native.so") if native else _SO
    if force or native or not os.path.isfile(so) or os.path.getmtime(so) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        flags = ["-O3", "-march=native"] if native else ["-O2"]
        subprocess.check_call(["gcc", *flags, "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC",
                               "-o", so, _SRC, "-lm"])
    return so


def lib(native=False):
    global _lib
    if _lib is None or (native and not getattr(_lib, "_native", False)):
        _lib = ctypes.CDLL(build(native=native))
        _lib._native = native
        fp = ctypes.POINTER(ctypes.c_float)
        _lib.dcnv3_oracle_forward_f32.argtypes = [fp] * 4 + [ctypes.POINTER(Params), ctypes.c_int]
        _lib.dcnv3_oracle_backward_f32.argtypes = [fp] * 7 + [ctypes.POINTER(Params), ctypes.c_int]
    return _lib


def max_threads():
    return int(lib().dcnv3_oracle_max_threads())


def _params(x, offset, kernel_size, strides, padding, dilation_rate, groups, group_channels,
            offset_scale):
    from .dcnv3_oracle import check_shapes, resolve_padding

    ph, pw = resolve_padding(kernel_size, padding)
    n, h, w, _ = x.shape
    mask_shape = offset.shape[:3] + (offset.shape[3] // 2,)
    _, _, ho, wo = check_shapes(x.shape, offset.shape, mask_shape, kernel_size, strides, (ph, pw),
                                dilation_rate, groups, group_channels)
    return Params(n, h, w, ho, wo, groups, group_channels, kernel_size[0], kernel_size[1],
                  strides[0], strides[1], ph, pw, dilation_rate[0], dilation_rate[1],
                  float(offset_scale))


def _fp(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def forward(x, offset, mask, kernel_size=(3, 3), strides=(1, 1), padding="SAME",
            dilation_rate=(1, 1), groups=4, group_channels=16, offset_scale=1.0, nthreads=0):
    x, offset, mask = (np.ascontiguousarray(a, dtype=np.float32) for a in (x, offset, mask))
    q = _params(x, offset, kernel_size, strides, padding, dilation_rate, groups, group_channels,
                offset_scale)
    out = np.empty((q.n, q.ho, q.wo, groups * group_channels), np.float32)
    rc = lib().dcnv3_oracle_forward_f32(_fp(x), _fp(offset), _fp(mask), _fp(out), ctypes.byref(q),
                                        nthreads)
    assert rc == 0
    return out


def backward(x, offset, mask, grad_out, kernel_size=(3, 3), strides=(1, 1), padding="SAME",
             dilation_rate=(1, 1), groups=4, group_channels=16, offset_scale=1.0, nthreads=0):
    x, offset, mask, grad_out = (np.ascontiguousarray(a, dtype=np.float32)
                                 for a in (x, offset, mask, grad_out))
    q = _params(x, offset, kernel_size, strides, padding, dilation_rate, groups, group_channels,
                offset_scale)
    gx, goff, gm = np.empty_like(x), np.empty_like(offset), np.empty_like(mask)
    rc = lib().dcnv3_oracle_backward_f32(_fp(x), _fp(offset), _fp(mask), _fp(grad_out), _fp(gx),
                                         _fp(goff), _fp(gm), ctypes.byref(q), nthreads)
    assert rc == 0
    return gx, goff, gm
