"""ctypes binding of the C ABI in include/dcnv3_b200.h (the thin shim of BASELINE north_star).

Tensors cross the boundary as DLPack `DLManagedTensor*` taken zero-copy from torch tensors; the
library validates dtype / shape / device / contiguity itself.  There is no CPU fallback: if the
shared library is missing, importing this module raises.
"""
import ctypes
import os

import torch
from torch.utils.dlpack import to_dlpack

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libdcnv3_b200.so")

F32, BF16 = 0, 1
FLAG_MASK_LOGITS = 1
FLAG_FORCE_GENERIC = 2
FLAG_WORKSPACE_ZEROED = 4
FLAG_REF_DTYPE = 8
FLAG_CHECK_WORKSPACE = 16

ERR_DTYPE, ERR_SHAPE, ERR_LAYOUT, ERR_DEVICE, ERR_CUDA, ERR_WORKSPACE, ERR_ARGUMENT = range(-1, -8, -1)


class Params(ctypes.Structure):
    """struct dcnv3_params"""
    _fields_ = [(k, ctypes.c_int32) for k in
                ("n", "h", "w", "ho", "wo", "groups", "group_channels", "kh", "kw", "sh", "sw", "ph",
                 "pw", "dh", "dw")] + [("offset_scale", ctypes.c_float), ("dtype", ctypes.c_int32),
                                       ("flags", ctypes.c_uint32)]


class DCNv3Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"dcnv3_b200 error {code}: {message}")
        self.code = code


def _load():
    if not os.path.isfile(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python -m iseg_b200.build` "
            "(there is no CPU / PyTorch fallback for the DCNv3 op)")
    lib = ctypes.CDLL(_LIB_PATH)
    vp, ci, cf, cu = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_uint
    pp = ctypes.POINTER(Params)
    lib.dcnv3_abi_version.restype = ci
    lib.dcnv3_last_error.restype = ctypes.c_char_p
    lib.dcnv3_build_info.restype = ctypes.c_char_p
    lib.dcnv3_check_params.argtypes = [pp]
    lib.dcnv3_launch_plan.argtypes = [pp, ctypes.POINTER(ctypes.c_int)]
    lib.dcnv3_forward.argtypes = [vp] * 4 + [pp, vp]
    lib.dcnv3_backward_workspace_bytes.argtypes = [pp]
    lib.dcnv3_backward_workspace_bytes.restype = ctypes.c_size_t
    lib.dcnv3_backward_workspace_zero_bytes.argtypes = [pp]
    lib.dcnv3_backward_workspace_zero_bytes.restype = ctypes.c_size_t
    lib.dcnv3_backward.argtypes = [vp] * 8 + [ctypes.c_size_t, pp, vp]
    lib.dcnv3_blend_supported.argtypes = [pp]
    lib.dcnv3_dcnv2_sample_workspace_bytes.argtypes = [ci] * 4
    lib.dcnv3_dcnv2_sample_workspace_bytes.restype = ctypes.c_size_t
    lib.dcnv3_dcnv2_sample_forward.argtypes = [vp] * 4 + [ci] * 7 + [vp]
    lib.dcnv3_dcnv2_sample_backward.argtypes = [vp] * 8 + [ctypes.c_size_t] + [ci] * 7 + [cu, vp]
    lib.dcnv3_deform_attn_workspace_bytes.argtypes = [ci] * 5
    lib.dcnv3_deform_attn_workspace_bytes.restype = ctypes.c_size_t
    lib.dcnv3_deform_attn_forward.argtypes = [vp] * 5 + [ci] * 7 + [vp]
    lib.dcnv3_deform_attn_backward.argtypes = [vp] * 10 + [ctypes.c_size_t] + [ci] * 7 + [cu, vp]
    lib.dcnv3_dwconv_ln_act.argtypes = [vp] * 6 + [ci] * 6 + [cf, ci, ci, vp]
    lib.dcnv3_layer_join.argtypes = [vp] * 7 + [ctypes.c_int64, ci, cf, ci, ci, vp]
    lib.dcnv3_forward_blend.argtypes = [vp] * 5 + [pp, vp]
    lib.dcnv3_backward_blend.argtypes = [vp] * 10 + [ctypes.c_size_t, pp, vp]
    scal = [ci] * 10 + [cf, cu, vp]
    lib.dcnv3_forward_dlpack.argtypes = [vp] * 4 + scal
    lib.dcnv3_backward_dlpack.argtypes = [vp] * 8 + scal
    lib.dcnv3_forward_host.argtypes = [vp] * 4 + [pp, ci]
    lib.dcnv3_forward_backward_host.argtypes = [vp] * 8 + [pp, ci]
    lib.dcnv3_forward_backward_host_async.argtypes = [vp] * 8 + [pp, ci, ci]
    lib.dcnv3_host_sync.argtypes = [ci]
    lib.dcnv3_set_kernel_timing.argtypes = [ci]
    lib.dcnv3_get_kernel_timing.argtypes = [ctypes.POINTER(ctypes.c_float)]
    lib.dcnv3_kernel_launch_count.restype = ctypes.c_uint64
    if lib.dcnv3_abi_version() != 1:
        raise ImportError("libdcnv3_b200.so ABI version mismatch")
    return lib


lib = _load()

_capsule_ptr = ctypes.pythonapi.PyCapsule_GetPointer
_capsule_ptr.restype = ctypes.c_void_p
_capsule_ptr.argtypes = [ctypes.py_object, ctypes.c_char_p]


def _dl(t):
    """(capsule, DLManagedTensor*) -- the capsule keeps the export alive for the call."""
    cap = to_dlpack(t)
    return cap, _capsule_ptr(cap, b"dltensor")


def check(rc):
    if rc != 0:
        msg = lib.dcnv3_last_error().decode()
        if rc == ERR_SHAPE:
            raise ValueError(f"dcnv3_b200: {msg}")
        if rc == ERR_DTYPE:
            raise TypeError(f"dcnv3_b200: {msg}")
        raise DCNv3Error(rc, msg)


def launch_plan(params):
    """dcnv3_launch_plan as a dict (CPU-only introspection of the tiling)."""
    buf = (ctypes.c_int * 25)()
    check(lib.dcnv3_launch_plan(ctypes.byref(params), buf))
    v = list(buf)
    keys = ("th", "tw", "bw", "bh", "halo_x", "halo_y", "ctas", "smem")
    return {"tiled": bool(v[0]), "forward": dict(zip(keys, v[1:9])), "gather": dict(zip(keys, v[9:17])),
            "scatter": dict(zip(("tj", "ring_lo", "ring_hi", "box_rows", "ctas", "smem", "threads", "merge"), v[17:25]))}


def launch_count():
    return int(lib.dcnv3_kernel_launch_count())


def make_params(x_shape, out_hw, kernel_size, strides, pad, dilation_rate, groups, group_channels,
                offset_scale, dtype, flags=0):
    n, h, w, _ = x_shape
    return Params(n, h, w, out_hw[0], out_hw[1], groups, group_channels, kernel_size[0],
                  kernel_size[1], strides[0], strides[1], pad[0], pad[1], dilation_rate[0],
                  dilation_rate[1], float(offset_scale), dtype, flags)


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


_ws_cache = {}


def _ws_key(device, nbytes):
    return (device.index, torch.cuda.current_stream(device).cuda_stream, int(nbytes))


def _workspace(device, nbytes, zero_bytes=None):
    """Backward workspace, zeroed once and kept per (device, stream, size): dcnv3_backward leaves its zero
    part (the first `zero_bytes`) zeroed, so later calls on the same stream skip the memset
    (DCNV3_FLAG_WORKSPACE_ZEROED); what lies behind the zero part is scratch.  Two shapes can share a cache
    entry (same total size): the entry remembers how long its known-zero prefix is and a call that needs a
    longer one re-zeroes the difference.
    A workspace first needed during CUDA-graph capture lives in the graph's private pool: it is handed
    out zeroed but never cached.  A failed backward drops its cache entry (see backward())."""
    zero_bytes = int(nbytes) if zero_bytes is None else int(zero_bytes)
    key = _ws_key(device, nbytes)
    ent = _ws_cache.get(key)
    if ent is None:
        ws = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        if not torch.cuda.is_current_stream_capturing():
            if len(_ws_cache) >= 64:
                _ws_cache.clear()
            _ws_cache[key] = [ws, zero_bytes]
        return ws
    ws, clean = ent
    if zero_bytes > clean:
        ws[clean:zero_bytes].zero_()
    ent[1] = zero_bytes
    return ws


def forward(x, offset, mask, kernel_size, strides, pad, dilation_rate, groups, group_channels,
            offset_scale, flags=0):
    """dcnv3_forward_dlpack on torch CUDA tensors; returns a fresh output tensor."""
    if not (x.is_cuda and offset.is_cuda and mask.is_cuda):
        raise DCNv3Error(ERR_DEVICE, "dcnv3_op needs CUDA tensors (no CPU fallback)")
    out = torch.empty((x.shape[0], offset.shape[1], offset.shape[2], groups * group_channels),
                      dtype=x.dtype, device=x.device)
    caps = [_dl(t) for t in (x, offset, mask, out)]
    with torch.cuda.device(x.device):
        rc = lib.dcnv3_forward_dlpack(
            *[c[1] for c in caps], kernel_size[0], kernel_size[1], strides[0], strides[1], pad[0],
            pad[1], dilation_rate[0], dilation_rate[1], groups, group_channels, float(offset_scale),
            flags, _stream(x))
    check(rc)
    return out


def backward(x, offset, mask, grad_out, kernel_size, strides, pad, dilation_rate, groups,
             group_channels, offset_scale, flags=0):
    """dcnv3_backward_dlpack; returns (grad_x, grad_offset, grad_mask)."""
    gx, goff, gm = torch.empty_like(x), torch.empty_like(offset), torch.empty_like(mask)
    dt = F32 if x.dtype == torch.float32 else BF16
    p = make_params(x.shape, offset.shape[1:3], kernel_size, strides, pad, dilation_rate, groups,
                    group_channels, offset_scale, dt, flags)
    ws_bytes = int(lib.dcnv3_backward_workspace_bytes(ctypes.byref(p)))
    ws = _workspace(x.device, ws_bytes, int(lib.dcnv3_backward_workspace_zero_bytes(ctypes.byref(p))))
    rc = None
    try:
        caps = [_dl(t) for t in (x, offset, mask, grad_out, gx, goff, gm, ws)]
        with torch.cuda.device(x.device):
            rc = lib.dcnv3_backward_dlpack(
                *[c[1] for c in caps], kernel_size[0], kernel_size[1], strides[0], strides[1], pad[0],
                pad[1], dilation_rate[0], dilation_rate[1], groups, group_channels, float(offset_scale),
                flags | FLAG_WORKSPACE_ZEROED, _stream(x))
    finally:
        if rc != 0:  # error or exception: the workspace may be dirty -- the next call starts from a fresh one
            _ws_cache.pop(_ws_key(x.device, ws_bytes), None)
    check(rc)
    return gx, goff, gm


def _blend_params(x, offset, kernel_size, strides, pad, dilation_rate, groups, group_channels, offset_scale, flags):
    dt = F32 if x.dtype == torch.float32 else BF16
    return make_params(x.shape, offset.shape[1:3], kernel_size, strides, pad, dilation_rate, groups,
                       group_channels, offset_scale, dt, flags)


def blend_supported(x, offset, kernel_size, strides, pad, dilation_rate, groups, group_channels,
                    offset_scale, flags=0):
    """Can the centre-feature-scale blend (reference dcn_v3.py:138-146) run fused inside the kernels for this
    configuration?  (Where the shared-memory tiled kernels run.)"""
    if x.dtype not in (torch.float32, torch.bfloat16) or not x.is_cuda:
        return False
    p = _blend_params(x, offset, kernel_size, strides, pad, dilation_rate, groups, group_channels, offset_scale, flags)
    return bool(lib.dcnv3_blend_supported(ctypes.byref(p)))


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def forward_blend(x, offset, mask, center_scale, kernel_size, strides, pad, dilation_rate, groups,
                  group_channels, offset_scale, flags=0):
    """dcnv3_forward_blend on torch CUDA tensors: core * (1 - s) + x * s, one launch."""
    for t in (x, offset, mask, center_scale):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == x.dtype and t.device == x.device):
            raise DCNv3Error(ERR_LAYOUT, "dcnv3_forward_blend needs dense CUDA tensors of one dtype on one device")
    n, ho, wo = offset.shape[:3]
    if tuple(center_scale.shape) != (n, ho, wo, groups) or tuple(x.shape[1:3]) != (ho, wo):
        raise ValueError("dcnv3_b200: center_scale must be [N, H, W, groups] and x must have the output's H, W")
    p = _blend_params(x, offset, kernel_size, strides, pad, dilation_rate, groups, group_channels, offset_scale, flags)
    out = torch.empty((n, ho, wo, groups * group_channels), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.dcnv3_forward_blend(_ptr(x), _ptr(offset), _ptr(mask), _ptr(center_scale), _ptr(out),
                                      ctypes.byref(p), _stream(x)))
    return out


def backward_blend(x, offset, mask, center_scale, grad_out, kernel_size, strides, pad, dilation_rate,
                   groups, group_channels, offset_scale, flags=0):
    """dcnv3_backward_blend; returns (grad_x, grad_offset, grad_mask, grad_center_scale)."""
    for t in (x, offset, mask, center_scale, grad_out):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == x.dtype and t.device == x.device):
            raise DCNv3Error(ERR_LAYOUT, "dcnv3_backward_blend needs dense CUDA tensors of one dtype on one device")
    gx, goff, gm, gs = (torch.empty_like(t) for t in (x, offset, mask, center_scale))
    p = _blend_params(x, offset, kernel_size, strides, pad, dilation_rate, groups, group_channels, offset_scale,
                      flags | FLAG_WORKSPACE_ZEROED)
    ws_bytes = int(lib.dcnv3_backward_workspace_bytes(ctypes.byref(p)))
    ws = _workspace(x.device, ws_bytes, int(lib.dcnv3_backward_workspace_zero_bytes(ctypes.byref(p))))
    rc = None
    try:
        with torch.cuda.device(x.device):
            rc = lib.dcnv3_backward_blend(_ptr(x), _ptr(offset), _ptr(mask), _ptr(center_scale), _ptr(grad_out),
                                          _ptr(gx), _ptr(goff), _ptr(gm), _ptr(gs), _ptr(ws), ws_bytes,
                                          ctypes.byref(p), _stream(x))
    finally:
        if rc != 0:
            _ws_cache.pop(_ws_key(x.device, ws_bytes), None)
    check(rc)
    return gx, goff, gm, gs


# ---- inference fast path of the layers around the op (include/dcnv3_b200.h: dcnv3_dwconv_ln_act, dcnv3_layer_join) ----
def fused_layers_usable(x):
    """The one-pass kernels serve dense CUDA fp32 / bf16 activations with channels % 4 == 0 (<= 4096), without
    autograd recording (they have no backward)."""
    return (x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and x.shape[-1] % 4 == 0
            and x.shape[-1] <= 4096 and not torch.is_grad_enabled())


def _dt(t):
    return F32 if t.dtype == torch.float32 else BF16


def _optr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def dwconv_ln_act(x, weight_kkc, bias, ln_weight, ln_bias, k, pad_lo, eps, gelu=True):
    """act(LayerNorm(DepthwiseConv2D(x) + bias)) on NHWC in one pass (reference dcn_v3.py:115-117)."""
    x = x.contiguous()
    n, h, w, c = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib.dcnv3_dwconv_ln_act(_ptr(x), _ptr(weight_kkc), _optr(bias), _ptr(ln_weight), _ptr(ln_bias), _ptr(out),
                                      n, h, w, c, int(k), int(pad_lo), float(eps), 1 if gelu else 0, _dt(x), _stream(x)))
    return out


def layer_join(y, residual, gamma, ln_weight, ln_bias, eps, mode, want_norm=True):
    """The joins of InternImageLayer (reference intern_image_layer.py:126-172) in one pass.
    mode 0: (residual + gamma * y, LayerNorm of that or None); mode 1: residual + gamma * LayerNorm(y); mode 2: LayerNorm(y)."""
    y = y.contiguous()
    c = y.shape[-1]
    rows = y.numel() // c
    residual = None if residual is None else residual.contiguous()
    out_sum = torch.empty_like(y)
    out_norm = torch.empty_like(y) if (mode == 0 and want_norm) else None
    with torch.cuda.device(y.device):
        check(lib.dcnv3_layer_join(_ptr(y), _optr(residual), _optr(gamma), _optr(ln_weight), _optr(ln_bias), _ptr(out_sum),
                                   _optr(out_norm), rows, c, float(eps), int(mode), _dt(y), _stream(y)))
    return (out_sum, out_norm) if mode == 0 else out_sum
