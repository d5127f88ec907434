"""`deform_attn_sample` -- the sampling + aggregation step of iSeg's deformable multi-head self-attention
(reference layers/deformable_multihead_self_attention.py:102-175 `_bilinear_sample`, then :233-235) as one op:

    out[n,h,w,hd,:] = sum_p attn[n,h,w,hd,p] * bilinear(value[n,:,:,hd,:], y[n,h,w,hd,p], x[n,h,w,hd,p])

A sibling of `dcnv3_op`: the same gather + bilinear + weighted-sum skeleton with that
function's own conventions (absolute pixel coordinates, neighbour indices clamped to the image, weights from the
fractional parts).  Hand-written CUDA behind the C ABI (`dcnv3_deform_attn_forward / _backward`); gradients for all
four inputs, `grad_value` bitwise reproducible; no CPU fallback.
"""
import ctypes

import torch

from ... import _cabi


def _check(value, y, x, attn):
    if value.dim() != 5 or y.dim() != 5 or y.shape != x.shape or y.shape != attn.shape or y.shape[:4] != value.shape[:4]:
        raise ValueError("deform_attn_sample: value [N,H,W,heads,C], y / x / attn [N,H,W,heads,P]")
    for t in (value, y, x, attn):
        if not t.is_cuda:
            raise _cabi.DCNv3Error(_cabi.ERR_DEVICE, "deform_attn_sample needs CUDA tensors (no CPU fallback)")
        if t.dtype != value.dtype or value.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("deform_attn_sample: float32 or bfloat16 tensors of one dtype")


class _DeformAttnSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, value, y, x, attn):
        value, y, x, attn = (t.contiguous() for t in (value, y, x, attn))
        n, h, w, heads, c = value.shape
        p = y.shape[-1]
        out = torch.empty_like(value)
        dt = _cabi.F32 if value.dtype == torch.float32 else _cabi.BF16
        with torch.cuda.device(value.device):
            _cabi.check(_cabi.lib.dcnv3_deform_attn_forward(
                _cabi._ptr(value), _cabi._ptr(y), _cabi._ptr(x), _cabi._ptr(attn), _cabi._ptr(out), n, h, w, heads, p, c, dt,
                _cabi._stream(value)))
        ctx.save_for_backward(value, y, x, attn)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        value, y, x, attn = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        n, h, w, heads, c = value.shape
        p = y.shape[-1]
        gv, gy, gx, ga = torch.empty_like(value), torch.empty_like(y), torch.empty_like(x), torch.empty_like(attn)
        dt = _cabi.F32 if value.dtype == torch.float32 else _cabi.BF16
        ws_bytes = int(_cabi.lib.dcnv3_deform_attn_workspace_bytes(n, h, w, heads, c))
        ws = _cabi._workspace(value.device, ws_bytes)  # all of it is the zero part
        rc = None
        try:
            with torch.cuda.device(value.device):
                rc = _cabi.lib.dcnv3_deform_attn_backward(
                    _cabi._ptr(value), _cabi._ptr(y), _cabi._ptr(x), _cabi._ptr(attn), _cabi._ptr(grad_out), _cabi._ptr(gv),
                    _cabi._ptr(gy), _cabi._ptr(gx), _cabi._ptr(ga), _cabi._ptr(ws), ws_bytes, n, h, w, heads, p, c, dt,
                    _cabi.FLAG_WORKSPACE_ZEROED, _cabi._stream(value))
        finally:
            if rc != 0:
                _cabi._ws_cache.pop(_cabi._ws_key(value.device, ws_bytes), None)
        _cabi.check(rc)
        return gv, gy, gx, ga


def deform_attn_sample(value, y, x, attn):
    """value [N,H,W,heads,C]; y, x (pixel coordinates), attn [N,H,W,heads,P]  ->  [N,H,W,heads,C]."""
    _check(value, y, x, attn)
    return _DeformAttnSample.apply(value, y, x, attn)
