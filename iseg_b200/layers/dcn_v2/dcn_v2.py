"""`DCNv2` -- torch re-statement of the reference Keras layer (reference layers/dcn_v2.py:15-300) around a CUDA sampler.

A sibling of the DCNv3 op.  `dcnv2_sample` is the stage of `_forward` between the offset
convolution and the contraction with the kernel (:137-247) as one op behind the C ABI (`dcnv3_dcnv2_sample_forward /
_backward`): per output pixel and tap, a mask-weighted bilinear sample of the zero-padded input, with that function's
own conventions (tap order, clipping of neighbours and coordinate to [0, H+1] x [0, W+1], weights from the clipped
values).  The two dense stages around it are stock torch: the offset convolution (:128-135, cuDNN) and the
[B, H*W, ks*C] x [ks*C, filters] contraction (:249-265, cuBLAS).  Weights keep the Keras layouts and names (kernel,
bias, offset_kernel, offset_bias).  Deterministic gradients; no CPU fallback.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _cabi


class _DCNv2Sample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, offsets, mask, kh, kw):
        x, offsets, mask = x.contiguous(), offsets.contiguous(), mask.contiguous()
        n, h, w, c = x.shape
        out = torch.empty((n, h, w, kh * kw, c), dtype=x.dtype, device=x.device)
        dt = _cabi.F32 if x.dtype == torch.float32 else _cabi.BF16
        with torch.cuda.device(x.device):
            _cabi.check(_cabi.lib.dcnv3_dcnv2_sample_forward(_cabi._ptr(x), _cabi._ptr(offsets), _cabi._ptr(mask), _cabi._ptr(out),
                                                             n, h, w, c, kh, kw, dt, _cabi._stream(x)))
        ctx.save_for_backward(x, offsets, mask)
        ctx.k = (kh, kw)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, offsets, mask = ctx.saved_tensors
        kh, kw = ctx.k
        grad_out = grad_out.contiguous()
        n, h, w, c = x.shape
        gx, goff, gm = torch.empty_like(x), torch.empty_like(offsets), torch.empty_like(mask)
        dt = _cabi.F32 if x.dtype == torch.float32 else _cabi.BF16
        ws_bytes = int(_cabi.lib.dcnv3_dcnv2_sample_workspace_bytes(n, h, w, c))
        ws = _cabi._workspace(x.device, ws_bytes)
        rc = None
        try:
            with torch.cuda.device(x.device):
                rc = _cabi.lib.dcnv3_dcnv2_sample_backward(
                    _cabi._ptr(x), _cabi._ptr(offsets), _cabi._ptr(mask), _cabi._ptr(grad_out), _cabi._ptr(gx), _cabi._ptr(goff),
                    _cabi._ptr(gm), _cabi._ptr(ws), ws_bytes, n, h, w, c, kh, kw, dt, _cabi.FLAG_WORKSPACE_ZEROED, _cabi._stream(x))
        finally:
            if rc != 0:
                _cabi._ws_cache.pop(_cabi._ws_key(x.device, ws_bytes), None)
        _cabi.check(rc)
        return gx, goff, gm, None, None


def dcnv2_sample(x, offsets, mask, kernel_size):
    """x [N,H,W,C]; offsets [N,H,W,ks,2] ((oy, ox) per tap); mask [N,H,W,ks] (after the sigmoid) -> [N,H,W,ks,C]
    (the reference's `map_all` before its reshape, dcn_v2.py:245-247)."""
    kh, kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
    ks = kh * kw
    if x.dim() != 4 or tuple(offsets.shape) != (*x.shape[:3], ks, 2) or tuple(mask.shape) != (*x.shape[:3], ks):
        raise ValueError("dcnv2_sample: x [N,H,W,C], offsets [N,H,W,kh*kw,2], mask [N,H,W,kh*kw]")
    for t in (x, offsets, mask):
        if not t.is_cuda:
            raise _cabi.DCNv3Error(_cabi.ERR_DEVICE, "dcnv2_sample needs CUDA tensors (no CPU fallback)")
        if t.dtype != x.dtype or x.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("dcnv2_sample: float32 or bfloat16 tensors of one dtype")
    return _DCNv2Sample.apply(x, offsets, mask, kh, kw)


class DCNv2(nn.Module):
    def __init__(self, filters, kernel_size, dilation_rate=1, use_bias=True, use_custom_offset=False, activation=None,
                 use_jit_compile=False, name=None, input_channels=None):
        super().__init__()
        self.filters = filters
        self.kernel_size = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        self.dilation = (dilation_rate, dilation_rate) if isinstance(dilation_rate, int) else tuple(dilation_rate)
        self.use_bias, self.use_custom_offset, self.name = use_bias, use_custom_offset, name
        self.activation = activation if activation is not None else (lambda t: t)
        self.built = False
        if input_channels is not None:
            self.build((None, None, None, input_channels))

    def build(self, input_shape):  # dcn_v2.py:61-113
        ic = int(input_shape[0][-1] if self.use_custom_offset else input_shape[-1])
        kh, kw = self.kernel_size
        self.kernel = nn.Parameter(torch.empty(kh, kw, ic, self.filters))
        nn.init.xavier_uniform_(self.kernel.view(kh * kw * ic, self.filters))  # glorot_uniform over (fan_in, fan_out)
        if self.use_bias:
            self.bias = nn.Parameter(torch.zeros(self.filters))
        self.offset_kernel = nn.Parameter(torch.zeros(kh, kw, ic, 3 * kh * kw))  # zero-initialised (:91-104)
        self.offset_bias = nn.Parameter(torch.zeros(3 * kh * kw))
        self.built = True

    def forward(self, inputs, training=None):
        x, offset = tuple(inputs) if self.use_custom_offset else (inputs, inputs)  # :275-278
        if not self.built:
            self.build([x.shape, offset.shape] if self.use_custom_offset else x.shape)
            self.to(device=x.device, dtype=x.dtype)
        kh, kw = self.kernel_size
        ks = kh * kw
        n, h, w, ic = x.shape
        pad = (self.dilation[0] * (kh - 1) // 2, self.dilation[1] * (kw - 1) // 2)
        off = F.conv2d(offset.permute(0, 3, 1, 2), self.offset_kernel.to(x.dtype).permute(3, 2, 0, 1), None, 1, pad,
                       self.dilation).permute(0, 2, 3, 1) + self.offset_bias.to(x.dtype)                      # :128-135
        oyox = off[..., :2 * ks].reshape(n, h, w, ks, 2)                                                    # :144-146
        mask = torch.sigmoid(off[..., 2 * ks:])                                                             # :148
        map_all = dcnv2_sample(x, oyox, mask, (kh, kw)).reshape(n, h * w, ks * ic)                          # :150-247
        out = torch.matmul(map_all, self.kernel.to(x.dtype).reshape(ks * ic, self.filters)).reshape(n, h, w, self.filters)
        if self.use_bias:
            out = out + self.bias.to(x.dtype)                                                               # :267-268
        return self.activation(out)

    call = forward
