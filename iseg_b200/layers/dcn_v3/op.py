"""`dcnv3_op` -- host-side mirror of the reference's operator entry point
(reference layers/dcn_v3/op.py:16-27): same name, argument order, argument meaning and error
behaviour, NHWC tensors in and out.  The arithmetic runs in the sm_100a kernels behind the C ABI;
the gradient the reference gets from TF autodiff is registered here as a torch.autograd.Function.
"""
import torch

from ... import _cabi


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


class _DCNv3Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, offset, mask, cfg):
        x, offset, mask = x.contiguous(), offset.contiguous(), mask.contiguous()
        ctx.cfg = cfg
        ctx.save_for_backward(x, offset, mask)
        return _cabi.forward(x, offset, mask, *cfg)

    @staticmethod
    def backward(ctx, grad_out):
        x, offset, mask = ctx.saved_tensors
        gx, goff, gm = _cabi.backward(x, offset, mask, grad_out.contiguous(), *ctx.cfg)
        return gx, goff, gm, None


class _DCNv3BlendFunction(torch.autograd.Function):
    """The op with the layer's centre-feature-scale blend (dcn_v3.py:138-146) inside the kernels."""

    @staticmethod
    def forward(ctx, x, offset, mask, center_scale, cfg):
        x, offset, mask, center_scale = (t.contiguous() for t in (x, offset, mask, center_scale))
        ctx.cfg = cfg
        ctx.save_for_backward(x, offset, mask, center_scale)
        return _cabi.forward_blend(x, offset, mask, center_scale, *cfg)

    @staticmethod
    def backward(ctx, grad_out):
        x, offset, mask, center_scale = ctx.saved_tensors
        gx, goff, gm, gs = _cabi.backward_blend(x, offset, mask, center_scale, grad_out.contiguous(), *ctx.cfg)
        return gx, goff, gm, gs, None


def dcnv3_op_center_scale(x, offset, mask, center_scale, kernel_size, strides, padding, dilation_rate, groups,
                          group_channels, offset_scale, mask_is_logits=False):
    """`dcnv3_op` followed by the layer's centre-feature-scale blend (reference dcn_v3.py:138-146):
    x_core * (1 - s) + x * s with s = center_scale [N,H,W,groups] broadcast over each group's channels.
    One forward launch and no extra backward launch where the tiled kernels run; elsewhere (other kernel sizes,
    32 channels per group, ...) the blend is applied with torch operations around `dcnv3_op`."""
    pad = _resolve_padding(kernel_size, padding)
    cfg = (tuple(int(k) for k in kernel_size), tuple(int(s) for s in strides), pad,
           tuple(int(d) for d in dilation_rate), int(groups), int(group_channels),
           float(offset_scale), _cabi.FLAG_MASK_LOGITS if mask_is_logits else 0)
    if tuple(x.shape[1:3]) == tuple(offset.shape[1:3]) and _cabi.blend_supported(x, offset, *cfg):
        if offset.dtype != x.dtype or mask.dtype != x.dtype or center_scale.dtype != x.dtype:
            raise TypeError("x, offset, mask and center_scale must have the same dtype")
        return _DCNv3BlendFunction.apply(x, offset, mask, center_scale, cfg)
    core = dcnv3_op(x, offset, mask, kernel_size, strides, padding, dilation_rate, groups, group_channels,
                    offset_scale, mask_is_logits=mask_is_logits)
    n, h, w, c = core.shape
    s = center_scale.unsqueeze(-1).expand(n, h, w, groups, group_channels).reshape(n, h, w, c)
    return core * (1 - s) + x * s


def dcnv3_op(x, offset, mask, kernel_size, strides, padding, dilation_rate, groups, group_channels,
             offset_scale, mask_is_logits=False, reference_dtype_math=False):
    """x [N,H,W,G*gc], offset [N,Ho,Wo,G*P*2], mask [N,Ho,Wo,G*P] -> [N,Ho,Wo,G*gc] (dtype of x).

    `mask_is_logits=True` (extension, not in the reference signature) fuses the softmax over the P
    taps that the layer applies just before the call (dcn_v3.py:120-123).
    `reference_dtype_math=True` (extension; bf16 tensors only) rounds every intermediate of the coordinate,
    weight and accumulation arithmetic to bf16 as the reference does under mixed_bfloat16 (op.py:62-87,
    utils.py:130-206); the default keeps that arithmetic in fp32."""
    pad = _resolve_padding(kernel_size, padding)
    if offset.dtype != x.dtype or mask.dtype != x.dtype:
        raise TypeError("x, offset and mask must have the same dtype")
    cfg = (tuple(int(k) for k in kernel_size), tuple(int(s) for s in strides), pad,
           tuple(int(d) for d in dilation_rate), int(groups), int(group_channels),
           float(offset_scale), (_cabi.FLAG_MASK_LOGITS if mask_is_logits else 0) |
           (_cabi.FLAG_REF_DTYPE if reference_dtype_math else 0))
    return _DCNv3Function.apply(x, offset, mask, cfg)
