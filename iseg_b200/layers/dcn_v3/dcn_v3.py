"""`DeformableConvolutionV3` -- torch re-statement of the reference Keras layer that wraps the op
(reference layers/dcn_v3/dcn_v3.py:16-150): same constructor arguments, `call(inputs, training)`
semantics, NHWC in / NHWC out, same sub-layer names (input_proj, dw_conv, dw_conv_norm, offset,
mask, output_proj, center_feature_scale_proj; dcn_v3.py:62-102) so weights can be moved by name.

Everything but the core op is stock dense / depthwise-conv / layer-norm work (cuBLAS / cuDNN through
torch); the core op is `dcnv3_op` with the mask soft-max fused into the kernel.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _cabi
from .op import dcnv3_op, dcnv3_op_center_scale

LAYER_NORM_EPSILON = 1e-6


class DeformableConvolutionV3(nn.Module):
    def __init__(self, filters=64, kernel_size=3, depthwise_kernel_size=None, strides=1,
                 padding="SAME", dilation_rate=1, groups=4, offset_scale=1.0, activation=F.gelu,
                 center_feature_scale=False, name=None, input_channels=None, fuse_softmax=True):
        super().__init__()
        assert filters % groups == 0, "filters must be divisible by groups"
        self.filters = filters
        self.kernel_size = kernel_size
        self.depthwise_kernel_size = depthwise_kernel_size or kernel_size
        self.strides = strides
        self.padding = padding
        self.dilation_rate = dilation_rate
        self.groups = groups
        self.filters_per_group = filters // groups
        self.offset_scale = offset_scale
        self.activation = activation if activation is not None else (lambda t: t)
        self.center_feature_scale = center_feature_scale
        self.name = name
        self.fuse_softmax = fuse_softmax
        self.built = False
        if input_channels is not None:
            self.build((None, None, None, input_channels))

    def build(self, input_shape):
        cin = int(input_shape[-1])
        k = self.depthwise_kernel_size
        gp = self.groups * self.kernel_size * self.kernel_size
        # Keras DepthwiseConv2D 'same' at stride 1 pads (k-1)//2 before and k//2 after: asymmetric for an even k
        self.dw_conv = nn.Conv2d(cin, cin, k, stride=1, groups=cin, bias=True, padding=0)
        self._dw_pad = ((k - 1) // 2, k // 2) if self.padding.lower() == "same" else (0, 0)
        self.dw_conv_norm = nn.LayerNorm(cin, eps=LAYER_NORM_EPSILON)
        self.offset = nn.Linear(cin, 2 * gp)
        self.mask = nn.Linear(cin, gp)
        for lin in (self.offset, self.mask):  # zero-initialised (dcn_v3.py:74-86)
            nn.init.zeros_(lin.weight)
            nn.init.zeros_(lin.bias)
        self.input_proj = nn.Linear(cin, cin)
        self.output_proj = nn.Linear(cin, self.filters)
        for lin in (self.input_proj, self.output_proj):  # Keras Dense defaults
            nn.init.xavier_uniform_(lin.weight)
            nn.init.zeros_(lin.bias)
        if self.center_feature_scale:
            self.center_feature_scale_proj = nn.Linear(cin, self.groups)
            nn.init.xavier_uniform_(self.center_feature_scale_proj.weight)
            nn.init.zeros_(self.center_feature_scale_proj.bias)
        self.built = True

    def load_reference_weights(self, weights):
        """weights: {"<sublayer>.<var>": array} with Keras variable layouts (Dense kernel [in,out],
        DepthwiseConv2D kernel [k,k,C,1])."""
        def t(a):
            return torch.as_tensor(a)
        with torch.no_grad():
            for name in ("input_proj", "output_proj", "offset", "mask", "center_feature_scale_proj"):
                if f"{name}.kernel" in weights:
                    lin = getattr(self, name)
                    lin.weight.copy_(t(weights[f"{name}.kernel"]).t())
                    lin.bias.copy_(t(weights[f"{name}.bias"]))
            self.dw_conv.weight.copy_(t(weights["dw_conv.depthwise_kernel"]).permute(2, 3, 0, 1))
            self.dw_conv.bias.copy_(t(weights["dw_conv.bias"]))
            self.dw_conv_norm.weight.copy_(t(weights["dw_conv_norm.gamma"]))
            self.dw_conv_norm.bias.copy_(t(weights["dw_conv_norm.beta"]))

    def forward(self, inputs, training=False):
        if not self.built:
            self.build(inputs.shape)
            self.to(device=inputs.device, dtype=inputs.dtype)
        n, h, w, c = inputs.shape
        x_proj = self.input_proj(inputs)  # dcn_v3.py:113
        lo, hi = self._dw_pad
        k = self.depthwise_kernel_size
        if (self.activation is F.gelu and self.padding.lower() == "same" and _cabi.fused_layers_usable(inputs)
                and self.dw_conv.weight.dtype == inputs.dtype):
            # inference: depthwise conv + LayerNorm + GELU in one pass over the tensor (:115-117)
            wt = self.dw_conv.weight.detach().permute(2, 3, 0, 1).reshape(k * k, c).contiguous()  # [k*k][C], tap-major
            x1 = _cabi.dwconv_ln_act(inputs, wt, self.dw_conv.bias, self.dw_conv_norm.weight, self.dw_conv_norm.bias,
                                     k, lo, LAYER_NORM_EPSILON, gelu=True)
        else:
            x1 = self.dw_conv(F.pad(inputs.permute(0, 3, 1, 2), (lo, hi, lo, hi))).permute(0, 2, 3, 1)  # :115
            x1 = self.activation(self.dw_conv_norm(x1))  # :116-117
        offset = self.offset(x1)  # :118
        mask = self.mask(x1)  # :120
        if not self.fuse_softmax:  # :121-123
            mask = torch.softmax(mask.reshape(n, h, w, self.groups, -1), dim=-1).reshape(n, h, w, -1)
        args = ([self.kernel_size] * 2, [self.strides] * 2, self.padding, [self.dilation_rate] * 2, self.groups,
                self.filters_per_group, self.offset_scale)
        if self.center_feature_scale:  # :125-146: the op and the blend x * (1 - cfs) + x_proj * cfs in one call
            cfs = self.center_feature_scale_proj(x1)  # [N, H, W, groups]
            x = dcnv3_op_center_scale(x_proj, offset, mask, cfs, *args, mask_is_logits=self.fuse_softmax)
        else:
            x = dcnv3_op(x_proj, offset, mask, *args, mask_is_logits=self.fuse_softmax)  # :125-136
        return self.output_proj(x)  # :148

    call = forward
