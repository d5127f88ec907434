"""InternImage backbone wiring around the B200 DCNv3 op (torch, NHWC end to end).

Host-side re-statement of reference backbones/intern_image/ (intern_image.py:16-135 model and presets
:137-187, intern_image_block.py:15-123, intern_image_layer.py:17-174, stem_layer.py:13-69,
dowmsample_layer.py:12-43, mlp_layer.py:10-59, utils/drops.py:8-22).  Only the DCNv3 core op is
hand-written CUDA (iseg_b200.layers.dcn_v3); the dense / conv / norm layers around it are stock torch
(cuBLAS / cuDNN) -- they are not on the hot path this repo accelerates.  Constructor arguments, the
three layer variants (pre-norm / post-norm / res-post-norm), endpoints and drop-path schedule follow
the reference so that weights map one to one by name.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _cabi
from ...layers.dcn_v3.dcn_v3 import DeformableConvolutionV3

LN_EPS = 1e-6


def _conv_same_s2(conv, x_nhwc):
    """Keras Conv2D(k=3, strides=2, padding='same') on NHWC: TF pads (0,1) when the size is even and
    (1,1) when it is odd -- not torch's symmetric padding=1."""
    _, h, w, _ = x_nhwc.shape
    ph = max((-(-h // 2) - 1) * 2 + 3 - h, 0)
    pw = max((-(-w // 2) - 1) * 2 + 3 - w, 0)
    x = x_nhwc.permute(0, 3, 1, 2)
    x = F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
    return conv(x).permute(0, 2, 3, 1)


def _layer_norm(norm, x):
    """LayerNorm over the channels: the one-pass kernel (`dcnv3_layer_join` mode 2) on the inference path, torch otherwise."""
    if not norm.training and _cabi.fused_layers_usable(x) and norm.weight.dtype == x.dtype:
        return _cabi.layer_join(x, None, None, norm.weight, norm.bias, norm.eps, 2)
    return norm(x)


def drop_path(x, drop_prob, training):
    """reference utils/drops.py:8-22 (per-sample, training only)."""
    if not training or drop_prob == 0.0:
        return x
    keep = 1.0 - drop_prob
    shape = (x.shape[0],) + (1,) * (x.dim() - 1)
    mask = torch.floor(keep + torch.rand(shape, dtype=x.dtype, device=x.device))
    return x / keep * mask


class StemLayer(nn.Module):
    def __init__(self, filters, activation, in_channels=3):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, filters // 2, 3, stride=2)
        self.norm1 = nn.LayerNorm(filters // 2, eps=LN_EPS)
        self.conv2 = nn.Conv2d(filters // 2, filters, 3, stride=2)
        self.norm2 = nn.LayerNorm(filters, eps=LN_EPS)
        self.activation = activation

    def forward(self, x):
        x = self.activation(_layer_norm(self.norm1, _conv_same_s2(self.conv1, x)))
        mid = x
        return _layer_norm(self.norm2, _conv_same_s2(self.conv2, x)), mid


class DownsampleLayer(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels * 2, 3, stride=2, bias=False)
        self.norm = nn.LayerNorm(channels * 2, eps=LN_EPS)

    def forward(self, x):
        return _layer_norm(self.norm, _conv_same_s2(self.conv, x))


class MLPLayer(nn.Module):
    def __init__(self, channels, hidden, activation, dropout_rate=0.0):
        super().__init__()
        self.fc1 = nn.Linear(channels, hidden)
        self.fc2 = nn.Linear(hidden, channels)
        self.dropout = nn.Dropout(dropout_rate)
        self.activation = activation

    def forward(self, x):
        return self.dropout(self.fc2(self.dropout(self.activation(self.fc1(x)))))


class InternImageLayer(nn.Module):
    def __init__(self, channels, groups, mlp_ratio=4, dropout_rate=0.0, drop_path_rate=0.0, activation=F.gelu,
                 use_post_norm=False, layer_scale=None, offset_scale=1.0, depthwise_kernel_size=None,
                 use_res_post_norm=False, center_feature_scale=False):
        super().__init__()
        self.use_post_norm, self.use_res_post_norm = use_post_norm, use_res_post_norm
        self.drop_path_rate = float(drop_path_rate)
        self.norm1 = nn.LayerNorm(channels, eps=LN_EPS)
        self.dcn = DeformableConvolutionV3(filters=channels, kernel_size=3, depthwise_kernel_size=depthwise_kernel_size,
                                           strides=1, padding="same", dilation_rate=1, groups=groups,
                                           offset_scale=offset_scale, activation=activation,
                                           center_feature_scale=center_feature_scale, input_channels=channels)
        self.norm2 = nn.LayerNorm(channels, eps=LN_EPS)
        self.mlp = MLPLayer(channels, int(channels * mlp_ratio), activation, dropout_rate)
        if layer_scale is not None:
            assert not use_res_post_norm, "use_res_post_norm and layer_scale can not be used at the same time"
            self.gamma1 = nn.Parameter(torch.ones(channels))  # reference initialises to ones regardless of the value
            self.gamma2 = nn.Parameter(torch.ones(channels))
        else:
            self.gamma1 = self.gamma2 = 1.0
        if use_res_post_norm:
            self.res_post_norm1 = nn.LayerNorm(channels, eps=LN_EPS)
            self.res_post_norm2 = nn.LayerNorm(channels, eps=LN_EPS)

    def _g(self, gamma, x):
        return gamma.to(x.dtype) if torch.is_tensor(gamma) else gamma

    def _gamma(self, gamma, x):
        return gamma.detach().to(x.dtype) if torch.is_tensor(gamma) else None

    def forward_fused(self, x, x_norm1=None, next_norm=None):
        """Inference fast path: the same arithmetic as `forward` in eval mode, with every
        join of the layer -- layer scale, residual add and the LayerNorm next to it -- done by one pass of
        `dcnv3_layer_join` instead of three element-wise kernels.  `x_norm1`: norm1(x) if the previous layer's last
        join already produced it; `next_norm`: the LayerNorm that will consume this layer's output first (the next
        pre-norm layer's norm1), fused into the last join.  Returns (out, next_norm(out) or None)."""
        j, eps = _cabi.layer_join, LN_EPS
        g1, g2 = self._gamma(self.gamma1, x), self._gamma(self.gamma2, x)
        if self.use_post_norm:  # :126-139
            z = j(self.dcn(x), x, g1, self.norm1.weight, self.norm1.bias, eps, 1)
            return j(self.mlp(z), z, g2, self.norm2.weight, self.norm2.bias, eps, 1), None
        if self.use_res_post_norm:  # :142-156
            n1 = j(x, None, None, self.norm1.weight, self.norm1.bias, eps, 2)
            z = j(self.dcn(n1), x, None, self.res_post_norm1.weight, self.res_post_norm1.bias, eps, 1)
            n2 = j(z, None, None, self.norm2.weight, self.norm2.bias, eps, 2)
            return j(self.mlp(n2), z, None, self.res_post_norm2.weight, self.res_post_norm2.bias, eps, 1), None
        # pre-norm, :158-172
        if x_norm1 is None:
            x_norm1 = j(x, None, None, self.norm1.weight, self.norm1.bias, eps, 2)
        z, n2 = j(self.dcn(x_norm1), x, g1, self.norm2.weight, self.norm2.bias, eps, 0)
        if next_norm is None:
            return j(self.mlp(n2), z, g2, None, None, eps, 0, want_norm=False)[0], None
        return j(self.mlp(n2), z, g2, next_norm.weight, next_norm.bias, eps, 0)

    def forward(self, x):
        if not self.training and _cabi.fused_layers_usable(x) and self.norm1.weight.dtype == x.dtype:
            return self.forward_fused(x)[0]
        t, dp, residual = self.training, self.drop_path_rate, x
        if self.use_post_norm:  # intern_image_layer.py:126-139
            x = drop_path(self.norm1(self.dcn(x)) * self._g(self.gamma1, x), dp, t)
            residual = x = residual + x
            x = drop_path(self.norm2(self.mlp(x)) * self._g(self.gamma2, x), dp, t)
        elif self.use_res_post_norm:  # :142-156
            x = drop_path(self.res_post_norm1(self.dcn(self.norm1(x))), dp, t)
            residual = x = residual + x
            x = drop_path(self.res_post_norm2(self.mlp(self.norm2(x))), dp, t)
        else:  # pre-norm, :158-172
            x = drop_path(self.dcn(self.norm1(x)) * self._g(self.gamma1, x), dp, t)
            residual = x = residual + x
            x = drop_path(self.mlp(self.norm2(x)) * self._g(self.gamma2, x), dp, t)
        return x + residual


class InternImageBlock(nn.Module):
    def __init__(self, channels, depth, groups, use_downsample=True, drop_path_rate=0.0, use_post_norm=False,
                 post_norm_block_ids=None, center_feature_scale=False, **layer_kw):
        super().__init__()
        rates = drop_path_rate if isinstance(drop_path_rate, (list, tuple)) else [drop_path_rate] * depth
        self.blocks = nn.ModuleList(
            InternImageLayer(channels, groups, drop_path_rate=rates[i], use_post_norm=use_post_norm,
                             center_feature_scale=center_feature_scale, **layer_kw) for i in range(depth))
        self.norm = nn.LayerNorm(channels, eps=LN_EPS) if (not use_post_norm or center_feature_scale) else None
        self.post_norm_block_ids = post_norm_block_ids
        if post_norm_block_ids is not None:
            self.post_norms = nn.ModuleList(nn.LayerNorm(channels, eps=LN_EPS) for _ in post_norm_block_ids)
        self.downsample = DownsampleLayer(channels) if use_downsample else None

    def forward(self, x):
        fused = (not self.training and _cabi.fused_layers_usable(x) and self.blocks[0].norm1.weight.dtype == x.dtype)
        pending = None  # norm1(x) of the coming layer, already produced by the previous layer's last join
        normed_is_final = False
        for i, blk in enumerate(self.blocks):
            level2 = self.post_norm_block_ids is not None and i in self.post_norm_block_ids
            if fused:
                # a pre-norm layer's first LayerNorm is fused into the join that ends the layer before it; the stage's
                # closing LayerNorm into the join that ends its last layer
                nxt = None
                if not (blk.use_post_norm or blk.use_res_post_norm) and not level2:
                    nxt = self.blocks[i + 1].norm1 if i + 1 < len(self.blocks) else self.norm
                x, normed = blk.forward_fused(x, pending, nxt)
                pending = normed if i + 1 < len(self.blocks) else None
                if i + 1 == len(self.blocks) and normed is not None:
                    x, normed_is_final = normed, True
                else:
                    normed_is_final = False
            else:
                x = blk(x)
                normed_is_final = False
            if level2:
                x = _layer_norm(self.post_norms[self.post_norm_block_ids.index(i)], x)
        if self.norm is not None and not (fused and normed_is_final):
            x = _layer_norm(self.norm, x)
        before = x
        if self.downsample is not None:
            x = self.downsample(x)
        return x, before


class InternImage(nn.Module):
    def __init__(self, stem_filters=64, depths=(3, 4, 18, 5), groups=(3, 6, 12, 24), mlp_ratio=4, dropout_rate=0.0,
                 drop_path_rate=0.2, drop_path_type="linear", activation=F.gelu, layer_scale=None, offset_scale=1.0,
                 use_post_norm=False, depthwise_kernel_size=None, use_level2_post_norm=False,
                 level2_post_norm_block_ids=None, use_res_post_norm=False, use_center_feature_scale=False,
                 return_endpoints=False, name=None, in_channels=3):
        super().__init__()
        self.name, self.return_endpoints = name, return_endpoints
        depths, groups = list(depths), list(groups)
        self.patch_embed = StemLayer(stem_filters, activation, in_channels)
        self.pos_drop = nn.Dropout(dropout_rate)
        if drop_path_type.lower() != "linear":
            raise ValueError(f"drop_path_type: {drop_path_type} not supported")
        dpr = [float(v) for v in np.linspace(0.0, drop_path_rate, sum(depths))]
        self.blocks = nn.ModuleList()
        for i, depth in enumerate(depths):
            self.blocks.append(InternImageBlock(
                stem_filters * 2 ** i, depth, groups[i], use_downsample=i < len(depths) - 1,
                drop_path_rate=dpr[sum(depths[:i]):sum(depths[:i + 1])], use_post_norm=use_post_norm,
                post_norm_block_ids=level2_post_norm_block_ids if (use_level2_post_norm and i == 2) else None,
                center_feature_scale=use_center_feature_scale, mlp_ratio=mlp_ratio, dropout_rate=dropout_rate,
                activation=activation, layer_scale=layer_scale, offset_scale=offset_scale,
                depthwise_kernel_size=depthwise_kernel_size, use_res_post_norm=use_res_post_norm))

    def forward(self, inputs):
        x, mid = self.patch_embed(inputs)
        x = self.pos_drop(x)
        endpoints = [mid]
        for blk in self.blocks:
            x, before = blk(x)
            endpoints.append(before)
        return endpoints if self.return_endpoints else x


# presets: tiny / small / huge exist in the reference (intern_image.py:137-187); base / large are the
# official OpenGVLab hyper-parameters through the same constructor
def intern_image_tiny(return_endpoints=False):
    return InternImage(64, [4, 4, 18, 4], [4, 8, 16, 32], 4.0, drop_path_rate=0.2, layer_scale=1.0, offset_scale=1.0,
                       use_post_norm=False, return_endpoints=return_endpoints, name="intern_image_tiny")


def intern_image_small(return_endpoints=False):
    return InternImage(80, [4, 4, 21, 4], [5, 10, 20, 40], 4.0, drop_path_rate=0.3, layer_scale=1.0, offset_scale=1.0,
                       use_post_norm=True, return_endpoints=return_endpoints, name="intern_image_small")


def intern_image_base(return_endpoints=False):
    return InternImage(112, [4, 4, 21, 4], [7, 14, 28, 56], 4.0, drop_path_rate=0.4, layer_scale=1.0, offset_scale=1.0,
                       use_post_norm=True, return_endpoints=return_endpoints, name="intern_image_base")


def intern_image_large(return_endpoints=False):
    return InternImage(160, [5, 5, 22, 5], [10, 20, 40, 80], 4.0, drop_path_rate=0.4, layer_scale=1.0, offset_scale=2.0,
                       use_post_norm=True, return_endpoints=return_endpoints, name="intern_image_large")


def intern_image_huge(return_endpoints=False):
    return InternImage(320, [6, 6, 32, 6], [10, 20, 40, 80], 4.0, drop_path_rate=0.5, layer_scale=None, offset_scale=1.0,
                       use_post_norm=False, depthwise_kernel_size=5, use_res_post_norm=True, use_level2_post_norm=True,
                       level2_post_norm_block_ids=[5, 11, 17, 23, 29], use_center_feature_scale=True,
                       return_endpoints=return_endpoints, name="intern_image_huge")
