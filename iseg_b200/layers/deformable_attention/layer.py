"""`DeformableMultiHeadSelfAttentionLayer` -- torch re-statement of the reference Keras layer around the sampler
(reference layers/deformable_multihead_self_attention.py:13-260): same constructor arguments and `call(inputs,
key=None, value=None)` semantics on NHWC tensors, same sub-layer names (value_proj, offset_proj, attn_proj).  The
1x1 convolutions / dense layers are stock torch; sampling and aggregation are `deform_attn_sample` (CUDA).
Not carried over: the NaN / Inf scrubbing and `check_numerics` calls the reference wraps around its tensors
(:182-186, :216-217, :241-242) -- debugging aids, identity on finite inputs."""
import torch
import torch.nn as nn

from .sample import deform_attn_sample


class DeformableMultiHeadSelfAttentionLayer(nn.Module):
    def __init__(self, filters=-1, num_heads=4, num_points=4, apply_linear=True, shared_qk=False, trainable=True,
                 use_dense_for_linear=False, offset_range_factor=8.0, use_jit_compile=False, name=None,
                 input_channels=None):
        super().__init__()
        self.filters, self.num_heads, self.num_points = filters, num_heads, num_points
        self.apply_linear, self.offset_range_factor, self.name = apply_linear, float(offset_range_factor), name
        self.built = False
        if input_channels is not None:
            self.build((None, None, None, input_channels))

    def build(self, input_shape):  # :66-87
        channels = int(input_shape[-1])
        value_filters = channels if self.filters == -1 else int(self.filters)
        if value_filters % self.num_heads != 0:
            raise ValueError(f"value filters ({value_filters}) must be divisible by num_heads ({self.num_heads}).")
        # Conv2D(1x1) and Dense are the same map on NHWC: one Linear either way
        if self.apply_linear:
            self.value_proj = nn.Linear(channels, value_filters)
        self.offset_proj = nn.Linear(channels, self.num_heads * self.num_points * 2)
        self.attn_proj = nn.Linear(channels, self.num_heads * self.num_points)
        self.built = True

    def forward(self, inputs, key=None, value=None, training=None):
        query = inputs
        value = query if value is None else value  # :254-259
        if not self.built:
            self.build(query.shape)
            self.to(device=query.device, dtype=query.dtype)
        n, h, w, _ = query.shape
        if self.apply_linear:
            value = self.value_proj(value)  # :189-190
        c_head = value.shape[-1] // self.num_heads
        offsets = torch.tanh(self.offset_proj(query).reshape(n, h, w, self.num_heads, self.num_points, 2))  # :196-205
        dy = offsets[..., 0] * (h / self.offset_range_factor)
        dx = offsets[..., 1] * (w / self.offset_range_factor)
        attn = torch.softmax(self.attn_proj(query).reshape(n, h, w, self.num_heads, self.num_points), dim=-1)  # :210-215
        ar = lambda k: torch.arange(k, device=query.device, dtype=query.dtype)  # noqa: E731
        y = (ar(h).reshape(1, h, 1, 1, 1) + dy).clamp(0.0, float(h - 1))  # :222-230
        x = (ar(w).reshape(1, 1, w, 1, 1) + dx).clamp(0.0, float(w - 1))
        out = deform_attn_sample(value.reshape(n, h, w, self.num_heads, c_head), y, x, attn)  # :233-235
        return out.reshape(n, h, w, self.num_heads * c_head)  # :238

    call = forward
