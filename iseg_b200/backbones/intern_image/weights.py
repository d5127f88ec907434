"""Name-based weight import for the InternImage backbone.

The reference restores checkpoints BY LAYER NAME (reference saver/h5_saver.py:38-49 `load_h5_weight_by_name`);
its InternImage sub-layers are named "<parent>/<child>" (backbones/intern_image/intern_image.py:73,111,
intern_image_block.py:75-99, intern_image_layer.py:59-116, mlp_layer.py:34-42, stem_layer.py:35-51,
dowmsample_layer.py:28-32, layers/dcn_v3/dcn_v3.py:62-102), "/" becoming "." under Keras 3
(utils/keras3_utils.py:23-29).  This module maps those names to the torch parameters of
`iseg_b200.backbones.intern_image.InternImage` and converts the Keras variable layouts:

    Dense   kernel [in, out]            -> Linear.weight [out, in]
    Conv2D  kernel [kh, kw, cin, cout]  -> Conv2d.weight [cout, cin, kh, kw]
    DepthwiseConv2D depthwise_kernel [kh, kw, C, 1] -> Conv2d(groups=C).weight [C, 1, kh, kw]
    LayerNormalization gamma / beta     -> LayerNorm.weight / bias

h5py is not available in this image, so the container format here is a flat `.npz` (or any mapping) keyed by
"<layer path>/<variable>"; keys may use "/" or ".", may carry the ":0" suffix and a leading model name.
"""
import numpy as np
import torch


def _dense(lin, prefix):
    return {f"{prefix}/kernel": (lin.weight, "dense"), f"{prefix}/bias": (lin.bias, None)}


def _norm(ln, prefix):
    return {f"{prefix}/gamma": (ln.weight, None), f"{prefix}/beta": (ln.bias, None)}


def _conv(conv, prefix):
    out = {f"{prefix}/kernel": (conv.weight, "conv")}
    if conv.bias is not None:
        out[f"{prefix}/bias"] = (conv.bias, None)
    return out


def reference_names(model):
    """{reference variable path: (torch parameter, layout tag)} for every parameter of `model`."""
    names = {}
    pe = model.patch_embed
    names.update(_conv(pe.conv1, "patch_embed/conv1")), names.update(_norm(pe.norm1, "patch_embed/norm1"))
    names.update(_conv(pe.conv2, "patch_embed/conv2")), names.update(_norm(pe.norm2, "patch_embed/norm2"))
    for i, blk in enumerate(model.blocks):
        b = f"block/{i}"
        for j, layer in enumerate(blk.blocks):
            p = f"{b}/layer/{j}"
            names.update(_norm(layer.norm1, f"{p}/norm1")), names.update(_norm(layer.norm2, f"{p}/norm2"))
            d, dp = layer.dcn, f"{p}/dcn"
            names[f"{dp}/dw_conv/depthwise_kernel"] = (d.dw_conv.weight, "depthwise")
            names[f"{dp}/dw_conv/bias"] = (d.dw_conv.bias, None)
            names.update(_norm(d.dw_conv_norm, f"{dp}/dw_conv_norm"))
            for sub in ("offset", "mask", "input_proj", "output_proj"):
                names.update(_dense(getattr(d, sub), f"{dp}/{sub}"))
            if d.center_feature_scale:
                names.update(_dense(d.center_feature_scale_proj, f"{dp}/center_feature_scale_proj"))
            names.update(_dense(layer.mlp.fc1, f"{p}/mlp/fc1")), names.update(_dense(layer.mlp.fc2, f"{p}/mlp/fc2"))
            if torch.is_tensor(layer.gamma1):
                names[f"{p}/gamma1"] = (layer.gamma1, None)
                names[f"{p}/gamma2"] = (layer.gamma2, None)
            if layer.use_res_post_norm:
                names.update(_norm(layer.res_post_norm1, f"{p}/res_post_norm1"))
                names.update(_norm(layer.res_post_norm2, f"{p}/res_post_norm2"))
        if blk.norm is not None:
            names.update(_norm(blk.norm, f"{b}/norm"))
        if blk.post_norm_block_ids is not None:
            for k, ln in enumerate(blk.post_norms):
                names.update(_norm(ln, f"{b}/post_norms/{k}"))
        if blk.downsample is not None:
            names.update(_conv(blk.downsample.conv, f"{b}/downsample/conv"))
            names.update(_norm(blk.downsample.norm, f"{b}/downsample/norm"))
    return names


def _to_torch(a, tag):
    t = torch.as_tensor(np.asarray(a))
    if tag == "dense":
        return t.t()
    if tag == "conv":
        return t.permute(3, 2, 0, 1)
    if tag == "depthwise":
        return t.permute(2, 3, 0, 1)
    return t


def _to_keras(t, tag):
    t = t.detach().cpu()
    if tag == "dense":
        t = t.t()
    elif tag == "conv":
        t = t.permute(2, 3, 1, 0)
    elif tag == "depthwise":
        t = t.permute(2, 3, 0, 1)
    return t.contiguous().float().numpy()


def _canonical(key, model_name):
    key = key[:-2] if key.endswith(":0") else key
    key = key.replace(".", "/")
    if model_name and key.startswith(model_name.replace(".", "/") + "/"):
        key = key[len(model_name) + 1:]
    return key


def export_reference_weights(model):
    """{reference variable path: numpy array in the Keras layout} (what a by-name checkpoint of the reference holds)."""
    return {k: _to_keras(p, tag) for k, (p, tag) in reference_names(model).items()}


def load_reference_weights(model, weights, strict=True):
    """Copies `weights` (a mapping or the path of an .npz) into `model` by reference variable name.
    Returns (loaded, missing, unexpected) name lists; with strict=True anything missing or unexpected raises."""
    if isinstance(weights, (str, bytes)):
        weights = dict(np.load(weights))
    table = reference_names(model)
    given = {_canonical(k, getattr(model, "name", None)): v for k, v in weights.items()}
    missing = [k for k in table if k not in given]
    unexpected = [k for k in given if k not in table]
    if strict and (missing or unexpected):
        raise KeyError(f"weights do not match the model: missing {missing[:5]}..., unexpected {unexpected[:5]}...")
    loaded = []
    with torch.no_grad():
        for k, (p, tag) in table.items():
            if k in given:
                t = _to_torch(given[k], tag)
                if tuple(t.shape) != tuple(p.shape):
                    raise ValueError(f"{k}: shape {tuple(t.shape)} does not fit {tuple(p.shape)}")
                p.copy_(t.to(p.dtype))
                loaded.append(k)
    return loaded, missing, unexpected
