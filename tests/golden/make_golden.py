"""Regenerates tests/golden/*.npz by executing the reference's own DCNv3 source.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

Every fixture holds the inputs and what the UNMODIFIED reference files
(/root/reference/layers/dcn_v3/op.py:16 dcnv3_op, utils.py:110 dcnv3_bilinear_sampler,
dcn_v3.py:16 DeformableConvolutionV3) produce for them when run over the torch-backed TF stand-in
(oracle/ref_runner.py); gradients are torch autograd through that same forward graph.  float64 cases
additionally carry central-difference checks at generation time (asserted below, not stored).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_runner  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def softmax_taps(z, groups):
    n, h, w, gp = z.shape
    z = z.reshape(n, h, w, groups, gp // groups)
    e = np.exp(z - z.max(-1, keepdims=True))
    return (e / e.sum(-1, keepdims=True)).reshape(n, h, w, gp)


def out_hw(h, w, k, s, d, padding):
    ph, pw = (k[0] // 2, k[1] // 2) if padding.upper() == "SAME" else (0, 0)
    hin, win = h + 2 * ph, w + 2 * pw
    return (hin - (d[0] * (k[0] - 1) + 1)) // s[0] + 1, (win - (d[1] * (k[1] - 1) + 1)) // s[1] + 1


# name: (N, H, W, G, gc, kernel, strides, dilation, padding, offset_scale, offset_sigma, dtype, seed)
OP_CASES = {
    "rand_5x7_g2c3_s1.7_f64": (2, 5, 7, 2, 3, (3, 3), (1, 1), (1, 1), "SAME", 1.7, 3.0, "f8", 0),
    "rand_5x7_g2c3_s1.7": (2, 5, 7, 2, 3, (3, 3), (1, 1), (1, 1), "SAME", 1.7, 3.0, "f4", 0),
    "smoke_17x17_g4c16": (1, 17, 17, 4, 16, (3, 3), (1, 1), (1, 1), "same", 1.0, 1.0, "f4", 1),
    "rect_9x20_g8c16": (1, 9, 20, 8, 16, (3, 3), (1, 1), (1, 1), "SAME", 1.0, 2.0, "f4", 2),
    "rect_20x9_g3c8_s2": (1, 20, 9, 3, 8, (3, 3), (1, 1), (1, 1), "SAME", 2.0, 1.5, "f4", 3),
    "far_offsets_12x12_g2c16": (1, 12, 12, 2, 16, (3, 3), (1, 1), (1, 1), "SAME", 1.0, 12.0, "f4", 4),
    "gc32_10x10_g2": (1, 10, 10, 2, 32, (3, 3), (1, 1), (1, 1), "SAME", 1.0, 1.0, "f4", 5),
    "gc32_33x34_g3": (1, 33, 34, 3, 32, (3, 3), (1, 1), (1, 1), "SAME", 1.0, 3.0, "f4", 13),  # 2 x 2 scatter tiles, odd G
    "valid_11x13_g2c4": (1, 11, 13, 2, 4, (3, 3), (1, 1), (1, 1), "VALID", 1.0, 1.0, "f4", 6),
    "k5_9x9_g2c4": (1, 9, 9, 2, 4, (5, 5), (1, 1), (1, 1), "SAME", 1.0, 1.0, "f4", 7),
    "k3x5_8x10_g1c8": (1, 8, 10, 1, 8, (3, 5), (1, 1), (1, 1), "SAME", 1.0, 1.0, "f4", 8),
    "stride2_12x12_g2c4": (1, 12, 12, 2, 4, (3, 3), (2, 2), (1, 1), "SAME", 1.0, 1.0, "f4", 9),
    "dil2_12x14_g2c4": (1, 12, 14, 2, 4, (3, 3), (1, 1), (2, 2), "SAME", 1.0, 1.0, "f4", 10),
    "tiny_1x1_g1c16": (1, 1, 1, 1, 16, (3, 3), (1, 1), (1, 1), "SAME", 1.0, 0.5, "f4", 11),
    "zero_offset_12x12_g4c16": (1, 12, 12, 4, 16, (3, 3), (1, 1), (1, 1), "SAME", 1.0, 0.0, "f4", 12),
}


def make_op_case(name, spec):
    n, h, w, g, gc, k, s, d, padding, scale, sigma, dt, seed = spec
    rng = np.random.default_rng(seed)
    ho, wo = out_hw(h, w, k, s, d, padding)
    p_ = k[0] * k[1]
    x = rng.standard_normal((n, h, w, g * gc)).astype(dt)
    offset = (sigma * rng.standard_normal((n, ho, wo, g * p_ * 2))).astype(dt)
    mask = softmax_taps(rng.standard_normal((n, ho, wo, g * p_)), g).astype(dt)
    grad_out = rng.standard_normal((n, ho, wo, g * gc)).astype(dt)
    kw = dict(kernel_size=k, strides=s, padding=padding, dilation_rate=d, groups=g,
              group_channels=gc, offset_scale=scale)
    out, gx, goff, gm = ref_runner.run_op(x, offset, mask, grad_out=grad_out, **kw)
    if dt == "f8":  # central differences on a few coordinates: autograd == derivative of the forward
        eps = 1e-6
        loss = lambda o: float((o * grad_out).sum())  # noqa: E731
        for arr, grad in ((x, gx), (offset, goff), (mask, gm)):
            for _ in range(6):
                idx = tuple(rng.integers(0, sz) for sz in arr.shape)
                keep = arr[idx]
                arr[idx] = keep + eps
                lp = loss(ref_runner.run_op(x, offset, mask, **kw))
                arr[idx] = keep - eps
                lm = loss(ref_runner.run_op(x, offset, mask, **kw))
                arr[idx] = keep
                assert abs((lp - lm) / (2 * eps) - grad[idx]) < 1e-6 * max(1.0, abs(grad[idx])), name
    np.savez_compressed(
        os.path.join(OUT, f"op_{name}.npz"), x=x, offset=offset, mask=mask, grad_out=grad_out,
        out=out, grad_x=gx, grad_offset=goff, grad_mask=gm,
        kernel_size=np.array(k), strides=np.array(s), dilation_rate=np.array(d),
        padding=np.array(padding), groups=np.array(g), group_channels=np.array(gc),
        offset_scale=np.array(scale, dtype="f8"))


# The reference under the mixed_bfloat16 policy: the same unmodified source run on bfloat16 tensors (every
# primitive then rounds to bf16, op.py:62-87 / utils.py:130-206).  Stored as uint16 bf16 bit patterns
# (tests/helpers.py load_op_case widens them to float32).
# name: (N, H, W, G, gc, offset_scale, offset_sigma, seed)   -- 3x3, stride 1, SAME (the InternImage configuration);
# offset scales are bf16-representable (a python scalar becomes a tensor of x.dtype in TF)
BF16_CASES = {
    "bf16_32x32_g4c16": (1, 32, 32, 4, 16, 1.0, 1.0, 20),
    "bf16_136x24_g2c16": (1, 136, 24, 2, 16, 1.0, 1.0, 21),   # coordinates beyond 128: bf16 step of 1 pixel
    "bf16_17x23_g2c16_s2": (2, 17, 23, 2, 16, 2.0, 2.0, 22),
    "bf16_12x12_g3c8_s0.5": (1, 12, 12, 3, 8, 0.5, 4.0, 23),
    "bf16_24x24_g8c16": (1, 24, 24, 8, 16, 1.0, 1.0, 24),
}


def make_bf16_case(name, spec):
    import torch

    n, h, w, g, gc, scale, sigma, seed = spec
    rng = np.random.default_rng(seed)
    ref = ref_runner.load()
    bf = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).bfloat16()  # noqa: E731
    x = bf(rng.standard_normal((n, h, w, g * gc)))
    offset = bf(sigma * rng.standard_normal((n, h, w, g * 18)))
    mask = bf(softmax_taps(rng.standard_normal((n, h, w, g * 9)), g))
    grad_out = bf(rng.standard_normal((n, h, w, g * gc)))
    tx, to, tm = (ref.tf.convert_to_tensor(t.clone()) for t in (x, offset, mask))
    tx.requires_grad_(True), to.requires_grad_(True), tm.requires_grad_(True)
    out = ref.dcnv3_op(tx, to, tm, [3, 3], [1, 1], "SAME", [1, 1], g, gc, scale)
    assert out.dtype == torch.bfloat16
    out.backward(grad_out)
    f = lambda t: t.detach().contiguous().view(torch.int16).numpy().view(np.uint16)  # noqa: E731  (bf16 bits)
    np.savez_compressed(
        os.path.join(OUT, f"op_{name}.npz"), bf16_bits=np.array(1), x=f(x), offset=f(offset), mask=f(mask), grad_out=f(grad_out),
        out=f(out), grad_x=f(tx.grad), grad_offset=f(to.grad), grad_mask=f(tm.grad),
        kernel_size=np.array((3, 3)), strides=np.array((1, 1)), dilation_rate=np.array((1, 1)),
        padding=np.array("SAME"), groups=np.array(g), group_channels=np.array(gc),
        offset_scale=np.array(scale, dtype="f8"))


def make_kats():
    """Known answers that are also derivable by hand."""
    h = w = 6
    g, gc, p_ = 2, 1, 9
    x = np.zeros((1, h, w, g * gc), np.float32)
    for r in range(h):
        for q in range(w):
            x[0, r, q, :] = 10 * r + q
    offset = np.zeros((1, h, w, g * p_ * 2), np.float32)
    mask = np.zeros((1, h, w, g, p_), np.float32)
    mask[..., 4] = 1
    mask = mask.reshape(1, h, w, g * p_)
    out = ref_runner.run_op(x, offset, mask, groups=g, group_channels=gc)
    assert abs(out[0, 1, 4, 0] - 32.125) < 1e-5 and abs(out[0, 0, 0, 0] - 1.375) < 1e-5
    assert abs(out[0, 5, 5, 0] - 42.625) < 1e-5
    # unit shift: +W_in/((W_in-2)*s) on channel 0 moves every sample one column to the right
    shift = np.float32((w + 2) / w)
    offset2 = offset.copy()
    offset2[..., 0::2] = shift
    out_shift = ref_runner.run_op(x, offset2, mask, groups=g, group_channels=gc)
    # out of range: +-1e3 => dead taps
    offset3 = np.full_like(offset, 1e3)
    offset3[..., 1::2] = -1e3
    out_dead = ref_runner.run_op(x, offset3, mask, groups=g, group_channels=gc)
    assert np.all(out_dead == 0)
    np.savez_compressed(os.path.join(OUT, "kat_ramp.npz"), x=x, offset=offset, mask=mask, out=out,
                        offset_shift=offset2, out_shift=out_shift, offset_dead=offset3,
                        out_dead=out_dead)
    # constant image, mask 1/9, zero offsets
    xc = np.full((1, 8, 8, 16), 2.5, np.float32)
    oc = np.zeros((1, 8, 8, 18), np.float32)
    mc = np.full((1, 8, 8, 9), 1 / 9, np.float32)
    outc = ref_runner.run_op(xc, oc, mc, groups=1, group_channels=16)
    np.savez_compressed(os.path.join(OUT, "kat_const.npz"), x=xc, offset=oc, mask=mc, out=outc)


def make_layer_case():
    """The boundary caller, dcn_v3.py:107-150, with non-zero offset/mask projections so that the
    data-dependent gather is exercised (the reference initialises them to zero, dcn_v3.py:74-86)."""
    import torch

    ref = ref_runner.load()
    tf = ref.tf
    for name, (c, g, k, dwk, cfs, hh, ww, scale, seed) in {
        "c64_g4_cfs": (64, 4, 3, None, True, 17, 17, 1.0, 0),
        "c32_g2_dw5": (32, 2, 3, 5, False, 9, 12, 2.0, 1),
    }.items():
        torch.manual_seed(seed)
        layer = ref.DeformableConvolutionV3(filters=c, kernel_size=k, depthwise_kernel_size=dwk,
                                            groups=g, offset_scale=scale, center_feature_scale=cfs)
        x = torch.randn(2, hh, ww, c)
        layer(tf.convert_to_tensor(x))  # build
        layer.offset.kernel = torch.randn_like(layer.offset.kernel) * 0.3
        layer.offset.bias = torch.randn_like(layer.offset.bias) * 0.5
        layer.mask.kernel = torch.randn_like(layer.mask.kernel) * 0.3
        layer.mask.bias = torch.randn_like(layer.mask.bias) * 0.3
        for sub in (layer.input_proj, layer.output_proj, layer.dw_conv):
            sub.bias = torch.randn_like(sub.bias) * 0.1
        layer.dw_norm.gamma = 1 + 0.1 * torch.randn_like(layer.dw_norm.gamma)
        layer.dw_norm.beta = 0.1 * torch.randn_like(layer.dw_norm.beta)
        y = layer(tf.convert_to_tensor(x))
        w = {
            "input_proj.kernel": layer.input_proj.kernel, "input_proj.bias": layer.input_proj.bias,
            "output_proj.kernel": layer.output_proj.kernel, "output_proj.bias": layer.output_proj.bias,
            "dw_conv.depthwise_kernel": layer.dw_conv.depthwise_kernel, "dw_conv.bias": layer.dw_conv.bias,
            "dw_conv_norm.gamma": layer.dw_norm.gamma, "dw_conv_norm.beta": layer.dw_norm.beta,
            "offset.kernel": layer.offset.kernel, "offset.bias": layer.offset.bias,
            "mask.kernel": layer.mask.kernel, "mask.bias": layer.mask.bias,
        }
        if cfs:
            layer.center_feature_scale_proj.bias = torch.rand(g)
            y = layer(tf.convert_to_tensor(x))
            w["center_feature_scale_proj.kernel"] = layer.center_feature_scale_proj.kernel
            w["center_feature_scale_proj.bias"] = layer.center_feature_scale_proj.bias
        np.savez_compressed(
            os.path.join(OUT, f"layer_{name}.npz"), x=x.numpy(), y=y.detach().numpy(),
            filters=np.array(c), groups=np.array(g), kernel_size=np.array(k),
            depthwise_kernel_size=np.array(dwk or 0), center_feature_scale=np.array(cfs),
            offset_scale=np.array(scale), **{f"w.{k_}": v.numpy() for k_, v in w.items()})


def make_sliding_indices():
    """Window starts of the reference's sliding-window inference: its own `_get_sliding_start_indexs_py`
    (utils/sliding_window_inference_utils.py:16-32, plain Python) evaluated over a grid of (length, crop) pairs."""
    import importlib.util
    import json
    import sys
    import types

    ref_runner.load()  # installs the tensorflow stand-in and the `iseg` package skeleton
    common = types.ModuleType("iseg.utils.common")
    common.isinstance_all = lambda xs, t: all(isinstance(v, t) for v in xs)
    sys.modules["iseg.utils.common"] = common
    path = os.path.join(ref_runner.REFERENCE_ROOT, "utils", "sliding_window_inference_utils.py")
    spec = importlib.util.spec_from_file_location("iseg.utils.sliding_window_inference_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cases = {}
    for crop in (24, 193, 257, 512, 513, 769, 1024):
        for length in sorted({crop, crop + 1, crop + 7, int(1.5 * crop), 2 * crop - 1, 2 * crop, 1024, 2048, 2049, 3000}):
            if length >= crop:
                cases[f"{length},{crop}"] = [int(v) for v in mod._get_sliding_start_indexs_py(length, crop, 2.0 / 3.0)]
    json.dump(cases, open(os.path.join(OUT, "sliding_indices.json"), "w"), indent=0, sort_keys=True)


# sibling op: the sampler of the reference's deformable multi-head self-attention.
# name: (N, H, W, heads, points, C_head, how far outside the image the coordinates may lie, dtype, seed)
DMSA_CASES = {
    "inside_9x11_h3p4c8": (2, 9, 11, 3, 4, 8, 0.0, "f4", 20),     # what the layer produces: clipped to the image (:229-230)
    "border_7x9_h2p5c5_f64": (1, 7, 9, 2, 5, 5, 2.5, "f8", 21),   # beyond the border: clamped neighbour indices (:128-131)
    "border_12x10_h4p4c16": (2, 12, 10, 4, 4, 16, 3.0, "f4", 22),
    "tiny_1x1_h1p2c4": (1, 1, 1, 1, 2, 4, 1.0, "f4", 23),
}


def make_dmsa_case(name, spec):
    n, h, w, heads, p, c, beyond, dt, seed = spec
    rng = np.random.default_rng(seed)
    dtype = np.float64 if dt == "f8" else np.float32
    value = rng.standard_normal((n, h, w, heads, c)).astype(dtype)
    y = rng.uniform(-beyond, h - 1 + beyond, (n, h, w, heads, p)).astype(dtype)
    x = rng.uniform(-beyond, w - 1 + beyond, (n, h, w, heads, p)).astype(dtype)
    y.reshape(-1)[::7] = np.round(y.reshape(-1)[::7])  # exact integers: floor(y) == y, weight 0 on the upper neighbour
    z = rng.standard_normal((n, h, w, heads, p))
    e = np.exp(z - z.max(-1, keepdims=True))
    attn = (e / e.sum(-1, keepdims=True)).astype(dtype)
    grad_out = rng.standard_normal((n, h, w, heads, c)).astype(dtype)
    out, gv, gy, gx, ga = ref_runner.run_deform_attn(value, y, x, attn, grad_out)
    np.savez_compressed(os.path.join(OUT, f"dmsa_{name}.npz"), value=value, y=y, x=x, attn=attn, grad_out=grad_out, out=out,
                        grad_value=gv, grad_y=gy, grad_x=gx, grad_attn=ga)


# sibling op: the reference's DCNv2 layer, build() + _forward() (layers/dcn_v2.py:61-113, :121-265).
# name: (N, H, W, C_in, filters, k, offset spread, dtype, seed)
DCNV2_CASES = {
    "k3_8x9_c6_o5": (2, 8, 9, 6, 5, 3, 1.5, "f4", 30),
    "k3_far_7x7_c4_o3_f64": (1, 7, 7, 4, 3, 3, 6.0, "f8", 31),    # offsets far beyond the image: every clip in play
    "k5_9x8_c3_o4": (1, 9, 8, 3, 4, 5, 2.0, "f4", 32),            # 5x5: the [0, H+1] clip range inside a pad-2 image
}


def make_dcnv2_case(name, spec):
    n, h, w, ic, oc, k, spread, dt, seed = spec
    rng = np.random.default_rng(seed)
    dtype = np.float64 if dt == "f8" else np.float32
    x = rng.standard_normal((n, h, w, ic)).astype(dtype)
    kernel = (rng.standard_normal((k, k, ic, oc)) * 0.3).astype(dtype)
    bias = rng.standard_normal(oc).astype(dtype)
    offset_kernel = (rng.standard_normal((k, k, ic, 3 * k * k)) * 0.15).astype(dtype)
    offset_bias = (rng.standard_normal(3 * k * k) * spread).astype(dtype)
    grad_out = rng.standard_normal((n, h, w, oc)).astype(dtype)
    out, gx, gk, gb, gok, gob = ref_runner.run_dcn_v2(x, kernel, bias, offset_kernel, offset_bias, grads_for=grad_out)
    np.savez_compressed(os.path.join(OUT, f"dcnv2_{name}.npz"), x=x, kernel=kernel, bias=bias, offset_kernel=offset_kernel,
                        offset_bias=offset_bias, grad_out=grad_out, out=out, grad_x=gx, grad_kernel=gk, grad_bias=gb,
                        grad_offset_kernel=gok, grad_offset_bias=gob)


if __name__ == "__main__":
    assert ref_runner.available(), "needs /root/reference"
    for nm, sp in OP_CASES.items():
        make_op_case(nm, sp)
        print("wrote", nm)
    for nm, sp in BF16_CASES.items():
        make_bf16_case(nm, sp)
        print("wrote", nm)
    make_kats()
    make_layer_case()
    make_sliding_indices()
    for nm, sp in DCNV2_CASES.items():
        make_dcnv2_case(nm, sp)
        print("wrote dcnv2", nm)
    for nm, sp in DMSA_CASES.items():
        make_dmsa_case(nm, sp)
        print("wrote dmsa", nm)
    print("done")
