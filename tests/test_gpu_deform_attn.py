"""Sibling gather op: the sampler of the reference's deformable multi-head self-attention
(layers/deformable_multihead_self_attention.py:102-175 + :233-235), CUDA path through the C ABI against the fixtures
made by the reference's own code and against the numpy oracle.  fp32 bar 1e-5, bf16 1e-2."""
import glob
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, rel_err
from oracle import deform_attn_oracle as D

pytestmark = pytest.mark.gpu


def run(value, y, x, attn, grad_out=None, dtype=torch.float32):
    from iseg_b200.layers.deformable_attention import deform_attn_sample
    t = [torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype) for a in (value, y, x, attn)]
    if grad_out is not None:
        for v in t:
            v.requires_grad_(True)
    out = deform_attn_sample(*t)
    if grad_out is None:
        return out.float().cpu().numpy()
    out.backward(torch.from_numpy(grad_out).to("cuda", dtype))
    return tuple(v.float().cpu().numpy() for v in (out.detach(), *(u.grad for u in t)))


@pytest.mark.parametrize("name", sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "dmsa_*.npz"))))
def test_golden(name):
    z = np.load(os.path.join(GOLDEN, f"dmsa_{name}.npz"))
    f = lambda a: a.astype(np.float32)  # noqa: E731  (the float64 fixture is run in fp32)
    got = run(f(z["value"]), f(z["y"]), f(z["x"]), f(z["attn"]), f(z["grad_out"]))
    tol = 1e-5
    for g, key in zip(got, ("out", "grad_value", "grad_y", "grad_x", "grad_attn")):
        assert rel_err(g, z[key]) <= tol, key
    if z["value"].dtype == np.float32:  # same products and sums as the reference: the forward agrees to the last bits
        assert rel_err(got[0], z["out"]) <= 3e-7


@pytest.mark.parametrize("shape, dtype", [((2, 33, 47, 4, 4, 32), torch.float32), ((1, 64, 64, 8, 8, 16), torch.float32),
                                          ((2, 20, 24, 3, 5, 7), torch.float32), ((2, 40, 40, 4, 4, 32), torch.bfloat16)])
def test_random_vs_oracle_and_reproducible(shape, dtype):
    n, h, w, heads, p, c = shape
    rng = np.random.default_rng(h)
    value = rng.standard_normal((n, h, w, heads, c)).astype(np.float32)
    y = rng.uniform(-2, h + 1, (n, h, w, heads, p)).astype(np.float32)
    x = rng.uniform(-2, w + 1, (n, h, w, heads, p)).astype(np.float32)
    attn = rng.uniform(0, 1, (n, h, w, heads, p)).astype(np.float32)
    go = (rng.standard_normal((n, h, w, heads, c)) * (10.0 ** rng.integers(-3, 4, (n, 1, 1, 1, 1)))).astype(np.float32)
    if dtype == torch.bfloat16:
        value, y, x, attn, go = (torch.from_numpy(a).bfloat16().float().numpy() for a in (value, y, x, attn, go))
    got = run(value, y, x, attn, go, dtype=dtype)
    ref = (D.forward(value, y, x, attn),) + D.backward(value, y, x, attn, go)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    for g, r, key in zip(got, ref, ("out", "grad_value", "grad_y", "grad_x", "grad_attn")):
        assert rel_err(g, r) <= tol, key
    again = run(value, y, x, attn, go, dtype=dtype)
    assert all(np.array_equal(a, b) for a, b in zip(got, again))  # integer accumulation: bitwise reproducible
    # an image's gradient does not depend on what else is in the batch (per-image fixed-point scale)
    alone = run(value[:1], y[:1], x[:1], attn[:1], go[:1], dtype=dtype)
    assert np.array_equal(alone[1], got[1][:1])


def test_layer_and_errors():
    from iseg_b200 import _cabi
    from iseg_b200.layers.deformable_attention import DeformableMultiHeadSelfAttentionLayer, deform_attn_sample
    torch.manual_seed(0)
    layer = DeformableMultiHeadSelfAttentionLayer(num_heads=4, num_points=4, input_channels=64).cuda()
    q = torch.randn(2, 18, 22, 64, device="cuda", requires_grad=True)
    out = layer(q)
    assert out.shape == (2, 18, 22, 64)
    out.square().mean().backward()
    assert torch.isfinite(q.grad).all() and layer.offset_proj.weight.grad.abs().sum() > 0
    # the layer against the same graph with the sampler replaced by the oracle
    with torch.no_grad():
        n, h, w, _ = q.shape
        value = layer.value_proj(q).reshape(n, h, w, 4, 16)
        off = torch.tanh(layer.offset_proj(q).reshape(n, h, w, 4, 4, 2))
        attn = torch.softmax(layer.attn_proj(q).reshape(n, h, w, 4, 4), -1)
        ar = lambda k: torch.arange(k, device="cuda", dtype=q.dtype)  # noqa: E731
        y = (ar(h).reshape(1, h, 1, 1, 1) + off[..., 0] * (h / 8.0)).clamp(0, h - 1)
        x = (ar(w).reshape(1, 1, w, 1, 1) + off[..., 1] * (w / 8.0)).clamp(0, w - 1)
        want = D.forward(*(t.cpu().numpy() for t in (value, y, x, attn))).reshape(n, h, w, 64)
    assert rel_err(out.detach().cpu().numpy(), want) <= 1e-5
    with pytest.raises(ValueError):
        deform_attn_sample(torch.zeros(1, 4, 4, 2, 8, device="cuda"), torch.zeros(1, 4, 4, 2, 3, device="cuda"),
                           torch.zeros(1, 4, 4, 2, 4, device="cuda"), torch.zeros(1, 4, 4, 2, 3, device="cuda"))
    with pytest.raises(_cabi.DCNv3Error):
        deform_attn_sample(*(torch.zeros(1, 4, 4, 2, 8),) * 1, *(torch.zeros(1, 4, 4, 2, 3),) * 3)
