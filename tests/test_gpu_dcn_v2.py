"""Sibling gather op: DCNv2 (reference layers/dcn_v2.py), CUDA sampler through the C ABI.
The layer -- offset convolution, sampler, contraction -- against the fixtures made by the reference's own build() +
_forward(), output and all five gradients; the sampler alone against the numpy oracle.  fp32 bar 1e-5 (the dense stages
around the sampler are cuDNN / cuBLAS: 5e-5 at the layer level), bf16 1e-2."""
import glob
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, rel_err
from oracle import dcnv2_oracle as D

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(GOLDEN, "dcnv2_*.npz"))))
def test_layer_matches_reference_layer(name):
    from iseg_b200.layers.dcn_v2 import DCNv2
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    z = np.load(os.path.join(GOLDEN, f"dcnv2_{name}.npz"))
    k, _, ic, oc = z["kernel"].shape
    layer = DCNv2(oc, k, input_channels=ic)
    with torch.no_grad():
        for pname in ("kernel", "bias", "offset_kernel", "offset_bias"):
            getattr(layer, pname).copy_(torch.from_numpy(z[pname].astype(np.float32)))
    layer = layer.cuda()
    x = torch.from_numpy(z["x"].astype(np.float32)).cuda().requires_grad_(True)
    out = layer(x)
    out.backward(torch.from_numpy(z["grad_out"].astype(np.float32)).cuda())
    assert rel_err(out.detach().cpu().numpy(), z["out"]) <= 5e-5
    got = dict(grad_x=x.grad, grad_kernel=layer.kernel.grad, grad_bias=layer.bias.grad,
               grad_offset_kernel=layer.offset_kernel.grad, grad_offset_bias=layer.offset_bias.grad)
    for key, g in got.items():
        assert rel_err(g.cpu().numpy(), z[key]) <= 5e-5, key


@pytest.mark.parametrize("shape, k, dtype", [((2, 33, 29, 16), 3, torch.float32), ((1, 20, 24, 7), 5, torch.float32),
                                             ((2, 40, 36, 32), 3, torch.bfloat16)])
def test_sampler_vs_oracle_and_reproducible(shape, k, dtype):
    from iseg_b200.layers.dcn_v2 import dcnv2_sample
    n, h, w, c = shape
    rng = np.random.default_rng(h)
    x = rng.standard_normal((n, h, w, c)).astype(np.float32)
    offs = rng.uniform(-3, 3, (n, h, w, k * k, 2)).astype(np.float32)
    mask = rng.uniform(0, 1, (n, h, w, k * k)).astype(np.float32)
    go = (rng.standard_normal((n, h, w, k * k, c)) * (10.0 ** rng.integers(-3, 4, (n, 1, 1, 1, 1)))).astype(np.float32)
    if dtype == torch.bfloat16:
        x, offs, mask, go = (torch.from_numpy(a).bfloat16().float().numpy() for a in (x, offs, mask, go))

    def run(x_, offs_, mask_, go_):
        t = [torch.from_numpy(a).to("cuda", dtype).requires_grad_(True) for a in (x_, offs_, mask_)]
        out = dcnv2_sample(*t, k)
        out.backward(torch.from_numpy(go_).to("cuda", dtype))
        return tuple(v.float().cpu().numpy() for v in (out.detach(), *(u.grad for u in t)))

    got = run(x, offs, mask, go)
    ref = (D.sample_forward(x, offs, mask, k, k),) + D.sample_backward(x, offs, mask, go, k, k)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    for g, r, key in zip(got, ref, ("out", "grad_x", "grad_offsets", "grad_mask")):
        assert rel_err(g, r) <= tol, key
    assert all(np.array_equal(a, b) for a, b in zip(got, run(x, offs, mask, go)))   # bitwise reproducible
    assert np.array_equal(run(x[:1], offs[:1], mask[:1], go[:1])[1], got[1][:1])     # batch invariant


def test_errors():
    from iseg_b200 import _cabi
    from iseg_b200.layers.dcn_v2 import dcnv2_sample
    x = torch.zeros(1, 4, 4, 8, device="cuda")
    with pytest.raises(ValueError):
        dcnv2_sample(x, torch.zeros(1, 4, 4, 9, 2, device="cuda"), torch.zeros(1, 4, 4, 8, device="cuda"), 3)
    with pytest.raises(ValueError):   # even kernels: the reference's clip range leaves the padded image
        dcnv2_sample(x, torch.zeros(1, 4, 4, 4, 2, device="cuda"), torch.zeros(1, 4, 4, 4, device="cuda"), 2)
    with pytest.raises(_cabi.DCNv3Error):
        dcnv2_sample(x.cpu(), torch.zeros(1, 4, 4, 9, 2), torch.zeros(1, 4, 4, 9), 3)
