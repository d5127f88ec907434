"""CPU restatement (numpy) of the sampling + aggregation of the reference's deformable multi-head self-attention.

TEST INFRASTRUCTURE ONLY (tests/, never the product path).  Follows, statement by statement,

    /root/reference/layers/deformable_multihead_self_attention.py
        :102-175  _bilinear_sample   (floor :124-125, neighbours :126-127, int cast + clip :128-131, weights from the
                                      unclamped fractional parts :133-136, four gathers :153-156, products :162-165,
                                      weighted sum left to right :167)
        :233-235  sampled * attn, summed over the points

in the dtype of its inputs (fp32 / fp64).  The backward is the analytic gradient of exactly that graph (floor / int
cast / clip carry no gradient, as in TF autodiff).  Pinned by tests/golden/dmsa_*.npz, which are produced by running the
reference's own unmodified `_bilinear_sample` over the torch-backed TF stand-in (oracle/ref_runner.py::run_deform_attn,
generator tests/golden/make_golden.py).
"""
import numpy as np


def _taps(y, x, h, w):
    y0, x0 = np.floor(y), np.floor(x)
    wy1, wx1 = y - y0, x - x0
    wy0, wx0 = 1.0 - wy1, 1.0 - wx1
    one = y.dtype.type(1)
    clip = lambda v, hi: np.clip(v.astype(np.int64), 0, hi)  # noqa: E731
    return (clip(y0, h - 1), clip(x0, w - 1), clip(y0 + one, h - 1), clip(x0 + one, w - 1), wy0, wy1, wx0, wx1)


def forward(value, y, x, attn):
    """value [N,H,W,heads,C]; y, x, attn [N,H,W,heads,P] -> [N,H,W,heads,C]."""
    n, h, w, heads, c = value.shape
    y0, x0, y1, x1, wy0, wy1, wx0, wx1 = _taps(y, x, h, w)
    ni = np.arange(n).reshape(n, 1, 1, 1, 1)
    hi = np.arange(heads).reshape(1, 1, 1, heads, 1)
    g = lambda yy, xx: value[ni, yy, xx, hi]  # noqa: E731  [N,H,W,heads,P,C]
    w00, w01, w10, w11 = (a[..., None] for a in (wy0 * wx0, wy0 * wx1, wy1 * wx0, wy1 * wx1))
    sampled = w00 * g(y0, x0) + w01 * g(y0, x1) + w10 * g(y1, x0) + w11 * g(y1, x1)
    out = np.zeros_like(value)
    for p in range(y.shape[-1]):  # sequential sum over the points
        out = out + sampled[..., p, :] * attn[..., p, None]
    return out


def backward(value, y, x, attn, grad_out):
    """Returns grad_value, grad_y, grad_x, grad_attn (float64 accumulation of the scatter)."""
    n, h, w, heads, c = value.shape
    y0, x0, y1, x1, wy0, wy1, wx0, wx1 = _taps(y, x, h, w)
    ni = np.arange(n).reshape(n, 1, 1, 1, 1)
    hi = np.arange(heads).reshape(1, 1, 1, heads, 1)
    g = lambda yy, xx: value[ni, yy, xx, hi].astype(np.float64)  # noqa: E731
    v00, v01, v10, v11 = g(y0, x0), g(y0, x1), g(y1, x0), g(y1, x1)
    go = grad_out.astype(np.float64)[..., None, :]            # [N,H,W,heads,1,C]
    wy0, wy1, wx0, wx1 = (a.astype(np.float64)[..., None] for a in (wy0, wy1, wx0, wx1))
    a = attn.astype(np.float64)[..., None]
    s = wy0 * wx0 * v00 + wy0 * wx1 * v01 + wy1 * wx0 * v10 + wy1 * wx1 * v11
    grad_attn = (go * s).sum(-1)
    grad_y = (a * go * ((v10 - v00) * wx0 + (v11 - v01) * wx1)).sum(-1)
    grad_x = (a * go * ((v01 - v00) * wy0 + (v11 - v10) * wy1)).sum(-1)
    gv = np.zeros(value.shape, np.float64)
    nn = np.broadcast_to(ni, y0.shape)
    hh = np.broadcast_to(hi, y0.shape)
    for yy, xx, wgt in ((y0, x0, wy0 * wx0), (y0, x1, wy0 * wx1), (y1, x0, wy1 * wx0), (y1, x1, wy1 * wx1)):
        np.add.at(gv, (nn, yy, xx, hh), a * go * wgt)
    dt = value.dtype
    return gv.astype(dt), grad_y.astype(dt), grad_x.astype(dt), grad_attn.astype(dt)
