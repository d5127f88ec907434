"""CPU restatement (numpy) of the sampling stage of the reference's DCNv2, and of the whole layer around it.

TEST INFRASTRUCTURE ONLY.  Follows /root/reference/layers/dcn_v2.py `_forward` (:121-265):
    :137-146  padded extents, split of the offset convolution's output into (oy, ox) pairs and mask logits
    :148      sigmoid
    :150-163  grid = pixel + (ph, pw) + patch offset (tap k row-major over (py, px), :110-111) [integers], + offsets
    :165-177  floor, +1, clip of both neighbours AND of the coordinate to [0, H+1] x [0, W+1]
    :193-211  weights from the clipped values: [d0y*d0x, d0y*d1x, d1y*d0x, d1y*d1x]
    :180-184  neighbour order (y1,x1) (y1,x0) (y0,x1) (y0,x0)
    :219      zero padding of x by (ph, pw)
    :227-247  per tap: gather, [1,4] x [4,C], x mask; stacked to map_all [B,H,W,ks*C]
    :249-271  contraction with the kernel, bias, activation
Pinned by tests/golden/dcnv2_*.npz: the reference's own build() + _forward() run over the torch-backed TF stand-in
(oracle/ref_runner.py::run_dcn_v2).  The gradient of the sampling stage is the analytic gradient of that graph
(tf.clip_by_value passes the gradient where the value lies inside the range, floor and int casts cut it).
"""
import numpy as np


def _taps(offs, h, w, kh, kw):
    """offs [N,H,W,ks,2] -> clipped neighbour indices (padded coordinates), the four distances, inside flags."""
    dt = offs.dtype
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    ks = kh * kw
    k = np.arange(ks)
    py, px = k // kw - ph, k % kw - pw
    gi = (np.arange(h).reshape(1, h, 1, 1) + ph + py.reshape(1, 1, 1, ks)).astype(dt)
    gj = (np.arange(w).reshape(1, 1, w, 1) + pw + px.reshape(1, 1, 1, ks)).astype(dt)
    gy, gx = gi + offs[..., 0], gj + offs[..., 1]
    hb, wb = dt.type(h + 1), dt.type(w + 1)
    fy, fx = np.floor(gy), np.floor(gx)
    y1, x1 = np.clip(fy + dt.type(1), 0, hb), np.clip(fx + dt.type(1), 0, wb)
    y0, x0 = np.clip(fy, 0, hb), np.clip(fx, 0, wb)
    gyc, gxc = np.clip(gy, 0, hb), np.clip(gx, 0, wb)
    iny = ((gy >= 0) & (gy <= hb)).astype(dt)
    inx = ((gx >= 0) & (gx <= wb)).astype(dt)
    return (y0.astype(np.int64), x0.astype(np.int64), y1.astype(np.int64), x1.astype(np.int64),
            gyc - y0, y1 - gyc, gxc - x0, x1 - gxc, iny, inx)


def sample_forward(x, offs, mask, kh, kw):
    """x [N,H,W,C], offs [N,H,W,ks,2] (oy, ox), mask [N,H,W,ks] (after the sigmoid) -> map_all [N,H,W,ks,C]."""
    n, h, w, c = x.shape
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    xp = np.pad(x, ((0, 0), (ph, ph), (pw, pw), (0, 0)))
    y0, x0, y1, x1, d0y, d1y, d0x, d1x, _, _ = _taps(offs, h, w, kh, kw)
    ni = np.arange(n).reshape(n, 1, 1, 1)
    g = lambda yy, xx: xp[ni, yy, xx]  # noqa: E731  [N,H,W,ks,C]
    e = lambda a: a[..., None]  # noqa: E731
    s = e(d0y * d0x) * g(y1, x1)
    s = s + e(d0y * d1x) * g(y1, x0)
    s = s + e(d1y * d0x) * g(y0, x1)
    s = s + e(d1y * d1x) * g(y0, x0)
    return s * e(mask)


def sample_backward(x, offs, mask, grad_out, kh, kw):
    """Gradients of sample_forward: grad_x, grad_offs [N,H,W,ks,2], grad_mask (float64 accumulation)."""
    n, h, w, c = x.shape
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    xp = np.pad(x, ((0, 0), (ph, ph), (pw, pw), (0, 0))).astype(np.float64)
    y0, x0, y1, x1, d0y, d1y, d0x, d1x, iny, inx = (a if a.dtype == np.int64 else a.astype(np.float64)
                                                    for a in _taps(offs, h, w, kh, kw))
    ni = np.arange(n).reshape(n, 1, 1, 1)
    v11, v10, v01, v00 = xp[ni, y1, x1], xp[ni, y1, x0], xp[ni, y0, x1], xp[ni, y0, x0]
    go = grad_out.astype(np.float64)
    m = mask.astype(np.float64)[..., None]
    e = lambda a: a[..., None]  # noqa: E731
    s = e(d0y * d0x) * v11 + e(d0y * d1x) * v10 + e(d1y * d0x) * v01 + e(d1y * d1x) * v00
    grad_mask = (go * s).sum(-1)
    gy = (go * ((e(d0x) * v11 + e(d1x) * v10) - (e(d0x) * v01 + e(d1x) * v00))).sum(-1) * m[..., 0] * iny
    gx = (go * ((e(d0y) * v11 + e(d1y) * v01) - (e(d0y) * v10 + e(d1y) * v00))).sum(-1) * m[..., 0] * inx
    gxp = np.zeros(xp.shape, np.float64)
    nn = np.broadcast_to(ni, y0.shape)
    for yy, xx, wgt in ((y1, x1, d0y * d0x), (y1, x0, d0y * d1x), (y0, x1, d1y * d0x), (y0, x0, d1y * d1x)):
        np.add.at(gxp, (nn, yy, xx), go * m * e(wgt))
    dt = x.dtype
    return (gxp[:, ph:ph + h, pw:pw + w].astype(dt), np.stack([gy, gx], -1).astype(dt), grad_mask.astype(dt))


def layer_forward(x, kernel, bias, offset_kernel, offset_bias):
    """The whole `_forward` (:121-265) with `offset = x`, stride 1, dilation 1, no activation."""
    n, h, w, ic = x.shape
    kh, kw, _, oc = kernel.shape
    ks = kh * kw
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    xp = np.pad(x, ((0, 0), (ph, ph), (pw, pw), (0, 0)))
    off = np.zeros((n, h, w, offset_kernel.shape[-1]), x.dtype)
    for a in range(kh):      # tf.nn.conv2d, SAME (:128-135)
        for b in range(kw):
            off += xp[:, a:a + h, b:b + w] @ offset_kernel[a, b]
    off = off + offset_bias
    oyox = off[..., :2 * ks].reshape(n, h, w, ks, 2)
    mask = 1.0 / (1.0 + np.exp(-off[..., 2 * ks:]))
    m = sample_forward(x, oyox, mask.astype(x.dtype), kh, kw).reshape(n, h * w, ks * ic)
    out = (m @ kernel.reshape(ks * ic, oc)).reshape(n, h, w, oc)
    return out if bias is None else out + bias
