"""CPU oracle for the DCNv3 core operator -- a numpy restatement of the reference algorithm.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this; the product path (iseg_b200/) never does and fails loudly
when its CUDA library is missing.

PARITY PINNING.  The reference (edwardyehuang/iSeg) ships no golden vectors or asserting tests for
this path (layers/dcn_v3/test_dcn_v3.py:17-35 only prints), and TensorFlow is not installable here,
so the reference cannot be run on real TF => "parity unpinned" against TF itself.  What pins this
oracle instead: tests/golden/*.npz hold outputs AND gradients produced by executing the reference's
own, unmodified op.py / utils.py / dcn_v3.py from /root/reference over a torch-backed stand-in for
the TF primitives (oracle/ref_runner.py, oracle/tf_shim/; generator tests/golden/make_golden.py).
tests/test_oracle.py checks every function below against those fixtures.

Three restatements live here, each citing the reference lines it follows:

  forward_literal   -- op.py:16-109 + utils.py:14-210 step by step (pad, reference points, dilation
                       grid, regroup transposes, the 9-iteration 4-corner gather loop), in x.dtype.
  forward           -- the same arithmetic per sampled point (n,h,w,g,p), vectorised; identical
                       floating-point operation order for the coordinates.
  backward          -- analytic gradient of the graph TF autodiff differentiates (floor / int cast /
                       clip have zero gradient; utils.py:146-166), scatter-add in a fixed order.

Layouts (NHWC, op.py:89-107): x[N,H,W,G*gc]; offset[N,Ho,Wo,(g*P+p)*2+{0,1}]; mask[N,Ho,Wo,g*P+p];
out[N,Ho,Wo,G*gc].  Tap p = i*kh + j with i the channel-0 ("x"/W) displacement (utils.py:77-101).
"""
import numpy as np


def resolve_padding(kernel_size, padding):
    """op.py:29-39 -- TypeError for a non-string, ValueError for anything but SAME/VALID."""
    if not isinstance(padding, str):
        raise TypeError("padding must be a string in 'SAME' or 'VALID'")
    p = padding.upper()
    if p == "SAME":
        return kernel_size[0] // 2, kernel_size[1] // 2
    if p == "VALID":
        return 0, 0
    raise ValueError("padding must be 'SAME' or 'VALID'")


def check_shapes(x_shape, offset_shape, mask_shape, kernel_size, strides, pad, dilation_rate,
                 groups, group_channels):
    """The consistency envelope the reference enforces implicitly through tf.reshape (op.py:83,
    utils.py:26-27,55): the reference-point grid must have exactly the offset's spatial size."""
    n, h, w, c = x_shape
    kh, kw = kernel_size
    sh, sw = strides
    dh, dw = dilation_rate
    ph, pw = pad
    hin, win = h + 2 * ph, w + 2 * pw
    ho = (hin - (dh * (kh - 1) + 1)) // sh + 1
    wo = (win - (dw * (kw - 1) + 1)) // sw + 1
    p_ = kh * kw
    if c != groups * group_channels:
        raise ValueError(f"channels {c} != groups*group_channels {groups * group_channels}")
    if tuple(offset_shape) != (n, ho, wo, groups * p_ * 2):
        raise ValueError(f"offset shape {tuple(offset_shape)} != {(n, ho, wo, groups * p_ * 2)}")
    if tuple(mask_shape) != (n, ho, wo, groups * p_):
        raise ValueError(f"mask shape {tuple(mask_shape)} != {(n, ho, wo, groups * p_)}")
    return hin, win, ho, wo


# --------------------------------------------------------------------------------------------------
# literal restatement
# --------------------------------------------------------------------------------------------------
def _reference_points(hin, win, kh, kw, dh, dw, sh, sw, dtype):
    """utils.py:14-58.  Returns [1,Ho,Wo,1,2] with (ref_y, ref_x) stacked in that order (:52)."""
    ho = (hin - (dh * (kh - 1) + 1)) // sh + 1
    wo = (win - (dw * (kw - 1) + 1)) // sw + 1
    y_start = np.float32((dh * (kh - 1)) // 2 + 0.5)
    x_start = np.float32((dw * (kw - 1)) // 2 + 0.5)
    # tf.linspace between start and start+(n-1)*stride: every point is start + i*stride exactly
    ys = (y_start + np.arange(ho, dtype=np.float32) * np.float32(sh)).astype(np.float32)
    xs = (x_start + np.arange(wo, dtype=np.float32) * np.float32(sw)).astype(np.float32)
    ref_y, ref_x = np.meshgrid(ys, xs, indexing="ij")
    ref_y = ref_y.astype(dtype).reshape(1, -1) / dtype.type(hin)
    ref_x = ref_x.astype(dtype).reshape(1, -1) / dtype.type(win)
    return np.stack([ref_y, ref_x], axis=-1).reshape(1, ho, wo, 1, 2)


def _dilation_grid(hin, win, kh, kw, dh, dw, groups, dtype):
    """utils.py:65-103.  [1,1,1,G*P,2]; channel 0 = W-direction displacement / W_in, varying with the
    slow tap index i; channel 1 = H-direction displacement / H_in, varying with j."""
    xs = -((dw * (kw - 1)) // 2) + np.arange(kw) * dw
    ys = -((dh * (kh - 1)) // 2) + np.arange(kh) * dh
    gx, gy = np.meshgrid(xs, ys, indexing="ij")  # [kw, kh]
    gx = gx.astype(dtype) / dtype.type(win)
    gy = gy.astype(dtype) / dtype.type(hin)
    grid = np.stack([gx, gy], axis=-1).reshape(-1, 1, 2)
    grid = np.tile(grid, (1, groups, 1)).transpose(1, 0, 2)
    return grid.reshape(1, 1, 1, groups * kh * kw, 2)


def forward_literal(x, offset, mask, kernel_size=(3, 3), strides=(1, 1), padding="SAME",
                    dilation_rate=(1, 1), groups=4, group_channels=16, offset_scale=1.0):
    """op.py:16-109 followed statement by statement; the sampler is utils.py:110-210."""
    ph, pw = resolve_padding(kernel_size, padding)
    kh, kw = kernel_size
    dh, dw = dilation_rate
    sh, sw = strides
    dtype = x.dtype
    s = dtype.type(offset_scale)
    xp = np.pad(x, [(0, 0), (ph, ph), (pw, pw), (0, 0)])  # op.py:46
    n, hin, win, c = xp.shape
    _, ho, wo, _ = offset.shape  # op.py:51
    p_ = kh * kw
    ref = _reference_points(hin, win, kh, kw, dh, dw, sh, sw, dtype)
    grid = _dilation_grid(hin, win, kh, kw, dh, dw, groups, dtype)
    norm = np.tile(np.array([win, hin]).reshape(1, 1, 1, 2), (1, 1, 1, groups * p_)).astype(dtype)
    loc = (ref + grid * s).reshape(1, ho, wo, groups * p_ * 2)  # op.py:82-83
    loc = loc + offset * s / norm  # op.py:85
    grids = 2 * loc - 1  # op.py:87
    img = xp.reshape(n, hin, win, groups, group_channels).transpose(0, 3, 1, 2, 4)
    img = img.reshape(n * groups, hin, win, group_channels)  # op.py:89-91
    grids = grids.reshape(n, ho * wo, groups, p_, 2).transpose(3, 0, 2, 1, 4)
    grids = grids.reshape(p_, n * groups, ho * wo, 2)  # op.py:93-95
    m = mask.reshape(n, ho * wo, groups, p_).transpose(3, 0, 2, 1)
    m = m.reshape(p_, n * groups, ho * wo, 1)  # op.py:97-99

    # ---- utils.py:110-210 ----
    max_y, max_x = hin - 1, win - 1
    gx, gy = grids[..., 0], grids[..., 1]
    xq = dtype.type(0.5) * ((gx + dtype.type(1.0)) * dtype.type(max_x - 1))  # :142
    yq = dtype.type(0.5) * ((gy + dtype.type(1.0)) * dtype.type(max_y - 1))  # :143
    x0 = np.floor(xq).astype(np.int32)
    y0 = np.floor(yq).astype(np.int32)
    x1, y1 = x0 + 1, y0 + 1
    x0, x1 = np.clip(x0, 0, max_x), np.clip(x1, 0, max_x)  # :152-155
    y0, y1 = np.clip(y0, 0, max_y), np.clip(y1, 0, max_y)
    dx0, dx1 = xq - x0.astype(dtype), x1.astype(dtype) - xq  # :163-166 (clipped corners)
    dy0, dy1 = yq - y0.astype(dtype), y1.astype(dtype) - yq
    deltas = np.stack([dx1 * dy1, dx1 * dy0, dx0 * dy1, dx0 * dy0], axis=1)  # wa wb wc wd :169-174
    all_x = np.stack([x0, x0, x1, x1], axis=1)  # :177
    all_y = np.stack([y0, y1, y0, y1], axis=1)  # :178
    nb, npts = n * groups, ho * wo
    deltas = deltas.reshape(p_, 4, nb * npts, 1)
    m = m.reshape(p_, nb * npts, 1)
    b = np.tile(np.arange(nb).reshape(1, nb, 1), (4, 1, npts)).reshape(4, -1)
    y = np.zeros((nb * npts, group_channels), dtype=dtype)
    for i in range(p_):  # :195-206
        vals = img[b, all_y[i].reshape(4, -1), all_x[i].reshape(4, -1)]  # GatherNd [4,M,gc]
        vals = vals * deltas[i]
        vals = vals[0] + vals[1] + vals[2] + vals[3]
        y = y + vals * m[i]
    out = y.reshape(n, groups, ho, wo, group_channels).transpose(0, 2, 3, 1, 4)
    return np.ascontiguousarray(out.reshape(n, ho, wo, groups * group_channels))  # op.py:105-107


# --------------------------------------------------------------------------------------------------
# per-point restatement (the form the CUDA kernels and oracle/dcnv3_oracle.c implement)
# --------------------------------------------------------------------------------------------------
class _Taps:
    """Coordinates, clipped corners and bilinear weights of every sampled point, [N,Ho,Wo,G,P]."""


def _taps(offset, hin, win, kernel_size, strides, dilation_rate, groups, offset_scale):
    dtype = offset.dtype
    kh, kw = kernel_size
    sh, sw = strides
    dh, dw = dilation_rate
    n, ho, wo, _ = offset.shape
    p_ = kh * kw
    s = dtype.type(offset_scale)
    off = offset.reshape(n, ho, wo, groups, p_, 2)
    hh = np.arange(ho, dtype=np.float32) * np.float32(sh) + np.float32((dh * (kh - 1)) // 2 + 0.5)
    ww = np.arange(wo, dtype=np.float32) * np.float32(sw) + np.float32((dw * (kw - 1)) // 2 + 0.5)
    ref0 = (hh.astype(dtype) / dtype.type(hin)).reshape(1, ho, 1, 1, 1)  # ref_y -> channel 0 (:52)
    ref1 = (ww.astype(dtype) / dtype.type(win)).reshape(1, 1, wo, 1, 1)
    pi, pj = np.divmod(np.arange(p_), kh)  # p = i*kh + j
    g0 = ((-((dw * (kw - 1)) // 2) + pi * dw).astype(dtype) / dtype.type(win)).reshape(1, 1, 1, 1, p_)
    g1 = ((-((dh * (kh - 1)) // 2) + pj * dh).astype(dtype) / dtype.type(hin)).reshape(1, 1, 1, 1, p_)
    loc0 = (ref0 + g0 * s) + off[..., 0] * s / dtype.type(win)  # op.py:82-85
    loc1 = (ref1 + g1 * s) + off[..., 1] * s / dtype.type(hin)
    t = _Taps()
    t.xq = dtype.type(0.5) * (((2 * loc0 - 1) + dtype.type(1.0)) * dtype.type(win - 2))
    t.yq = dtype.type(0.5) * (((2 * loc1 - 1) + dtype.type(1.0)) * dtype.type(hin - 2))
    fx = np.floor(t.xq).astype(np.int64)
    fy = np.floor(t.yq).astype(np.int64)
    t.x0, t.x1 = np.clip(fx, 0, win - 1), np.clip(fx + 1, 0, win - 1)
    t.y0, t.y1 = np.clip(fy, 0, hin - 1), np.clip(fy + 1, 0, hin - 1)
    t.dx0, t.dx1 = t.xq - t.x0.astype(dtype), t.x1.astype(dtype) - t.xq
    t.dy0, t.dy1 = t.yq - t.y0.astype(dtype), t.y1.astype(dtype) - t.yq
    return t


def _corner_values(xp, t, groups, group_channels):
    """I_a..I_d [N,Ho,Wo,G,P,gc] for corners (y0,x0) (y1,x0) (y0,x1) (y1,x1) -- utils.py:177-178."""
    n, hin, win, _ = xp.shape
    img = xp.reshape(n, hin, win, groups, group_channels)
    nn = np.arange(n).reshape(n, 1, 1, 1, 1)
    gg = np.arange(groups).reshape(1, 1, 1, groups, 1)
    return (img[nn, t.y0, t.x0, gg], img[nn, t.y1, t.x0, gg],
            img[nn, t.y0, t.x1, gg], img[nn, t.y1, t.x1, gg])


def forward(x, offset, mask, kernel_size=(3, 3), strides=(1, 1), padding="SAME",
            dilation_rate=(1, 1), groups=4, group_channels=16, offset_scale=1.0):
    """out[n,h,w,g,:] = sum_p mask * (wa*Ia + wb*Ib + wc*Ic + wd*Id), taps added in order p=0..P-1
    (utils.py:195-206)."""
    ph, pw = resolve_padding(kernel_size, padding)
    hin, win, ho, wo = check_shapes(x.shape, offset.shape, mask.shape, kernel_size, strides,
                                    (ph, pw), dilation_rate, groups, group_channels)
    n = x.shape[0]
    p_ = kernel_size[0] * kernel_size[1]
    xp = np.pad(x, [(0, 0), (ph, ph), (pw, pw), (0, 0)])
    t = _taps(offset, hin, win, kernel_size, strides, dilation_rate, groups, offset_scale)
    ia, ib, ic, id_ = _corner_values(xp, t, groups, group_channels)
    wa, wb = (t.dx1 * t.dy1)[..., None], (t.dx1 * t.dy0)[..., None]
    wc, wd = (t.dx0 * t.dy1)[..., None], (t.dx0 * t.dy0)[..., None]
    s_ = ((ia * wa + ib * wb) + ic * wc) + id_ * wd
    m = mask.reshape(n, ho, wo, groups, p_, 1)
    out = np.zeros((n, ho, wo, groups, group_channels), dtype=x.dtype)
    for p in range(p_):
        out = out + s_[..., p, :] * m[..., p, :]
    return out.reshape(n, ho, wo, groups * group_channels)


def backward(x, offset, mask, grad_out, kernel_size=(3, 3), strides=(1, 1), padding="SAME",
             dilation_rate=(1, 1), groups=4, group_channels=16, offset_scale=1.0, accumulate=None):
    """(grad_x, grad_offset, grad_mask).  `accumulate=np.float64` keeps every per-tap quantity in
    x.dtype (as the reference computes it) but sums the grad_x scatter in float64 -- the yardstick for
    long collision chains, where an x.dtype running sum is itself ~sqrt(n)*eps off.  d/d offset flows only through dx0 = xq - x0f etc.
    (utils.py:163-166); chain factor d xq / d offset = (W_in-2)*s/W_in (op.py:85-87, utils.py:142).
    Contributions to the zero ring of the padded image are dropped (gradient of tf.pad, op.py:46)."""
    ph, pw = resolve_padding(kernel_size, padding)
    hin, win, ho, wo = check_shapes(x.shape, offset.shape, mask.shape, kernel_size, strides,
                                    (ph, pw), dilation_rate, groups, group_channels)
    dtype = x.dtype
    n = x.shape[0]
    p_ = kernel_size[0] * kernel_size[1]
    xp = np.pad(x, [(0, 0), (ph, ph), (pw, pw), (0, 0)])
    t = _taps(offset, hin, win, kernel_size, strides, dilation_rate, groups, offset_scale)
    ia, ib, ic, id_ = _corner_values(xp, t, groups, group_channels)
    go = grad_out.reshape(n, ho, wo, groups, 1, group_channels)
    m = mask.reshape(n, ho, wo, groups, p_)
    wa, wb, wc, wd = t.dx1 * t.dy1, t.dx1 * t.dy0, t.dx0 * t.dy1, t.dx0 * t.dy0
    da, db = (go * ia).sum(-1), (go * ib).sum(-1)
    dc, dd = (go * ic).sum(-1), (go * id_).sum(-1)
    grad_mask = ((wa * da + wb * db) + wc * dc) + wd * dd
    g_xq = m * (t.dy1 * (dc - da) + t.dy0 * (dd - db))
    g_yq = m * (t.dx1 * (db - da) + t.dx0 * (dd - dc))
    s = dtype.type(offset_scale)
    fx = dtype.type(win - 2) * s / dtype.type(win)
    fy = dtype.type(hin - 2) * s / dtype.type(hin)
    grad_offset = np.stack([g_xq * fx, g_yq * fy], axis=-1).reshape(offset.shape).astype(dtype)
    grad_mask = grad_mask.reshape(mask.shape).astype(dtype)

    gxp = np.zeros((n, hin, win, groups, group_channels), dtype=accumulate or dtype)
    nn = np.broadcast_to(np.arange(n).reshape(n, 1, 1, 1, 1), t.x0.shape)
    gg = np.broadcast_to(np.arange(groups).reshape(1, 1, 1, groups, 1), t.x0.shape)
    gs = go * m[..., None]  # [N,Ho,Wo,G,P,gc]
    for (yy, xx, w_) in ((t.y0, t.x0, wa), (t.y1, t.x0, wb), (t.y0, t.x1, wc), (t.y1, t.x1, wd)):
        np.add.at(gxp, (nn, yy, xx, gg), gs * w_[..., None])
    grad_x = gxp[:, ph:hin - ph, pw:win - pw].reshape(x.shape)
    return np.ascontiguousarray(grad_x).astype(dtype), grad_offset, grad_mask


# --------------------------------------------------------------------------------------------------
# the reference under mixed_bfloat16: every primitive rounds to bfloat16 (op.py:62-87, utils.py:130-206)
# --------------------------------------------------------------------------------------------------
def rb(a):
    """float32 -> nearest bfloat16 (ties to even) -> float32, the rounding every bf16 TF / torch primitive
    applies to its result."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = r.view(np.float32).reshape(a.shape)
    return np.where(np.isfinite(a), out, a)


def _taps_bf16(offset, hin, win, kernel_size, strides, dilation_rate, groups, offset_scale):
    """_taps with x.dtype = bfloat16: the same operations in the same order, each result rounded with rb().
    Python scalars become tensors of x.dtype first (offset_scale -> rb(offset_scale)), ints cast to x.dtype
    (H_in, W_in, the clipped corner indices, utils.py:158-161) are rounded too."""
    f32 = np.float32
    kh, kw = kernel_size
    sh, sw = strides
    dh, dw = dilation_rate
    n, ho, wo, _ = offset.shape
    p_ = kh * kw
    s = rb(f32(offset_scale))
    off = offset.reshape(n, ho, wo, groups, p_, 2).astype(f32)
    hh = np.arange(ho, dtype=f32) * f32(sh) + f32((dh * (kh - 1)) // 2 + 0.5)  # fp32 linspace (utils.py:35-36)
    ww = np.arange(wo, dtype=f32) * f32(sw) + f32((dw * (kw - 1)) // 2 + 0.5)
    hin_b, win_b = rb(f32(hin)), rb(f32(win))
    ref0 = rb(rb(hh) / hin_b).reshape(1, ho, 1, 1, 1)  # utils.py:40-50; ref_y -> channel 0 (:52)
    ref1 = rb(rb(ww) / win_b).reshape(1, 1, wo, 1, 1)
    pi, pj = np.divmod(np.arange(p_), kh)
    g0 = rb(rb((-((dw * (kw - 1)) // 2) + pi * dw).astype(f32)) / win_b).reshape(1, 1, 1, 1, p_)
    g1 = rb(rb((-((dh * (kh - 1)) // 2) + pj * dh).astype(f32)) / hin_b).reshape(1, 1, 1, 1, p_)
    loc0 = rb(rb(ref0 + rb(g0 * s)) + rb(rb(off[..., 0] * s) / win_b))  # op.py:82-85
    loc1 = rb(rb(ref1 + rb(g1 * s)) + rb(rb(off[..., 1] * s) / hin_b))
    t = _Taps()
    t.xq = rb(f32(0.5) * rb(rb(rb(rb(f32(2) * loc0) - f32(1)) + f32(1)) * rb(f32(win - 2))))  # op.py:87, utils.py:142
    t.yq = rb(f32(0.5) * rb(rb(rb(rb(f32(2) * loc1) - f32(1)) + f32(1)) * rb(f32(hin - 2))))
    fx = np.floor(t.xq).astype(np.int64)
    fy = np.floor(t.yq).astype(np.int64)
    t.x0, t.x1 = np.clip(fx, 0, win - 1), np.clip(fx + 1, 0, win - 1)
    t.y0, t.y1 = np.clip(fy, 0, hin - 1), np.clip(fy + 1, 0, hin - 1)
    t.dx0, t.dx1 = rb(t.xq - rb(t.x0.astype(f32))), rb(rb(t.x1.astype(f32)) - t.xq)  # utils.py:158-166
    t.dy0, t.dy1 = rb(t.yq - rb(t.y0.astype(f32))), rb(rb(t.y1.astype(f32)) - t.yq)
    return t


def forward_bf16(x, offset, mask, kernel_size=(3, 3), strides=(1, 1), padding="SAME",
                 dilation_rate=(1, 1), groups=4, group_channels=16, offset_scale=1.0):
    """The forward as the reference computes it when x.dtype is bfloat16.  Inputs: float32 arrays holding
    bf16-representable values.  Weights (utils.py:169-172), the four corner products (:201), their sum (:202,
    one rounding: the reduction accumulates in fp32), the mask product (:204) and the running output (:206)
    are each rounded to bf16.  Pinned against tests/golden/op_bf16_*.npz (the unmodified reference run on
    bf16 tensors through the shim)."""
    ph, pw = resolve_padding(kernel_size, padding)
    hin, win, ho, wo = check_shapes(x.shape, offset.shape, mask.shape, kernel_size, strides,
                                    (ph, pw), dilation_rate, groups, group_channels)
    n = x.shape[0]
    p_ = kernel_size[0] * kernel_size[1]
    xp = np.pad(x.astype(np.float32), [(0, 0), (ph, ph), (pw, pw), (0, 0)])
    t = _taps_bf16(offset, hin, win, kernel_size, strides, dilation_rate, groups, offset_scale)
    ia, ib, ic, id_ = _corner_values(xp, t, groups, group_channels)
    wa, wb = rb(t.dx1 * t.dy1)[..., None], rb(t.dx1 * t.dy0)[..., None]
    wc, wd = rb(t.dx0 * t.dy1)[..., None], rb(t.dx0 * t.dy0)[..., None]
    s_ = rb(((rb(ia * wa) + rb(ib * wb)) + rb(ic * wc)) + rb(id_ * wd))
    m = mask.reshape(n, ho, wo, groups, p_, 1).astype(np.float32)
    out = np.zeros((n, ho, wo, groups, group_channels), dtype=np.float32)
    for p in range(p_):
        out = rb(out + rb(s_[..., p, :] * m[..., p, :]))
    return out.reshape(n, ho, wo, groups * group_channels)


def backward_bf16_coords(x, offset, mask, grad_out, kernel_size=(3, 3), strides=(1, 1), padding="SAME",
                         dilation_rate=(1, 1), groups=4, group_channels=16, offset_scale=1.0):
    """Gradients with the sampling cells and bilinear weights of the bf16 reference (_taps_bf16) and fp32
    gradient arithmetic -- what the kernels compute under DCNV3_FLAG_REF_DTYPE.  (What TF / XLA autodiff does
    in bf16 -- every backward primitive rounded to bf16, bf16 scatter accumulation -- is an implementation
    detail that is not reproduced; tests record the distance to the shim's bf16 autograd instead.)"""
    ph, pw = resolve_padding(kernel_size, padding)
    hin, win, ho, wo = check_shapes(x.shape, offset.shape, mask.shape, kernel_size, strides,
                                    (ph, pw), dilation_rate, groups, group_channels)
    f32 = np.float32
    n = x.shape[0]
    p_ = kernel_size[0] * kernel_size[1]
    xp = np.pad(x.astype(f32), [(0, 0), (ph, ph), (pw, pw), (0, 0)])
    t = _taps_bf16(offset, hin, win, kernel_size, strides, dilation_rate, groups, offset_scale)
    ia, ib, ic, id_ = _corner_values(xp, t, groups, group_channels)
    go = grad_out.reshape(n, ho, wo, groups, 1, group_channels).astype(f32)
    m = mask.reshape(n, ho, wo, groups, p_).astype(f32)
    wa, wb, wc, wd = t.dx1 * t.dy1, t.dx1 * t.dy0, t.dx0 * t.dy1, t.dx0 * t.dy0
    da, db = (go * ia).sum(-1), (go * ib).sum(-1)
    dc, dd = (go * ic).sum(-1), (go * id_).sum(-1)
    grad_mask = ((wa * da + wb * db) + wc * dc) + wd * dd
    g_xq = m * (t.dy1 * (dc - da) + t.dy0 * (dd - db))
    g_yq = m * (t.dx1 * (db - da) + t.dx0 * (dd - dc))
    s = f32(offset_scale)
    fx, fy = f32(win - 2) * s / f32(win), f32(hin - 2) * s / f32(hin)
    grad_offset = np.stack([g_xq * fx, g_yq * fy], axis=-1).reshape(offset.shape).astype(f32)
    gxp = np.zeros((n, hin, win, groups, group_channels), dtype=np.float64)
    nn = np.broadcast_to(np.arange(n).reshape(n, 1, 1, 1, 1), t.x0.shape)
    gg = np.broadcast_to(np.arange(groups).reshape(1, 1, 1, groups, 1), t.x0.shape)
    gs = go * m[..., None]
    for (yy, xx, w_) in ((t.y0, t.x0, wa), (t.y1, t.x0, wb), (t.y0, t.x1, wc), (t.y1, t.x1, wd)):
        np.add.at(gxp, (nn, yy, xx, gg), gs * w_[..., None])
    grad_x = gxp[:, ph:hin - ph, pw:win - pw].reshape(x.shape)
    return np.ascontiguousarray(grad_x).astype(f32), grad_offset, grad_mask.reshape(mask.shape).astype(f32)


def mask_softmax(logits, groups):
    """layers/dcn_v3/dcn_v3.py:120-123 -- softmax over the P taps of each group."""
    n, h, w, gp = logits.shape
    z = logits.reshape(n, h, w, groups, gp // groups)
    z = z - z.max(-1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(-1, keepdims=True)).reshape(n, h, w, gp).astype(logits.dtype)


def sampled_points(offset_shape):
    """One sampled point = one (n,h,w,g,p) tuple; offset has 2 values per point."""
    n, ho, wo, c = offset_shape
    return n * ho * wo * (c // 2)
