"""Parity of the CUDA path (through the C ABI) against the oracle and the golden vectors.

Bar: fp32  max|d|/max|ref| <= 1e-5 ; bf16 <= 1e-2 against the fp32 oracle evaluated
on the bf16-rounded inputs ; gradients bit-identical run to run."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import GOLDEN, golden_op_cases, load_op_case, make_inputs, rel_err
from oracle import c_oracle, dcnv3_oracle as O

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-5
TOL_BF16 = 1e-2


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import iseg_b200
    from iseg_b200 import _cabi
    return iseg_b200, _cabi


def cuda(*arrs, dtype=torch.float32):
    return [torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype) for a in arrs]


def run_op(ops, x, off, m, go=None, dtype=torch.float32, mask_is_logits=False, reference_dtype_math=False, **kw):
    iseg, _ = ops
    tx, to, tm = cuda(x, off, m, dtype=dtype)
    if go is not None:
        tx.requires_grad_(True), to.requires_grad_(True), tm.requires_grad_(True)
    out = iseg.dcnv3_op(tx, to, tm, list(kw.get("kernel_size", (3, 3))), list(kw.get("strides", (1, 1))),
                        kw.get("padding", "SAME"), list(kw.get("dilation_rate", (1, 1))), kw["groups"],
                        kw["group_channels"], kw.get("offset_scale", 1.0), mask_is_logits=mask_is_logits,
                        reference_dtype_math=reference_dtype_math)
    if go is None:
        return out.float().cpu().numpy()
    out.backward(cuda(go, dtype=dtype)[0])
    return tuple(t.float().cpu().numpy() for t in (out.detach(), tx.grad, to.grad, tm.grad))


@pytest.mark.parametrize("name", [n for n in golden_op_cases() if not n.endswith("f64")])
def test_golden_fp32(ops, name):
    z, kw = load_op_case(name)
    out, gx, goff, gm = run_op(ops, z["x"], z["offset"], z["mask"], z["grad_out"], **kw)
    assert rel_err(out, z["out"]) <= TOL_F32
    assert rel_err(gx, z["grad_x"]) <= TOL_F32
    assert rel_err(goff, z["grad_offset"]) <= TOL_F32
    assert rel_err(gm, z["grad_mask"]) <= TOL_F32


@pytest.mark.parametrize("name", golden_op_cases(bf16=True))
def test_golden_bf16_reference_dtype(ops, name):
    """The reference run on bfloat16 tensors (tests/golden/op_bf16_*, made by executing its unmodified source
    in bf16).  With reference_dtype_math=True the kernels round every intermediate as the reference does: the
    forward matches bit for bit or to one bf16 ulp, all four results within the 1e-2 bar of north_star."""
    z, kw = load_op_case(name)
    out, gx, goff, gm = run_op(ops, z["x"], z["offset"], z["mask"], z["grad_out"], dtype=torch.bfloat16,
                               reference_dtype_math=True, **kw)
    assert rel_err(out, z["out"]) <= TOL_BF16
    assert (out != z["out"]).mean() < 1e-3   # essentially bit-exact (exp / division free path: identical roundings)
    assert rel_err(goff, z["grad_offset"]) <= TOL_BF16
    assert rel_err(gm, z["grad_mask"]) <= TOL_BF16
    # grad_x: the fixture's own gradient is a bf16 scatter-add (torch autograd in bf16 standing in for TF's; every
    # one of the up to ~100 additions into a cell rounds to 8 bits), so the fixture itself is only good to about
    # 1e-2; against the same cells / weights accumulated exactly the kernel is within the bar
    assert rel_err(gx, z["grad_x"]) <= 2 * TOL_BF16
    rx, roff, rm = O.backward_bf16_coords(z["x"], z["offset"], z["mask"], z["grad_out"], **kw)
    assert rel_err(gx, rx) <= TOL_BF16 and rel_err(goff, roff) <= TOL_BF16 and rel_err(gm, rm) <= TOL_BF16
    # the default bf16 mode keeps coordinates in fp32: a different (documented) result
    out32 = run_op(ops, z["x"], z["offset"], z["mask"], dtype=torch.bfloat16, **kw)
    assert rel_err(out32, c_oracle.forward(z["x"], z["offset"], z["mask"], **kw)) <= TOL_BF16


@pytest.mark.parametrize("case", [(2, 64, 64, 8, 16, 1.0, 1.0), (1, 160, 160, 10, 16, 1.0, 2.0), (1, 193, 193, 4, 16, 1.0, 1.0)])
def test_random_bf16_reference_dtype_vs_oracle(ops, case):
    n, h, w, g, gc, sigma, scale = case
    x, off, m, go = make_inputs(n, h, w, g, gc, sigma=sigma, seed=3 + h)
    rnd = lambda a: torch.from_numpy(a).bfloat16().float().numpy()  # noqa: E731
    x, off, m, go = rnd(x), rnd(off), rnd(m), rnd(go)
    kw = dict(groups=g, group_channels=gc, offset_scale=scale)
    out, gx, goff, gm = run_op(ops, x, off, m, go, dtype=torch.bfloat16, reference_dtype_math=True, **kw)
    ref = O.forward_bf16(x, off, m, **kw)
    assert rel_err(out, ref) <= TOL_BF16 and (out != ref).mean() < 1e-3
    rx, roff, rm = O.backward_bf16_coords(x, off, m, go, **kw)
    assert rel_err(gx, rx) <= TOL_BF16 and rel_err(goff, roff) <= TOL_BF16 and rel_err(gm, rm) <= TOL_BF16


def test_known_answers(ops):
    z = np.load(f"{GOLDEN}/kat_ramp.npz")
    kw = dict(groups=2, group_channels=1)
    out = run_op(ops, z["x"], z["offset"], z["mask"], **kw)
    assert abs(out[0, 1, 4, 0] - 32.125) < 1e-4 and abs(out[0, 5, 5, 0] - 42.625) < 1e-4
    assert rel_err(out, z["out"]) <= TOL_F32
    assert rel_err(run_op(ops, z["x"], z["offset_shift"], z["mask"], **kw), z["out_shift"]) <= TOL_F32
    res = run_op(ops, z["x"], z["offset_dead"], z["mask"], np.ones_like(z["out"]), **kw)
    assert all(not r.any() for r in res)  # dead taps: output and all three gradients exactly 0
    zc = np.load(f"{GOLDEN}/kat_const.npz")
    assert rel_err(run_op(ops, zc["x"], zc["offset"], zc["mask"], groups=1, group_channels=16), zc["out"]) <= TOL_F32


CASES = [  # n, h, w, G, gc, sigma, offset_scale
    (2, 128, 128, 4, 16, 1.0, 1.0),    # BASELINE config 1
    (2, 64, 64, 8, 16, 1.0, 1.0),
    (3, 32, 32, 16, 16, 2.0, 1.0),
    (2, 16, 16, 32, 16, 1.0, 1.0),
    (1, 49, 97, 4, 16, 4.0, 1.0),      # odd, non-square (sliding-window tile shapes), wide offsets
    (1, 40, 40, 10, 16, 1.0, 2.0),     # InternImage-L style: G=10, offset_scale 2
    (1, 20, 20, 40, 32, 1.0, 2.0),     # gc = 32 (InternImage-H): tiled kernels on half groups
    (2, 40, 33, 5, 32, 2.0, 1.0),      # gc = 32, odd group count, several tiles per image
    (1, 70, 64, 10, 32, 4.0, 1.0),     # gc = 32, wide offsets (out-of-box taps, landings beyond the ring)
    (1, 24, 24, 7, 16, 1.0, 1.0),      # odd group count (InternImage-B): trailing half-empty chunk
    (2, 40, 33, 5, 16, 2.0, 1.0),      # InternImage-S stage 1 style (G=5)
    (1, 9, 11, 3, 5, 1.5, 1.0),        # gc not a multiple of 4 -> scalar path
    (1, 20, 20, 80, 16, 1.0, 2.0),     # InternImage-L stage 4: G = 80, offset_scale 2
    (3, 64, 64, 2, 16, 8.0, 1.0),      # most landings beyond the scatter ring (64-bit side path)
]


@pytest.mark.parametrize("case", CASES)
def test_random_fp32_vs_c_oracle(ops, case):
    n, h, w, g, gc, sigma, scale = case
    x, off, m, go = make_inputs(n, h, w, g, gc, sigma=sigma, seed=h * 131 + g)
    off.reshape(-1)[::97] *= 50.0  # ~1 % outliers: exercise clipping / dead taps
    kw = dict(groups=g, group_channels=gc, offset_scale=scale)
    out, gx, goff, gm = run_op(ops, x, off, m, go, **kw)
    ref_out = c_oracle.forward(x, off, m, **kw)
    rx, roff, rm = c_oracle.backward(x, off, m, go, **kw)
    assert rel_err(out, ref_out) <= TOL_F32
    assert rel_err(gx, rx) <= TOL_F32
    assert rel_err(goff, roff) <= TOL_F32
    assert rel_err(gm, rm) <= TOL_F32


@pytest.mark.parametrize("case", CASES[:4] + CASES[5:11] + [(1, 20, 300, 4, 16, 1.0, 1.0), (1, 130, 70, 6, 16, 3.0, 2.0),
                                                           (1, 20, 20, 80, 16, 1.0, 2.0)])
def test_random_bf16(ops, case):
    n, h, w, g, gc, sigma, scale = case
    x, off, m, go = make_inputs(n, h, w, g, gc, sigma=sigma, seed=7 + h)
    rnd = lambda a: torch.from_numpy(a).bfloat16().float().numpy()  # noqa: E731
    x, off, m, go = rnd(x), rnd(off), rnd(m), rnd(go)
    kw = dict(groups=g, group_channels=gc, offset_scale=scale)
    out, gx, goff, gm = run_op(ops, x, off, m, go, dtype=torch.bfloat16, **kw)
    ref_out = c_oracle.forward(x, off, m, **kw)
    rx, roff, rm = c_oracle.backward(x, off, m, go, **kw)
    assert rel_err(out, ref_out) <= TOL_BF16
    assert rel_err(gx, rx) <= TOL_BF16
    assert rel_err(goff, roff) <= TOL_BF16
    assert rel_err(gm, rm) <= TOL_BF16


def test_fuzz_shapes_offsets_vs_c_oracle(ops):
    """Seeded fuzzing over what the reference's signature admits: odd / non-square sizes, any group count,
    gc in {4,8,16,32}, offset_scale, offset spread from 0 to far beyond the staged halo, raw (un-normalised,
    signed) masks, VALID padding, 1x1..5x5 kernels.  Tiled and generic paths are both hit."""
    rng = np.random.default_rng(2024)
    for trial in range(24):
        k = int(rng.choice([3, 3, 3, 3, 1, 5]))
        padding = str(rng.choice(["SAME", "SAME", "SAME", "VALID"]))
        g = int(rng.integers(1, 11))
        gc = int(rng.choice([16, 16, 16, 4, 8, 32]))
        n, h, w = int(rng.integers(1, 4)), int(rng.integers(k, 45)), int(rng.integers(k, 45))
        sigma = float(rng.choice([0.0, 0.5, 1.0, 3.0, 10.0]))
        scale = float(rng.choice([1.0, 1.0, 2.0, 0.5]))
        ho, wo = (h, w) if padding == "SAME" else (h - k + 1, w - k + 1)
        p = k * k
        x = rng.standard_normal((n, h, w, g * gc), dtype=np.float32)
        off = (sigma * rng.standard_normal((n, ho, wo, g * p * 2), dtype=np.float32)).astype(np.float32)
        m = rng.standard_normal((n, ho, wo, g * p), dtype=np.float32)  # raw mask: any sign / magnitude
        if trial % 3 == 0:
            m = O.mask_softmax(m, g)
        if trial % 5 == 0:
            m *= 7.0  # beyond the fixed-point weight range of the fast path -> exact side path
        go = rng.standard_normal((n, ho, wo, g * gc), dtype=np.float32) * float(rng.choice([1.0, 1e-6, 1e4]))
        kw = dict(kernel_size=(k, k), padding=padding, groups=g, group_channels=gc, offset_scale=scale)
        out, gx, goff, gm = run_op(ops, x, off, m, go, **kw)
        ref_out = c_oracle.forward(x, off, m, **kw)
        _, roff, rm = c_oracle.backward(x, off, m, go, **kw)
        # grad_x yardstick: fp32 per-tap arithmetic (as the reference) with an exact (float64) scatter sum --
        # with large cancelling contributions an fp32 running sum is itself several 1e-5 off
        rx, _, _ = O.backward(x, off, m, go, accumulate=np.float64, **kw)
        tag = (trial, n, h, w, g, gc, k, padding, sigma, scale)
        for a, b, name in ((out, ref_out, "out"), (gx, rx, "grad_x"), (goff, roff, "grad_offset"), (gm, rm, "grad_mask")):
            assert rel_err(a, b) <= TOL_F32 or not np.abs(b).max() > 0, (name, rel_err(a, b), tag)


@pytest.mark.parametrize("shape", [(1, 20, 300, 4), (1, 130, 70, 6), (2, 97, 49, 2), (1, 257, 33, 3), (1, 33, 257, 2)])
@pytest.mark.parametrize("scale,sigma", [(1.0, 1.0), (2.0, 1.5), (0.5, 6.0)])
def test_wide_and_tall_images(ops, shape, scale, sigma):
    """Strongly non-square images: the reference pairs output rows with input columns, so the
    scatter tiles see very different numbers of home pixels per tile, partial tiles on both axes, rings
    clipped by the image and -- at sigma = 6 -- many landings beyond the ring (64-bit side path)."""
    n, h, w, g = shape
    x, off, m, go = make_inputs(n, h, w, g, 16, sigma=sigma, seed=h * 7 + w)
    kw = dict(groups=g, group_channels=16, offset_scale=scale)
    out, gx, goff, gm = run_op(ops, x, off, m, go, **kw)
    ref_out = c_oracle.forward(x, off, m, **kw)
    rx, roff, rm = c_oracle.backward(x, off, m, go, **kw)
    assert rel_err(out, ref_out) <= TOL_F32
    assert rel_err(gx, rx) <= TOL_F32
    assert rel_err(goff, roff) <= TOL_F32
    assert rel_err(gm, rm) <= TOL_F32
    again = run_op(ops, x, off, m, go, **kw)
    assert np.array_equal(gx, again[1])


def test_workspace_stays_zeroed_across_generic_and_tiled(ops):
    """Both backward paths share one cached, zeroed-once workspace (DCNV3_FLAG_WORKSPACE_ZEROED): whichever
    ran last must leave it all-zero, including after far landings and hot cells."""
    _, cabi = ops
    n, h, w, g, gc = 2, 40, 70, 4, 16
    x, off, m, go = make_inputs(n, h, w, g, gc, sigma=5.0, seed=11)
    rx, roff, rm = c_oracle.backward(x, off, m, go, groups=g, group_channels=gc)
    t = [torch.from_numpy(a).cuda() for a in (x, off, m, go)]
    cfg = ((3, 3), (1, 1), (1, 1), (1, 1), g, gc, 1.0)
    for flags in (cabi.FLAG_FORCE_GENERIC, 0, cabi.FLAG_FORCE_GENERIC, 0, 0):
        gx, goff, gm = cabi.backward(*t, *cfg, flags=flags)
        assert rel_err(gx.cpu().numpy(), rx) <= TOL_F32, flags
        assert rel_err(goff.cpu().numpy(), roff) <= TOL_F32 and rel_err(gm.cpu().numpy(), rm) <= TOL_F32
    import ctypes
    prm = cabi.make_params(x.shape, (h, w), *cfg[:4], g, gc, 1.0, cabi.F32)
    zb = int(cabi.lib.dcnv3_backward_workspace_zero_bytes(ctypes.byref(prm)))
    ws = cabi._workspace(t[0].device, int(cabi.lib.dcnv3_backward_workspace_bytes(ctypes.byref(prm))), zb)  # the cached one
    torch.cuda.synchronize()
    assert int(ws[:zb].count_nonzero()) == 0  # header and per-image maxima included (behind zb: scratch, no contract)


def test_backward_bitwise_reproducible(ops):
    x, off, m, go = make_inputs(2, 48, 40, 5, 32, sigma=2.0, seed=4)   # 32 channels per group: half groups
    kw = dict(groups=5, group_channels=32)
    a = run_op(ops, x, off, m, go, **kw)
    for _ in range(2):
        b = run_op(ops, x, off, m, go, **kw)
        assert all(np.array_equal(p, q) for p, q in zip(a, b))
    x, off, m, go = make_inputs(4, 64, 64, 8, 16, sigma=2.0, seed=3)
    kw = dict(groups=8, group_channels=16)
    a = run_op(ops, x, off, m, go, **kw)
    for _ in range(3):
        b = run_op(ops, x, off, m, go, **kw)
        assert all(np.array_equal(p, q) for p, q in zip(a, b))
    # heavy collisions: every tap of every pixel lands in the same cell (~37k contributions per
    # channel).  An fp32 running sum of that length carries ~sqrt(n)*eps error itself, so the yardstick
    # here is the oracle with fp32 per-tap arithmetic (as the reference) but a float64 scatter sum.
    x, off, m, go = x[:1], off[:1], m[:1], go[:1]
    # target cell inside a 32x32 scatter tile, in the ring just across a tile border, and at the image corner
    for target in (30.0, 33.0, 1.0):
        off2 = np.zeros_like(off)
        off2[..., 0::2] = (target - np.arange(64, dtype=np.float32)).reshape(1, 64, 1, 1) * 66 / 64
        off2[..., 1::2] = (target - np.arange(64, dtype=np.float32)).reshape(1, 1, 64, 1) * 66 / 64
        a = run_op(ops, x, off2, m, go, **kw)
        b = run_op(ops, x, off2, m, go, **kw)
        assert all(np.array_equal(p, q) for p, q in zip(a, b))
        rx, roff, rm = c_oracle.backward(x, off2, m, go, **kw)
        assert rel_err(a[2], roff) <= TOL_F32 and rel_err(a[3], rm) <= TOL_F32
        dx, _, _ = O.backward(x, off2, m, go, accumulate=np.float64, **kw)
        assert rel_err(a[1], dx) <= 1e-6        # fixed-point accumulation: exact up to the final rounding
        assert rel_err(rx, dx) <= 1e-4          # (the fp32 running sum of the C oracle is the loose one)
    # the same on an image that is a single scatter tile (hot cells are converted by the scatter CTA itself)
    x, off, m, go = make_inputs(2, 32, 32, 3, 16, sigma=1.0, seed=5)
    kw = dict(groups=3, group_channels=16)
    off2 = np.zeros_like(off)
    off2[..., 0::2] = (12.0 - np.arange(32, dtype=np.float32)).reshape(1, 32, 1, 1) * 34 / 32
    off2[..., 1::2] = (20.0 - np.arange(32, dtype=np.float32)).reshape(1, 1, 32, 1) * 34 / 32
    a = run_op(ops, x, off2, m, go, **kw)
    b = run_op(ops, x, off2, m, go, **kw)
    assert all(np.array_equal(p, q) for p, q in zip(a, b))
    dx, _, _ = O.backward(x, off2, m, go, accumulate=np.float64, **kw)
    assert rel_err(a[1], dx) <= 1e-6
    c = run_op(ops, x, off, m, go, **kw)   # the workspace is left clean: an ordinary call right after
    rx, _, _ = c_oracle.backward(x, off, m, go, **kw)
    assert rel_err(c[1], rx) <= TOL_F32


@pytest.mark.parametrize("shape,sigma", [((2, 24, 20, 4, 16), 1.0), ((1, 70, 45, 5, 16), 4.0), ((2, 40, 36, 3, 32), 2.0)])
def test_fused_softmax_matches_layer_semantics(ops, shape, sigma):
    """(second case: several scatter tiles, odd group count, landings beyond the ring)"""
    n, h, w, g, gc = shape
    x, off, _, go = make_inputs(n, h, w, g, gc, sigma=sigma, seed=11)
    logits = np.random.default_rng(5).standard_normal((n, h, w, g * 9)).astype(np.float32) * 2
    kw = dict(groups=g, group_channels=gc)
    out, gx, goff, gl = run_op(ops, x, off, logits, go, mask_is_logits=True, **kw)
    mask = O.mask_softmax(logits, g)
    assert rel_err(out, c_oracle.forward(x, off, mask, **kw)) <= TOL_F32
    rx, roff, rm = c_oracle.backward(x, off, mask, go, **kw)
    mm, gg = mask.reshape(n, h, w, g, 9), rm.reshape(n, h, w, g, 9)
    rl = (mm * (gg - (mm * gg).sum(-1, keepdims=True))).reshape(n, h, w, g * 9)  # softmax Jacobian
    assert rel_err(gx, rx) <= TOL_F32 and rel_err(goff, roff) <= TOL_F32 and rel_err(gl, rl) <= TOL_F32


def test_empty_and_tiny(ops):
    iseg, _ = ops
    e = lambda *s: torch.zeros(*s, device="cuda")  # noqa: E731
    out = iseg.dcnv3_op(e(0, 8, 8, 64), e(0, 8, 8, 72), e(0, 8, 8, 36), [3, 3], [1, 1], "SAME", [1, 1], 4, 16, 1.0)
    assert out.shape == (0, 8, 8, 64)
    x, off, m, go = make_inputs(1, 1, 1, 1, 16, seed=2)
    kw = dict(groups=1, group_channels=16)
    out = run_op(ops, x, off, m, **kw)
    assert rel_err(out, c_oracle.forward(x, off, m, **kw)) <= TOL_F32 or not c_oracle.forward(x, off, m, **kw).any()


def test_shape_and_dtype_errors(ops):
    iseg, cabi = ops
    x = torch.zeros(1, 8, 8, 64, device="cuda")
    off = torch.zeros(1, 8, 8, 72, device="cuda")
    m = torch.zeros(1, 8, 8, 36, device="cuda")
    with pytest.raises(ValueError):
        iseg.dcnv3_op(x, off[:, :7], m[:, :7], [3, 3], [1, 1], "SAME", [1, 1], 4, 16, 1.0)
    with pytest.raises(ValueError):
        iseg.dcnv3_op(x, off, m, [3, 3], [1, 1], "SAME", [1, 1], 8, 16, 1.0)
    with pytest.raises(TypeError):
        iseg.dcnv3_op(x, off.half(), m, [3, 3], [1, 1], "SAME", [1, 1], 4, 16, 1.0)
    with pytest.raises(TypeError):
        iseg.dcnv3_op(x.half(), off.half(), m.half(), [3, 3], [1, 1], "SAME", [1, 1], 4, 16, 1.0)


def test_raw_pointer_and_host_entry_points(ops):
    _, cabi = ops
    n, h, w, g, gc = 2, 20, 28, 4, 16
    x, off, m, go = make_inputs(n, h, w, g, gc, seed=21)
    kw = dict(groups=g, group_channels=gc)
    ref_out = c_oracle.forward(x, off, m, **kw)
    rx, roff, rm = c_oracle.backward(x, off, m, go, **kw)
    p = cabi.make_params(x.shape, (h, w), (3, 3), (1, 1), (1, 1), (1, 1), g, gc, 1.0, cabi.F32)
    # host buffers in, host buffers out (what a CPU-tensor caller binds)
    out, gx, goff, gm = np.empty_like(ref_out), np.empty_like(x), np.empty_like(off), np.empty_like(m)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    cabi.check(cabi.lib.dcnv3_forward_backward_host(vp(x), vp(off), vp(m), vp(go), vp(out), vp(gx), vp(goff),
                                                    vp(gm), ctypes.byref(p), 0))
    assert rel_err(out, ref_out) <= TOL_F32 and rel_err(gx, rx) <= TOL_F32
    assert rel_err(goff, roff) <= TOL_F32 and rel_err(gm, rm) <= TOL_F32
    out2 = np.empty_like(ref_out)
    cabi.check(cabi.lib.dcnv3_forward_host(vp(x), vp(off), vp(m), vp(out2), ctypes.byref(p), 0))
    assert np.array_equal(out, out2)
    # raw device pointers
    tx, to, tm = cuda(x, off, m)
    tout = torch.empty(n, h, w, g * gc, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    cabi.check(cabi.lib.dcnv3_forward(tx.data_ptr(), to.data_ptr(), tm.data_ptr(), tout.data_ptr(), ctypes.byref(p), st))
    assert np.array_equal(tout.cpu().numpy(), out)
    # too-small workspace is refused
    rc = cabi.lib.dcnv3_backward(tx.data_ptr(), to.data_ptr(), tm.data_ptr(), tout.data_ptr(), tx.data_ptr(),
                                 to.data_ptr(), tm.data_ptr(), tx.data_ptr(), 16, ctypes.byref(p), st)
    assert rc == cabi.ERR_WORKSPACE


def test_host_numpy_binding(ops):
    """iseg_b200/bindings/host_numpy.py -- the tested part of the reference-side binding of INTEGRATION.md (the
    TF lines around it are `tf.numpy_function` + `tf.custom_gradient`)."""
    from iseg_b200.bindings.host_numpy import dcnv3_op_numpy, dcnv3_op_with_grads_numpy
    n, h, w, g, gc = 2, 26, 35, 6, 16
    x, off, m, go = make_inputs(n, h, w, g, gc, sigma=1.5, seed=31)
    args = ([3, 3], [1, 1], "same", [1, 1], g, gc, 1.0)
    kw = dict(groups=g, group_channels=gc)
    keep = [a.copy() for a in (x, off, m, go)]
    out = dcnv3_op_numpy(x, off, m, *args)
    assert rel_err(out, c_oracle.forward(x, off, m, **kw)) <= TOL_F32
    out2, gx, goff, gm = dcnv3_op_with_grads_numpy(x, off, m, go, *args)
    rx, roff, rm = c_oracle.backward(x, off, m, go, **kw)
    assert np.array_equal(out, out2)
    assert rel_err(gx, rx) <= TOL_F32 and rel_err(goff, roff) <= TOL_F32 and rel_err(gm, rm) <= TOL_F32
    assert all(np.array_equal(a, b) for a, b in zip(keep, (x, off, m, go)))  # inputs are never written
    logits = np.random.default_rng(0).standard_normal(m.shape).astype(np.float32)
    out3 = dcnv3_op_numpy(x, off, logits, *args, mask_is_logits=True)
    assert rel_err(out3, c_oracle.forward(x, off, O.mask_softmax(logits, g), **kw)) <= TOL_F32
    with pytest.raises(TypeError):
        dcnv3_op_numpy(x, off, m, [3, 3], [1, 1], 1, [1, 1], g, gc, 1.0)        # op.py:29-30
    with pytest.raises(ValueError):
        dcnv3_op_numpy(x, off, m, [3, 3], [1, 1], "full", [1, 1], g, gc, 1.0)   # op.py:32-39
    with pytest.raises(ValueError):
        dcnv3_op_numpy(x, off[:, :5], m[:, :5], *args)                          # op.py:83 reshape


def test_dirty_workspace_is_detected_on_request(ops):
    """DCNV3_FLAG_WORKSPACE_ZEROED is a promise; DCNV3_FLAG_CHECK_WORKSPACE verifies it (debug aid)."""
    _, cabi = ops
    n, h, w, g, gc = 1, 20, 20, 4, 16
    x, off, m, go = make_inputs(n, h, w, g, gc, seed=4)
    tx, to, tm, tgo = cuda(x, off, m, go)
    gx, goff, gm = torch.empty_like(tx), torch.empty_like(to), torch.empty_like(tm)
    p = cabi.make_params(x.shape, (h, w), (3, 3), (1, 1), (1, 1), (1, 1), g, gc, 1.0, cabi.F32,
                         cabi.FLAG_WORKSPACE_ZEROED | cabi.FLAG_CHECK_WORKSPACE)
    nb = int(cabi.lib.dcnv3_backward_workspace_bytes(ctypes.byref(p)))
    ws = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    args = [t.data_ptr() for t in (tx, to, tm, tgo, gx, goff, gm, ws)]
    assert cabi.lib.dcnv3_backward(*args, nb, ctypes.byref(p), st) == 0
    torch.cuda.synchronize()
    zb = int(cabi.lib.dcnv3_backward_workspace_zero_bytes(ctypes.byref(p)))
    assert 0 < zb <= nb and int(ws[:zb].count_nonzero()) == 0
    rx, _, _ = c_oracle.backward(x, off, m, go, groups=g, group_channels=gc)
    assert rel_err(gx.cpu().numpy(), rx) <= TOL_F32
    # the scratch behind the zero part may hold anything (it does now) and the promise still holds
    assert cabi.lib.dcnv3_backward(*args, nb, ctypes.byref(p), st) == 0
    ws[zb // 2] = 1  # something scribbled on the zero part
    assert cabi.lib.dcnv3_backward(*args, nb, ctypes.byref(p), st) == cabi.ERR_WORKSPACE
    assert b"not all-zero" in cabi.lib.dcnv3_last_error()


def test_layer_matches_reference_layer(ops):
    iseg, _ = ops
    for name in ("c64_g4_cfs", "c32_g2_dw5"):
        z = np.load(f"{GOLDEN}/layer_{name}.npz")
        w = {k[2:]: z[k] for k in z.files if k.startswith("w.")}
        for fuse in (True, False):
            layer = iseg.DeformableConvolutionV3(
                filters=int(z["filters"]), kernel_size=int(z["kernel_size"]),
                depthwise_kernel_size=int(z["depthwise_kernel_size"]) or None, groups=int(z["groups"]),
                offset_scale=float(z["offset_scale"]), center_feature_scale=bool(z["center_feature_scale"]),
                input_channels=z["x"].shape[-1], fuse_softmax=fuse)
            layer.load_reference_weights(w)
            layer = layer.cuda()
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
            with torch.no_grad():
                y = layer(torch.from_numpy(z["x"]).cuda()).cpu().numpy()
            assert rel_err(y, z["y"]) <= 5e-5, (name, fuse)  # dense / conv / LN around the op are cuBLAS / cuDNN


FULL_CASES = [  # dtype, (n, h, w, g, gc), offset_scale -- BASELINE.json configs at their full sizes
    (torch.float32, (16, 128, 128, 4, 16), 1.0),   # config 2: InternImage-T at a 512 crop, batch 16, stages 1-4
    (torch.float32, (16, 64, 64, 8, 16), 1.0),
    (torch.float32, (16, 32, 32, 16, 16), 1.0),
    (torch.float32, (16, 16, 16, 32, 16), 1.0),
    (torch.bfloat16, (16, 128, 128, 4, 16), 1.0),
    (torch.bfloat16, (16, 64, 64, 8, 16), 1.0),
    (torch.bfloat16, (16, 32, 32, 16, 16), 1.0),
    (torch.bfloat16, (16, 16, 16, 32, 16), 1.0),
    (torch.bfloat16, (4, 160, 160, 10, 16), 2.0),  # config 4: InternImage-L at a 640 crop, stage 1 (4 images) ...
    (torch.bfloat16, (16, 20, 20, 80, 16), 2.0),   # ... and stage 4 (G = 80)
    (torch.bfloat16, (2, 193, 193, 4, 16), 1.0),   # config 5: 769x769 sliding-window tiles, stage 1
    (torch.bfloat16, (16, 40, 40, 40, 32), 2.0),   # config 4's gc = 32 variant: C1280 / G40 (InternImage-H stage 3 at 640)
    (torch.float32, (4, 80, 80, 20, 32), 1.0),     # InternImage-H stage 2 shape, fp32
]


@pytest.mark.parametrize("dtype,shape,scale", FULL_CASES)
def test_full_size_vs_c_oracle(ops, dtype, shape, scale):
    """Full-size configurations against the oracle itself (not only through properties): the C restatement
    does the 9.4 M sampled points of stage 1 at batch 16 in about a second."""
    n, h, w, g, gc = shape
    x, off, m, go = make_inputs(n, h, w, g, gc, sigma=1.0, seed=h + g)
    tol = TOL_F32
    if dtype == torch.bfloat16:
        rnd = lambda a: torch.from_numpy(a).bfloat16().float().numpy()  # noqa: E731
        x, off, m, go = rnd(x), rnd(off), rnd(m), rnd(go)
        tol = TOL_BF16
    kw = dict(groups=g, group_channels=gc, offset_scale=scale)
    out, gx, goff, gm = run_op(ops, x, off, m, go, dtype=dtype, **kw)
    assert rel_err(out, c_oracle.forward(x, off, m, **kw)) <= tol
    rx, roff, rm = c_oracle.backward(x, off, m, go, **kw)
    assert rel_err(gx, rx) <= tol and rel_err(goff, roff) <= tol and rel_err(gm, rm) <= tol


@pytest.mark.parametrize("flags", ["tiled", "generic"])
def test_grad_x_batch_invariant_and_wide_range(ops, flags):
    """The fixed-point scale of grad_x comes from each image's own max|grad_out|: an image's gradient is
    bit-identical whether it is alone or shares the batch with images whose grad_out is 1e6 times larger or
    smaller, and every image keeps fp32-like accuracy relative to its own magnitude."""
    _, cabi = ops
    n, h, w, g, gc = 4, 40, 48, 4, 16
    x, off, m, go = make_inputs(n, h, w, g, gc, sigma=1.5, seed=77)
    scales = np.array([1.0, 1e6, 1e-6, 3e3], np.float32).reshape(n, 1, 1, 1)
    go = go * scales
    cfg = ((3, 3), (1, 1), (1, 1), (1, 1), g, gc, 1.0)
    fl = cabi.FLAG_FORCE_GENERIC if flags == "generic" else 0
    t = [torch.from_numpy(a).cuda() for a in (x, off, m, go)]
    gx, goff, gm = (v.cpu().numpy() for v in cabi.backward(*t, *cfg, flags=fl))
    rx, roff, rm = c_oracle.backward(x, off, m, go, groups=g, group_channels=gc)
    for i in range(n):
        assert rel_err(gx[i], rx[i]) <= TOL_F32, i      # per image: relative to the image's own maximum
        alone = [torch.from_numpy(a[i:i + 1]).cuda() for a in (x, off, m, go)]
        gxi = cabi.backward(*alone, *cfg, flags=fl)[0].cpu().numpy()
        assert np.array_equal(gxi[0], gx[i]), i          # batch invariance, bit for bit
    assert rel_err(goff, roff) <= TOL_F32 and rel_err(gm, rm) <= TOL_F32


@pytest.mark.parametrize("flags", ["tiled", "generic"])
@pytest.mark.parametrize("bad", [float("nan"), float("inf")])
def test_non_finite_grad_out_is_not_swallowed(ops, flags, bad):
    """A NaN / Inf in grad_out cannot be carried by integer accumulation; instead of disappearing (fmaxf skips
    NaN, float -> int of NaN is 0) it turns that image's grad_x into NaN, so found-inf checks fire.  Other
    images of the batch are untouched, and the workspace is left clean."""
    _, cabi = ops
    n, h, w, g, gc = 3, 36, 36, 4, 16
    x, off, m, go = make_inputs(n, h, w, g, gc, seed=5)
    go[1, 7, 9, 3] = bad
    cfg = ((3, 3), (1, 1), (1, 1), (1, 1), g, gc, 1.0)
    fl = cabi.FLAG_FORCE_GENERIC if flags == "generic" else 0
    t = [torch.from_numpy(a).cuda() for a in (x, off, m, go)]
    gx, goff, gm = (v.cpu().numpy() for v in cabi.backward(*t, *cfg, flags=fl))
    assert np.isnan(gx[1]).all()
    assert not np.isfinite(goff[1, 7, 9]).all() or not np.isfinite(gm[1, 7, 9]).all()
    go[1, 7, 9, 3] = 0.0
    rx, _, _ = c_oracle.backward(x, off, m, go, groups=g, group_channels=gc)
    assert rel_err(gx[0], rx[0]) <= TOL_F32 and rel_err(gx[2], rx[2]) <= TOL_F32
    t[3] = torch.from_numpy(go).cuda()
    gx2 = cabi.backward(*t, *cfg, flags=fl)[0].cpu().numpy()  # same cached workspace: must have been left zeroed
    assert rel_err(gx2, rx) <= TOL_F32


def test_two_threads_two_streams(ops):
    """Re-entrancy: two host threads drive the library concurrently on their own streams (the reference's
    MirroredStrategy runs one host thread per replica); results are those of the sequential run."""
    import threading
    iseg, _ = ops
    cases = [(2, 48, 40, 4, 16, 11), (3, 33, 57, 6, 16, 12)]
    want, got, errs = {}, {}, []
    data = {}
    for i, (n, h, w, g, gc, seed) in enumerate(cases):
        data[i] = (make_inputs(n, h, w, g, gc, seed=seed), dict(groups=g, group_channels=gc))
        want[i] = run_op(ops, *data[i][0], **data[i][1])

    def worker(i):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                for _ in range(20):
                    got[i] = run_op(ops, *data[i][0], **data[i][1])
                torch.cuda.current_stream().synchronize()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=worker, args=(i,)) for i in data]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for i in data:
        assert all(np.array_equal(a, b) for a, b in zip(got[i], want[i])), i


@pytest.mark.parametrize("dtype,shape,scale", [
    (torch.float32, (16, 128, 128, 4, 16), 1.0),   # BASELINE config 2, stage 1, batch 16
    (torch.bfloat16, (16, 128, 128, 4, 16), 1.0),
    (torch.bfloat16, (4, 160, 160, 10, 16), 2.0),  # config 4: InternImage-L stage 1 at a 640 crop, offset_scale 2
    (torch.bfloat16, (2, 193, 193, 4, 16), 1.0),   # config 5: one 769x769 sliding-window tile, stage 1
])
def test_full_size_properties(ops, dtype, shape, scale):
    """BASELINE configs at full size: properties that need no oracle.  out is linear in x and in mask, so
    with L = <out, go>:
        <x, grad_x> = L          (adjoint identity ties forward and backward scatter)
        <mask, grad_mask> = L
    and the forward is linear in x."""
    iseg, _ = ops
    n, h, w, g, gc = shape
    gen = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=gen)  # noqa: E731
    x, x2, off, go = r(n, h, w, g * gc), r(n, h, w, g * gc), r(n, h, w, g * 18), r(n, h, w, g * gc)
    mask = torch.softmax(r(n, h, w, g, 9), -1).reshape(n, h, w, g * 9)
    x, x2, off, go, mask = (t.to(dtype) for t in (x, x2, off, go, mask))
    args = ([3, 3], [1, 1], "SAME", [1, 1], g, gc, scale)
    xr, orq, mr = x.clone().requires_grad_(), off.clone().requires_grad_(), mask.clone().requires_grad_()
    out = iseg.dcnv3_op(xr, orq, mr, *args)
    out.backward(go)
    L = (out.double() * go.double()).sum().item()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    scale = (out.double().abs() * go.double().abs()).sum().item()
    assert abs((x.double() * xr.grad.double()).sum().item() - L) <= tol * scale
    assert abs((mask.double() * mr.grad.double()).sum().item() - L) <= tol * scale
    out2 = iseg.dcnv3_op(x2, off, mask, *args)
    out12 = iseg.dcnv3_op((x.float() * 0.5 + x2.float() * 2).to(dtype), off, mask, *args)
    lin = out.detach().float() * 0.5 + out2.float() * 2
    assert (out12.float() - lin).abs().max().item() <= (1e-5 if dtype == torch.float32 else 6e-2) * lin.abs().max().item()
    # determinism at full size
    xr2, or2, mr2 = x.clone().requires_grad_(), off.clone().requires_grad_(), mask.clone().requires_grad_()
    iseg.dcnv3_op(xr2, or2, mr2, *args).backward(go)
    assert torch.equal(xr.grad, xr2.grad) and torch.equal(orq.grad, or2.grad) and torch.equal(mr.grad, mr2.grad)


# ---- centre-feature-scale blend fused into the kernels (reference dcn_v3.py:138-146) -------------------------
def _blend_reference(x, off, m, cs, go, g, gc, offset_scale=1.0, logits=False):
    """fp32 numpy restatement of `x_core * (1 - cfs) + x_proj * cfs` around the oracle's op, and its gradients:
    the core sees grad_out * (1 - s); x also receives grad_out * s directly; d s = sum_c grad_out * (x - core)."""
    n, h, w, c = x.shape
    s = np.repeat(cs, gc, axis=-1).astype(np.float32)                       # [N,H,W,G] -> [N,H,W,C]
    mm = m
    if logits:
        z = m.reshape(n, h, w, g, 9)
        e = np.exp(z - z.max(-1, keepdims=True))
        mm = (e / e.sum(-1, keepdims=True)).reshape(n, h, w, g * 9).astype(np.float32)
    core = c_oracle.forward(x, off, mm, groups=g, group_channels=gc, offset_scale=offset_scale)
    out = core * (np.float32(1) - s) + x * s
    gcore = (go * (np.float32(1) - s)).astype(np.float32)
    gx, goff, gm = c_oracle.backward(x, off, mm, gcore, groups=g, group_channels=gc, offset_scale=offset_scale)
    gx = gx + go * s
    if logits:  # soft-max Jacobian (dcn_v3.py:120-123)
        gmr, mr = gm.reshape(n, h, w, g, 9), mm.reshape(n, h, w, g, 9)
        gm = (mr * (gmr - (mr * gmr).sum(-1, keepdims=True))).reshape(n, h, w, g * 9)
    gs = (go * (x - core)).reshape(n, h, w, g, gc).sum(-1)
    return out, gx, goff, gm, gs


@pytest.mark.parametrize("shape, dtype, logits, scale", [
    ((2, 40, 40, 4, 16), torch.float32, False, 1.0),     # several scatter tiles (ring hand-over, merge kernel)
    ((3, 17, 23, 3, 16), torch.float32, True, 1.0),      # single tile, odd group count (phantom group), fused soft-max
    ((2, 64, 48, 8, 16), torch.bfloat16, False, 1.0),
    ((1, 48, 40, 10, 16), torch.float32, False, 2.0),    # offset_scale 2: narrow ring, the per-tap scatter walk
    ((2, 40, 36, 3, 32), torch.float32, True, 1.0),      # 32 channels per group (InternImage-H): half groups, fused soft-max
    ((2, 24, 40, 5, 32), torch.bfloat16, False, 1.0),    # ... bf16, odd group count (phantom half groups)
])
def test_center_scale_blend_fused(ops, shape, dtype, logits, scale):
    iseg, cabi = ops
    n, h, w, g, gc = shape
    x, off, m, go = make_inputs(n, h, w, g, gc, sigma=1.5, seed=21)
    rng = np.random.default_rng(5)
    cs = rng.uniform(-0.5, 1.5, size=(n, h, w, g)).astype(np.float32)       # no sigmoid in the reference: any real
    if logits:
        m = rng.standard_normal(m.shape).astype(np.float32)
    if dtype == torch.bfloat16:  # the yardstick sees the bf16-rounded inputs
        x, off, m, go, cs = (torch.from_numpy(a).to(torch.bfloat16).float().numpy() for a in (x, off, m, go, cs))
    tx, to, tm, ts = cuda(x, off, m, cs, dtype=dtype)
    for t in (tx, to, tm, ts):
        t.requires_grad_(True)
    assert cabi.blend_supported(tx, to, (3, 3), (1, 1), (1, 1), (1, 1), g, gc, scale)
    before = cabi.launch_count()
    out = iseg.dcnv3_op_center_scale(tx, to, tm, ts, [3, 3], [1, 1], "SAME", [1, 1], g, gc, scale, mask_is_logits=logits)
    assert cabi.launch_count() - before == 1  # the blend costs no launch of its own
    out.backward(cuda(go, dtype=dtype)[0])
    got = [t.float().cpu().numpy() for t in (out.detach(), tx.grad, to.grad, tm.grad, ts.grad)]
    ref = _blend_reference(x, off, m, cs, go, g, gc, offset_scale=scale, logits=logits)
    tol = TOL_F32 if dtype == torch.float32 else TOL_BF16
    for name, a, b in zip(("out", "grad_x", "grad_offset", "grad_mask", "grad_center_scale"), got, ref):
        assert rel_err(a, b) <= tol, name
    # bitwise reproducible, blend included
    tx.grad = None
    out2 = iseg.dcnv3_op_center_scale(tx, to, tm, ts, [3, 3], [1, 1], "SAME", [1, 1], g, gc, scale, mask_is_logits=logits)
    out2.backward(cuda(go, dtype=dtype)[0])
    assert torch.equal(out2, out) and np.array_equal(tx.grad.float().cpu().numpy(), got[1])


def test_center_scale_blend_unfused_configurations(ops):
    """Where the tiled kernels do not run (here 8 channels per group) there is no fused blend: it is then applied
    around the op with torch operations, same values; and the C ABI says so instead of computing something else."""
    iseg, cabi = ops
    n, h, w, g, gc = 1, 12, 12, 2, 8
    x, off, m, go = make_inputs(n, h, w, g, gc, seed=8)
    cs = np.random.default_rng(2).uniform(0, 1, size=(n, h, w, g)).astype(np.float32)
    tx, to, tm, ts = cuda(x, off, m, cs)
    assert not cabi.blend_supported(tx, to, (3, 3), (1, 1), (1, 1), (1, 1), g, gc, 1.0)
    out = iseg.dcnv3_op_center_scale(tx, to, tm, ts, [3, 3], [1, 1], "SAME", [1, 1], g, gc, 1.0)
    ref = _blend_reference(x, off, m, cs, go, g, gc)[0]
    assert rel_err(out.cpu().numpy(), ref) <= TOL_F32
    with pytest.raises(cabi.DCNv3Error):
        cabi.forward_blend(tx, to, tm, ts, (3, 3), (1, 1), (1, 1), (1, 1), g, gc, 1.0)
