"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/dcnv3_b200.h declares, validates parameters like the reference's reshapes would, and the host
wrapper keeps the reference's error behaviour.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cabi():
    from iseg_b200 import build
    build.build()
    from iseg_b200 import _cabi
    return _cabi


def declared_functions():
    src = open(os.path.join(ROOT, "include", "dcnv3_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dcnv3_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(cabi):
    names = declared_functions()
    assert len(names) >= 12
    for name in names:
        assert hasattr(cabi.lib, name), f"{name} declared in include/dcnv3_b200.h but not exported"


def test_version_and_build_info(cabi):
    assert cabi.lib.dcnv3_abi_version() == 1
    assert b"sm_100a" in cabi.lib.dcnv3_build_info()


def test_param_validation_mirrors_reference_reshape(cabi):
    mk = lambda **kw: cabi.make_params(  # noqa: E731
        kw.get("x", (2, 8, 9, 64)), kw.get("out", (8, 9)), kw.get("k", (3, 3)), kw.get("s", (1, 1)),
        kw.get("pad", (1, 1)), kw.get("d", (1, 1)), kw.get("g", 4), kw.get("gc", 16), 1.0,
        kw.get("dtype", 0))
    chk = lambda p: cabi.lib.dcnv3_check_params(ctypes.byref(p))  # noqa: E731
    assert chk(mk()) == 0
    assert chk(mk(out=(7, 9))) == cabi.ERR_SHAPE          # offset grid != reference-point grid
    assert b"op.py:83" in cabi.lib.dcnv3_last_error()
    assert chk(mk(pad=(0, 0), out=(6, 7))) == 0            # VALID
    assert chk(mk(s=(2, 2), out=(4, 5))) == 0              # (10-3)//2+1, (11-3)//2+1
    assert chk(mk(d=(2, 2), out=(6, 7))) == 0              # dilation 2 with SAME pad 1
    assert chk(mk(k=(9, 9), pad=(4, 4))) == cabi.ERR_ARGUMENT
    assert chk(mk(dtype=7)) == cabi.ERR_DTYPE
    assert chk(mk(x=(2, 1, 1, 64), pad=(0, 0), out=(1, 1))) == cabi.ERR_SHAPE  # kernel > input
    assert cabi.lib.dcnv3_backward_workspace_bytes(ctypes.byref(mk())) > 0
    assert cabi.lib.dcnv3_backward_workspace_bytes(ctypes.byref(mk(out=(1, 1)))) == 0


def test_compute_entry_points_fail_loudly_without_cuda(cabi):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = cabi.make_params((1, 4, 4, 16), (4, 4), (3, 3), (1, 1), (1, 1), (1, 1), 1, 16, 1.0, 0)
    buf = (ctypes.c_float * 4096)()
    rc = cabi.lib.dcnv3_forward_host(buf, buf, buf, buf, ctypes.byref(p), 0)
    assert rc == cabi.ERR_CUDA and cabi.lib.dcnv3_last_error()


def test_op_error_conventions(cabi):
    from iseg_b200 import dcnv3_op
    x = torch.zeros(1, 4, 4, 16)
    off = torch.zeros(1, 4, 4, 18)
    m = torch.zeros(1, 4, 4, 9)
    with pytest.raises(TypeError):  # reference op.py:29-30
        dcnv3_op(x, off, m, [3, 3], [1, 1], 1, [1, 1], 1, 16, 1.0)
    with pytest.raises(ValueError):  # reference op.py:38-39
        dcnv3_op(x, off, m, [3, 3], [1, 1], "full", [1, 1], 1, 16, 1.0)
    with pytest.raises(cabi.DCNv3Error):  # CPU tensors: no fallback
        dcnv3_op(x, off, m, [3, 3], [1, 1], "same", [1, 1], 1, 16, 1.0)


def test_layer_signature_matches_reference():
    import inspect
    from iseg_b200 import DeformableConvolutionV3
    sig = inspect.signature(DeformableConvolutionV3.__init__)
    ref_args = ["filters", "kernel_size", "depthwise_kernel_size", "strides", "padding",
                "dilation_rate", "groups", "offset_scale", "activation", "center_feature_scale", "name"]
    assert list(sig.parameters)[1:1 + len(ref_args)] == ref_args  # dcn_v3.py:18-31
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["filters"], d["kernel_size"], d["strides"], d["padding"], d["groups"]) == (64, 3, 1, "SAME", 4)
    with pytest.raises(AssertionError):
        DeformableConvolutionV3(filters=10, groups=4)  # dcn_v3.py:34
    layer = DeformableConvolutionV3(filters=32, groups=2, center_feature_scale=True, input_channels=32)
    names = {n.split(".")[0] for n, _ in layer.named_parameters()}
    assert names == {"dw_conv", "dw_conv_norm", "offset", "mask", "input_proj", "output_proj",
                     "center_feature_scale_proj"}  # dcn_v3.py:62-102
    assert not layer.offset.weight.any() and not layer.mask.weight.any()  # zero init, :74-86


def test_launch_plans_fit_the_hardware_for_any_shape(cabi):
    """dcnv3_launch_plan (no GPU): whatever the image shape, group count, dtype and offset_scale, a tiled
    plan must fit an SM's shared memory (227 KB per CTA, two forward / gather CTAs per SM) and keep its
    geometry consistent; strongly non-square images used to fail at launch time (the reference pairs output
    rows with input columns, so a 16-row tile spans 16*W/H columns)."""
    import numpy as np
    rng = np.random.default_rng(0)
    shapes = [(128, 128), (64, 64), (32, 32), (16, 16), (160, 160), (80, 80), (40, 40), (20, 20), (193, 193), (97, 97),
              (49, 49), (25, 25), (256, 512), (512, 256), (20, 300), (300, 20), (257, 33), (33, 257), (3, 3), (3, 1000),
              (1000, 3), (1024, 2048), (7, 64)]
    shapes += [(int(rng.integers(3, 700)), int(rng.integers(3, 700))) for _ in range(300)]
    n_tiled = 0
    for h, w in shapes:
        for scale in (0.25, 0.5, 1.0, 2.0, 4.0):
            for dtype, g in ((cabi.F32, 4), (cabi.BF16, 5), (cabi.F32, 1), (cabi.BF16, 8)):
                p = cabi.make_params((2, h, w, g * 16), (h, w), (3, 3), (1, 1), (1, 1), (1, 1), g, 16, scale, dtype)
                plan = cabi.launch_plan(p)
                if not plan["tiled"]:
                    continue
                n_tiled += 1
                f, ga, sc = plan["forward"], plan["gather"], plan["scatter"]
                for k in (f, ga):
                    assert 1 <= k["th"] <= 16 and 1 <= k["tw"] <= 128, (h, w, scale, k)
                    assert 2 <= k["bw"] <= min(w + 2, 256) and 2 <= k["bh"] <= min(h + 2, 256), (h, w, scale, k)
                    assert k["halo_x"] >= 1 and k["halo_y"] >= 1 and k["ctas"] >= 1
                # forward: box + (when the group count allows TMA / cp.async staging) eight per-warp side slots, two
                # CTAs per SM; the gather kernel always has its result slots
                slots = 8 * (3456 if dtype == cabi.F32 else 1792) if g % (2 if dtype == cabi.F32 else 4) == 0 else 0
                assert f["bw"] * f["bh"] * 128 <= (82 if slots else 100) * 1024, (h, w, scale, f)
                assert f["smem"] == f["bw"] * f["bh"] * 128 + slots, (h, w, scale, f)
                assert 2 * (f["smem"] + 1024) <= 227 * 1024, (h, w, scale, f)
                assert ga["bw"] * ga["bh"] * 128 <= 82 * 1024 and 2 * (ga["smem"] + 1024) <= 227 * 1024, (h, w, scale, ga)
                pitch = sc["tj"] + 9
                assert sc["tj"] in (16, 32) and 1 <= sc["ring_lo"] <= 4 and 2 <= sc["ring_hi"] <= 5
                assert 1 <= sc["box_rows"] <= pitch and sc["smem"] <= 227 * 1024 and sc["ctas"] >= 1
                assert sc["merge"] == (h > sc["tj"] or w > sc["tj"])
    assert n_tiled > 4000
    # the shapes InternImage runs (square-ish, offset_scale 1 or 2): always tiled, full 16 x 16 tiles, and
    # at offset_scale 1 a forward reach of 3 offset units (halo 4)
    for h, w, g, scale in ((128, 128, 4, 1.0), (32, 32, 16, 1.0), (193, 193, 4, 1.0), (160, 160, 10, 2.0),
                           (256, 512, 4, 1.0), (25, 49, 32, 1.0)):
        p = cabi.make_params((16, h, w, g * 16), (h, w), (3, 3), (1, 1), (1, 1), (1, 1), g, 16, scale, cabi.F32)
        plan = cabi.launch_plan(p)
        assert plan["tiled"], (h, w, plan)
        if scale == 1.0:
            # 16 x 16 tiles; 8 x 16 where that would leave fewer than four waves of CTAs on 148 SMs x 2
            few = 16 * ((g + 1) // 2) * -(-h // 16) * -(-w // 16) < 4 * 296
            assert plan["forward"]["th"] == (8 if few else 16) and plan["forward"]["tw"] == 16, (h, w, plan)
            assert plan["gather"]["th"] == plan["forward"]["th"]
        if scale == 1.0:
            assert plan["scatter"]["ring_lo"] == 4 and plan["scatter"]["ring_hi"] == 5
            assert plan["forward"]["halo_x"] == (4 if h == w else 3)  # (a 2:1 image gives up a little reach to fit 82 KB)
    # 32 channels per group (InternImage-H): the tiled kernels run every group as two 16-channel half groups, so the
    # plan is that of twice the groups at 16 channels -- except that the side inputs are not staged (72-byte runs)
    for dtype in (cabi.F32, cabi.BF16):
        p32 = cabi.make_params((16, 40, 40, 40 * 32), (40, 40), (3, 3), (1, 1), (1, 1), (1, 1), 40, 32, 2.0, dtype)
        p16 = cabi.make_params((16, 40, 40, 80 * 16), (40, 40), (3, 3), (1, 1), (1, 1), (1, 1), 80, 16, 2.0, dtype)
        a, b = cabi.launch_plan(p32), cabi.launch_plan(p16)
        assert a["tiled"] and b["tiled"] and a["scatter"] == b["scatter"]
        assert a["gather"]["ctas"] == b["gather"]["ctas"] and a["forward"]["ctas"] <= b["forward"]["ctas"]  # (100 KB box)
        assert a["forward"]["smem"] == a["forward"]["bw"] * a["forward"]["bh"] * 128   # no side slots
        ws = lambda p: int(cabi.lib.dcnv3_backward_workspace_bytes(ctypes.byref(p)))  # noqa: E731
        assert ws(p32) >= ws(p16)   # (>=: the generic kernels' workspace is the floor in both)
    # other kernel sizes / strides / channel counts are served by the generic kernels
    p = cabi.make_params((2, 32, 32, 64), (32, 32), (5, 5), (1, 1), (2, 2), (1, 1), 4, 16, 1.0, cabi.F32)
    assert not cabi.launch_plan(p)["tiled"]
    p = cabi.make_params((2, 32, 32, 32), (32, 32), (3, 3), (1, 1), (1, 1), (1, 1), 4, 8, 1.0, cabi.F32)
    assert not cabi.launch_plan(p)["tiled"]
