"""InternImage wiring (torch re-statement of reference backbones/intern_image/) around the CUDA op."""
import pytest
import torch

from iseg_b200.backbones.intern_image import intern_image_base, intern_image_small, intern_image_tiny
from iseg_b200.backbones.intern_image.intern_image import _conv_same_s2


def test_presets_match_reference_hyperparameters():
    m = intern_image_tiny()
    assert [len(b.blocks) for b in m.blocks] == [4, 4, 18, 4]          # intern_image.py:139-143
    assert [b.blocks[0].dcn.groups for b in m.blocks] == [4, 8, 16, 32]
    assert [b.blocks[0].dcn.filters for b in m.blocks] == [64, 128, 256, 512]
    assert all(b.blocks[0].dcn.filters_per_group == 16 for b in m.blocks)
    assert m.blocks[3].downsample is None and m.blocks[0].downsample is not None
    s = intern_image_small()
    assert s.blocks[0].blocks[0].use_post_norm and s.blocks[0].norm is None  # intern_image_block.py:81-84
    rates = [l.drop_path_rate for b in m.blocks for l in b.blocks]
    assert rates[0] == 0.0 and abs(rates[-1] - 0.2) < 1e-9 and rates == sorted(rates)


def test_weights_load_by_reference_name_roundtrip(tmp_path):
    """checkpoints of the reference are restored by layer name (saver/h5_saver.py:38-49)."""
    import numpy as np
    from iseg_b200.backbones.intern_image.weights import export_reference_weights, load_reference_weights, reference_names
    torch.manual_seed(0)
    src = intern_image_tiny()
    for p in src.parameters():
        torch.nn.init.normal_(p, std=0.1)
    w = export_reference_weights(src)
    # names and Keras layouts as the reference declares them
    assert w["block/2/layer/17/dcn/offset/kernel"].shape == (256, 2 * 16 * 9)          # Dense [in, out]
    assert w["block/0/layer/0/dcn/dw_conv/depthwise_kernel"].shape == (3, 3, 64, 1)   # DepthwiseConv2D
    assert w["patch_embed/conv1/kernel"].shape == (3, 3, 3, 32) and "block/3/downsample/conv/kernel" not in w
    assert w["block/1/downsample/conv/kernel"].shape == (3, 3, 128, 256) and "block/1/downsample/conv/bias" not in w
    assert "block/0/layer/3/gamma1" in w and "block/0/norm/gamma" in w
    assert len(w) == len(list(src.parameters())) == len(reference_names(src))
    # Keras-3 style keys ("." separators, ":0" suffix, model-name prefix) through an .npz on disk
    path = str(tmp_path / "ckpt.npz")
    np.savez(path, **{"intern_image_tiny." + k.replace("/", ".") + ":0": v for k, v in w.items()})
    dst = intern_image_tiny()
    loaded, missing, unexpected = load_reference_weights(dst, path)
    assert not missing and not unexpected and len(loaded) == len(w)
    for a, b in zip(src.parameters(), dst.parameters()):
        assert torch.equal(a, b)
    w.pop("block/0/norm/beta")
    with pytest.raises(KeyError):
        load_reference_weights(intern_image_tiny(), w)
    assert load_reference_weights(intern_image_tiny(), w, strict=False)[1] == ["block/0/norm/beta"]
    huge_like = export_reference_weights(
        __import__("iseg_b200.backbones.intern_image.intern_image", fromlist=["InternImage"]).InternImage(
            32, [1, 1, 2, 1], [2, 4, 8, 16], layer_scale=None, use_res_post_norm=True, use_level2_post_norm=True,
            level2_post_norm_block_ids=[1], use_center_feature_scale=True, depthwise_kernel_size=5))
    assert "block/2/post_norms/0/gamma" in huge_like and "block/0/layer/0/res_post_norm1/beta" in huge_like
    assert huge_like["block/0/layer/0/dcn/center_feature_scale_proj/kernel"].shape == (32, 2)


def test_tf_same_padding_stride2_shapes():
    conv = torch.nn.Conv2d(3, 4, 3, stride=2)
    for h, w in ((512, 512), (769, 769), (15, 18)):
        y = _conv_same_s2(conv, torch.zeros(1, h, w, 3))
        assert y.shape[1:3] == (-(-h // 2), -(-w // 2))  # 769 -> 385 -> 193


@pytest.mark.gpu
def test_backbone_forward_backward_on_gpu():
    torch.manual_seed(0)
    m = intern_image_tiny(return_endpoints=True).cuda().eval()  # eval: no drop-path randomness
    for blk in m.blocks:  # make the deformable gather data dependent (the reference zero-initialises these)
        for layer in blk.blocks:
            torch.nn.init.normal_(layer.dcn.offset.weight, std=0.05)
            torch.nn.init.normal_(layer.dcn.mask.weight, std=0.05)
    x = torch.randn(2, 128, 160, 3, device="cuda", requires_grad=True)
    ends = m(x)
    assert [tuple(e.shape[1:]) for e in ends] == [(64, 80, 32), (32, 40, 64), (16, 20, 128), (8, 10, 256), (4, 5, 512)]
    ends[-1].square().mean().backward()
    assert torch.isfinite(x.grad).all() and x.grad.abs().sum() > 0
    g1 = m.blocks[0].blocks[0].dcn.offset.weight.grad.clone()
    m.zero_grad(); x.grad = None
    m(x)[-1].square().mean().backward()
    # (the DCNv3 op is bitwise reproducible -- tests/test_gpu_parity.py; cuDNN / cuBLAS around it need not be)
    g2 = m.blocks[0].blocks[0].dcn.offset.weight.grad
    assert torch.allclose(g1, g2, rtol=1e-3, atol=1e-6 * g1.abs().max().item())


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_huge_variant_32_channels_per_group(dtype):
    """InternImage-H wiring in small (intern_image.py:261: 32 channels per group, 5x5 depthwise branch, res-post-norm,
    level-2 post-norms, centre-feature-scale): forward + backward; its DCNv3 layers run the tiled kernels on half
    groups (the C ABI's launch plan says so), centre-feature-scale blend included."""
    from iseg_b200 import _cabi
    from iseg_b200.backbones.intern_image.intern_image import InternImage
    torch.manual_seed(3)
    m = InternImage(64, [1, 1, 3, 1], [2, 4, 8, 16], 4.0, layer_scale=None, offset_scale=1.0, use_post_norm=False,
                    depthwise_kernel_size=5, use_res_post_norm=True, use_level2_post_norm=True,
                    level2_post_norm_block_ids=[1], use_center_feature_scale=True).cuda().to(dtype).eval()
    for blk in m.blocks:
        for layer in blk.blocks:
            assert layer.dcn.filters_per_group == 32
            torch.nn.init.normal_(layer.dcn.offset.weight, std=0.05)
            torch.nn.init.normal_(layer.dcn.mask.weight, std=0.05)
    p = _cabi.make_params((2, 24, 32, 64), (24, 32), (3, 3), (1, 1), (1, 1), (1, 1), 2, 32, 1.0,
                          _cabi.F32 if dtype == torch.float32 else _cabi.BF16)
    assert _cabi.launch_plan(p)["tiled"]
    x = torch.randn(2, 96, 128, 3, device="cuda", dtype=dtype, requires_grad=True)
    y = m(x)
    assert y.shape == (2, 3, 4, 512) and torch.isfinite(y.float()).all()
    y.float().square().mean().backward()
    assert torch.isfinite(x.grad.float()).all() and x.grad.abs().sum() > 0
    assert all(torch.isfinite(q.grad.float()).all() for q in m.parameters() if q.grad is not None)


@pytest.mark.gpu
def test_base_variant_odd_groups_bf16():
    m = intern_image_base().cuda().to(torch.bfloat16).eval()
    with torch.no_grad():
        y = m(torch.randn(1, 64, 64, 3, device="cuda", dtype=torch.bfloat16))
    assert y.shape == (1, 2, 2, 896) and torch.isfinite(y.float()).all()


@pytest.mark.gpu
def test_sliding_window_with_intern_image_tiles_matches_reference_rule():
    """BASELINE config 5 in small: sliding-window inference with InternImage-T tiles of odd size (193 -> stage
    shapes 49 / 25 / 13 / 7, like the 769 -> 193 / 97 / 49 / 25 of the real configuration).  The tiled driver must
    give what the reference's sequential rule gives (core_inference.py:230-304: every tile's logits zero-padded to
    the full image, summed in tile order, divided by the count map), with the window starts of the reference's own
    function (tests/golden/sliding_indices.json)."""
    import json
    import os

    from iseg_b200.distribution import inference_with_sliding_window, sliding_window_tiles
    torch.manual_seed(1)
    m = intern_image_tiny(return_endpoints=True).cuda().eval()
    for blk in m.blocks:
        for layer in blk.blocks:
            torch.nn.init.normal_(layer.dcn.offset.weight, std=0.05)
            torch.nn.init.normal_(layer.dcn.mask.weight, std=0.05)
    head = torch.nn.Linear(64, 5).cuda()

    def model_fn(tile):  # [N, h, w, 3] -> [N, h, w, 5]: stage-1 endpoint (stride 4), linear head, nearest upsampling
        with torch.no_grad():
            f = head(m(tile)[1])
        up = f.repeat_interleave(4, dim=1).repeat_interleave(4, dim=2)
        return up[:, :tile.shape[1], :tile.shape[2]]

    height, width, crop = 300, 420, 193
    image = torch.randn(1, height, width, 3, device="cuda")
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sliding_indices.json")))
    tiles = sliding_window_tiles(height, width, crop, crop)
    ys, xs = sorted({t[0] for t in tiles}), sorted({t[1] for t in tiles})
    assert ys == cases.get(f"{height},{crop}", ys) and len(tiles) == len(ys) * len(xs) and len(tiles) >= 4
    out = inference_with_sliding_window(model_fn, image, crop_h=crop, crop_w=crop)
    # the reference's rule, literally: pad to full size, add in order, divide by the count map
    acc = torch.zeros(1, height, width, 5, device="cuda")
    cnt = torch.zeros(1, height, width, 1, device="cuda")
    for (y, x, h, w) in tiles:
        lg = model_fn(image[:, y:y + h, x:x + w])
        acc += torch.nn.functional.pad(lg, (0, 0, x, width - x - w, y, height - y - h))
        cnt += torch.nn.functional.pad(torch.ones(1, h, w, 1, device="cuda"), (0, 0, x, width - x - w, y, height - y - h))
    ref = acc / cnt
    assert torch.isfinite(out).all() and (out - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()

    # every window has the same shape: the tile network as ONE CUDA graph, replayed per tile -- same numbers
    from iseg_b200.backbones.intern_image import GraphedInference

    class TileNet(torch.nn.Module):
        def forward(self, tile):
            return model_fn(tile)

    graphed = GraphedInference(TileNet().eval())
    with torch.no_grad():
        out_g = inference_with_sliding_window(graphed, image, crop_h=crop, crop_w=crop)
    assert len(graphed._graphs) == 1 and torch.equal(out_g, out)


@pytest.mark.gpu
def test_graphed_inference_matches_eager():
    """Product-level CUDA graph of a whole backbone forward (and of one stage): same numbers as eager, one
    graph launch per call, re-captured per input shape."""
    from iseg_b200 import _cabi
    from iseg_b200.backbones.intern_image import GraphedInference
    torch.manual_seed(2)
    m = intern_image_tiny(return_endpoints=True).cuda().eval()
    for blk in m.blocks:
        for layer in blk.blocks:
            torch.nn.init.normal_(layer.dcn.offset.weight, std=0.05)
            torch.nn.init.normal_(layer.dcn.mask.weight, std=0.05)
    gm = GraphedInference(m)
    for shape in ((2, 96, 128, 3), (1, 193, 193, 3)):
        x = torch.randn(*shape, device="cuda")
        with torch.no_grad():
            want = m(x)
            got = gm(x)
            n0 = _cabi.launch_count()
            got2 = gm(torch.randn(*shape, device="cuda"))
            assert _cabi.launch_count() == n0  # replay: no launches through the C ABI, the graph holds them
        assert all(torch.equal(a, b) for a, b in zip(want, got))
        assert not torch.equal(got2[-1], got[-1])
    stage = GraphedInference(m.blocks[2])  # one stage: 18 layers of 32x32-sized work at a 512 crop
    f = torch.randn(2, 24, 32, 256, device="cuda")
    with torch.no_grad():
        a, b = m.blocks[2](f), stage(f)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


@pytest.mark.gpu
@pytest.mark.parametrize("c", [72, 200, 448, 896, 1280])   # 1 / 2 / 4 / 8 pieces per lane in registers, row buffer beyond
@pytest.mark.parametrize("dtype, tol", [(torch.float32, 2e-5), (torch.bfloat16, 3e-2)])
def test_fused_layer_kernels_match_torch(dtype, tol, c):
    """dcnv3_dwconv_ln_act and dcnv3_layer_join against the torch operations they replace
    (reference layers/dcn_v3/dcn_v3.py:115-117, backbones/intern_image/intern_image_layer.py:126-172)."""
    import torch.nn.functional as F
    from iseg_b200 import _cabi
    torch.manual_seed(3)
    torch.backends.cudnn.allow_tf32 = False
    n, h, w = 2, 13, 17
    rnd = lambda *s: torch.randn(*s, device="cuda").to(dtype)  # noqa: E731
    x, r, gamma, lw, lb = rnd(n, h, w, c), rnd(n, h, w, c), rnd(c), rnd(c), rnd(c)
    eps = 1e-6
    with torch.no_grad():
        for k, lo in ((3, 1), (5, 2), (4, 1)):   # (even kernel: Keras pads (k-1)//2 before, k//2 after)
            wt, bias = rnd(c, 1, k, k) * 0.3, rnd(c)
            want = F.conv2d(F.pad(x.permute(0, 3, 1, 2), (lo, k - 1 - lo, lo, k - 1 - lo)), wt, bias, groups=c).permute(0, 2, 3, 1)
            want = F.gelu(F.layer_norm(want, (c,), lw, lb, eps))
            got = _cabi.dwconv_ln_act(x, wt.permute(2, 3, 0, 1).reshape(k * k, c).contiguous(), bias, lw, lb, k, lo, eps)
            assert _rel(got, want) <= tol, k
        z = r + x * gamma
        s0, n0 = _cabi.layer_join(x, r, gamma, lw, lb, eps, 0)
        assert _rel(s0, z) <= tol and _rel(n0, F.layer_norm(z, (c,), lw, lb, eps)) <= tol
        assert _rel(_cabi.layer_join(x, r, None, None, None, eps, 0, want_norm=False)[0], r + x) <= tol
        assert _rel(_cabi.layer_join(x, r, gamma, lw, lb, eps, 1), r + F.layer_norm(x, (c,), lw, lb, eps) * gamma) <= tol
        assert _rel(_cabi.layer_join(x, None, None, lw, lb, eps, 2), F.layer_norm(x, (c,), lw, lb, eps)) <= tol
    with pytest.raises(ValueError):
        _cabi.layer_join(rnd(4, 6), rnd(4, 6), None, rnd(6), rnd(6), eps, 1)   # channels % 4 != 0


@pytest.mark.gpu
def test_fused_inference_path_matches_unfused_layers():
    """under torch.no_grad() in eval mode the layers run the one-pass kernels (joins and the
    DCNv3 layer's depthwise-conv branch); with autograd recording they run the torch operations.  Same numbers for
    the three layer variants of the reference, with and without layer scale, and for a whole stage."""
    from iseg_b200 import _cabi
    from iseg_b200.backbones.intern_image.intern_image import InternImageBlock, InternImageLayer
    torch.manual_seed(4)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    def randomise(mod):
        for name, p in mod.named_parameters():
            torch.nn.init.normal_(p, std=0.05 if ("offset" in name or "mask" in name) else 0.3)
        return mod.cuda().eval()

    x = torch.randn(2, 24, 20, 64, device="cuda")
    variants = [dict(layer_scale=1.0), dict(), dict(use_post_norm=True, layer_scale=1.0),
                dict(use_res_post_norm=True, center_feature_scale=True, depthwise_kernel_size=5)]
    for kw in variants:
        layer = randomise(InternImageLayer(64, 4, **kw))
        want = layer(x)                       # autograd recording on: torch operations
        n0 = _cabi.launch_count()
        with torch.no_grad():
            got = layer(x)
        assert _cabi.launch_count() - n0 >= 4  # op + depthwise branch + two joins at least
        assert _rel(got, want) <= 5e-5, kw
    for kw in (dict(layer_scale=1.0), dict(use_post_norm=True, layer_scale=1.0)):
        stage = randomise(InternImageBlock(64, 3, 4, use_downsample=True, **kw))
        want = stage(x)
        with torch.no_grad():
            got = stage(x)
        assert _rel(got[0], want[0]) <= 1e-4 and _rel(got[1], want[1]) <= 1e-4, kw
    m = randomise(intern_image_tiny(return_endpoints=True))
    img = torch.randn(1, 96, 128, 3, device="cuda")
    want = m(img)
    with torch.no_grad():
        got = m(img)
    assert all(_rel(a, b) <= 5e-4 for a, b in zip(got, want))
    mb = randomise(intern_image_tiny()).to(torch.bfloat16)
    imgb = img.to(torch.bfloat16)
    wantb = mb(imgb)
    with torch.no_grad():
        gotb = mb(imgb)
    assert torch.isfinite(gotb.float()).all() and _rel(gotb, wantb) <= 0.1
