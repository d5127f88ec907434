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
def test_base_variant_odd_groups_bf16():
    m = intern_image_base().cuda().to(torch.bfloat16).eval()
    with torch.no_grad():
        y = m(torch.randn(1, 64, 64, 3, device="cuda", dtype=torch.bfloat16))
    assert y.shape == (1, 2, 2, 896) and torch.isfinite(y.float()).all()
