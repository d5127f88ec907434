#!/usr/bin/env python
"""Secondary measurements for BASELINE.json configs 3-5 (bench.py is configs[1], the headline):

  cfg3  InternImage-B backbone forward at 512x512, batch 32, bf16, with the CUDA DCNv3 op swapped in
        (eager and replayed from a CUDA graph; the dense / conv / norm layers around the op are stock torch)
  cfg4  InternImage-L DCNv3 core op fwd+bwd at a 640x640 crop (stages 160^2xC160/G10 x5, 80^2xC320/G20 x5,
        40^2xC640/G40 x22, 20^2xC1280/G80 x5, offset_scale 2), bf16, batch 16 per GPU, whole images sharded
  cfg5  sliding-window inference, InternImage-T on a 1024x2048 image with 769x769 windows (8 tiles), tiles
        dealt over the ranks and combined by one all-reduce

    python tools/config_bench.py [cfg3] [cfg4] [cfg5]          (torchrun for more than one GPU)

Prints one JSON line per config on rank 0.  Synthetic data, random-init weights (offset / mask projections
are given small random weights so that the gather is data dependent; the reference zero-initialises them).
"""
import ctypes
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iseg_b200 import _cabi as cabi  # noqa: E402
from iseg_b200.backbones.intern_image import GraphedInference, intern_image_base, intern_image_tiny  # noqa: E402
from iseg_b200.distribution import BatchShardStrategy, inference_with_sliding_window  # noqa: E402

P, GC = 9, 16


def timed(fn, steps, warmup, strategy):
    for _ in range(warmup):
        fn()
    if strategy.world_size > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    if strategy.world_size > 1:
        t = torch.tensor([ms], device=strategy.device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def randomise_offsets(model):
    for blk in model.blocks:
        for layer in blk.blocks:
            torch.nn.init.normal_(layer.dcn.offset.weight, std=0.05)
            torch.nn.init.normal_(layer.dcn.mask.weight, std=0.05)


def cfg3(strategy):
    torch.manual_seed(0)
    batch = 32
    model = intern_image_base().to(strategy.device).to(torch.bfloat16).eval()
    randomise_offsets(model)
    x = torch.randn(batch, 512, 512, 3, device=strategy.device, dtype=torch.bfloat16)
    with torch.no_grad():
        n0 = cabi.launch_count()
        y = model(x)
        ours = cabi.launch_count() - n0
        eager = timed(lambda: model(x), 10, 3, strategy)
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            model(x)
            torch.cuda.synchronize()
            with torch.cuda.graph(graph, stream=side):
                yg = model(x)
        graphed = timed(graph.replay, 10, 3, strategy)
        same = torch.equal(y, yg)
    return {"config": "cfg3 InternImage-B backbone forward, 512x512, bf16, DCNv3 op = libdcnv3_b200",
            "batch_per_gpu": batch, "n_gpus": strategy.world_size, "eager_ms": eager, "cuda_graph_ms": graphed,
            "images_per_s": batch * strategy.world_size / (graphed * 1e-3), "graph_output_equals_eager": same,
            "dcnv3_kernels_per_forward": ours, "output_shape": list(y.shape)}


def cfg4(strategy):
    stages = [(160, 160, 160, 10, 5), (80, 80, 320, 20, 5), (40, 40, 640, 40, 22), (20, 20, 1280, 80, 5)]
    batch, dev = 16, strategy.device
    gen = torch.Generator(device=dev).manual_seed(strategy.rank)
    r = lambda *s: torch.randn(*s, device=dev, generator=gen)  # noqa: E731
    layers = []
    for h, w, c, g, depth in stages:  # one tensor set per stage keeps the -L working set small
        x, off, go = r(batch, h, w, c).bfloat16(), r(batch, h, w, g * 18).bfloat16(), r(batch, h, w, c).bfloat16()
        m = torch.softmax(r(batch, h, w, g, P), -1).reshape(batch, h, w, g * P).bfloat16()
        outs = [torch.empty_like(t) for t in (x, x, off, m)]
        p = cabi.make_params(x.shape, (h, w), (3, 3), (1, 1), (1, 1), (1, 1), g, GC, 2.0, cabi.BF16,
                             cabi.FLAG_WORKSPACE_ZEROED)
        wsb = int(cabi.lib.dcnv3_backward_workspace_bytes(ctypes.byref(p)))
        ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
        vp = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
        layers.append((depth, p, wsb, [vp(t) for t in (x, off, m)], vp(outs[0]), vp(go), [vp(t) for t in outs[1:]],
                       vp(ws), (x, off, m, go, outs, ws)))
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step():
        for depth, p, wsb, ins, out, go, grads, ws, _ in layers:
            for _ in range(depth):
                cabi.check(cabi.lib.dcnv3_forward(*ins, out, ctypes.byref(p), st))
        for depth, p, wsb, ins, out, go, grads, ws, _ in reversed(layers):
            for _ in range(depth):
                cabi.check(cabi.lib.dcnv3_backward(*ins, go, *grads, ws, wsb, ctypes.byref(p), st))

    ms = timed(step, 10, 3, strategy)
    points = sum(batch * h * w * g * P * d for h, w, _, g, d in stages)
    nbytes = sum(batch * h * w * (5 * c + 9 * g * P) * 2 * d for h, w, c, g, d in stages)
    return {"config": "cfg4 InternImage-L DCNv3 core op fwd+bwd, 37 layers at a 640x640 crop, bf16, offset_scale 2",
            "batch_per_gpu": batch, "n_gpus": strategy.world_size, "ms_per_step": ms,
            "points_per_s": points * strategy.world_size / (ms * 1e-3),
            "algo_hbm_gbs_per_gpu": nbytes / (ms * 1e-3) * 1e-9}


def cfg5(strategy):
    torch.manual_seed(0)
    model = intern_image_tiny().to(strategy.device).to(torch.bfloat16).eval()
    randomise_offsets(model)
    head = torch.nn.Conv2d(512, 19, 1).to(strategy.device).to(torch.bfloat16)  # Cityscapes-sized logit head

    class TileNet(torch.nn.Module):  # backbone features -> per-pixel logits at tile resolution
        def __init__(self):
            super().__init__()
            self.model, self.head = model, head

        def forward(self, tile):
            f = self.model(tile)
            logits = self.head(f.permute(0, 3, 1, 2))
            return F.interpolate(logits.float(), size=tile.shape[1:3], mode="bilinear", align_corners=False).permute(0, 2, 3, 1)

    net = TileNet().eval()
    graphed = GraphedInference(net)  # every 769x769 window has the same shape: one CUDA graph, replayed per tile
    img = torch.randn(1, 1024, 2048, 3, device=strategy.device, dtype=torch.bfloat16)
    with torch.no_grad():
        out = inference_with_sliding_window(net, img, 769, 769, strategy)
        ms = timed(lambda: inference_with_sliding_window(net, img, 769, 769, strategy), 10, 3, strategy)
        out_g = inference_with_sliding_window(graphed, img, 769, 769, strategy)
        ms_g = timed(lambda: inference_with_sliding_window(graphed, img, 769, 769, strategy), 10, 3, strategy)
    return {"config": "cfg5 sliding-window inference, InternImage-T, 1024x2048, 769x769 windows (8 tiles), bf16",
            "n_gpus": strategy.world_size, "eager_ms_per_image": ms, "ms_per_image": ms_g, "images_per_s": 1e3 / ms_g,
            "launch": "one CUDA graph of the tile network (GraphedInference), replayed per tile",
            "graph_output_equals_eager": bool(torch.equal(out, out_g)),
            "output_shape": list(out.shape), "finite": bool(torch.isfinite(out_g).all().item())}


def siblings(strategy):
    """The two sibling gather ops at a representative shape each: forward + backward through the torch
    ops, CUDA events, algorithmic bytes (every tensor read or written once) / time against the measured HBM peak.  They
    run generic thread-per-point kernels and are not tuned; this is the measurement, not a target."""
    from iseg_b200.layers.dcn_v2 import dcnv2_sample
    from iseg_b200.layers.deformable_attention import deform_attn_sample
    dev = strategy.device
    gen = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s: torch.randn(*s, device=dev, generator=gen)  # noqa: E731
    out = []
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError, ValueError):
        peak = None
    # deformable attention: batch 16, 64x64, 8 heads x 32 channels, 4 points
    n, h, w, heads, pts, c = 16, 64, 64, 8, 4, 32
    value = r(n, h, w, heads, c).requires_grad_()
    y = (torch.rand(n, h, w, heads, pts, device=dev, generator=gen) * (h - 1)).requires_grad_()
    x = (torch.rand(n, h, w, heads, pts, device=dev, generator=gen) * (w - 1)).requires_grad_()
    attn = torch.softmax(r(n, h, w, heads, pts), -1).requires_grad_()
    go = r(n, h, w, heads, c)

    def da():
        for t in (value, y, x, attn):
            t.grad = None
        deform_attn_sample(value, y, x, attn).backward(go)

    ms = timed(da, 20, 3, strategy)
    nbytes = 4 * (n * h * w * heads * (4 * c + 6 * pts))  # value, out, grad_out, grad_value + 3 point tensors and their gradients
    out.append({"config": "sibling: deform_attn_sample fwd+bwd, batch 16, 64x64, 8 heads x 32 ch, 4 points, fp32", "ms": ms,
                "algo_gbs": nbytes / (ms * 1e-3) * 1e-9, "frac_of_hbm_peak": None if peak is None else nbytes / (ms * 1e-3) * 1e-9 / peak})
    # DCNv2 sampler: batch 16, 64x64, 128 channels, 3x3
    n, h, w, c, k = 16, 64, 64, 128, 3
    xx = r(n, h, w, c).requires_grad_()
    offs = (r(n, h, w, k * k, 2)).requires_grad_()
    mask = torch.sigmoid(r(n, h, w, k * k)).requires_grad_()
    go2 = r(n, h, w, k * k, c)

    def d2():
        for t in (xx, offs, mask):
            t.grad = None
        dcnv2_sample(xx, offs, mask, k).backward(go2)

    ms = timed(d2, 20, 3, strategy)
    nbytes = 4 * (n * h * w * (2 * c + 2 * k * k * c + 6 * k * k))  # x, grad_x, map_all, its gradient, offsets / mask and gradients
    out.append({"config": "sibling: dcnv2_sample fwd+bwd, batch 16, 64x64, 128 channels, 3x3, fp32", "ms": ms,
                "algo_gbs": nbytes / (ms * 1e-3) * 1e-9, "frac_of_hbm_peak": None if peak is None else nbytes / (ms * 1e-3) * 1e-9 / peak})
    return out


def main():
    which = [a for a in sys.argv[1:] if a.startswith("cfg") or a == "siblings"] or ["cfg3", "cfg4", "cfg5"]
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    strategy = BatchShardStrategy()
    t0 = time.time()
    for name in which:
        res = {"cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5, "siblings": siblings}[name](strategy)
        torch.cuda.empty_cache()
        if strategy.rank == 0:
            for line in (res if isinstance(res, list) else [res]):
                print(json.dumps(line))
    if strategy.rank == 0:
        print(f"# {time.time() - t0:.1f} s", file=sys.stderr)
    if strategy.world_size > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
