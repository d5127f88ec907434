"""Sliding-window tiling, sharded over ranks.

Window placement follows the reference exactly: stride int(2/3 * crop), one extra window flush with the
far edge when the strided windows do not reach it (reference utils/sliding_window_inference_utils.py:
16-32); the reference then runs the tiles SEQUENTIALLY on one device, pads each tile's logits back to
full size, sums them and divides by a count map (core_inference.py:265-301).  Tiles are independent
until that final sum, so here they are dealt to the ranks and combined by one reduction.
"""
import torch


def get_sliding_start_indexs(length, crop_length):
    stride = int(2.0 / 3.0 * crop_length)
    times = (length - crop_length) // stride + 1
    idx = [stride * i for i in range(times)]
    if length - (times - 1) * stride > crop_length:
        idx.append(length - crop_length)
    return idx


def sliding_window_tiles(height, width, crop_h, crop_w):
    """[(y0, x0, h, w)] in the reference's row-major order; the window is clipped to the image."""
    ch, cw = min(crop_h, height), min(crop_w, width)
    return [(y, x, ch, cw) for y in get_sliding_start_indexs(height, ch) for x in get_sliding_start_indexs(width, cw)]


def shard_tiles(tiles, world_size, rank):
    """Round-robin deal: rank r gets tiles r, r+world, ... (1024x2048 with 769^2 windows -> 8 tiles, one
    per GPU of an 8-GPU box)."""
    return tiles[rank::world_size]


def stitch(tile_logits, tiles, height, width):
    """Sum of zero-padded tile logits divided by the count map (core_inference.py:276,299-301).
    tile_logits[i]: [N, h, w, C] for tiles[i].  Call on every rank with its own tiles, then all-reduce
    `acc` and `cnt` (or pass the gathered lists on one rank)."""
    n, _, _, c = tile_logits[0].shape
    acc = tile_logits[0].new_zeros((n, height, width, c))
    cnt = tile_logits[0].new_zeros((1, height, width, 1))
    for t, (y, x, h, w) in zip(tile_logits, tiles):
        acc[:, y:y + h, x:x + w] += t
        cnt[:, y:y + h, x:x + w] += 1
    return acc, cnt


def inference_with_sliding_window(model_fn, image, crop_h=769, crop_w=769, strategy=None):
    """Sliding-window inference with the tiles dealt over the ranks of `strategy`
    (reference core_inference.py:230-304 runs them sequentially on one device).

    image: [N, H, W, C] on this rank's device (every rank holds the full image, as every replica does in
    the reference); model_fn maps a [N, h, w, C] tile to [N, h, w, K] logits.  Every rank returns the full
    [N, H, W, K] result: per-rank partial sums and the count map are combined by ONE all-reduce.

    All windows of an image have the same shape, so a model_fn wrapped in
    iseg_b200.backbones.intern_image.GraphedInference runs as one CUDA graph replayed per tile (batch-1 tiles are
    launch bound: 66 -> 20 ms per 1024x2048 image with InternImage-T on one B200, tools/config_bench.py cfg5)."""
    import torch.distributed as dist

    n, height, width, _ = image.shape
    tiles = sliding_window_tiles(height, width, crop_h, crop_w)
    world = strategy.world_size if strategy is not None else 1
    rank = strategy.rank if strategy is not None else 0
    mine = shard_tiles(tiles, world, rank)
    logits = [model_fn(image[:, y:y + h, x:x + w]) for (y, x, h, w) in mine]
    if logits:
        acc, cnt = stitch(logits, mine, height, width)
    else:  # more ranks than tiles: contribute zeros of the right shape (K from a 1x1 probe is avoided)
        k = model_fn(image[:, :tiles[0][2], :tiles[0][3]]).shape[-1]
        acc = image.new_zeros((n, height, width, k))
        cnt = image.new_zeros((1, height, width, 1))
    if world > 1:
        dist.all_reduce(acc)
        dist.all_reduce(cnt)
    return acc / cnt
