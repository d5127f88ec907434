"""Host-side logic of the batch-sharding path, world_size 2 on the gloo backend (no GPU)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from iseg_b200.distribution import (get_sliding_start_indexs, shard_range, shard_tiles, sliding_window_tiles)


def test_shard_range_covers_batch():
    for n in (0, 1, 7, 16, 33):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_sliding_window_indices_match_reference_rule():
    # reference utils/sliding_window_inference_utils.py:16-32, 1024x2048 with 769 windows
    assert get_sliding_start_indexs(1024, 769) == [0, 255]
    assert get_sliding_start_indexs(2048, 769) == [0, 512, 1024, 1279]
    assert get_sliding_start_indexs(512, 512) == [0]
    tiles = sliding_window_tiles(1024, 2048, 769, 769)
    assert len(tiles) == 8
    assert sorted(sum((shard_tiles(tiles, 8, r) for r in range(8)), [])) == sorted(tiles)
    assert [len(shard_tiles(tiles, 4, r)) for r in range(4)] == [2, 2, 2, 2]


def test_sliding_window_indices_match_reference_function():
    """tests/golden/sliding_indices.json: the reference's own `_get_sliding_start_indexs_py` over 67 (length, crop)
    pairs (tests/golden/make_golden.py::make_sliding_indices)."""
    import json
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sliding_indices.json")))
    assert len(cases) > 60
    for key, want in cases.items():
        length, crop = (int(v) for v in key.split(","))
        assert get_sliding_start_indexs(length, crop) == want, key


def test_sliding_window_inference_single_rank_matches_sequential_rule():
    from iseg_b200.distribution import inference_with_sliding_window
    torch.manual_seed(0)
    img = torch.randn(1, 40, 70, 3)
    wgt = torch.randn(3, 5)
    model = lambda t: t @ wgt  # noqa: E731  (pointwise: every window must agree where they overlap)
    out = inference_with_sliding_window(model, img, crop_h=24, crop_w=24)
    assert torch.allclose(out, img @ wgt, atol=1e-5)


def _worker(rank, world, port, results):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from iseg_b200.distribution import BatchShardStrategy, all_reduce_values
    st = BatchShardStrategy(backend="gloo", device=torch.device("cpu"))
    torch.manual_seed(0)
    full = torch.randn(5, 3, 4, 8)  # 5 images over 2 ranks: 3 + 2
    local = st.run(lambda t: t * 2 + 1, full)
    a, b = shard_range(5, world, rank)
    assert local.shape[0] == b - a
    gathered = st.gather(local, total=5)
    ok = torch.equal(gathered, full * 2 + 1)  # bit-exact: images never interact
    s = all_reduce_values(torch.tensor([float(local.shape[0])]))
    ok = ok and s.item() == 5.0
    # sliding-window tiles dealt over the two ranks, one all-reduce at the end
    from iseg_b200.distribution import inference_with_sliding_window
    img = torch.randn(1, 40, 70, 3)
    wgt = torch.randn(3, 4)
    out = inference_with_sliding_window(lambda t: t @ wgt, img, crop_h=24, crop_w=24, strategy=st)
    ok = ok and torch.allclose(out, img @ wgt, atol=1e-5)
    results[rank] = bool(ok)
    st.close()


def test_two_rank_shard_run_gather_gloo():
    world = 2
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, 29611, results), nprocs=world, join=True)
        assert all(results[r] for r in range(world))
