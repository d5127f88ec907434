"""Batch sharding over several GPUs: every rank runs the DCNv3 kernels on its own images; outputs and
gradients gathered over NCCL must be BIT-identical to the single-GPU result.  Images never interact, and the
fixed-point scale of grad_x is per image, so this holds whatever else shares a rank's shard -- the images here
carry grad_out magnitudes spread over 1e-4 .. 1e4 on purpose.  Skipped unless at least two GPUs are visible.
Recorded runs on 2 and 8 GPUs: profiles/r02_multi_gpu.md."""
import os

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, results):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import iseg_b200
    from iseg_b200.distribution import BatchShardStrategy
    st = BatchShardStrategy()
    ok = True
    for dtype, (n, h, w, g, gc) in ((torch.float32, (11, 40, 36, 4, 16)), (torch.bfloat16, (9, 33, 70, 8, 16))):
        gen = torch.Generator().manual_seed(0)
        x = torch.randn(n, h, w, g * gc, generator=gen)
        off = torch.randn(n, h, w, g * 18, generator=gen) * 2
        mask = torch.softmax(torch.randn(n, h, w, g, 9, generator=gen), -1).reshape(n, h, w, g * 9)
        go = torch.randn(n, h, w, g * gc, generator=gen)
        go = go * (10.0 ** torch.linspace(-4, 4, n)).reshape(n, 1, 1, 1)  # per-image magnitudes 1e-4 .. 1e4
        x, off, mask, go = (t.to(dtype) for t in (x, off, mask, go))
        args = ([3, 3], [1, 1], "SAME", [1, 1], g, gc, 1.0)

        def fwd_bwd(x_, off_, mask_, go_):
            x_, off_, mask_ = (t.to(st.device).requires_grad_() for t in (x_, off_, mask_))
            out = iseg_b200.dcnv3_op(x_, off_, mask_, *args)
            out.backward(go_.to(st.device))
            return out.detach(), x_.grad, off_.grad, mask_.grad

        local = fwd_bwd(*st.shard(x, off, mask, go))
        gathered = [st.gather(t, total=n) for t in local]
        if rank == 0:
            single = fwd_bwd(x, off, mask, go)
            ok = ok and all(torch.equal(a, b) for a, b in zip(gathered, single))
    if rank == 0:
        results["ok"] = ok
    st.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_shard_vs_single_bit_exact(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, 29621 + world, results), nprocs=world, join=True)
        assert results["ok"]
