"""Batch sharding of whole images over the GPUs of one box -- the role the reference gives to
`tf.distribute.MirroredStrategy` (reference distribution/distribution_utils.py:74-88, selected by
`get_distribution_strategy` :98-121) re-done as one process per GPU on torch.distributed / NCCL.

The DCNv3 op never mixes images (batch index is a pure gather coordinate, reference
layers/dcn_v3/utils.py:191-198), so the op itself needs no collective: every rank runs the kernels on
its own images.  NCCL is used only where a full tensor is wanted on every rank -- the counterpart of
`strategy.experimental_local_results` + `tf.concat` in reference core_predict.py:136-153 -- and for the
scalar reductions of `all_reduce_values` (reference distribution_utils.py:158-169).
"""
import os

import torch
import torch.distributed as dist


def shard_range(n, world_size, rank):
    """[start, stop) of the images rank `rank` owns: n // world each, remainder to the low ranks."""
    base, rem = divmod(int(n), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class BatchShardStrategy:
    """One process per GPU; mirrors the small surface of a tf.distribute strategy that iSeg touches:
    `num_replicas_in_sync`, `scope()`, `run(fn, args)` on the local shard, plus `gather`."""

    def __init__(self, backend=None, device=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if device is None:
            device = torch.device("cuda", self.local_rank) if torch.cuda.is_available() else torch.device("cpu")
        self.device = device
        if device.type == "cuda":
            torch.cuda.set_device(device)
        if self.world_size > 1 and not dist.is_initialized():
            backend = backend or ("nccl" if device.type == "cuda" else "gloo")
            kw = {"device_id": device} if backend == "nccl" else {}
            dist.init_process_group(backend, **kw)

    @property
    def num_replicas_in_sync(self):
        return self.world_size

    def scope(self):
        import contextlib
        return contextlib.nullcontext()

    def shard(self, *tensors):
        """Slices this rank's images (dim 0) out of full-batch tensors."""
        out = []
        for t in tensors:
            a, b = shard_range(t.shape[0], self.world_size, self.rank)
            out.append(t[a:b])
        return out[0] if len(out) == 1 else tuple(out)

    def run(self, fn, *full_batch_tensors):
        """fn on the local shard (no collective)."""
        local = self.shard(*full_batch_tensors)
        return fn(*local) if isinstance(local, tuple) else fn(local)

    def gather(self, local_out, total):
        return all_gather_outputs(local_out, total, self.world_size, self.rank)

    def close(self):
        if dist.is_initialized():
            dist.destroy_process_group()


def get_distribution_strategy(gpu_memory_growth=True, cuda_visible_devices=None, use_tpu=False, tpu_name=None,
                              use_one_device_strategy=False):
    """Same arguments as the reference (distribution_utils.py:98-104); TPU is not a target here."""
    if use_tpu:
        raise NotImplementedError("TPU strategies are outside the B200 hot path")
    if cuda_visible_devices is not None:
        os.environ["CUDA_VISIBLE_DEVICES"] = str(cuda_visible_devices)
    if use_one_device_strategy:
        os.environ.setdefault("WORLD_SIZE", "1")
    return BatchShardStrategy()


def all_gather_outputs(local_out, total, world_size=None, rank=None):
    """All ranks get the [total, ...] tensor made of every rank's shard in rank order.  Shards may
    differ by one image (shard_range); they are padded to the largest for `all_gather_into_tensor`."""
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if world_size == 1:
        return local_out
    counts = [b - a for a, b in (shard_range(total, world_size, r) for r in range(world_size))]
    cmax = max(counts)
    pad = local_out
    if local_out.shape[0] < cmax:
        pad = torch.cat([local_out, local_out.new_zeros((cmax - local_out.shape[0],) + tuple(local_out.shape[1:]))])
    buf = local_out.new_empty((world_size * cmax,) + tuple(local_out.shape[1:]))
    dist.all_gather_into_tensor(buf, pad.contiguous())
    if all(c == cmax for c in counts):
        return buf
    return torch.cat([buf[r * cmax:r * cmax + c] for r, c in enumerate(counts)])


def all_reduce_values(values, reduce_op="sum"):
    """reference distribution_utils.py:158-169."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return values
    op = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[reduce_op]
    single = torch.is_tensor(values)
    out = []
    for v in ([values] if single else values):
        v = v.clone()
        dist.all_reduce(v, op=op)
        out.append(v)
    return out[0] if single else out
