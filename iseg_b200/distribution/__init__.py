from .distribution_utils import (  # noqa: F401
    BatchShardStrategy, all_gather_outputs, all_reduce_values, get_distribution_strategy, shard_range,
)
from .sliding_window import get_sliding_start_indexs, sliding_window_tiles, shard_tiles  # noqa: F401
