from .distribution_utils import (  # noqa: F401
    BatchShardStrategy, all_gather_outputs, all_reduce_values, get_distribution_strategy, shard_range,
)
from .sliding_window import (  # noqa: F401
    get_sliding_start_indexs, inference_with_sliding_window, shard_tiles, sliding_window_tiles, stitch,
)
