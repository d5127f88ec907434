from .intern_image import (  # noqa: F401
    InternImage, InternImageBlock, InternImageLayer, intern_image_base, intern_image_huge, intern_image_large,
    intern_image_small, intern_image_tiny,
)
from .graphed import GraphedInference  # noqa: F401,E402
from .weights import export_reference_weights, load_reference_weights, reference_names  # noqa: F401,E402
