from .op import dcnv3_op  # noqa: F401
from .dcn_v3 import DeformableConvolutionV3  # noqa: F401
