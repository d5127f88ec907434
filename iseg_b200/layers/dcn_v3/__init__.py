from .op import dcnv3_op, dcnv3_op_center_scale  # noqa: F401
from .dcn_v3 import DeformableConvolutionV3  # noqa: F401
