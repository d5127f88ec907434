"""iseg_b200 -- B200-native DCNv3 core operator behind iSeg's own `dcnv3_op` / `DeformableConvolutionV3`
interfaces (reference: edwardyehuang/iSeg layers/dcn_v3).  Hand-written sm_100a CUDA reached through
a C ABI (include/dcnv3_b200.h); no CPU fallback."""
from . import _cabi  # noqa: F401  (raises if libdcnv3_b200.so has not been built)
from .layers.dcn_v3.op import dcnv3_op  # noqa: F401
from .layers.dcn_v3.dcn_v3 import DeformableConvolutionV3  # noqa: F401

__all__ = ["dcnv3_op", "DeformableConvolutionV3"]
