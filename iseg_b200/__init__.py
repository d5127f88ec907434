"""iseg_b200 -- B200-native DCNv3 core operator behind iSeg's own `dcnv3_op` / `DeformableConvolutionV3`
interfaces (reference: edwardyehuang/iSeg layers/dcn_v3).  Hand-written sm_100a CUDA reached through
a C ABI (include/dcnv3_b200.h); there is no CPU fallback: the first use of the op loads
lib/libdcnv3_b200.so and raises ImportError if it has not been built (`python iseg_b200/build.py`)."""

__all__ = ["dcnv3_op", "dcnv3_op_center_scale", "DeformableConvolutionV3"]


def __getattr__(name):  # resolved on first use so that `iseg_b200.build` can run before the library exists
    if name == "dcnv3_op":
        from .layers.dcn_v3.op import dcnv3_op
        return dcnv3_op
    if name == "dcnv3_op_center_scale":
        from .layers.dcn_v3.op import dcnv3_op_center_scale
        return dcnv3_op_center_scale
    if name == "DeformableConvolutionV3":
        from .layers.dcn_v3.dcn_v3 import DeformableConvolutionV3
        return DeformableConvolutionV3
    if name == "_cabi":
        import importlib
        return importlib.import_module("._cabi", __name__)
    raise AttributeError(name)
