from .dcn_v2 import DCNv2, dcnv2_sample  # noqa: F401
