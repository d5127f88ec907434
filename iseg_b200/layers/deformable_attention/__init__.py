from .sample import deform_attn_sample  # noqa: F401
from .layer import DeformableMultiHeadSelfAttentionLayer  # noqa: F401
