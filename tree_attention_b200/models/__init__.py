from .decoder import TreeDecodeSession  # noqa: F401
from .tree_attention import TreeAttention, TreeSelfAttention  # noqa: F401
