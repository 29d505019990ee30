from .local import attention_partial, decode_attention, flash_res_lse  # noqa: F401
from .reference import (  # noqa: F401
    attention_bwd_ref,
    attention_partial_ref,
    attention_ref,
    merge_many,
    merge_pair,
    merge_tree,
    sharded_attention_ref,
)
