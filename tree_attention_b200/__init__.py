"""tree_attention_b200 -- a Blackwell-native (sm_100a) Tree Attention framework.

Public API (names and signatures of ``/root/reference/model.py`` plus ``tree_attention``)::

    setup(rank, world_size)            cleanup()
    make_data(shape, rank, device)     flash_res_lse(q, k, v, softmax_scale=1.0, is_causal=False)
    tree_decode(q, k, v, rank, world_size, device)
    tree_attention(q, k, v, *, group=None, causal=False, softmax_scale=None, ...)
"""
from .ops.local import attention_partial, flash_res_lse  # noqa: F401
from .parallel.runtime import cleanup, get_runtime, setup  # noqa: F401
from .parallel.tree import (allreduce_sum, combine_partials, tree_attention, tree_decode, zigzag_shard,  # noqa: F401
                            zigzag_unshard)
from .utils.data import make_data  # noqa: F401

__version__ = "0.1.0"
__all__ = [
    "setup", "cleanup", "get_runtime", "make_data", "flash_res_lse", "attention_partial",
    "tree_decode", "tree_attention", "combine_partials", "allreduce_sum", "zigzag_shard", "zigzag_unshard",
]
