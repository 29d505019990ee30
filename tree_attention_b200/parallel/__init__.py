from .runtime import cleanup, get_runtime, setup  # noqa: F401
from .tree import combine_partials, tree_attention, tree_decode  # noqa: F401
