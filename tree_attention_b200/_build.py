"""In-tree build + loader for the native extension ``tree_attention_b200._C``.

* ``build()`` compiles every ``csrc/*.cu`` for ``sm_100a`` (``-gencode arch=compute_100a,code=sm_100a
  -lineinfo``) plus the torch bindings with ninja via ``torch.utils.cpp_extension`` and leaves
  ``csrc/build/_C.so`` in the tree, so the binary travels to the GPU box with the repo snapshot.
* ``load()`` imports that ``.so`` directly (no ninja, no JIT cache) when its recorded source hash
  matches; otherwise it rebuilds if ``nvcc`` is present and fails loudly if not.

The reference (``/root/reference/model.py``) has no native code to build.
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import shutil
import sys
from pathlib import Path
from types import ModuleType
from typing import Optional

_PKG = Path(__file__).resolve().parent
CSRC = _PKG / "csrc"
BUILD_DIR = CSRC / "build"
SO_PATH = BUILD_DIR / "_C.so"
HASH_PATH = BUILD_DIR / "_C.hash"

CUDA_SOURCES = ["decode_simt.cu", "decode_tc_sm100.cu", "decode_swap_sm100.cu", "combine.cu", "umma_probe.cu", "attn_fwd_sm100.cu", "attn_bwd_sm100.cu",
                "quant.cu", "reduce.cu"]
CPP_SOURCES = ["bindings.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "--use_fast_math", "-std=c++17",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]
CXX_FLAGS = ["-O2", "-std=c++17"]

_module: Optional[ModuleType] = None


def sources() -> list[Path]:
    out = [CSRC / s for s in CUDA_SOURCES + CPP_SOURCES if (CSRC / s).exists()]
    return out


def source_hash() -> str:
    h = hashlib.sha256()
    files = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.cpp")))
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS + CXX_FLAGS).encode())
    return h.hexdigest()


def build(verbose: bool = True, force: bool = False) -> ModuleType:
    """Compile the extension in-tree (cross-compiles for sm_100a; no GPU needed)."""
    global _module
    from torch.utils import cpp_extension

    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    want = source_hash()
    if not force and SO_PATH.exists() and HASH_PATH.exists() and HASH_PATH.read_text().strip() == want:
        return load()
    os.environ.setdefault("MAX_JOBS", str(min(8, os.cpu_count() or 4)))
    mod = cpp_extension.load(
        name="_C",
        sources=[str(s) for s in sources()],
        extra_cflags=CXX_FLAGS,
        extra_cuda_cflags=NVCC_FLAGS,
        extra_include_paths=[str(CSRC)],
        build_directory=str(BUILD_DIR),
        verbose=verbose,
        with_cuda=True,
    )
    HASH_PATH.write_text(want)
    _module = mod
    return mod


def _import_so() -> ModuleType:
    import torch  # noqa: F401  (libtorch symbols must be loaded first)

    spec = importlib.util.spec_from_file_location("tree_attention_b200._C", str(SO_PATH))
    if spec is None or spec.loader is None:
        raise ImportError(f"cannot load {SO_PATH}")
    # the extension was compiled with TORCH_EXTENSION_NAME=_C, i.e. exports PyInit__C
    spec = importlib.util.spec_from_file_location("_C", str(SO_PATH))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules.setdefault("tree_attention_b200._C", mod)
    return mod


def load(required: bool = True) -> Optional[ModuleType]:
    """Return the native module, building it if sources changed and nvcc is available."""
    global _module
    if _module is not None:
        return _module
    have_so = SO_PATH.exists()
    fresh = have_so and HASH_PATH.exists() and HASH_PATH.read_text().strip() == source_hash()
    if have_so and (fresh or shutil.which("nvcc") is None or os.environ.get("TREE_ATTN_NO_REBUILD")):
        if not fresh:
            import warnings

            warnings.warn(
                f"{SO_PATH} was built from different sources than the ones in {CSRC} and cannot be rebuilt here "
                "(nvcc missing or TREE_ATTN_NO_REBUILD set): loading the STALE binary", RuntimeWarning, stacklevel=2)
        _module = _import_so()
        return _module
    if shutil.which("nvcc") is not None:
        return build(verbose=bool(os.environ.get("TREE_ATTN_VERBOSE_BUILD")))
    if required:
        raise ImportError(
            "tree_attention_b200 native extension is not built and nvcc is unavailable; "
            "run `python -c 'import __graft_entry__ as g; g.build()'` on a machine with CUDA 12.9"
        )
    return None


def available() -> bool:
    try:
        return load(required=False) is not None
    except Exception:
        return False
