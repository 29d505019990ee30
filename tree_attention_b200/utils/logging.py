"""Logging facade.

The reference logs through ``loguru`` (``/root/reference/model.py:5``) without declaring the
dependency and logs inside the hot op (D14).  We use loguru when it is
importable and fall back to the standard library otherwise; nothing in ``ops/`` or the CUDA
path ever logs inside a timed region.
"""
from __future__ import annotations

import logging as _logging
import os
import sys

try:  # pragma: no cover - exercised implicitly
    from loguru import logger as _loguru_logger

    _HAVE_LOGURU = True
except Exception:  # pragma: no cover
    _loguru_logger = None
    _HAVE_LOGURU = False


class _StdLogger:
    def __init__(self) -> None:
        self._l = _logging.getLogger("tree_attention_b200")
        if not self._l.handlers:
            h = _logging.StreamHandler(sys.stderr)
            h.setFormatter(_logging.Formatter("%(asctime)s | %(levelname)-7s | %(message)s"))
            self._l.addHandler(h)
        self._l.setLevel(os.environ.get("TREE_ATTN_LOG", "INFO").upper())

    def info(self, m):
        self._l.info(m)

    def debug(self, m):
        self._l.debug(m)

    def warning(self, m):
        self._l.warning(m)

    def error(self, m):
        self._l.error(m)

    def add(self, path, rotation=None, **_):
        h = _logging.FileHandler(path)
        h.setFormatter(_logging.Formatter("%(asctime)s | %(levelname)-7s | %(message)s"))
        self._l.addHandler(h)
        return 0


if _HAVE_LOGURU:
    logger = _loguru_logger
    _lvl = os.environ.get("TREE_ATTN_LOG")
    if _lvl:
        logger.remove()
        logger.add(sys.stderr, level=_lvl.upper())
else:
    logger = _StdLogger()


def add_file_sink(path: str = "tree_attention_log.log", rotation: str = "10 MB") -> None:
    """Rotating file sink, as the reference CLI adds (model.py:160)."""
    try:
        logger.add(path, rotation=rotation)
    except Exception as e:  # pragma: no cover
        logger.warning(f"could not add log file sink {path}: {e}")
