from .logging import logger, add_file_sink  # noqa: F401
