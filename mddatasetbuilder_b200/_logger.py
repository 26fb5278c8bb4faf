"""Logger with the reference's format string (reference _logger.py:7-14; coloredlogs is optional)."""
import logging

from ._version import version as __version__

logger = logging.getLogger(__name__)
_FMT = f"%(asctime)s - MDDatasetBuilder {__version__} - %(levelname)s: %(message)s"
try:  # pragma: no cover - cosmetic
    import coloredlogs

    coloredlogs.install(fmt=_FMT, level=logging.INFO, milliseconds=True, logger=logger)
except Exception:
    if not logger.handlers:
        _h = logging.StreamHandler()
        _h.setFormatter(logging.Formatter(_FMT))
        logger.addHandler(_h)
    logger.setLevel(logging.INFO)
