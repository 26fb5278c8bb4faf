"""No-op stand-in for coloredlogs (reference _logger.py:9). Test infrastructure only."""
import logging


def install(level=logging.INFO, logger=None, fmt=None, **kw):
    if logger is not None:
        logger.setLevel(level)
        if not logger.handlers:
            h = logging.StreamHandler()
            if fmt:
                h.setFormatter(logging.Formatter(fmt))
            logger.addHandler(h)
