"""Package logger (same name and default behaviour as the reference's fbpinns/util/logger.py:22-41)."""
import logging
import sys

logger = logging.getLogger("fbpinns")
if not logger.handlers:
    _h = logging.StreamHandler(sys.stdout)
    _h.setFormatter(logging.Formatter("[%(levelname)s] %(asctime)s - %(message)s", "%Y-%m-%d %H:%M:%S"))
    logger.addHandler(_h)
    logger.setLevel(logging.INFO)
    logger.propagate = False
