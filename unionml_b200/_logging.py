"""Logger ``unionml_b200`` (the reference logs through ``logging.getLogger("unionml")``, ``unionml/_logging.py:3-7``)."""
import logging

logger = logging.getLogger("unionml_b200")
if not logger.handlers:
    _h = logging.StreamHandler()
    _h.setFormatter(logging.Formatter("%(asctime)s %(name)s %(levelname)s %(message)s"))
    logger.addHandler(_h)
logger.setLevel(logging.INFO)
