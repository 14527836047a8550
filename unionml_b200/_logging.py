"""Package logger.  The reference logs through ``logging.getLogger("unionml")`` at INFO with a stream handler
(``unionml/_logging.py``); callback failures on the predict path are reported here and never propagated."""
import logging


def _make_logger() -> logging.Logger:
    log = logging.getLogger("unionml_b200")
    log.setLevel(logging.INFO)
    if not any(isinstance(h, logging.StreamHandler) for h in log.handlers):
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter("%(asctime)s %(name)s %(levelname)s %(message)s"))
        log.addHandler(handler)
    return log


logger = _make_logger()
