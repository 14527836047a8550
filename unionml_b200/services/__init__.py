"""Service shims that call the same device path as ``Model.predict`` (ref. ``unionml/services/__init__.py``)."""
from unionml_b200.services.runnable import PredictRunnable, create_runnable  # noqa: F401
