"""Import shim (test infrastructure): the reference imports minknow_api at module import time."""
__version__ = "6.0.0"
