"""Import shim (test infrastructure): Bottleneck is not installed in the build container.

Only `move_sum` is used by the reference hot path (boss/runs/reference.py:233-234,259-260).
The arithmetic lives in oracle/move_sum.py.
"""
from oracle.move_sum import move_sum  # noqa: F401
