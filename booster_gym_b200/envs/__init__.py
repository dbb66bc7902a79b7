from .t1 import T1  # noqa: F401  (the reference resolves the task class by name: utils/runner.py:27)
