from . import AnyOf, AllOf, Event, Timeout, Process, Condition  # noqa: F401  (runner/check.py:5 imports AnyOf)
