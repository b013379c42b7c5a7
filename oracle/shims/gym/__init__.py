"""gym stub (oracle side only): rl_env/WRSN.py:3-4,21,31-32 needs gym.Env and spaces.Box."""
from . import spaces  # noqa: F401


class Env:
    pass
