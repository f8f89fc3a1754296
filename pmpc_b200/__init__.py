"""pmpc_b200 — B200-native drop-in for the ``solve`` path of StanfordASL/pmpc."""
from .handle import DeviceSolver  # noqa: F401
