"""Import shim (test infrastructure): lets the reference env import without gymnasium.

Only the names the reference's env-side modules touch are provided
(reference: src/environment/envs/multi_env.py:5, single_env.py:17-19). No arithmetic lives here.
"""
from . import spaces  # noqa: F401


class Env:  # pragma: no cover - placeholder base class
    pass
