"""Import shim (test infrastructure) for pettingzoo.ParallelEnv (reference: multi_env.py:6)."""


class ParallelEnv:  # pragma: no cover - placeholder base class
    pass
