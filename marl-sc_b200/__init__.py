"""marlsc_b200 - B200-native batched implementation of marl-sc's inventory-environment hot path
(env reset/step, rollout buffer, GAE). See DESIGN.md. Import as ``import marlsc_b200``
(alias module at the repo root; the directory name ``marl-sc_b200`` is not importable)."""

__version__ = "0.1.0"

from . import config, registry, seeds  # noqa: F401
from .config import load_algorithm_config, load_environment_config  # noqa: F401
