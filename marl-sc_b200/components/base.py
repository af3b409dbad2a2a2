"""Component base classes (reference: src/environment/components/base.py:7-23).

In this implementation a registered component is a *device component spec*: it validates its
parameters on the host and contributes an enum plus parameter tables to the ``marlsc_env_spec_t``
the fused step kernel is built from (``spec_fields``). Samplers additionally keep a host-side
``sample`` that replays the reference's NumPy stream draw for draw.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Dict, Optional

import numpy as np


class DeviceComponent(ABC):
    """Anything that contributes fields to the kernel spec."""

    def spec_fields(self) -> Dict[str, Any]:
        return {}


class StochasticComponent(DeviceComponent):
    @abstractmethod
    def reset(self, rng: Optional[np.random.Generator] = None):
        ...
