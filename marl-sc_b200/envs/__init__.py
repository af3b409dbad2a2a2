from .batched_env import BatchedInventoryEnv, DeviceLines, DeviceOrders, HostRollout
from .multi_env import InventoryEnvironment
from .single_env import CentralizedEnvWrapper

__all__ = ["BatchedInventoryEnv", "CentralizedEnvWrapper", "DeviceLines", "DeviceOrders", "HostRollout", "InventoryEnvironment"]
