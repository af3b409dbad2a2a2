from .batched_env import BatchedInventoryEnv, DeviceOrders, HostRollout
from .multi_env import InventoryEnvironment
from .single_env import CentralizedEnvWrapper

__all__ = ["BatchedInventoryEnv", "CentralizedEnvWrapper", "DeviceOrders", "HostRollout", "InventoryEnvironment"]
