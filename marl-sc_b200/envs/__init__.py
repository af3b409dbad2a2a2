from .batched_env import BatchedInventoryEnv, DeviceOrders, HostRollout
from .multi_env import InventoryEnvironment

__all__ = ["BatchedInventoryEnv", "DeviceOrders", "HostRollout", "InventoryEnvironment"]
