from .batched_env import BatchedInventoryEnv, DeviceOrders
from .multi_env import InventoryEnvironment

__all__ = ["BatchedInventoryEnv", "DeviceOrders", "InventoryEnvironment"]
