from .base import DeviceComponent, StochasticComponent
from .demand_allocator import BaseDemandAllocator, GreedyDemandAllocator
from .demand_sampler import (BaseDemandSampler, EmpiricalDemandSampler, Order, PoissonDemandSampler,
                             ReplayDemandSampler)
from .lead_time_sampler import BaseLeadTimeSampler, FixedLeadTimeSampler, StochasticLeadTimeSampler
from .lost_sales_handler import (BaseLostSalesHandler, ClosestLostSalesHandler, CostLostSalesHandler,
                                 ShipmentLostSalesHandler)
from .reward_calculator import BaseRewardCalculator, CostRewardCalculator
