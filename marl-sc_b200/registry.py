"""Name -> class registries for the five pluggable step components
(reference: src/environment/registry.py:15-308: same ``register_*`` / ``get_*`` pairs, same
registered names, same "Unknown ...: X. Available: [...]" errors)."""
from __future__ import annotations

from typing import Dict, Optional, Type

from .config.schema import EnvironmentConfig
from .context import EnvironmentContext, create_environment_context

DEMAND_SAMPLER_REGISTRY: Dict[str, Type] = {}
DEMAND_ALLOCATOR_REGISTRY: Dict[str, Type] = {}
LEAD_TIME_SAMPLER_REGISTRY: Dict[str, Type] = {}
LOST_SALES_HANDLER_REGISTRY: Dict[str, Type] = {}
REWARD_CALCULATOR_REGISTRY: Dict[str, Type] = {}


def _build(registry: Dict[str, Type], label: str, field: str, env_config: EnvironmentConfig,
           context: Optional[EnvironmentContext]):
    component_config = getattr(env_config.components, field)
    kind = component_config.type
    if kind not in registry:
        raise ValueError(f"Unknown {label}: {kind}. Available: {list(registry.keys())}")
    if context is None:
        if env_config.data_source.type == "real_world":
            raise ValueError("context must be provided when using real_world data source. "
                             "Preprocessing requires a seed that should be spawned in InventoryEnvironment.__init__().")
        context = create_environment_context(env_config)
    return registry[kind](context, component_config)


def _check_component(kind: str, name: str, cls: Type) -> None:
    """A registered component is a device component spec: the step runs in CUDA kernels, so a class cannot bring its
    own arithmetic (the reference's ``allocate`` / ``calculate_lost_sales`` / ``calculate`` hooks) - it contributes an
    enum plus parameter tables through ``spec_fields()`` and may only select among the behaviours the kernels implement."""
    if not isinstance(cls, type) or not callable(getattr(cls, "spec_fields", None)):
        raise TypeError(f"{kind} '{name}': {getattr(cls, '__name__', cls)} must be a class with a spec_fields() method "
                        "(subclass marlsc_b200.components.base.DeviceComponent or one of the registered components); "
                        "host-side callables cannot run inside the CUDA step")


def register_demand_sampler(name: str, sampler_class: Type):
    _check_component("demand sampler", name, sampler_class)
    DEMAND_SAMPLER_REGISTRY[name] = sampler_class


def get_demand_sampler(env_config: EnvironmentConfig, context: Optional[EnvironmentContext] = None):
    return _build(DEMAND_SAMPLER_REGISTRY, "demand sampler", "demand_sampler", env_config, context)


def register_demand_allocator(name: str, allocator_class: Type):
    _check_component("demand allocator", name, allocator_class)
    DEMAND_ALLOCATOR_REGISTRY[name] = allocator_class


def get_demand_allocator(env_config: EnvironmentConfig, context: Optional[EnvironmentContext] = None):
    return _build(DEMAND_ALLOCATOR_REGISTRY, "demand allocator", "demand_allocator", env_config, context)


def register_lead_time_sampler(name: str, sampler_class: Type):
    _check_component("lead time sampler", name, sampler_class)
    LEAD_TIME_SAMPLER_REGISTRY[name] = sampler_class


def get_lead_time_sampler(env_config: EnvironmentConfig, context: Optional[EnvironmentContext] = None):
    return _build(LEAD_TIME_SAMPLER_REGISTRY, "lead time sampler", "lead_time_sampler", env_config, context)


def register_lost_sales_handler(name: str, handler_class: Type):
    _check_component("lost sales handler", name, handler_class)
    LOST_SALES_HANDLER_REGISTRY[name] = handler_class


def get_lost_sales_handler(env_config: EnvironmentConfig, context: Optional[EnvironmentContext] = None):
    return _build(LOST_SALES_HANDLER_REGISTRY, "lost sales handler", "lost_sales_handler", env_config, context)


def register_reward_calculator(name: str, calculator_class: Type):
    _check_component("reward calculator", name, calculator_class)
    REWARD_CALCULATOR_REGISTRY[name] = calculator_class


def get_reward_calculator(env_config: EnvironmentConfig, context: Optional[EnvironmentContext] = None):
    return _build(REWARD_CALCULATOR_REGISTRY, "reward calculator", "reward_calculator", env_config, context)


from .components import (ClosestLostSalesHandler, CostLostSalesHandler, CostRewardCalculator,  # noqa: E402
                         EmpiricalDemandSampler, FixedLeadTimeSampler, GreedyDemandAllocator,
                         PoissonDemandSampler, ReplayDemandSampler, ShipmentLostSalesHandler,
                         StochasticLeadTimeSampler)

register_demand_sampler("poisson", PoissonDemandSampler)
register_demand_sampler("empirical", EmpiricalDemandSampler)
register_demand_sampler("replay", ReplayDemandSampler)
register_demand_allocator("greedy", GreedyDemandAllocator)
register_lead_time_sampler("fixed", FixedLeadTimeSampler)
register_lead_time_sampler("stochastic", StochasticLeadTimeSampler)
register_lost_sales_handler("closest", ClosestLostSalesHandler)
register_lost_sales_handler("shipment", ShipmentLostSalesHandler)
register_lost_sales_handler("cost", CostLostSalesHandler)
register_reward_calculator("cost", CostRewardCalculator)
