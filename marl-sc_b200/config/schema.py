"""Pydantic models for the environment and PPO-family algorithm YAML files.

Field names, literals and cross-field rules follow the reference so its shipped YAML files load
unmodified (reference: src/config/schema.py - initial inventory :15-69, cost structure :77-181,
components :188-430, action space :541-580, features :598-639, EnvironmentConfig :646-899,
shared/PPO/IPPO/MAPPO/CPPO :986-1225). Tune / GRU / CNN network schemas are out of scope for this
hot path: only MLP actor/critic descriptions are accepted.

One deliberate extension: ``EnvironmentConfig.allow_region_mismatch`` (default False) lifts the
``n_regions == n_warehouses`` rule (reference schema.py:670-675) that the env code itself never relies
on, so the 10 warehouse x 50 region network of BASELINE.json can be described; ``allow_empirical_frame`` lets the
``empirical`` demand sampler run on a caller-supplied demand frame with a ``custom`` data source.
"""
from __future__ import annotations

from typing import Any, Dict, List, Literal, Optional, Union

from pydantic import (BaseModel, ConfigDict, Field, NonNegativeFloat, NonNegativeInt, PositiveFloat,
                      PositiveInt, field_validator, model_validator)
from typing_extensions import Annotated


class _Strict(BaseModel):
    model_config = ConfigDict(extra="forbid")


def _rect(v, what: str):
    """Non-empty rectangular 2-D list or None."""
    if v is None:
        return v
    if not v or any(not row for row in v):
        raise ValueError(f"{what} must be a non-empty 2D list")
    if len({len(row) for row in v}) != 1:
        raise ValueError(f"{what} must be rectangular (all rows same length)")
    return v


def _is_num(x) -> bool:
    return isinstance(x, (int, float)) and not isinstance(x, bool)


# ---------------------------------------------------------------- initial inventory
class InitialInventoryUniform(_Strict):
    type: Literal["uniform"]
    params: Dict[Literal["min", "max"], NonNegativeInt]

    @model_validator(mode="after")
    def _bounds(self):
        lo, hi = self.params.get("min"), self.params.get("max")
        if lo is None or hi is None:
            raise ValueError("uniform params must contain min and max")
        if lo > hi:
            raise ValueError("uniform params must satisfy min <= max")
        return self


class InitialInventoryCustom(_Strict):
    type: Literal["custom"]
    params: Dict[Literal["values"], Union[NonNegativeInt, List[List[NonNegativeInt]]]]

    @field_validator("params", mode="after")
    @classmethod
    def _shape(cls, v):
        val = v["values"]
        if isinstance(val, bool):
            raise ValueError("custom params.values must be an int, not a bool")
        if not isinstance(val, int):
            _rect(val, "custom params.values")
        return v


class InitialInventoryZero(_Strict):
    type: Literal["zero"]
    params: Optional[None] = None


InitialInventoryConfig = Union[InitialInventoryUniform, InitialInventoryCustom, InitialInventoryZero]


# ---------------------------------------------------------------- cost structure
class ShipmentCostConfig(_Strict):
    outbound_fixed: Optional[List[List[NonNegativeFloat]]] = None
    outbound_variable: Optional[List[List[NonNegativeFloat]]] = None
    inbound_fixed: Optional[List[List[NonNegativeFloat]]] = None
    inbound_variable: Optional[List[List[NonNegativeFloat]]] = None

    @model_validator(mode="after")
    def _shapes(self):
        for name in ("outbound_fixed", "outbound_variable", "inbound_fixed", "inbound_variable"):
            _rect(getattr(self, name), f"shipment_cost.{name}")
        for side in ("outbound", "inbound"):
            var, fix = getattr(self, f"{side}_variable"), getattr(self, f"{side}_fixed")
            if var is None:
                continue
            if fix is None:
                raise ValueError(f"shipment_cost.{side}_variable cannot be specified without shipment_cost.{side}_fixed")
            if len(var) != len(fix) or len(var[0]) != len(fix[0]):
                raise ValueError(f"shipment_cost.{side}_variable must have the same shape as {side}_fixed")
        return self


class CostStructureConfig(_Strict):
    holding_cost: Union[PositiveFloat, List[PositiveFloat]]
    penalty_cost: Union[NonNegativeFloat, List[NonNegativeFloat]]
    shipment_cost: ShipmentCostConfig
    sku_weights: Optional[List[PositiveFloat]] = None
    distances: Optional[List[List[NonNegativeFloat]]] = None


# ---------------------------------------------------------------- components
class DemandSamplerPoisson(_Strict):
    type: Literal["poisson"]
    params: Dict[Literal["lambda_orders", "probability_skus", "lambda_quantity"],
                 Union[PositiveFloat, List[PositiveFloat], List[List[PositiveFloat]]]]

    @field_validator("params", mode="after")
    @classmethod
    def _shapes(cls, v):
        for k in ("lambda_orders", "probability_skus", "lambda_quantity"):
            if k not in v:
                raise ValueError(f"poisson params must contain '{k}'")
        lo, ps, lq = v["lambda_orders"], v["probability_skus"], v["lambda_quantity"]
        for i, p in enumerate(ps if isinstance(ps, list) else [ps]):
            if _is_num(p) and p > 1:
                raise ValueError(f"probability_skus[{i}] must be <= 1.0, got {p}")
        scal = [_is_num(x) for x in (lo, ps, lq)]
        if all(scal):
            return v
        if any(scal):
            raise ValueError("poisson params must be either all scalars or all arrays; cannot mix scalar and array parameters")
        if isinstance(lo[0], list) or isinstance(ps[0], list):
            raise ValueError("poisson params.lambda_orders / probability_skus must be 1D lists in array mode")
        if not isinstance(lq[0], list):
            raise ValueError("poisson params.lambda_quantity must be a 2D list when using array mode")
        if len(lo) != len(ps):
            raise ValueError("poisson params.lambda_orders and probability_skus must have the same length")
        _rect(lq, "poisson params.lambda_quantity")
        if len(lq) != len(lo):
            raise ValueError(f"poisson params.lambda_quantity must have {len(lo)} rows")
        return v


class DemandSamplerEmpirical(_Strict):
    type: Literal["empirical"]
    params: Optional[None] = None


class DemandSamplerReplay(_Strict):
    """Extension: demand supplied as pre-sampled order tensors (parity / benchmark runs)."""
    type: Literal["replay"]
    params: Optional[Dict[str, Any]] = None


DemandSamplerConfig = Union[DemandSamplerPoisson, DemandSamplerEmpirical, DemandSamplerReplay]


class DemandAllocatorGreedy(_Strict):
    type: Literal["greedy"]
    params: Dict[Literal["max_splits"], Union[Literal["default"], NonNegativeInt]]


class DemandAllocatorLP(_Strict):
    type: Literal["lp"]
    params: Optional[Dict[str, Any]] = None


DemandAllocatorConfig = Union[DemandAllocatorGreedy, DemandAllocatorLP]


class DeviationConfig(_Strict):
    type: Literal["uniform"]
    max_deviation: Union[NonNegativeInt, List[NonNegativeInt]]


class FixedLeadTimeParams(_Strict):
    expected_lead_times: List[List[PositiveInt]]

    @field_validator("expected_lead_times", mode="after")
    @classmethod
    def _shape(cls, v):
        return _rect(v, "expected_lead_times")


class StochasticLeadTimeParams(FixedLeadTimeParams):
    deviation: DeviationConfig


class LeadTimeSamplerFixed(_Strict):
    type: Literal["fixed"]
    params: FixedLeadTimeParams


class LeadTimeSamplerStochastic(_Strict):
    type: Literal["stochastic"]
    params: StochasticLeadTimeParams


LeadTimeSamplerConfig = Union[LeadTimeSamplerFixed, LeadTimeSamplerStochastic]


class LostSalesClosest(_Strict):
    type: Literal["closest"]
    params: Optional[None] = None


class LostSalesShipment(_Strict):
    type: Literal["shipment"]
    params: Optional[None] = None


class LostSalesCost(_Strict):
    type: Literal["cost"]
    params: Dict[Literal["alpha"], NonNegativeFloat]


LostSalesHandlerConfig = Union[LostSalesClosest, LostSalesShipment, LostSalesCost]

COST_TYPES = ["holding_cost", "penalty_cost", "outbound_shipment_cost", "inbound_shipment_cost"]


class RewardCostParams(_Strict):
    scope: Literal["team", "agent"]
    scale_factor: PositiveFloat
    cost_weights: List[NonNegativeFloat] = Field(min_length=1)

    @field_validator("cost_weights")
    @classmethod
    def _weights(cls, v):
        if any(x > 1.0 for x in v):
            raise ValueError("cost_weights must be in [0.0, 1.0]")
        if abs(float(sum(v)) - 1.0) > 1e-6:
            raise ValueError(f"cost_weights must sum to 1.0 (got {float(sum(v))})")
        return v


class RewardCalculatorCost(_Strict):
    type: Literal["cost"]
    params: RewardCostParams

    @model_validator(mode="after")
    def _n_weights(self):
        if len(self.params.cost_weights) != len(COST_TYPES):
            raise ValueError(f"reward_calculator.params.cost_weights must have length {len(COST_TYPES)} "
                             f"(one weight per cost type: {COST_TYPES})")
        return self


RewardCalculatorConfig = Union[RewardCalculatorCost]


class ComponentsConfig(_Strict):
    demand_sampler: DemandSamplerConfig = Field(..., discriminator="type")
    demand_allocator: DemandAllocatorConfig = Field(..., discriminator="type")
    lead_time_sampler: LeadTimeSamplerConfig = Field(..., discriminator="type")
    lost_sales_handler: LostSalesHandlerConfig = Field(..., discriminator="type")
    reward_calculator: RewardCalculatorConfig = Field(..., discriminator="type")


# ---------------------------------------------------------------- data source / action space / features
class DataSourceCustom(_Strict):
    type: Literal["custom"]


class DataSourceOther(BaseModel):
    """``synthetic`` / ``real_world`` sources need the reference's raw CSVs and pickled models, which
    are not part of this hot path; accepted for schema compatibility."""
    model_config = ConfigDict(extra="allow")
    type: Literal["synthetic", "real_world"]


DataSourceConfig = Union[DataSourceCustom, DataSourceOther]


class ActionSpaceDirectParams(_Strict):
    max_order_quantities: List[PositiveInt]


class ActionSpaceDirect(_Strict):
    type: Literal["direct"]
    params: ActionSpaceDirectParams


class ActionSpaceDemandCenteredParams(_Strict):
    max_quantity_adjustment: List[PositiveInt]


class ActionSpaceDemandCentered(_Strict):
    type: Literal["demand_centered"]
    params: ActionSpaceDemandCenteredParams


class ActionSpaceBaseStockParams(_Strict):
    max_stock_level: List[PositiveInt]


class ActionSpaceBaseStock(_Strict):
    type: Literal["base_stock"]
    params: ActionSpaceBaseStockParams


ActionSpaceConfig = Union[ActionSpaceDirect, ActionSpaceDemandCentered, ActionSpaceBaseStock]

_AGGREGATE_PARENTS = [
    ("inventory", "inventory_aggregate"), ("pipeline", "pipeline_aggregate"),
    ("incoming_demand_home", "incoming_demand_home_aggregate"),
    ("units_shipped_away", "units_shipped_away_aggregate"),
    ("rolling_demand_mean", "rolling_demand_mean_aggregate"),
    ("demand_forecast", "demand_forecast_aggregate")]


class FeatureConfig(_Strict):
    inventory: bool = True
    pipeline: bool = True
    incoming_demand_home: bool = True
    units_shipped_home: bool = True
    units_shipped_away: bool = True
    stockout: bool = True
    rolling_demand_mean: bool = True
    demand_forecast: bool = True
    days_of_supply: bool = False
    net_inventory_position: bool = False
    demand_variability: bool = False
    demand_history: bool = False
    inventory_aggregate: bool = True
    pipeline_aggregate: bool = True
    incoming_demand_home_aggregate: bool = True
    units_shipped_away_aggregate: bool = True
    rolling_demand_mean_aggregate: bool = True
    demand_forecast_aggregate: bool = True

    @model_validator(mode="after")
    def _rules(self):
        if not self.inventory:
            raise ValueError("inventory must always be enabled")
        if not self.pipeline:
            raise ValueError("pipeline must always be enabled")
        for parent, agg in _AGGREGATE_PARENTS:
            if getattr(self, agg) and not getattr(self, parent):
                raise ValueError(f"'{agg}' cannot be enabled when '{parent}' is disabled")
        return self


# ---------------------------------------------------------------- environment
class EnvironmentConfig(_Strict):
    n_warehouses: PositiveInt
    n_skus: PositiveInt
    n_regions: PositiveInt
    episode_length: PositiveInt
    max_wh_capacities: List[PositiveFloat]
    action_space: ActionSpaceConfig = Field(..., discriminator="type")
    initial_inventory: InitialInventoryConfig = Field(..., discriminator="type")
    cost_structure: CostStructureConfig
    components: ComponentsConfig
    data_source: DataSourceConfig = Field(..., discriminator="type")
    features: FeatureConfig = Field(default_factory=FeatureConfig)
    allow_region_mismatch: bool = False
    # extension: the empirical sampler replays a demand frame handed in through env_meta["preprocessed_data"]
    # (marlsc_b200.data) with cost tables given in the config, because the raw CSVs the reference's real_world data
    # source preprocesses are not part of its repository
    allow_empirical_frame: bool = False

    @model_validator(mode="after")
    def _shape_checks(self):
        W, S, R = self.n_warehouses, self.n_skus, self.n_regions
        if R != W and not self.allow_region_mismatch:
            raise ValueError(f"n_regions ({R}) must equal n_warehouses ({W}) "
                             f"(home region assumption: each warehouse is assigned to exactly one region)")

        def need(cond, msg):
            if not cond:
                raise ValueError(msg)

        ap = self.action_space.params
        for attr in ("max_order_quantities", "max_quantity_adjustment", "max_stock_level"):
            if hasattr(ap, attr):
                need(len(getattr(ap, attr)) == S,
                     f"action_space.params.{attr} must have length n_skus={S}, got {len(getattr(ap, attr))}")
        need(len(self.max_wh_capacities) == W, f"max_wh_capacities must have length n_warehouses={W}, "
                                              f"got {len(self.max_wh_capacities)}")
        if isinstance(self.initial_inventory, InitialInventoryCustom):
            val = self.initial_inventory.params["values"]
            if isinstance(val, list):
                need(len(val) == W, f"initial_inventory.custom params.values must have {W} rows (n_warehouses), got {len(val)}")
                need(all(len(r) == S for r in val), f"initial_inventory.custom params.values rows must have length n_skus={S}")
        cs = self.cost_structure
        for nm in ("holding_cost", "penalty_cost"):
            v = getattr(cs, nm)
            need(not isinstance(v, list) or len(v) == S, f"{nm} list must have length n_skus={S}")
        sc = cs.shipment_cost
        for nm, cols, cname in (("outbound_fixed", R, "n_regions"), ("outbound_variable", R, "n_regions"),
                                ("inbound_fixed", S, "n_skus"), ("inbound_variable", S, "n_skus")):
            m = getattr(sc, nm)
            if m is not None:
                need(len(m) == W, f"shipment_cost.{nm} must have {W} rows (n_warehouses), got {len(m)}")
                need(all(len(r) == cols for r in m), f"shipment_cost.{nm} must have {cols} columns ({cname}) in every row")
        ds = self.components.demand_sampler
        if isinstance(ds, DemandSamplerPoisson) and isinstance(ds.params["lambda_orders"], list):
            lo, ps, lq = ds.params["lambda_orders"], ds.params["probability_skus"], ds.params["lambda_quantity"]
            need(len(lo) == R, f"demand_sampler.poisson params.lambda_orders must have length n_regions={R}, got {len(lo)}")
            need(len(ps) == R, f"demand_sampler.poisson params.probability_skus must have length n_regions={R}, got {len(ps)}")
            need(len(lq) == R, f"demand_sampler.poisson params.lambda_quantity must have {R} rows (n_regions), got {len(lq)}")
            need(all(len(r) == S for r in lq), f"demand_sampler.poisson params.lambda_quantity rows must have length n_skus={S}")
        lt = self.components.lead_time_sampler
        elt = lt.params.expected_lead_times
        need(len(elt) == W, f"lead_time_sampler params.expected_lead_times must have {W} rows (n_warehouses), got {len(elt)}")
        need(all(len(r) == S for r in elt), f"lead_time_sampler params.expected_lead_times rows must have length n_skus={S}")
        if isinstance(lt, LeadTimeSamplerStochastic):
            md = lt.params.deviation.max_deviation
            need(not isinstance(md, list) or len(md) == S,
                 f"lead_time_sampler deviation.max_deviation list must have length n_skus={S}")
        need(cs.sku_weights is None or len(cs.sku_weights) == S, f"sku_weights list must have length n_skus={S}")
        if cs.distances is not None:
            _rect(cs.distances, "distances")
            need(len(cs.distances) == W, f"distances must have {W} rows (n_warehouses), got {len(cs.distances)}")
            need(all(len(r) == R for r in cs.distances), f"distances must have {R} columns (n_regions) in every row")
        return self

    @model_validator(mode="after")
    def _post_checks(self):
        alloc = self.components.demand_allocator
        if isinstance(alloc, DemandAllocatorGreedy):
            ms = alloc.params["max_splits"]
            if ms == "default":
                alloc.params["max_splits"] = self.n_warehouses - 1
            elif ms >= self.n_warehouses:
                raise ValueError(f"demand_allocator.greedy max_splits must be < n_warehouses={self.n_warehouses}")
        if (isinstance(self.components.demand_sampler, DemandSamplerEmpirical) and self.data_source.type != "real_world"
                and not self.allow_empirical_frame):
            raise ValueError("demand_sampler.type='empirical' requires data_source.type='real_world', "
                             f"got data_source.type='{self.data_source.type}'")
        if self.data_source.type == "custom":
            sc = self.cost_structure.shipment_cost
            for nm in ("outbound_fixed", "outbound_variable", "inbound_fixed", "inbound_variable"):
                if getattr(sc, nm) is None:
                    raise ValueError(f"shipment_cost.{nm} must be specified when using custom data source. ")
            if self.cost_structure.sku_weights is None:
                raise ValueError("sku_weights must be specified when using custom data source. ")
            if self.cost_structure.distances is None:
                raise ValueError("distances must be specified when using custom data source. ")
        return self


# ---------------------------------------------------------------- algorithm (PPO family)
ActivationName = Literal["relu", "tanh", "sigmoid", "elu", "selu", "gelu", "swish", "mish",
                         "hard_swish", "hard_sigmoid"]


class MLPConfig(_Strict):
    hidden_sizes: List[PositiveInt] = Field(default_factory=lambda: [256])
    activation: ActivationName = "relu"
    output_activation: Optional[ActivationName] = None
    output_activation_mu: Optional[ActivationName] = None
    output_activation_sigma: Optional[ActivationName] = None
    output_dim: Optional[PositiveInt] = None


class NetworkConfig(_Strict):
    type: Literal["mlp"]          # gru / cnn builders are outside this hot path
    config: MLPConfig = Field(default_factory=MLPConfig)


class ActorCriticConfig(_Strict):
    shared_layers: Optional[NetworkConfig] = None
    actor: NetworkConfig
    critic: NetworkConfig
    use_mu_sigma_head: bool = False

    @model_validator(mode="after")
    def _heads(self):
        c = self.critic.config
        if c.output_activation_mu is not None or c.output_activation_sigma is not None:
            raise ValueError("output_activation_mu / output_activation_sigma are only valid for the actor")
        a = self.actor.config
        if not self.use_mu_sigma_head and (a.output_activation_mu is not None or a.output_activation_sigma is not None):
            raise ValueError("output_activation_mu / output_activation_sigma need use_mu_sigma_head")
        if self.use_mu_sigma_head and a.output_activation is not None:
            raise ValueError("output_activation must not be set on the actor when use_mu_sigma_head is true")
        return self


class SharedAlgorithmConfig(_Strict):
    num_iterations: PositiveInt
    checkpoint_freq: PositiveInt
    batch_size: PositiveInt
    num_epochs: PositiveInt
    num_minibatches: PositiveInt
    learning_rate: Union[PositiveFloat, List[List[Union[int, float]]]]
    num_env_runners: NonNegativeInt = 0
    num_envs_per_env_runner: NonNegativeInt = 1
    num_cpus_per_env_runner: PositiveInt = 1
    eval_interval: PositiveInt = 1
    num_eval_episodes: PositiveInt = 1
    evaluation_parallel_to_training: bool = False

    @field_validator("learning_rate", mode="after")
    @classmethod
    def _schedule(cls, v):
        if _is_num(v):
            return v
        if len(v) < 2 or any(len(p) != 2 for p in v):
            raise ValueError("learning_rate schedule must be a list of at least two [timestep, lr] pairs")
        if v[0][0] != 0:
            raise ValueError(f"learning_rate schedule must start at timestep 0, got timestep {v[0][0]}")
        if any(v[i][0] <= v[i - 1][0] for i in range(1, len(v))):
            raise ValueError("learning_rate schedule timesteps must be strictly increasing")
        if any(p[1] < 0 for p in v):
            raise ValueError("learning_rate schedule values must be non-negative")
        return v

    @model_validator(mode="after")
    def _relations(self):
        if self.batch_size < self.num_minibatches:
            raise ValueError(f"batch_size ({self.batch_size}) must be >= num_minibatches ({self.num_minibatches})")
        if self.batch_size % self.num_minibatches != 0:
            raise ValueError(f"batch_size ({self.batch_size}) must be divisible by num_minibatches ({self.num_minibatches})")
        if self.checkpoint_freq > self.num_iterations:
            raise ValueError(f"checkpoint_freq ({self.checkpoint_freq}) must be <= num_iterations ({self.num_iterations})")
        return self


class PPOConfig(_Strict):
    use_gae: Optional[bool] = True
    lam: Optional[NonNegativeFloat] = 0.95
    gamma: Optional[NonNegativeFloat] = 0.99
    use_kl_loss: Optional[bool] = False
    grad_clip: Optional[NonNegativeFloat] = None
    entropy_coeff: Optional[NonNegativeFloat] = 0.01
    vf_loss_coeff: Optional[NonNegativeFloat] = 1
    clip_param: Optional[PositiveFloat] = 0.2
    vf_clip_param: Optional[NonNegativeFloat] = 10
    logstd_init: Optional[float] = 0
    logstd_floor: Optional[float] = -2.0

    @model_validator(mode="after")
    def _ranges(self):
        if self.clip_param > 1.0:
            raise ValueError("clip_param should typically be <= 1.0")
        if not 0.0 <= self.lam <= 1.0:
            raise ValueError("lam must be in [0.0, 1.0]")
        if not 0.0 <= self.gamma <= 1.0:
            raise ValueError("gamma must be in [0.0, 1.0]")
        return self


ObsNormalization = Literal["off", "meanstd", "meanstd_custom", "meanstd_grouped", "ratio"]


class _MultiAgentSpecific(PPOConfig):
    obs_normalization: ObsNormalization = "off"
    parameter_sharing: bool = False
    hysteretic_beta: Optional[float] = None
    warmstart_weights_path: Optional[str] = None
    actor_obs_type: Literal["local", "global"] = "local"
    networks: ActorCriticConfig

    @field_validator("hysteretic_beta", mode="after")
    @classmethod
    def _beta(cls, v):
        if v is not None and not 0.0 < v <= 1.0:
            raise ValueError("hysteretic_beta must be in (0.0, 1.0]")
        return v

    @model_validator(mode="after")
    def _shared_layers(self):
        if self.networks.shared_layers is not None and self.actor_obs_type != self.critic_obs_type:
            raise ValueError("Shared layers require actor_obs_type and critic_obs_type to match")
        return self


class IPPOSpecificConfig(_MultiAgentSpecific):
    critic_obs_type: Literal["local", "global"] = "local"


class MAPPOSpecificConfig(_MultiAgentSpecific):
    critic_obs_type: Literal["local", "global"] = "global"


class CPPOSpecificConfig(PPOConfig):
    obs_normalization: ObsNormalization = "off"
    warmstart_weights_path: Optional[str] = None
    networks: ActorCriticConfig


class IPPOConfig(_Strict):
    name: Literal["ippo"]
    shared: SharedAlgorithmConfig
    algorithm_specific: IPPOSpecificConfig


class MAPPOConfig(_Strict):
    name: Literal["mappo"]
    shared: SharedAlgorithmConfig
    algorithm_specific: MAPPOSpecificConfig


class CPPOConfig(_Strict):
    name: Literal["cppo"]
    shared: SharedAlgorithmConfig
    algorithm_specific: CPPOSpecificConfig


AlgorithmConfig = Annotated[Union[IPPOConfig, MAPPOConfig, CPPOConfig], Field(discriminator="name")]
