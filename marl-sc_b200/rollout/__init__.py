from .baselines import AdaptiveBaseStock, base_stock_levels, baseline_rollout, constant_actions, random_actions
from .collector import Rollout, RolloutCollector, shard_envs
from .gae import compute_gae, standardize_
from .obs_stats import compute_obs_statistics
from .policy import ActorCritic, mlp
from .ppo import PPOLearner

__all__ = ["AdaptiveBaseStock", "constant_actions", "random_actions", "base_stock_levels", "baseline_rollout", "Rollout", "RolloutCollector", "shard_envs", "compute_gae", "compute_obs_statistics", "standardize_", "ActorCritic", "mlp", "PPOLearner"]
