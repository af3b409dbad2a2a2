from .loader import (ConfigError, ConfigFileError, ConfigValidationError, algorithm_config_from_dict,
                     environment_config_from_dict, load_algorithm_config, load_environment_config,
                     load_feature_config, load_yaml, validate_config)
from .schema import (AlgorithmConfig, CPPOConfig, EnvironmentConfig, FeatureConfig, IPPOConfig, MAPPOConfig)

__all__ = ["ConfigError", "ConfigFileError", "ConfigValidationError", "algorithm_config_from_dict",
           "environment_config_from_dict", "load_algorithm_config", "load_environment_config",
           "load_feature_config", "load_yaml", "validate_config", "AlgorithmConfig", "CPPOConfig",
           "EnvironmentConfig", "FeatureConfig", "IPPOConfig", "MAPPOConfig"]
