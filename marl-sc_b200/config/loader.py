"""YAML -> validated config objects (reference: src/config/loader.py:17-29, 32-133, 223-244, 290-315).

Same entry points and error types as the reference loader; ``synthetic`` / ``real_world`` data
generation and tune configs are outside this hot path.
"""
from __future__ import annotations

import copy
from pathlib import Path
from typing import Any, Dict, Optional, Union

import yaml
from pydantic import TypeAdapter, ValidationError

from .schema import AlgorithmConfig, CPPOConfig, EnvironmentConfig, FeatureConfig, IPPOConfig, MAPPOConfig


class ConfigError(Exception):
    """Base class for configuration problems."""


class ConfigFileError(ConfigError):
    """The file is missing, not a file, or not parseable YAML."""


class ConfigValidationError(ConfigError):
    """The content does not satisfy the schema."""


def load_yaml(path: Union[str, Path]) -> Dict[str, Any]:
    p = Path(path)
    if not p.exists():
        raise ConfigFileError(f"Config file not found: {path}")
    if not p.is_file():
        raise ConfigFileError(f"Path is not a file: {path}")
    try:
        with open(p, "r", encoding="utf-8") as fh:
            data = yaml.safe_load(fh)
    except yaml.YAMLError as exc:
        raise ConfigFileError(f"Error parsing YAML file {path}: {exc}")
    except OSError as exc:
        raise ConfigFileError(f"Error reading config file {path}: {exc}")
    return {} if data is None else data


def validate_config(config_dict: Dict[str, Any], schema):
    try:
        return TypeAdapter(schema).validate_python(config_dict)
    except ValidationError as exc:
        lines = [" -> ".join(str(loc) for loc in err["loc"]) + f": {err['msg']}" for err in exc.errors()]
        raise ConfigValidationError("Configuration validation failed:\n" + "\n".join(lines))


def load_feature_config(path: Union[str, Path]) -> FeatureConfig:
    d = load_yaml(path)
    return validate_config(d.get("features", d), FeatureConfig)


def _migrate_env_config(d: Dict[str, Any]) -> Dict[str, Any]:
    """Legacy layout: top-level ``max_order_quantities`` (scalar or list) -> direct action space."""
    if d.get("action_space") is not None:
        return d
    mq = d.pop("max_order_quantities", None)
    if mq is None:
        return d
    n = d.get("n_skus", 1)
    vals = [int(mq)] * n if isinstance(mq, (int, float)) else [int(x) for x in mq]
    d["action_space"] = {"type": "direct", "params": {"max_order_quantities": vals}}
    return d


def _resolve_relative(ref: str, yaml_path: Optional[Path]) -> Path:
    p = Path(ref)
    if p.is_absolute() or p.exists() or yaml_path is None:
        return p
    for anc in yaml_path.resolve().parents:       # the reference resolves against its repo root (cwd)
        if (anc / p).exists():
            return anc / p
    return p


def environment_config_from_dict(env: Dict[str, Any], yaml_path: Optional[Path] = None) -> EnvironmentConfig:
    d = copy.deepcopy(env.get("environment", env))
    d = _migrate_env_config(d)
    fpath = d.pop("feature_config_path", None)
    if fpath is not None:
        d["features"] = load_feature_config(_resolve_relative(fpath, yaml_path)).model_dump()
    if d.get("data_source", {}).get("type") == "synthetic":
        raise ConfigValidationError(
            "data_source.type='synthetic' needs the reference's raw CSVs and fitted generator models, "
            "which are not part of this hot path; use a 'custom' data source")
    return validate_config(d, EnvironmentConfig)


def load_environment_config(path: Union[str, Path], seed_manager=None) -> EnvironmentConfig:
    """Same call as the reference (loader.py:117); ``seed_manager`` is accepted for signature parity."""
    return environment_config_from_dict(load_yaml(path), Path(path))


def load_algorithm_config(path: Union[str, Path]) -> Union[IPPOConfig, MAPPOConfig, CPPOConfig]:
    d = load_yaml(path)
    return validate_config(d.get("algorithm", d), AlgorithmConfig)


def algorithm_config_from_dict(d: Dict[str, Any]):
    return validate_config(d.get("algorithm", d), AlgorithmConfig)
