"""Attribute-dict configuration with the reference's three-level merge.

Mirrors omnisafe/utils/config.py:L27-262 and tools.py:L246-269: `Config` (nested attribute access,
`recurisve_update`, `todict`), `get_default_kwargs_yaml(algo, env_id, algo_type)` = YAML `defaults`
block (+) optional `<env_id>:` block, and `recursive_check_config` (unknown keys raise KeyError).
"""
from __future__ import annotations

import json
import os
from typing import Any

import yaml

_CFG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'configs')


class Config:
    def __init__(self, **kwargs: Any) -> None:
        for key, value in kwargs.items():
            self[key] = value

    def __setitem__(self, key: str, value: Any) -> None:
        setattr(self, key, Config.dict2config(value) if isinstance(value, dict) else value)

    def __getitem__(self, key: str) -> Any:
        return getattr(self, key)

    def __contains__(self, key: str) -> bool:
        return key in self.__dict__

    def items(self):
        return self.__dict__.items()

    @staticmethod
    def dict2config(d: dict) -> 'Config':
        cfg = Config()
        for k, v in d.items():
            cfg[k] = v
        return cfg

    def todict(self) -> dict:
        return {k: (v.todict() if isinstance(v, Config) else v) for k, v in self.__dict__.items()}

    def tojson(self) -> str:
        return json.dumps(self.todict(), indent=4)

    def recurisve_update(self, update_args: dict) -> None:  # (sic) reference spelling kept
        for key, value in update_args.items():
            if key in self and isinstance(self[key], Config) and isinstance(value, dict):
                self[key].recurisve_update(value)
            elif key in self and isinstance(self[key], Config) and isinstance(value, Config):
                self[key].recurisve_update(value.todict())
            else:
                self[key] = value

    def __repr__(self) -> str:
        return f'Config({self.todict()})'


def recursive_check_config(config: dict, default_config, exclude_keys: tuple = ()) -> None:
    """Unknown keys in a custom config raise KeyError (omnisafe/utils/tools.py:L246-269)."""
    assert isinstance(config, dict), 'custom_cfgs must be a dict!'
    for key in config:
        if key not in default_config and key not in exclude_keys:
            raise KeyError(f'Invalid key: {key}')
        if isinstance(config[key], dict) and key != 'env_cfgs':
            recursive_check_config(config[key], default_config[key])


def get_default_kwargs_yaml(algo: str, env_id: str, algo_type: str = 'on-policy') -> Config:
    path = os.path.join(_CFG_DIR, algo_type, f'{algo}.yaml')
    with open(path, encoding='utf-8') as fh:
        kwargs = yaml.safe_load(fh)
    cfg = Config.dict2config(kwargs['defaults'])
    env_spec = kwargs.get(env_id)
    if env_spec is not None:
        cfg.recurisve_update(env_spec)
    return cfg


def check_all_configs(cfgs: Config) -> None:
    """Range / type checks the hot path relies on (subset of omnisafe/utils/config.py:L265-408)."""
    a, t = cfgs.algo_cfgs, cfgs.train_cfgs
    assert isinstance(a.update_iters, int) and a.update_iters > 0, 'update_iters must be a positive int'
    assert isinstance(a.steps_per_epoch, int) and a.steps_per_epoch > 0
    assert isinstance(a.batch_size, int) and a.batch_size > 0
    assert 0.0 <= a.gamma <= 1.0 and 0.0 <= a.lam <= 1.0 and 0.0 <= a.lam_c <= 1.0
    assert a.adv_estimation_method in ('gae', 'gae-rtg', 'vtrace', 'plain')
    assert a.penalty_coef >= 0.0
    assert isinstance(t.vector_env_nums, int) and t.vector_env_nums >= 1
    assert isinstance(t.parallel, int) and t.parallel >= 1
    assert t.total_steps >= a.steps_per_epoch, 'total_steps must cover at least one epoch'
    # upstream keys the fused path does not implement: refuse the settings that would change behaviour, say so for the rest
    m = getattr(cfgs, 'model_cfgs', None)
    if m is not None and getattr(m, 'exploration_noise_anneal', False):
        raise NotImplementedError('model_cfgs.exploration_noise_anneal=True is not implemented on the fused path '
                                  '(the Gaussian actor keeps its learned log_std)')
    if getattr(a, 'fvp_sample_freq', 1) != 1:
        # the reference subsamples the env-major batch (obs[::freq]); the slabs are time-major, so a stride would pick a
        # different sample set
        raise NotImplementedError('algo_cfgs.fvp_sample_freq != 1 is not implemented on the fused path')
    lg = getattr(cfgs, 'logger_cfgs', None)
    if lg is not None and (getattr(lg, 'use_wandb', False)):
        import warnings
        warnings.warn('logger_cfgs.use_wandb is ignored by omnisafe_b200 (progress.csv / config.json / torch_save only)', stacklevel=2)
