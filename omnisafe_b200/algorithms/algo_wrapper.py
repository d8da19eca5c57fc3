"""`omnisafe_b200.Agent` -- the `omnisafe.Agent(algo, env_id, train_terminal_cfgs, custom_cfgs)`
entry point (mirrors omnisafe/algorithms/algo_wrapper.py:L36-269: config merge, checks,
`distributed.fork`, `registry.get(algo)(env_id, cfgs)`, `learn()`)."""
from __future__ import annotations

import sys

from omnisafe_b200.algorithms import ALGORITHM2TYPE, registry
from omnisafe_b200.envs import support_envs
from omnisafe_b200.utils import distributed
from omnisafe_b200.utils.config import (Config, check_all_configs, get_default_kwargs_yaml,
                                        recursive_check_config)


class AlgoWrapper:
    def __init__(self, algo: str, env_id: str, train_terminal_cfgs: dict | None = None,
                 custom_cfgs: dict | None = None) -> None:
        self.algo, self.env_id = algo, env_id
        self.train_terminal_cfgs, self.custom_cfgs = train_terminal_cfgs, custom_cfgs
        self._evaluator = None
        self.cfgs = self._init_config()
        self._init_checks()
        self._init_algo()

    def _init_config(self) -> Config:
        assert self.algo in ALGORITHM2TYPE, f'{self.algo} doesn\'t exist. Please choose from {list(ALGORITHM2TYPE)}.'
        self.algo_type = ALGORITHM2TYPE[self.algo]
        cfgs = get_default_kwargs_yaml(self.algo, self.env_id, self.algo_type)
        cfgs.recurisve_update({'exp_name': f'{self.algo}-{{{self.env_id}}}', 'env_id': self.env_id, 'algo': self.algo})
        if self.custom_cfgs:
            recursive_check_config(self.custom_cfgs, cfgs, exclude_keys=('algo', 'env_id'))
            cfgs.recurisve_update(self.custom_cfgs)
        if self.train_terminal_cfgs:
            recursive_check_config(self.train_terminal_cfgs, cfgs.train_cfgs)
            cfgs.train_cfgs.recurisve_update(self.train_terminal_cfgs)
        epochs = cfgs.train_cfgs.total_steps // cfgs.algo_cfgs.steps_per_epoch
        cfgs.train_cfgs.recurisve_update({'epochs': epochs})
        return cfgs

    def _init_checks(self) -> None:
        assert isinstance(self.algo, str), 'algo must be a string!'
        assert isinstance(self.cfgs.train_cfgs.parallel, int), 'parallel must be an integer!'
        assert self.cfgs.train_cfgs.parallel > 0, 'parallel must be greater than 0!'
        assert self.env_id in support_envs(), (
            f"{self.env_id} doesn't exist. omnisafe_b200 accelerates {support_envs()}; "
            'use upstream omnisafe for simulator-backed environments.')

    def _init_algo(self) -> None:
        check_all_configs(self.cfgs)
        if distributed.fork(self.cfgs.train_cfgs.parallel, device=self.cfgs.train_cfgs.device):
            sys.exit()
        self.agent = registry.get(self.algo)(env_id=self.env_id, cfgs=self.cfgs)

    def learn(self) -> tuple[float, float, float]:
        return self.agent.learn()

    def evaluate(self, *_, **__):
        raise NotImplementedError('evaluation / rendering stay with the upstream Evaluator, which loads '
                                  "this run's torch_save/epoch-k.pt (omnisafe/evaluator.py:L113-178)")
