"""INTEGRATION.md §2 executed: the import swap applied to the mounted reference, then the reference's OWN
`omnisafe.Agent` (config loading, key checks, registry lookup) constructs the omnisafe_b200 class -- up to its device
check (there is no GPU in the build container and the path has no CPU fallback).  Skipped where /root/reference is
absent (GPU box)."""
import os

import pytest

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'omnisafe')), reason='reference tree not mounted')


def test_reference_agent_constructs_the_accelerated_class():
    from oracle import ref_shim

    ref_shim.install()
    import omnisafe
    from omnisafe.algorithms import registry as ref_registry

    import omnisafe_b200.integration as integ
    from omnisafe_b200.algorithms import on_policy as mine

    upstream = ref_registry.get('PPOLag')
    swapped = integ.install()
    assert {'PPOLag', 'CPO', 'TRPOLag', 'FOCOPS', 'PPO', 'TRPO', 'PCPO', 'RCPO', 'PDO'} <= set(swapped)
    assert ref_registry.get('PPOLag') is mine.PPOLag and ref_registry.get('PPOLag') is not upstream
    assert omnisafe.algorithms.on_policy.PPOLag is mine.PPOLag
    # same constructor contract as BaseAlgo (algorithms/base_algo.py:L34-53): (env_id, cfgs)
    import inspect
    assert list(inspect.signature(mine.PPOLag.__init__).parameters)[1:] == ['env_id', 'cfgs']
    # the reference's Agent: upstream PPOLag.yaml + custom_cfgs key checking + registry.get(algo)(env_id, cfgs)
    custom = {'train_cfgs': {'vector_env_nums': 8, 'total_steps': 8 * 16 * 2},
              'algo_cfgs': {'steps_per_epoch': 8 * 16, 'batch_size': 32, 'update_iters': 2},
              'logger_cfgs': {'use_tensorboard': False, 'use_wandb': False, 'log_dir': '/tmp/osb_dropin'}}
    with pytest.raises(RuntimeError, match='omnisafe_b200 runs this path as sm_100a CUDA kernels only'):
        omnisafe.Agent('PPOLag', 'SyntheticBox-v0', custom_cfgs=custom)          # default device 'cpu' -> OUR class refuses
    # an unknown custom key is still rejected by the reference's own checker before our class is reached
    with pytest.raises(KeyError):
        omnisafe.Agent('PPOLag', 'SyntheticBox-v0', custom_cfgs={'algo_cfgs': {'no_such_key': 1}})
    # the accelerated Lagrange keeps the reference signature update_lagrange_multiplier(Jc: float)
    from omnisafe.common.lagrange import Lagrange as RefLagrange
    from omnisafe_b200.common.lagrange import Lagrange
    assert list(inspect.signature(Lagrange.update_lagrange_multiplier).parameters) == \
        list(inspect.signature(RefLagrange.update_lagrange_multiplier).parameters)
