"""Drop the accelerated classes into an upstream `omnisafe` checkout.

    import omnisafe, omnisafe_b200.integration
    omnisafe_b200.integration.install()
    omnisafe.Agent('PPOLag', 'SyntheticBox-v0', custom_cfgs={'train_cfgs': {'device': 'cuda:0', ...}, ...}).learn()

After `install()` the reference's own `omnisafe.Agent` (algorithms/algo_wrapper.py:L56-269) -- its config loading,
key checking, `distributed.fork` and `registry.get(algo)(env_id, cfgs)` (algo_wrapper.py:L149-170) -- constructs the
omnisafe_b200 classes: same class names, so the upstream YAMLs and `ALGORITHM2TYPE` keep applying
(algorithms/__init__.py:L69-85).  The registry refuses duplicate registrations (registry.py:L56-57), so the entries
are REPLACED in its table rather than registered again.
"""
from __future__ import annotations

import sys


def accelerated_classes() -> dict:
    from omnisafe_b200.algorithms import on_policy as mine  # noqa: PLC0415
    from omnisafe_b200.algorithms import registry as my_registry  # noqa: PLC0415

    return {name: my_registry.get(name) for name in my_registry.REGISTRY.names() if my_registry.get(name).__module__ == mine.__name__}


def install(omnisafe_module=None) -> list[str]:
    """Swap every accelerated on-policy class into `omnisafe`'s registry and namespace; register the HBM-resident
    `SyntheticBox-v0` id with the reference's env table so that `AlgoWrapper._init_checks` (algo_wrapper.py:L140-147)
    accepts it.  Returns the swapped class names."""
    omnisafe = omnisafe_module or sys.modules.get('omnisafe')
    if omnisafe is None:
        import omnisafe  # noqa: PLC0415
    from omnisafe.algorithms import registry as ref_registry  # noqa: PLC0415
    from omnisafe.envs.core import CMDP, ENV_REGISTRY, support_envs  # noqa: PLC0415

    swapped = []
    for name, cls in accelerated_classes().items():
        if name in ref_registry.REGISTRY._module_dict:       # only names the reference knows: its YAMLs / type table apply
            ref_registry.REGISTRY._module_dict[name] = cls
            for mod_name in ('omnisafe.algorithms.on_policy', 'omnisafe.algorithms'):
                mod = sys.modules.get(mod_name)
                if mod is not None and hasattr(mod, name):
                    setattr(mod, name, cls)
            swapped.append(name)
    if 'SyntheticBox-v0' not in support_envs():
        class SyntheticBoxPlaceholder(CMDP):     # the accelerated adapter steps this env inside its CUDA kernels
            _support_envs = ['SyntheticBox-v0']  # noqa: RUF012
            need_auto_reset_wrapper = False
            need_time_limit_wrapper = False
            need_evaluation = False

            def __init__(self, env_id, **kwargs):
                raise RuntimeError('SyntheticBox-v0 is stepped in-kernel by omnisafe_b200; for CPU runs of the unmodified '
                                   'reference use tests/golden/make_golden.py::RefSyntheticBox')

            step = reset = set_seed = render = close = None
            max_episode_steps = None

        SyntheticBoxPlaceholder.__abstractmethods__ = frozenset()
        ENV_REGISTRY.register(SyntheticBoxPlaceholder)
    return sorted(swapped)
