"""CPU, build container only: a checkpoint written by omnisafe_b200 (same classes, CPU tensors, no kernel
launch) is loaded by the UNMODIFIED reference Evaluator (`omnisafe/evaluator.py:L113-178`) and played
in the reference's own wrapper stack.  Skipped where /root/reference does not exist (GPU box)."""
import json
import os
import sys
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.skipif(not os.path.isdir('/root/reference/omnisafe'), reason='reference tree not present')


def test_reference_evaluator_loads_our_checkpoint(tmp_path):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import make_golden  # noqa: F401  (installs the reference shim and registers RefSyntheticBox)
    from omnisafe.evaluator import Evaluator

    from omnisafe_b200.common.logger import Logger
    from omnisafe_b200.common.normalizer import Normalizer
    from omnisafe_b200.models import ConstraintActorCritic
    from omnisafe_b200.utils.config import get_default_kwargs_yaml
    from oracle import actor_critic as oac

    O, A = 12, 3
    cfgs = get_default_kwargs_yaml('PPOLag', 'SyntheticBox-v0', 'on-policy')
    cfgs.recurisve_update({'exp_name': 'PPOLag-{SyntheticBox-v0}', 'env_id': 'SyntheticBox-v0', 'algo': 'PPOLag',
                           'env_cfgs': {'obs_dim': O, 'act_dim': A, 'max_episode_steps': 8, 'term_prob': 0.0},
                           'logger_cfgs': {'log_dir': str(tmp_path)}, 'train_cfgs': {'epochs': 1}})
    ac = ConstraintActorCritic(O, A, cfgs.model_cfgs, epochs=1, device='cpu')
    theta = oac.init_theta(O, A, seed=3)
    ac.load_flat(theta)
    norm = Normalizer((O,), clip=5.0, device='cpu')
    norm.mean.copy_(torch.linspace(-0.2, 0.2, O)); norm.std.fill_(1.5); norm.sumsq.fill_(2.25 * 99); norm.count[0] = 100
    logger = Logger(str(tmp_path), cfgs.exp_name, seed=0, config=cfgs)
    logger.setup_torch_saver({'pi': ac.actor_state_dict, 'obs_normalizer': norm})
    logger.torch_save()
    logger.close()
    assert json.load(open(os.path.join(logger.log_dir, 'config.json')))['algo'] == 'PPOLag'

    ev = Evaluator()
    ev.load_saved(save_dir=logger.log_dir, model_name='epoch-0.pt')
    # the reference actor now holds OUR parameters: same deterministic action as the oracle forward
    obs = torch.linspace(-1, 1, O).reshape(1, O)
    with torch.no_grad():
        act_ref = ev._actor.predict(obs, deterministic=True).numpy()
    nets = oac.unflatten(torch.as_tensor(theta), O, A)
    np.testing.assert_allclose(act_ref, oac.mlp(nets['actor'], obs).numpy(), rtol=1e-6, atol=1e-6)
    # and the wrapper stack normalises with OUR statistics
    w = ev._env
    while not hasattr(w, '_obs_normalizer'):
        w = w._env
    np.testing.assert_allclose(w._obs_normalizer.mean.numpy(), norm.mean.numpy())
    np.testing.assert_allclose(w._obs_normalizer.std.numpy(), norm.std.numpy())
    rets, costs = ev.evaluate(num_episodes=2)
    assert len(rets) == 2 and np.isfinite(rets).all() and np.isfinite(costs).all()
