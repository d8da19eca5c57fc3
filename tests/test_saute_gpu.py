"""GPU parity of the Saute / Simmer / EarlyTerminated adapters: the safety-state augmentation inside the fused
rollout kernels vs the fixtures produced by unmodified PPOSaute / PPOSimmerPID rollouts on the synthetic env
(adapter/saute_adapter.py:L135-217, adapter/simmer_adapter.py:L97-131), and the budget controller vs the reference's."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model_cfgs():
    net = NS(hidden_sizes=[64, 64], activation='tanh', lr=3e-4)
    return NS(actor=net, critic=net, actor_type='gaussian_learning', linear_lr_decay=True,
              weight_initialization_mode='kaiming_uniform')


def _rollout(dev, g, adapter_cls, precision, extra_algo=None, control=None):
    from omnisafe_b200.common.buffer import VectorOnPolicyBuffer
    from omnisafe_b200.models import ConstraintActorCritic

    N, T, O, A = int(g['N']), int(g['T']), int(g['O']), int(g['A'])
    algo = dict(obs_normalize=True, reward_normalize=False, cost_normalize=False, safety_budget=float(g['algo_safety_budget']),
                saute_gamma=float(g['algo_saute_gamma']), max_ep_len=float(g['algo_max_ep_len']),
                unsafe_reward=float(g['algo_unsafe_reward']))
    algo.update(extra_algo or {})
    cfgs = NS(algo_cfgs=NS(**algo), logger_cfgs=NS(window_lens=10), control_cfgs=control,
              env_cfgs=dict(obs_dim=O, act_dim=A, max_episode_steps=int(g['tmax']), term_prob=float(g['term_prob'])))
    ad = adapter_cls('SyntheticBox-v0', N, int(g['seed']), cfgs, device=dev)
    ad.precision = precision
    assert ad.obs_dim == O + 1
    agent = ConstraintActorCritic(O + 1, A, _model_cfgs(), epochs=1, device=dev)
    agent.load_flat(g['theta'])
    buf = VectorOnPolicyBuffer(O + 1, A, T, float(g['gamma']), float(g['lam']), float(g['lam_c']), 'gae', 0.0, True, True,
                               num_envs=N, device=dev)
    ad.rollout(T, agent, buf, eps=torch.as_tensor(g['eps']).to(dev))
    torch.cuda.synchronize()
    return ad, buf


@pytest.mark.parametrize('precision', [0, 2])      # exact fp32 FMA tiles / split-bf16 tensor-core tiles (persistent kernel)
def test_saute_rollout_golden(cuda, golden_dir, precision):
    from omnisafe_b200.adapter.saute_adapter import SauteAdapter

    g = np.load(os.path.join(golden_dir, 'rollout_pposaute.npz'))
    ad, buf = _rollout(cuda, g, SauteAdapter, precision)
    T, N, O = int(g['T']), int(g['N']), int(g['O'])
    sl = {k: v.cpu().numpy() for k, v in buf.data.items() if v is not None}
    t = dict(rtol=2e-5, atol=2e-5)
    assert sl['obs'].shape == (T, N, O + 1)
    np.testing.assert_allclose(sl['obs'], g['slab_obs'], **t)              # last column = the safety state z
    np.testing.assert_allclose(sl['act'], g['slab_act'], **t)
    np.testing.assert_allclose(sl['reward'], g['slab_reward'], **t)
    assert (sl['reward'] == np.float32(g['algo_unsafe_reward'])).mean() > 0.2      # the unsafe branch is exercised
    assert np.array_equal(sl['cost'], g['slab_cost'])
    np.testing.assert_allclose(sl['value_r'], g['slab_value_r'], **t)
    buf.finish_paths()
    torch.cuda.synchronize()
    np.testing.assert_allclose(buf.data['adv_r'].cpu().numpy(), g['slab_adv_r'], rtol=1e-4, atol=5e-5)   # bootstrap on z = 1 finals
    meta, ring = ad.ep_meta.cpu().numpy(), ad.ep_ring.cpu().numpy()
    cnt, head = int(meta[0]), int(meta[1])
    order = [(head - cnt + i) % 10 for i in range(cnt)]
    np.testing.assert_allclose(ring[0][order], g['win_ret'], rtol=1e-5, atol=1e-5)     # episode returns keep the raw reward
    assert np.isfinite(ad.ep_budget_mean())


@pytest.mark.parametrize('precision', [0, 2])
def test_simmer_rollout_golden(cuda, golden_dir, precision):
    from omnisafe_b200.adapter.simmer_adapter import SimmerAdapter

    g = np.load(os.path.join(golden_dir, 'rollout_pposimmer.npz'))
    control = NS(kp=0.0005, ki=0.00001, kd=0.0, polyak=0.995)
    ad, buf = _rollout(cuda, g, SimmerAdapter, precision, extra_algo=dict(upper_budget=float(g['algo_upper_budget'])), control=control)
    sl = {k: v.cpu().numpy() for k, v in buf.data.items() if v is not None}
    t = dict(rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(sl['obs'], g['slab_obs'], **t)              # z starts the epoch at safety_budget / upper_budget
    np.testing.assert_allclose(sl['reward'], g['slab_reward'], **t)
    np.testing.assert_allclose(sl['value_r'], g['slab_value_r'], **t)


def test_simmer_controller_golden(golden_dir):
    """SimmerPIDAgent (product code, host side like the reference's) == the reference controller over recorded cost
    sequences (two gain settings, the clamp to the budget bound)."""
    from omnisafe_b200.common.simmer_agent import SimmerPIDAgent

    c = np.load(os.path.join(golden_dir, 'simmer_controller.npz'))
    scale = float(c['scale']) if 'scale' in c else (1 - 0.999 ** 1000) / (1 - 0.999) / 1000
    i = 0
    while f'budget_{i}' in c:
        kp, ki, kd, polyak = [float(x) for x in c[f'cfg_{i}']]
        agent = SimmerPIDAgent(NS(kp=kp, ki=ki, kd=kd, polyak=polyak), torch.ones(3, 1) * 25.0 * scale)
        budget = torch.ones(3, 1) * 15.0 * scale
        hist = []
        for cost in c['costs']:
            budget = agent.act(budget, torch.as_tensor(cost) * scale)
            hist.append(budget.numpy().copy())
        np.testing.assert_allclose(np.stack(hist), c[f'budget_{i}'], rtol=1e-6, atol=1e-7)
        i += 1
    assert i >= 1


def _early_rollout(dev, N, T, O, A, seed, theta, eps, tmax, term_prob, cost_limit, precision):
    from omnisafe_b200.adapter.early_terminated_adapter import EarlyTerminatedAdapter
    from omnisafe_b200.common.buffer import VectorOnPolicyBuffer
    from omnisafe_b200.models import ConstraintActorCritic

    cfgs = NS(algo_cfgs=NS(obs_normalize=True, reward_normalize=False, cost_normalize=False, cost_limit=cost_limit),
              logger_cfgs=NS(window_lens=10), env_cfgs=dict(obs_dim=O, act_dim=A, max_episode_steps=tmax, term_prob=term_prob))
    ad = EarlyTerminatedAdapter('SyntheticBox-v0', N, seed, cfgs, device=dev)
    ad.precision = precision
    agent = ConstraintActorCritic(O, A, _model_cfgs(), epochs=1, device=dev)
    agent.load_flat(theta)
    buf = VectorOnPolicyBuffer(O, A, T, 0.99, 0.95, 0.95, 'gae', 0.0, True, True, num_envs=N, device=dev)
    ad.rollout(T, agent, buf, eps=torch.as_tensor(eps).to(dev))
    torch.cuda.synchronize()
    return ad, buf


@pytest.mark.parametrize('precision', [0, 2])
def test_early_terminated_rollout_golden(cuda, golden_dir, precision):
    """One env, as upstream requires: the slabs of an unmodified PPOEarlyTerminated rollout (20 early terminations, the
    accumulator carried across ordinary episode ends)."""
    g = np.load(os.path.join(golden_dir, 'rollout_ppoearly.npz'))
    N, T, O, A = int(g['N']), int(g['T']), int(g['O']), int(g['A'])
    ad, buf = _early_rollout(cuda, N, T, O, A, int(g['seed']), g['theta'], g['eps'], int(g['tmax']), float(g['term_prob']),
                             float(g['algo_cost_limit']), precision)
    sl = {k: v.cpu().numpy() for k, v in buf.data.items() if v is not None}
    t = dict(rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(sl['obs'], g['slab_obs'], **t)
    np.testing.assert_allclose(sl['act'], g['slab_act'], **t)
    assert np.array_equal(sl['reward'] == 0, g['slab_reward'] == 0) and (sl['reward'] == 0).sum() >= 10
    np.testing.assert_allclose(sl['reward'], g['slab_reward'], **t)
    assert np.array_equal(sl['cost'], g['slab_cost'])
    np.testing.assert_allclose(sl['value_r'], g['slab_value_r'], **t)
    meta, ring = ad.ep_meta.cpu().numpy(), ad.ep_ring.cpu().numpy()
    cnt, head = int(meta[0]), int(meta[1])
    order = [(head - cnt + i) % 10 for i in range(cnt)]
    np.testing.assert_allclose(ring[0][order], g['win_ret'], rtol=1e-5, atol=1e-5)
    assert np.array_equal(ring[1][order], g['win_cost']) and np.array_equal(ring[2][order], g['win_len'])
    buf.finish_paths()
    torch.cuda.synchronize()
    np.testing.assert_allclose(buf.data['adv_r'].cpu().numpy(), g['slab_adv_r'], rtol=1e-4, atol=5e-5)


@pytest.mark.parametrize('precision', [0, 2])
def test_early_terminated_many_envs_vs_oracle(cuda, precision):
    """The per-env generalisation (256 envs, both ends in one step included) against the oracle."""
    from oracle import actor_critic as oac
    from oracle import rollout as orollout
    from oracle.normalizer import Normalizer as ONormalizer
    from oracle.synthetic_env import SyntheticBoxEnv as OEnv

    N, T, O, A, seed, tmax, tp, limit = 256, 40, 17, 6, 3, 6, 0.03, 1.0
    theta = oac.init_theta(O, A, seed=4)
    eps = np.random.default_rng(5).standard_normal((T, N, A)).astype(np.float32)
    ad, buf = _early_rollout(cuda, N, T, O, A, seed, theta, eps, tmax, tp, limit, precision)
    env = OEnv(N, O, A, max_episode_steps=tmax, seed=seed, term_prob=tp)
    ref = orollout.rollout_epoch(env, ONormalizer((O,)), theta, T, eps, early={'cost_limit': limit, 'acc': np.zeros(N, np.float32)})
    sl = {k: v.cpu().numpy() for k, v in buf.data.items() if v is not None}
    t = dict(rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(sl['obs'], ref['obs'], **t)
    np.testing.assert_allclose(sl['reward'], ref['rew'], **t)
    assert np.array_equal(sl['flags'], ref['flags'])
    both = ((ref['flags'] & 2) != 0) & (ref['rew'] == 0)
    assert (ref['rew'] == 0).sum() > 100 and both.sum() > 0            # early terminations, some on a time-limit step
    np.testing.assert_allclose(sl['value_r'], ref['val_r'], **t)


@pytest.mark.parametrize('algo', ['PPOSaute', 'TRPOSaute', 'PPOSimmerPID', 'TRPOSimmerPID', 'PPOEarlyTerminated', 'TRPOEarlyTerminated'])
def test_saute_family_trains_and_logs(cuda, tmp_path, algo):
    import omnisafe_b200

    cfg = {'seed': 0,
           'train_cfgs': {'device': 'cuda', 'vector_env_nums': 256, 'total_steps': 256 * 32 * 2},
           'algo_cfgs': dict({'steps_per_epoch': 256 * 32, 'batch_size': 2048, 'update_iters': 2},
                             **({'cost_limit': 3.0} if 'Early' in algo else {'max_ep_len': 16})),
           'logger_cfgs': {'log_dir': str(tmp_path), 'use_tensorboard': False, 'save_model_freq': 1},
           'env_cfgs': {'obs_dim': 17, 'act_dim': 6, 'max_episode_steps': 16}}
    agent = omnisafe_b200.Agent(algo, 'SyntheticBox-v0', custom_cfgs=cfg)
    ep_ret, ep_cost, ep_len = agent.learn()
    assert np.isfinite(ep_ret) and np.isfinite(ep_cost) and (ep_len <= 16 if 'Early' in algo else ep_len == 16)
    assert agent.agent._actor_critic.obs_dim == (17 if 'Early' in algo else 18)
