"""CPU tests of the host-side logic: config surface, registry, Agent checks, and the N > 1 path
(world_size-2 gloo): global steps_per_epoch partitioning, env-id sharding, averaging helpers."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def test_config_merge_and_unknown_keys():
    from omnisafe_b200.utils.config import get_default_kwargs_yaml, recursive_check_config

    cfg = get_default_kwargs_yaml('PPOLag', 'Some-Other-Env', 'on-policy')
    assert cfg.algo_cfgs.steps_per_epoch == 20000 and cfg.algo_cfgs.batch_size == 64      # upstream defaults
    assert cfg.lagrange_cfgs.lambda_lr == 0.035 and cfg.model_cfgs.actor.hidden_sizes == [64, 64]
    cfg2 = get_default_kwargs_yaml('PPOLag', 'SyntheticBox-v0', 'on-policy')                 # env block merged
    assert cfg2.train_cfgs.vector_env_nums == 4096 and cfg2.algo_cfgs.batch_size == 16384
    with pytest.raises(KeyError):
        recursive_check_config({'algo_cfgs': {'not_a_key': 1}}, cfg)
    recursive_check_config({'algo_cfgs': {'clip': 0.1}, 'env_cfgs': {'anything': 1}}, cfg)
    cfg.recurisve_update({'algo_cfgs': {'clip': 0.1}})
    assert cfg.algo_cfgs.clip == 0.1 and cfg.algo_cfgs.gamma == 0.99
    for algo in ('TRPOLag', 'CPO', 'FOCOPS'):
        c = get_default_kwargs_yaml(algo, 'x', 'on-policy')
        assert c.algo_cfgs.use_cost is True
    assert get_default_kwargs_yaml('CPO', 'x').algo_cfgs.cg_iters == 15


def test_registry_and_agent_checks():
    import omnisafe_b200
    from omnisafe_b200.algorithms import ALGORITHM2TYPE, registry

    assert ALGORITHM2TYPE['PPOLag'] == 'on-policy' and registry.get('CPO').__name__ == 'CPO'
    with pytest.raises(KeyError):
        registry.get('NoSuchAlgo')
    with pytest.raises(KeyError):
        registry.register(registry.get('PPOLag'))          # duplicate names are rejected
    with pytest.raises(AssertionError):
        omnisafe_b200.Agent('NoSuchAlgo', 'SyntheticBox-v0')
    with pytest.raises(AssertionError):
        omnisafe_b200.Agent('PPOLag', 'NoSuchEnv-v0')
    with pytest.raises(KeyError):
        omnisafe_b200.Agent('PPOLag', 'SyntheticBox-v0', custom_cfgs={'algo_cfgs': {'bogus': 1}})
    # device='cpu' is refused loudly: the path has no CPU fallback
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        omnisafe_b200.Agent('PPOLag', 'SyntheticBox-v0', custom_cfgs={'train_cfgs': {'device': 'cpu'}})


def test_param_layout_matches_reference_counts():
    from omnisafe_b200.models import param_layout

    lay = param_layout(60, 8)    # actor 8 592, each critic 8 129
    assert lay['actor']['size'] == 8592 and lay['reward_critic']['size'] == 8129 and lay['total'] == 24850
    assert list(lay['actor']['entries'])[0] == 'log_std'


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from omnisafe_b200.utils import distributed

    distributed.init_process_group('cpu')      # gloo
    assert distributed.world_size() == world and distributed.get_rank() == rank
    # steps_per_epoch is global: local T = steps // world // envs (policy_gradient.py:L70-77)
    T = distributed.local_steps(2 * 16 * 8, 16)
    avg = distributed.dist_avg(torch.tensor(float(rank + 1)))
    sums = torch.tensor([1.0 + rank, 10.0, 2.0 * rank, 5.0], dtype=torch.float64)
    distributed.all_reduce_(sums)
    try:
        distributed.local_steps(2 * 16 * 8 + 1, 16)
        bad = False
    except AssertionError:
        bad = True
    q.put((rank, T, float(avg), sums.tolist(), rank * 16, bad))


def test_two_rank_gloo_host_logic():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, T, avg, sums, env_off, bad in res:
        assert T == 8 and abs(avg - 1.5) < 1e-6 and bad
        assert sums == [3.0, 20.0, 2.0, 10.0]          # SUM over ranks of the epoch statistics vector
        assert env_off == rank * 16                    # rank r owns global envs [r*N, (r+1)*N)


def test_cpo_step_direction_ieee_fallbacks():
    """after an inexact CG solve A = q - r^2/s can be slightly negative and s <= 0; the reference
    (second_order/cpo.py:L271-337, fp32 tensors) then gets NaN, the NaN comparison is False and lambda_b_star is
    chosen / the step is NaN (rejected by the line search) -- it never raises."""
    import math
    from types import SimpleNamespace as NS

    import torch

    from omnisafe_b200.algorithms.on_policy import CPO

    me = NS(_cfgs=NS(algo_cfgs=NS(target_kl=0.01)))
    x, p = torch.ones(4), torch.full((4,), 0.5)
    # case 1/2 with A < 0: lambda_a is NaN -> f_a >= f_b is False -> lambda_b_star
    q, kl = 0.3, 0.01
    step, lam, nu = CPO._step_direction(me, 2, q, x, -1e-3, 0.5, q, p, 0.1, 1.0, -0.2)
    lam_b = max(math.sqrt(q / (2 * kl)), 0.1 / (-0.2 + 1e-8))
    assert abs(lam - lam_b) < 1e-9 and torch.isfinite(step).all()
    # B == 0: A / B is inf, not ZeroDivisionError
    step, lam, nu = CPO._step_direction(me, 1, q, x, 0.2, 0.0, q, p, 0.1, 1.0, 0.2)
    assert math.isfinite(lam) and torch.isfinite(step).all()
    # case 0 with s < 0: NaN step (the line search then rejects it) instead of `math domain error`
    step, lam, nu = CPO._step_direction(me, 0, q, x, 0.2, -0.1, q, p, 0.1, -0.5, 0.2)
    assert nu != nu and torch.isnan(step).all()
    # well-conditioned values still give the closed form
    step, lam, nu = CPO._step_direction(me, 3, 0.5, x, 0.1, -0.1, 0.5, p, 0.1, 1.0, -0.2)
    assert abs(float(step[0]) - math.sqrt(2 * kl / (0.5 + 1e-8))) < 1e-7


def test_simmer_controller_matches_reference_sequences():
    """SimmerPIDAgent (host-side product code) == the reference controller over recorded cost sequences (two gain settings,
    the clamp to the budget bound included) -- common/simmer_agent.py:L132-186; fixture from the unmodified reference."""
    from types import SimpleNamespace as NS

    import numpy as np
    import torch

    from omnisafe_b200.common.simmer_agent import SimmerPIDAgent

    c = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'simmer_controller.npz'))
    scale = float(c['scale'])
    for i in range(2):
        kp, ki, kd, polyak = (float(v) for v in c[f'cfg_{i}'])
        agent = SimmerPIDAgent(NS(kp=kp, ki=ki, kd=kd, polyak=polyak), torch.ones(3, 1) * 25.0 * scale)
        budget = torch.ones(3, 1) * 15.0 * scale
        hist = []
        for cost in c['costs']:
            budget = agent.act(budget, torch.as_tensor(cost) * scale)
            hist.append(budget.numpy().copy())
        np.testing.assert_array_equal(np.stack(hist), c[f'budget_{i}'])


def test_unimplemented_upstream_options_are_refused():
    """Options the fused path does not implement must not validate silently."""
    import pytest

    from omnisafe_b200.utils.config import check_all_configs, get_default_kwargs_yaml

    cfgs = get_default_kwargs_yaml('PPOLag', 'SyntheticBox-v0')
    check_all_configs(cfgs)
    cfgs.model_cfgs.exploration_noise_anneal = True
    with pytest.raises(NotImplementedError):
        check_all_configs(cfgs)
    cfgs = get_default_kwargs_yaml('CPO', 'SyntheticBox-v0')
    cfgs.algo_cfgs.fvp_sample_freq = 2
    with pytest.raises(NotImplementedError):
        check_all_configs(cfgs)


def test_saute_per_step_budget_is_the_reference_fp32_value():
    """saute_adapter.py:L62-68: a python-double expression multiplied into an fp32 tensor."""
    import torch

    from omnisafe_b200.adapter.saute_adapter import per_step_budget

    for budget, g, L in ((25.0, 0.999, 1000), (2.0, 0.9, 8), (1.0, 0.9, 8)):
        ref = (budget * (1 - g ** L) / (1 - g) / L * torch.ones(1, 1))[0, 0].item()
        assert per_step_budget(budget, g, L) == ref
