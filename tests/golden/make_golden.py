"""Generate the golden fixtures under tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shim.py) in the build container.

    python tests/golden/make_golden.py

The fixtures pin the CPU oracle (tests/test_oracle_golden.py); the GPU parity tests then compare
the CUDA path with the oracle.  /root/reference does not exist on the GPU box, so this script is
never run there -- only its committed outputs travel.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
import omnisafe  # noqa: E402,F401
from gymnasium.spaces import Box  # noqa: E402  (shim)
from omnisafe.common.buffer import VectorOnPolicyBuffer  # noqa: E402
from omnisafe.common.normalizer import Normalizer  # noqa: E402
from omnisafe.envs.core import CMDP, env_register  # noqa: E402
from omnisafe.utils.math import conjugate_gradients, discount_cumsum  # noqa: E402

from oracle.synthetic_env import SyntheticBoxEnv as OracleEnv  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def gen_discount_cumsum():
    rng = np.random.default_rng(1)
    out = {}
    for i, (n, d) in enumerate([(5, 0.9), (5, 0.99), (5, 0.999), (1, 0.5), (37, 0.9405), (200, 0.99)]):
        x = np.arange(1, 6, dtype=np.float64) if n == 5 else rng.standard_normal(n)
        y = discount_cumsum(torch.as_tensor(x.astype(np.float32)), d).numpy()
        out[f'x{i}'] = x.astype(np.float32); out[f'd{i}'] = np.float64(d); out[f'y{i}'] = y
    out['n'] = 6
    np.savez(os.path.join(OUT, 'discount_cumsum.npz'), **out)


def gen_buffer_gae(estimator='gae', fname='buffer_gae.npz', seed=2):
    """Drive the reference VectorOnPolicyBuffer with random data + random path ends."""
    rng = np.random.default_rng(seed)
    N, T, O, A = 6, 48, 3, 2
    gamma, lam, lam_c, pen = 0.99, 0.95, 0.9, 0.05
    buf = VectorOnPolicyBuffer(
        obs_space=Box(-1, 1, (O,)), act_space=Box(-1, 1, (A,)), size=T, gamma=gamma, lam=lam,
        lam_c=lam_c, advantage_estimator=estimator, penalty_coefficient=pen,
        standardized_adv_r=True, standardized_adv_c=True, num_envs=N)
    rew = rng.random((T, N)).astype(np.float32)
    cost = (rng.random((T, N)) < 0.2).astype(np.float32)
    val_r = rng.standard_normal((T, N)).astype(np.float32)
    val_c = rng.standard_normal((T, N)).astype(np.float32)
    obs = rng.standard_normal((T, N, O)).astype(np.float32)
    act = rng.standard_normal((T, N, A)).astype(np.float32)
    logp = rng.standard_normal((T, N)).astype(np.float32)
    flags = np.zeros((T, N), np.uint8)
    flags[rng.random((T, N)) < 0.06] |= 1
    flags[rng.random((T, N)) < 0.06] |= 2
    boot_r = rng.standard_normal((T, N)).astype(np.float32)
    boot_c = rng.standard_normal((T, N)).astype(np.float32)
    for t in range(T):
        buf.store(obs=torch.as_tensor(obs[t]), act=torch.as_tensor(act[t]),
                  reward=torch.as_tensor(rew[t]), cost=torch.as_tensor(cost[t]),
                  value_r=torch.as_tensor(val_r[t]), value_c=torch.as_tensor(val_c[t]),
                  logp=torch.as_tensor(logp[t]))
        for i in range(N):
            if flags[t, i] or t == T - 1:
                term = bool(flags[t, i] & 1)
                lr = torch.zeros(1) if term else torch.as_tensor(boot_r[t, i:i + 1])
                lc = torch.zeros(1) if term else torch.as_tensor(boot_c[t, i:i + 1])
                buf.finish_path(lr, lc, i)
    raw = {k: np.stack([b.data[k].numpy().copy() for b in buf.buffers], 1)  # -> [T, N]
           for k in ('adv_r', 'adv_c', 'target_value_r', 'target_value_c', 'discounted_ret')}
    data = buf.get()
    np.savez(os.path.join(OUT, fname), estimator=estimator, rew=rew, cost=cost, val_r=val_r, val_c=val_c,
             obs=obs, act=act, logp=logp, flags=flags, boot_r=boot_r, boot_c=boot_c,
             gamma=gamma, lam=lam, lam_c=lam_c, pen=pen,
             **{'raw_' + k: v for k, v in raw.items()},
             **{'get_' + k: v.numpy() for k, v in data.items()})


def gen_normalizer():
    rng = np.random.default_rng(3)
    norm = Normalizer((5,), clip=5)
    batches = [rng.standard_normal((n, 5)).astype(np.float32) * s + m
               for n, s, m in [(16, 1.0, 0.0), (3, 2.0, 1.0), (16, 0.5, -1.0), (1, 1.0, 0.0), (16, 3.0, 2.0)]]
    out = {'nb': len(batches)}
    for i, b in enumerate(batches):
        y = norm.normalize(torch.as_tensor(b)).numpy()
        out[f'x{i}'] = b; out[f'y{i}'] = y
        out[f'mean{i}'] = norm.mean.numpy().copy(); out[f'std{i}'] = norm.std.numpy().copy()
    np.savez(os.path.join(OUT, 'normalizer.npz'), **out)


@env_register
class RefSyntheticBox(CMDP):
    """The synthetic Box env as an ordinary reference CMDP (torch-CPU wrapper around the numpy
    spec in oracle/synthetic_env.py) so the unmodified reference classes can roll it out."""

    _support_envs = ['SyntheticBox-v0']  # noqa: RUF012
    need_auto_reset_wrapper = False
    need_time_limit_wrapper = False
    need_evaluation = False

    def __init__(self, env_id, num_envs=1, device='cpu', **kw):
        super().__init__(env_id)
        self._num_envs = num_envs
        # keep the env-spec keys only (the Evaluator also passes render_mode / camera / size kwargs)
        self._kw = {k: v for k, v in kw.items() if k in ('obs_dim', 'act_dim', 'max_episode_steps', 'term_prob', 'cost_threshold')}
        self._env = None
        O, A = kw.get('obs_dim', 60), kw.get('act_dim', 8)
        self._observation_space = Box(-10.0, 10.0, (O,))
        self._action_space = Box(-1.0, 1.0, (A,))
        self._seed = 0

    def set_seed(self, seed):
        self._seed = seed
        self._env = OracleEnv(self._num_envs, seed=seed, **self._kw)

    def reset(self, seed=None, options=None):
        if self._env is None:
            self.set_seed(self._seed)
        obs = torch.as_tensor(self._env.reset())
        return (obs[0] if self._num_envs == 1 else obs), {}     # a single env is unbatched (the reference adds Unsqueeze)

    def step(self, action):
        if action.dim() == 1:      # the Evaluator drives a single env with an unbatched action
            action = action.unsqueeze(0)
        nobs, rew, cost, term, trunc, final, fin = self._env.step(action.numpy())
        info = {}
        if fin.any():
            info['final_observation'] = torch.as_tensor(final.copy())
            info['_final_observation'] = torch.as_tensor(fin)
        out = (torch.as_tensor(nobs), torch.as_tensor(rew), torch.as_tensor(cost), torch.as_tensor(term), torch.as_tensor(trunc))
        if self._num_envs == 1:
            out = tuple(x[0] for x in out)
            info = {k: v[0] for k, v in info.items()}
        return (*out, info)

    def render(self):
        return None

    def close(self):
        pass

    @property
    def max_episode_steps(self):
        return self._kw.get('max_episode_steps', 64)


def _build_algo(algo, N, T, O, A, seed, extra_algo=None, tmax=8, term_prob=0.05, epochs=2, extra_lagrange=None):
    from omnisafe.utils.config import get_default_kwargs_yaml
    from omnisafe.utils.tools import recursive_check_config
    from omnisafe.algorithms import registry

    cfgs = get_default_kwargs_yaml(algo, 'SyntheticBox-v0', 'on-policy')
    custom = {
        'seed': seed,
        'train_cfgs': {'vector_env_nums': N, 'total_steps': N * T * epochs, 'torch_threads': 1},
        'algo_cfgs': {'steps_per_epoch': N * T, 'batch_size': 32, 'update_iters': 2, **(extra_algo or {})},
        'logger_cfgs': {'use_tensorboard': False, 'use_wandb': False, 'log_dir': '/tmp/osb_golden_runs',
                        'window_lens': 10},
        'env_cfgs': {'obs_dim': O, 'act_dim': A, 'max_episode_steps': tmax, 'term_prob': term_prob},
    }
    if extra_lagrange:
        custom['lagrange_cfgs'] = dict(extra_lagrange)
    env_cfgs = custom.pop('env_cfgs')   # CPO.yaml has no env_cfgs default block: bypass the key check
    recursive_check_config(custom, cfgs)
    cfgs.recurisve_update(custom)
    cfgs.recurisve_update({'env_cfgs': env_cfgs})
    cfgs.recurisve_update({'exp_name': f'{algo}-golden', 'env_id': 'SyntheticBox-v0', 'algo': algo})
    cfgs.train_cfgs.recurisve_update({'epochs': epochs})
    return registry.get(algo)(env_id='SyntheticBox-v0', cfgs=cfgs)


def _flat_theta(ac):
    parts = []
    for net in (ac.actor, ac.reward_critic, ac.cost_critic):
        parts += [p.detach().reshape(-1) for p in net.parameters()]
    return torch.cat(parts).numpy().copy()


def gen_rollout(name='PPOLag', fname='rollout_ppolag.npz', seed=5, epochs_rolled=1, extra_algo=None, N=8, T=24):
    """Epoch(s) of the unmodified OnPolicyAdapter.rollout + buffer on the synthetic env.  PDO's defaults
    switch RewardNormalize / CostNormalize on (PDO.yaml:L44-46): the slabs then hold normalised values."""
    import torch.distributions.normal as tdn
    import torch.distributions.utils as tdu

    O, A = 12, 3
    algo = _build_algo(name, N, T, O, A, seed, epochs=2, extra_algo=extra_algo)
    theta = _flat_theta(algo._actor_critic)
    drawn = []
    orig = tdn._standard_normal

    def rec(shape, dtype, device):
        e = orig(shape, dtype, device)
        drawn.append(e.clone())
        return e

    tdn._standard_normal = rec
    try:
        for e in range(epochs_rolled):
            if e > 0:
                algo._buf.get()          # drain the buffer; normaliser states carry over to the next epoch
            algo._env.rollout(steps_per_epoch=T, agent=algo._actor_critic, buffer=algo._buf,
                              logger=algo._logger)
    finally:
        tdn._standard_normal = orig
    eps = np.stack([d.numpy() for d in drawn if tuple(d.shape) == (N, A)])
    assert eps.shape[0] == T * epochs_rolled, eps.shape
    bufs = algo._buf.buffers
    fields = ('obs', 'act', 'reward', 'cost', 'value_r', 'value_c', 'logp', 'adv_r', 'adv_c',
              'target_value_r', 'target_value_c', 'discounted_ret')
    data = {k: np.stack([b.data[k].numpy().copy() for b in bufs], 1) for k in fields}
    window = {k: np.array(list(algo._logger._data[k]), np.float32)
              for k in ('Metrics/EpRet', 'Metrics/EpCost', 'Metrics/EpLen')}
    norm = algo._env._env  # walk the wrapper stack to ObsNormalize
    while not hasattr(norm, '_obs_normalizer'):
        norm = norm._env
    nz = norm._obs_normalizer
    extra = {}
    w = algo._env._env
    while w is not None:
        for key, attr in (('rnorm', '_reward_normalizer'), ('cnorm', '_cost_normalizer')):
            if hasattr(w, attr):
                z = getattr(w, attr)
                extra.update({f'{key}_mean': z.mean.numpy(), f'{key}_std': z.std.numpy(), f'{key}_count': int(z._count)})
        w = getattr(w, '_env', None)
    got = algo._buf.get()
    extra.update({'algo_' + k: v for k, v in (extra_algo or {}).items()})
    np.savez(os.path.join(OUT, fname), N=N, T=T, O=O, A=A, seed=seed, theta=theta, epochs_rolled=epochs_rolled, **extra,
             eps=eps, tmax=8, term_prob=0.05, gamma=0.99, lam=0.95, lam_c=0.95,
             norm_mean=nz.mean.numpy(), norm_std=nz.std.numpy(), norm_count=int(nz._count),
             win_ret=window['Metrics/EpRet'], win_cost=window['Metrics/EpCost'],
             win_len=window['Metrics/EpLen'],
             **{'slab_' + k: v for k, v in data.items()},
             **{'get_' + k: v.numpy() for k, v in got.items()})
    return algo


def _record_randperm(fn):
    perms = []
    orig = torch.randperm

    def rec(n, *a, **k):
        out = orig(n, *a, **k)
        perms.append(out.clone())
        return out

    torch.randperm = rec
    try:
        ret = fn()
    finally:
        torch.randperm = orig
    return ret, perms


def _last(logger, key):
    v = logger._data[key]
    return np.array(list(v), np.float64)


def gen_update_ppolag(algo):
    """PPOLag._update of the unmodified reference on the rollout above (2 passes x 6 minibatches)."""
    theta0 = _flat_theta(algo._actor_critic)
    data = {k: v.numpy().copy() for k, v in algo._buf.get().items()}
    lam0 = float(algo._lagrange.lagrangian_multiplier.item())
    Jc = algo._logger.get_stats('Metrics/EpCost')[0]
    _, perms = _record_randperm(algo._update)
    B = data['obs'].shape[0]
    perms = np.stack([p.numpy() for p in perms if p.numel() == B])
    lg = algo._logger
    np.savez(os.path.join(OUT, 'update_ppolag.npz'), theta0=theta0, theta1=_flat_theta(algo._actor_critic),
             lam0=lam0, lam1=float(algo._lagrange.lagrangian_multiplier.item()), Jc=Jc, perms=perms,
             batch_size=32, update_iters=2, cost_limit=25.0, lambda_lr=0.035,
             loss_pi=_last(lg, 'Loss/Loss_pi'), loss_r=_last(lg, 'Loss/Loss_reward_critic'),
             loss_c=_last(lg, 'Loss/Loss_cost_critic'), kl=_last(lg, 'Train/KL'),
             stop_iter=_last(lg, 'Train/StopIter'), ratio=_last(lg, 'Train/PolicyRatio'),
             **{'data_' + k: v for k, v in data.items()})


def gen_update_focops():
    """FOCOPS._update of the unmodified reference (incl. its [b,1] x [b] broadcast in the loss)."""
    N, T, O, A, seed = 8, 24, 12, 3, 11
    algo = _build_algo('FOCOPS', N, T, O, A, seed, tmax=8, term_prob=0.05)
    theta0 = _flat_theta(algo._actor_critic)
    algo._env.rollout(steps_per_epoch=T, agent=algo._actor_critic, buffer=algo._buf, logger=algo._logger)
    data = {k: v.numpy().copy() for k, v in algo._buf.get().items()}
    lam0 = float(algo._lagrange.lagrangian_multiplier.item())
    Jc = algo._logger.get_stats('Metrics/EpCost')[0]
    _, perms = _record_randperm(algo._update)
    B = data['obs'].shape[0]
    perms = np.stack([p.numpy() for p in perms if p.numel() == B])
    lg = algo._logger
    np.savez(os.path.join(OUT, 'update_focops.npz'), N=N, T=T, O=O, A=A, theta0=theta0,
             theta1=_flat_theta(algo._actor_critic), lam0=lam0,
             lam1=float(algo._lagrange.lagrangian_multiplier.item()), Jc=Jc, perms=perms, batch_size=32,
             update_iters=2, cost_limit=25.0, lambda_lr=0.035, upper_bound=2.0, focops_lam=1.5, focops_eta=0.02,
             kl=_last(lg, 'Train/KL'), stop_iter=_last(lg, 'Train/StopIter'), loss_pi=_last(lg, 'Loss/Loss_pi'),
             **{'data_' + k: v for k, v in data.items()})


def gen_update_p3o():
    """P3O._update of the unmodified reference with the relu penalty active (cost_limit below Jc)."""
    N, T, O, A, seed = 8, 24, 12, 3, 13
    algo = _build_algo('P3O', N, T, O, A, seed, extra_algo={'cost_limit': 1.0, 'kappa': 0.5}, tmax=8, term_prob=0.05)
    theta0 = _flat_theta(algo._actor_critic)
    algo._env.rollout(steps_per_epoch=T, agent=algo._actor_critic, buffer=algo._buf, logger=algo._logger)
    data = {k: v.numpy().copy() for k, v in algo._buf.get().items()}
    Jc = algo._logger.get_stats('Metrics/EpCost')[0]
    _, perms = _record_randperm(algo._update)
    B = data['obs'].shape[0]
    perms = np.stack([p.numpy() for p in perms if p.numel() == B])
    lg = algo._logger
    np.savez(os.path.join(OUT, 'update_p3o.npz'), N=N, T=T, O=O, A=A, theta0=theta0,
             theta1=_flat_theta(algo._actor_critic), Jc=Jc, perms=perms, batch_size=32, update_iters=2,
             cost_limit=1.0, kappa=0.5, kl=_last(lg, 'Train/KL'), stop_iter=_last(lg, 'Train/StopIter'),
             loss_pi=_last(lg, 'Loss/Loss_pi'), loss_pi_cost=_last(lg, 'Loss/Loss_pi_cost'),
             **{'data_' + k: v for k, v in data.items()})


def gen_cpo(name='CPO', fname='update_cpo.npz', seed=7, cost_limit=2.0):
    """CPO / PCPO: Fisher-vector product, CG solve and one full actor+critic update of the reference."""
    from omnisafe.utils.math import conjugate_gradients as ref_cg

    N, T, O, A = 8, 24, 12, 3
    algo = _build_algo(name, N, T, O, A, seed, extra_algo={'cost_limit': cost_limit}, tmax=8, term_prob=0.05)
    theta0 = _flat_theta(algo._actor_critic)
    algo._env.rollout(steps_per_epoch=T, agent=algo._actor_critic, buffer=algo._buf, logger=algo._logger)
    data = {k: v.numpy().copy() for k, v in algo._buf.get().items()}
    obs = torch.as_tensor(data['obs'])
    algo._fvp_obs = obs[:: algo._cfgs.algo_cfgs.fvp_sample_freq]
    g = torch.Generator().manual_seed(0)
    P = sum(p.numel() for p in algo._actor_critic.actor.parameters())
    vec = torch.randn(P, generator=g) * 0.1
    fv = algo._fvp(vec).detach().numpy().copy()
    bvec = torch.randn(P, generator=g) * 0.01
    xcg = ref_cg(algo._fvp, bvec, algo._cfgs.algo_cfgs.cg_iters).detach().numpy().copy()
    ep_cost = algo._logger.get_stats('Metrics/EpCost')[0]
    _, perms = _record_randperm(algo._update)
    B = data['obs'].shape[0]
    perms = np.stack([p.numpy() for p in perms if p.numel() == B])
    lg = algo._logger
    misc = {k.split('/')[1]: _last(lg, k) for k in (
        'Misc/AcceptanceStep', 'Misc/Alpha', 'Misc/FinalStepNorm', 'Misc/xHx', 'Misc/H_inv_g',
        'Misc/gradient_norm', 'Misc/cost_gradient_norm', 'Misc/Lambda_star', 'Misc/Nu_star',
        'Misc/OptimCase', 'Misc/A', 'Misc/B', 'Misc/q', 'Misc/r', 'Misc/s')}
    np.savez(os.path.join(OUT, fname), N=N, T=T, O=O, A=A, seed=seed, theta0=theta0,
             theta1=_flat_theta(algo._actor_critic), vec=vec.numpy(), fvp=fv, bvec=bvec.numpy(), xcg=xcg,
             ep_cost=ep_cost, cost_limit=cost_limit, perms=perms, batch_size=32, update_iters=2,
             cg_damping=0.1, cg_iters=15, target_kl=0.01, kl=_last(lg, 'Train/KL'),
             **{'misc_' + k: v for k, v in misc.items()}, **{'data_' + k: v for k, v in data.items()})


def gen_update_trpo(name='TRPOLag', fname='update_trpolag.npz', seed=17, extra=None, lagrange=None):
    """TRPOLag / OnCRPO / RCPO ._update of the unmodified reference: natural direction, line search, critics."""
    N, T, O, A = 8, 24, 12, 3
    algo = _build_algo(name, N, T, O, A, seed, extra_algo=extra, tmax=8, term_prob=0.05, extra_lagrange=lagrange)
    theta0 = _flat_theta(algo._actor_critic)
    algo._env.rollout(steps_per_epoch=T, agent=algo._actor_critic, buffer=algo._buf, logger=algo._logger)
    data = {k: v.numpy().copy() for k, v in algo._buf.get().items()}
    ep_cost = algo._logger.get_stats('Metrics/EpCost')[0]
    lam0 = float(algo._lagrange.lagrangian_multiplier.item()) if hasattr(algo, '_lagrange') else 0.0
    _, perms = _record_randperm(algo._update)
    lam1 = float(algo._lagrange.lagrangian_multiplier.item()) if hasattr(algo, '_lagrange') else 0.0
    B = data['obs'].shape[0]
    perms = np.stack([p.numpy() for p in perms if p.numel() == B])
    lg = algo._logger
    keys = ['Misc/Alpha', 'Misc/FinalStepNorm', 'Misc/xHx', 'Misc/H_inv_g', 'Misc/gradient_norm']
    if name != 'RCPO':
        keys.append('Misc/AcceptanceStep')
    misc = {k.split('/')[1]: _last(lg, k) for k in keys}
    np.savez(os.path.join(OUT, fname), name=name, N=N, T=T, O=O, A=A, seed=seed, theta0=theta0,
             theta1=_flat_theta(algo._actor_critic), ep_cost=ep_cost, lam0=lam0, lam1=lam1, perms=perms, batch_size=32,
             update_iters=2, kl=_last(lg, 'Train/KL'), **{'extra_' + k: v for k, v in (extra or {}).items()},
             **{'lagrange_' + k: v for k, v in (lagrange or {}).items()},
             **{'misc_' + k: v for k, v in misc.items()}, **{'data_' + k: v for k, v in data.items()})


def gen_update_first_order(name, fname, seed, extra=None, lagrange=None):
    """IPO / CPPOPID / PDO ._update of the unmodified reference (penalty resp. multiplier from the logged Jc)."""
    N, T, O, A = 8, 24, 12, 3
    algo = _build_algo(name, N, T, O, A, seed, extra_algo=extra, tmax=8, term_prob=0.05, extra_lagrange=lagrange)
    theta0 = _flat_theta(algo._actor_critic)
    algo._env.rollout(steps_per_epoch=T, agent=algo._actor_critic, buffer=algo._buf, logger=algo._logger)
    data = {k: v.numpy().copy() for k, v in algo._buf.get().items()}
    ep_cost = algo._logger.get_stats('Metrics/EpCost')[0]
    _, perms = _record_randperm(algo._update)
    B = data['obs'].shape[0]
    perms = np.stack([p.numpy() for p in perms if p.numel() == B])
    lg = algo._logger
    if name == 'IPO':
        lam1 = float(_last(lg, 'Misc/Penalty')[-1])
    else:
        lam = algo._lagrange.lagrangian_multiplier
        lam1 = float(lam.item() if hasattr(lam, 'item') else lam)
    np.savez(os.path.join(OUT, fname), name=name, N=N, T=T, O=O, A=A, seed=seed, theta0=theta0,
             theta1=_flat_theta(algo._actor_critic), ep_cost=ep_cost, lam1=lam1, perms=perms, batch_size=32,
             update_iters=2, kl=_last(lg, 'Train/KL'), stop_iter=_last(lg, 'Train/StopIter'),
             loss_pi=_last(lg, 'Loss/Loss_pi'), **{'extra_' + k: v for k, v in (extra or {}).items()},
             **{'lagrange_' + k: v for k, v in (lagrange or {}).items()}, **{'data_' + k: v for k, v in data.items()})


def gen_simmer():
    """SimmerAdapter.control_budget + SimmerPIDAgent (adapter/simmer_adapter.py:L113-131, common/simmer_agent.py:L91-170)
    over a cost sequence, two controller settings; then one PPOSimmerPID rollout whose epoch starts at the
    controlled relative budget."""
    from omnisafe.common.simmer_agent import SimmerPIDAgent
    from omnisafe.utils.config import Config

    rng = np.random.default_rng(8)
    costs = np.concatenate([np.linspace(0, 40, 20), 25 + 10 * rng.standard_normal(30), np.linspace(40, 0, 20)]).astype(np.float32)
    out = {'costs': costs}
    gam, L = 0.999, 1000
    scale = (1 - gam ** L) / (1 - gam) / L
    for i, (kp, ki, kd, polyak) in enumerate([(0.0005, 0.00001, 0.0, 0.995), (0.05, 0.002, 0.3, 0.5)]):
        bound = torch.ones(3, 1) * 25.0 * scale
        agent = SimmerPIDAgent(cfgs=Config.dict2config({'kp': kp, 'ki': ki, 'kd': kd, 'polyak': polyak}), budget_bound=bound)
        budget = torch.ones(3, 1) * 15.0 * scale
        hist = []
        for c in costs:
            budget = agent.act(safety_budget=budget, observation=torch.as_tensor(c) * scale)
            hist.append(budget.numpy().copy())
        out[f'budget_{i}'] = np.stack(hist)
        out[f'cfg_{i}'] = np.array([kp, ki, kd, polyak])
    out['scale'] = scale
    np.savez(os.path.join(OUT, 'simmer_controller.npz'), **out)


def gen_update_cup():
    """CUP._update of the unmodified reference: the PPO stage on adv_r, then the cost-projection stage."""
    N, T, O, A, seed = 8, 24, 12, 3, 29
    algo = _build_algo('CUP', N, T, O, A, seed, tmax=8, term_prob=0.05, extra_lagrange={'cost_limit': 1.0})
    theta0 = _flat_theta(algo._actor_critic)
    algo._env.rollout(steps_per_epoch=T, agent=algo._actor_critic, buffer=algo._buf, logger=algo._logger)
    data = {k: v.numpy().copy() for k, v in algo._buf.get().items()}
    algo._buf.get = lambda: {k: torch.as_tensor(v) for k, v in data.items()}      # CUP reads the buffer twice
    Jc = algo._logger.get_stats('Metrics/EpCost')[0]
    _, perms = _record_randperm(algo._update)
    B = data['obs'].shape[0]
    perms = np.stack([p.numpy() for p in perms if p.numel() == B])
    lg = algo._logger
    np.savez(os.path.join(OUT, 'update_cup.npz'), N=N, T=T, O=O, A=A, seed=seed, theta0=theta0,
             theta1=_flat_theta(algo._actor_critic), Jc=Jc, lam1=float(algo._lagrange.lagrangian_multiplier.item()),
             perms=perms, batch_size=32, update_iters=2, cost_limit=1.0, stop_iter=_last(lg, 'Train/StopIter'),
             second_stop_iter=_last(lg, 'Train/SecondStepStopIter'), kl=_last(lg, 'Train/KL'),
             **{'data_' + k: v for k, v in data.items()})


def gen_pid():
    """PIDLagrangian.pid_update (common/pid_lagrange.py:L95-125) over cost sequences that exercise the
    integral clamp, the delayed derivative (deque roll-over) and the three normalisation modes."""
    from omnisafe.common.pid_lagrange import PIDLagrangian

    rng = np.random.default_rng(5)
    base = dict(pid_kp=0.1, pid_ki=0.01, pid_kd=0.01, pid_d_delay=10, pid_delta_p_ema_alpha=0.95,
                pid_delta_d_ema_alpha=0.95, sum_norm=True, diff_norm=False, penalty_max=100.0,
                lagrangian_multiplier_init=0.001, cost_limit=25.0)
    cfgs = [base, dict(base, diff_norm=True), dict(base, sum_norm=False, penalty_max=0.3, pid_kd=0.5, pid_d_delay=3),
            dict(base, pid_d_delay=1, pid_ki=0.1)]
    costs = np.concatenate([np.linspace(0, 60, 25), 25 + 20 * rng.standard_normal(40), np.linspace(60, 0, 25)])
    out = {'costs': costs, 'n_cfgs': len(cfgs)}
    for i, cfg in enumerate(cfgs):
        pid = PIDLagrangian(**cfg)
        lam = []
        for c in costs:
            pid.pid_update(float(c))
            lam.append(pid.lagrangian_multiplier)
        out[f'lam_{i}'] = np.asarray(lam, np.float64)
        for k, v in cfg.items():
            out[f'cfg_{i}_{k}'] = v
    np.savez(os.path.join(OUT, 'pid_lagrange.npz'), **out)


if __name__ == '__main__':
    torch.set_num_threads(1)
    gen_discount_cumsum()
    gen_buffer_gae()
    gen_buffer_gae('gae-rtg', 'buffer_gae_rtg.npz', seed=12)
    gen_buffer_gae('plain', 'buffer_plain.npz', seed=13)
    gen_buffer_gae('vtrace', 'buffer_vtrace.npz', seed=14)
    gen_normalizer()
    algo = gen_rollout()
    gen_update_ppolag(algo)
    gen_rollout('PDO', 'rollout_pdo.npz', seed=9, epochs_rolled=2)
    gen_rollout('PPOSaute', 'rollout_pposaute.npz', seed=31,
                extra_algo={'safety_budget': 2.0, 'saute_gamma': 0.9, 'max_ep_len': 8, 'unsafe_reward': -0.5})
    gen_update_focops()
    gen_update_p3o()
    gen_cpo()
    gen_cpo('PCPO', 'update_pcpo.npz', seed=11, cost_limit=1.0)
    gen_update_trpo('TRPOLag', 'update_trpolag.npz', seed=17)
    gen_update_trpo('OnCRPO', 'update_oncrpo.npz', seed=19, extra={'cost_limit': 1.0, 'distance': 0.5})
    gen_update_trpo('RCPO', 'update_rcpo.npz', seed=27, lagrange={'cost_limit': 1.0})
    gen_update_first_order('IPO', 'update_ipo.npz', 21, {'cost_limit': 6.0, 'kappa': 0.5})
    gen_update_first_order('CPPOPID', 'update_cppopid.npz', 22, lagrange={'cost_limit': 1.0})
    gen_update_first_order('PDO', 'update_pdo.npz', 23, lagrange={'cost_limit': 1.0})
    gen_update_cup()
    gen_simmer()
    gen_rollout('PPOSimmerPID', 'rollout_pposimmer.npz', seed=33,
                extra_algo={'safety_budget': 1.0, 'upper_budget': 2.0, 'saute_gamma': 0.9, 'max_ep_len': 8, 'unsafe_reward': -0.5})
    gen_pid()
    print('golden fixtures written to', OUT)
