"""Oracle: one epoch of OnPolicyAdapter.rollout on the synthetic env (numpy + torch-CPU).

TEST INFRASTRUCTURE ONLY.  Follows adapter/onpolicy_adapter.py:L58-136 (rollout loop, per-env
done handling, bootstrap rules), envs/wrapper.py:L231-241 (ObsNormalize.step: final observations
are pushed/normalised before the batch), envs/wrapper.py:L510-514 (ActionScale) and
common/buffer/vector_onpolicy_buffer.py:L96-99 (store), vectorised over envs.
"""
from __future__ import annotations

import numpy as np

from oracle import actor_critic as ac
from oracle.gae import FLAG_TERMINATED, FLAG_TRUNCATED

F32 = np.float32


def action_scale(act, lo=-1.0, hi=1.0, mn=-1.0, mx=1.0):
    """old_min + (old_max - old_min) * (action - min) / (max - min), fp32 left to right."""
    act = np.asarray(act, F32)
    t = (act - F32(mn)).astype(F32)
    t = (F32(hi - lo) * t).astype(F32)
    t = (t / F32(mx - mn)).astype(F32)
    return (F32(lo) + t).astype(F32)


def rollout_epoch(env, norm, theta, T, eps, obs_normalize=True, window=None, ep_state=None, saute=None, early=None):
    """Returns time-major slabs (dict of [T, N, ...] float32 arrays + uint8 flags).

    `eps` is the [T, N, A] standard-normal stream; `window` (list) receives (EpRet, EpCost, EpLen)
    of finished episodes in (step, env) order, like the reference Logger deque.

    `saute` = {'budget': per-step safety budget, 'gamma': saute_gamma, 'unsafe_reward': r} switches on the
    SauteAdapter semantics (adapter/saute_adapter.py:L135-217): the network sees [normalised obs | safety
    state z], z starts at 1, z <- (z - cost / budget) / gamma after every step, the stored reward becomes
    `unsafe_reward` once z <= 0, z returns to 1 when the episode ends (so final observations carry z = 1,
    as in the reference); episode returns keep the original reward.  'z0' (Simmer, adapter/simmer_adapter.py:L97-111)
    is the value z takes at the epoch's reset, the relative safety budget; episode ends still reset it to 1.  `theta` is then sized for O + 1 inputs.

    `early` = {'cost_limit': c, 'acc': float32 array [N]} switches on EarlyTerminatedAdapter.step
    (adapter/early_terminated_adapter.py:L56-98; upstream single env, here per env): the accumulated cost -- never cleared by
    ordinary episode ends -- exceeding the limit stores reward 0, terminates the episode and resets the env (a second reset
    when the env's own episode ended in the same step); the normaliser sees the pre-reset state and then the reset one."""
    N, O, A = env.N, env.O, env.A
    On = O + (1 if saute else 0)                     # network input width
    z = np.full((N, 1), F32(saute.get('z0', 1.0)) if saute else F32(1), F32)   # Simmer starts an epoch at the relative budget
    aug = (lambda x, zz: np.concatenate([x, zz], axis=-1).astype(F32)) if saute else (lambda x, zz: x)
    sl = {
        'obs': np.zeros((T, N, On), F32), 'act': np.zeros((T, N, A), F32),
        'logp': np.zeros((T, N), F32), 'rew': np.zeros((T, N), F32), 'cost': np.zeros((T, N), F32),
        'val_r': np.zeros((T, N), F32), 'val_c': np.zeros((T, N), F32),
        'boot_r': np.zeros((T, N), F32), 'boot_c': np.zeros((T, N), F32),
        'flags': np.zeros((T, N), np.uint8),
    }
    ep_ret = np.zeros(N, F32); ep_cost = np.zeros(N, F32); ep_len = np.zeros(N, F32)
    raw = env.reset()
    obs = aug(norm.normalize(raw) if obs_normalize else raw, z)
    for t in range(T):
        act, v_r, v_c, logp = ac.step(theta, obs, eps[t], On, A)
        nraw, rew, cost, term, trunc, final_raw, fin = env.step(action_scale(act))
        if early is not None:
            early['acc'] = (early['acc'] + cost).astype(F32)
            ex = early['acc'] > F32(early['cost_limit'])
            if ex.any():
                rew = np.where(ex, F32(0), rew).astype(F32)
                only = ex & ~fin                                  # the env itself did not end: its next obs is the pre-reset state
                final_raw = np.where(only[:, None], nraw, final_raw).astype(F32)
                with np.errstate(over='ignore'):
                    env.episode = np.where(ex, env.episode + np.uint32(1), env.episode).astype(np.uint32)
                env.ep_step = np.where(ex, 0, env.ep_step).astype(np.int32)
                env.s = np.where(ex[:, None], env._reset_values(env.episode), env.s).astype(F32)
                nraw = env.s.copy()
                term = term | ex
                fin = fin | ex
                early['acc'] = np.where(ex, F32(0), early['acc']).astype(F32)
        final_norm = np.zeros((N, O), F32)
        if fin.any():
            final_norm[fin] = norm.normalize(final_raw[fin]) if obs_normalize else final_raw[fin]
        nobs = norm.normalize(nraw) if obs_normalize else nraw
        ep_ret = (ep_ret + rew).astype(F32); ep_cost = (ep_cost + cost).astype(F32); ep_len += 1
        if saute:
            z = ((z - (cost[:, None] / F32(saute['budget'])).astype(F32)).astype(F32) / F32(saute['gamma'])).astype(F32)
            safe = (z[:, 0] > 0).astype(F32)
            rew = (safe * rew + (F32(1) - safe) * F32(saute['unsafe_reward'])).astype(F32)
            done = (term | trunc).astype(F32)[:, None]
            z = (z * (F32(1) - done) + done).astype(F32)
        final_norm = aug(final_norm, z)
        nobs = aug(nobs, z)
        sl['obs'][t] = obs; sl['act'][t] = act; sl['logp'][t] = logp
        sl['rew'][t] = rew; sl['cost'][t] = cost; sl['val_r'][t] = v_r; sl['val_c'][t] = v_c
        sl['flags'][t] = term.astype(np.uint8) * FLAG_TERMINATED + trunc.astype(np.uint8) * FLAG_TRUNCATED
        obs = nobs
        epoch_end = t == T - 1
        need_final = trunc & ~term
        need_next = (~term) & (~trunc) & epoch_end
        if need_final.any():
            br, bc = ac.values(theta, final_norm, On, A)
            sl['boot_r'][t][need_final] = br[need_final]; sl['boot_c'][t][need_final] = bc[need_final]
        if np.any(need_next):
            br, bc = ac.values(theta, obs, On, A)
            sl['boot_r'][t][need_next] = br[need_next]; sl['boot_c'][t][need_next] = bc[need_next]
        for i in np.nonzero(fin)[0]:
            if window is not None:
                window.append((float(ep_ret[i]), float(ep_cost[i]), float(ep_len[i])))
        ep_ret[fin] = 0; ep_cost[fin] = 0; ep_len[fin] = 0
    return sl


def normalize_rows(norm, slab):
    """RewardNormalize / CostNormalize (envs/wrapper.py:L280-423) over a finished [T, N] slab: row t is
    pushed into `norm` (oracle Normalizer with shape ()) and normalised, step after step as the wrapper
    does inside `env.step`."""
    import numpy as np
    return np.stack([norm.normalize(slab[t]) for t in range(slab.shape[0])]).astype(np.float32)
