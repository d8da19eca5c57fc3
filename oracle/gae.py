"""Oracle: discounted cumulative sums, dual GAE and advantage statistics (numpy).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np

FLAG_TERMINATED = 1
FLAG_TRUNCATED = 2


def discount_cumsum(x, discount: float) -> np.ndarray:
    """y_t = x_t + discount * y_{t+1}, carried in float64.

    Follows omnisafe/utils/math.py:L59-82: the input is upcast to float64, `discount` is a python
    double, multiply and add are rounded separately.
    """
    y = np.array(x, dtype=np.float64, copy=True)
    cum = y[-1] if len(y) else 0.0
    for i in range(len(y) - 2, -1, -1):
        cum = y[i] + np.float64(discount) * cum
        y[i] = cum
    return y


def finish_path(rew, cost, val_r, val_c, last_r, last_c, gamma, lam, lam_c, pen=0.0, estimator='gae'):
    """One path of OnPolicyBuffer.finish_path (onpolicy_buffer.py:L148-203; estimators 'gae' L299-303,
    'gae-rtg' L305-310, 'vtrace' L312-326, 'plain' L328-331).

    Inputs are float32 1-D arrays of one path; returns float32 (adv_r, tv_r, adv_c, tv_c, ret).
    """
    f32 = np.float32
    rewards = np.concatenate([np.asarray(rew, f32), np.asarray([last_r], f32)])
    values_r = np.concatenate([np.asarray(val_r, f32), np.asarray([last_r], f32)])
    costs = np.concatenate([np.asarray(cost, f32), np.asarray([last_c], f32)])
    values_c = np.concatenate([np.asarray(val_c, f32), np.asarray([last_c], f32)])
    ret = discount_cumsum(rewards, gamma)[:-1].astype(f32)
    rewards = (rewards - f32(pen) * costs).astype(f32)

    def vtrace(values, rews):
        """_calculate_v_trace (onpolicy_buffer.py:L338-405) with policy == behaviour probabilities, so
        rho = c = 1 exactly (L312-326): sequential fp32 arithmetic as torch executes it."""
        n = len(rews) - 1
        v_s = values[:-1].copy()
        last = values[-1]
        g32 = f32(gamma)
        for i in range(n - 1, -1, -1):
            delta = f32(f32(rews[i] + f32(g32 * values[i + 1])) - values[i])
            v_s[i] = f32(v_s[i] + f32(delta + f32(g32 * f32(last - values[i + 1]))))
            last = v_s[i]
        v_next = np.concatenate([v_s[1:], values[-1:]])
        adv = ((rews[:-1] + (g32 * v_next).astype(f32)).astype(f32) - values[:-1]).astype(f32)
        return adv, v_s

    def adv_and_target(values, rews, lam_):
        if estimator == 'vtrace':
            return vtrace(values, rews)
        deltas = ((rews[:-1] + f32(gamma) * values[1:]).astype(f32) - values[:-1]).astype(f32)
        if estimator == 'plain':
            return deltas, discount_cumsum(rews, gamma)[:-1].astype(f32)
        adv = discount_cumsum(deltas, gamma * lam_)
        if estimator == 'gae-rtg':
            return adv.astype(f32), discount_cumsum(rews, gamma)[:-1].astype(f32)
        assert estimator == 'gae'
        target = adv + values[:-1].astype(np.float64)
        return adv.astype(f32), target.astype(f32)

    adv_r, tv_r = adv_and_target(values_r, rewards, lam)
    adv_c, tv_c = adv_and_target(values_c, costs, lam_c)
    return adv_r, tv_r, adv_c, tv_c, ret


def dual_gae_slab(rew, cost, val_r, val_c, flags, boot_r, boot_c, gamma, lam, lam_c, pen=0.0):
    """Dual GAE over time-major [T, N] slabs, vectorised over envs, sequential over time.

    Equivalent to calling finish_path for every path (a path ends where flags != 0 or at t == T-1;
    bootstrap 0 if terminated else boot[t, i]) -- the arithmetic per element is identical, so the
    result is bit-identical to the per-path restatement (tests check this).
    Returns dict of float32 [T, N]: adv_r, adv_c, tv_r, tv_c, disc_ret.
    """
    f32, f64 = np.float32, np.float64
    rew = np.asarray(rew, f32); cost = np.asarray(cost, f32)
    val_r = np.asarray(val_r, f32); val_c = np.asarray(val_c, f32)
    flags = np.asarray(flags, np.uint8)
    T, N = rew.shape
    g, glr, glc = f64(gamma), f64(gamma * lam), f64(gamma * lam_c)
    adv_r = np.zeros((T, N), f32); adv_c = np.zeros((T, N), f32)
    tv_r = np.zeros((T, N), f32); tv_c = np.zeros((T, N), f32)
    ret = np.zeros((T, N), f32)
    Ar = np.zeros(N, f64); Ac = np.zeros(N, f64); Ag = np.zeros(N, f64)
    for t in range(T - 1, -1, -1):
        end = (flags[t] != 0) | (t == T - 1)
        term = (flags[t] & FLAG_TERMINATED) != 0
        if t + 1 < T:
            nr_mid, nc_mid = val_r[t + 1], val_c[t + 1]
        else:
            nr_mid = nc_mid = np.zeros(N, f32)
        nr = np.where(end, np.where(term, f32(0), boot_r[t]), nr_mid).astype(f32)
        nc = np.where(end, np.where(term, f32(0), boot_c[t]), nc_mid).astype(f32)
        rp = (rew[t] - (f32(pen) * cost[t]).astype(f32)).astype(f32)
        dr = ((rp + (f32(gamma) * nr).astype(f32)).astype(f32) - val_r[t]).astype(f32)
        dc = ((cost[t] + (f32(gamma) * nc).astype(f32)).astype(f32) - val_c[t]).astype(f32)
        Ar = np.where(end, dr.astype(f64), dr.astype(f64) + glr * Ar)
        Ac = np.where(end, dc.astype(f64), dc.astype(f64) + glc * Ac)
        Ag = np.where(end, rew[t].astype(f64) + g * nr.astype(f64), rew[t].astype(f64) + g * Ag)
        adv_r[t] = Ar.astype(f32); adv_c[t] = Ac.astype(f32)
        tv_r[t] = (Ar + val_r[t].astype(f64)).astype(f32)
        tv_c[t] = (Ac + val_c[t].astype(f64)).astype(f32)
        ret[t] = Ag.astype(f32)
    return {'adv_r': adv_r, 'adv_c': adv_c, 'tv_r': tv_r, 'tv_c': tv_c, 'disc_ret': ret}


def dual_gae_per_path(rew, cost, val_r, val_c, flags, boot_r, boot_c, gamma, lam, lam_c, pen=0.0, estimator='gae'):
    """Per-path restatement: split every env column into paths and call finish_path on each (the form the
    reference executes; also the oracle of the 'gae-rtg' / 'plain' estimators)."""
    T, N = rew.shape
    out = {k: np.zeros((T, N), np.float32) for k in ('adv_r', 'adv_c', 'tv_r', 'tv_c', 'disc_ret')}
    for i in range(N):
        start = 0
        for t in range(T):
            if flags[t, i] != 0 or t == T - 1:
                term = (flags[t, i] & FLAG_TERMINATED) != 0
                lr = 0.0 if term else boot_r[t, i]
                lc = 0.0 if term else boot_c[t, i]
                sl = slice(start, t + 1)
                a_r, t_r, a_c, t_c, ret = finish_path(
                    rew[sl, i], cost[sl, i], val_r[sl, i], val_c[sl, i], lr, lc,
                    gamma, lam, lam_c, pen, estimator)
                out['adv_r'][sl, i] = a_r; out['tv_r'][sl, i] = t_r
                out['adv_c'][sl, i] = a_c; out['tv_c'][sl, i] = t_c
                out['disc_ret'][sl, i] = ret
                start = t + 1
    return out


def adv_statistics(adv_r, adv_c):
    """mean / population-std as dist_statistics_scalar does (utils/distributed.py:L382-388)."""
    f32 = np.float32
    x = np.asarray(adv_r, f32).reshape(-1)
    n = f32(x.size)
    mean = f32(x.sum(dtype=f32) / n)
    std = f32(np.sqrt(((x - mean) ** 2).sum(dtype=f32) / n))
    xc = np.asarray(adv_c, f32).reshape(-1)
    cmean = f32(xc.sum(dtype=f32) / f32(xc.size))
    return mean, std, cmean


def standardize(adv_r, adv_c, standardized_r=True, standardized_c=True):
    """VectorOnPolicyBuffer.get() epilogue (vector_onpolicy_buffer.py:L131-136)."""
    mean, std, cmean = adv_statistics(adv_r, adv_c)
    out_r = ((adv_r - mean) / (std + np.float32(1e-8))).astype(np.float32) if standardized_r else adv_r
    out_c = (adv_c - cmean).astype(np.float32) if standardized_c else adv_c
    return out_r, out_c
