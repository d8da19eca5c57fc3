"""Oracle: the synthetic HBM-resident Box CMDP, restated on the CPU (numpy, bit-exact spec).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference ships no multi-env synthetic environment (SURVEY.md §8c "Fixtures"); this module is
the *specification* of the env that `omnisafe_b200/csrc/rollout.cu` implements in-kernel.  It obeys
the reference's vector-env contract (omnisafe/envs/core.py:L37-182; N > 1 envs must auto-reset and
time-limit themselves and expose `final_observation`, envs/wrapper.py:L51,L130,L233).

Spec (all arithmetic fp32, every op rounded separately; integers are uint32 with wrap-around):
  hash      h = mix(seed ^ a*0x9E3779B1); h = mix(h ^ b*0x85EBCA77); h = mix(h ^ c*0xC2B2AE3D)
            mix = lowbias32
  reset     s_j = (hash(seed, gid, episode, j) >> 8) * 2^-23 - 1            (uniform in [-1, 1))
  step      a   = clip(action, -1, 1)
            s'_j = clip(0.95*s_j + 0.1*a[j mod A] + b_j, -10, 10),  b_j = 0.02*(((7j+3) mod 5) - 2)
            reward = 1 - (sum_j s'_j^2) / O   with the fixed tree: 8 interleaved partial sums
                     p_q = sum_{j = q (mod 8), ascending} s'_j^2, then ((p0+p1)+(p2+p3))+((p4+p5)+(p6+p7))
            cost   = 1 if s'_0 > cost_threshold else 0
            terminated = term_threshold != 0 and hash(seed ^ 0xA5A5A5A5, gid, gstep, 0xFFFF) < term_threshold
            truncated  = ep_step + 1 >= max_episode_steps
            on terminated|truncated: final_observation = s', episode += 1, s = reset(episode), ep_step = 0
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
F32 = np.float32


def _mix32(x):
    x = x.astype(U32)
    x = x ^ (x >> U32(16))
    x = x * U32(0x7FEB352D)
    x = x ^ (x >> U32(15))
    x = x * U32(0x846CA68B)
    x = x ^ (x >> U32(16))
    return x


def hash4(seed, a, b, c):
    with np.errstate(over='ignore'):
        seed = np.asarray(seed, U32); a = np.asarray(a, U32); b = np.asarray(b, U32); c = np.asarray(c, U32)
        h = _mix32(seed ^ (a * U32(0x9E3779B1)))
        h = _mix32(h ^ (b * U32(0x85EBCA77)))
        h = _mix32(h ^ (c * U32(0xC2B2AE3D)))
    return h


def u32_to_unit(h):
    return ((h >> U32(8)).astype(F32) * F32(1.0 / 8388608.0) - F32(1.0)).astype(F32)


def env_bias(obs_dim: int) -> np.ndarray:
    j = np.arange(obs_dim)
    k = ((7 * j + 3) % 5 - 2).astype(F32)
    return (F32(0.02) * k).astype(F32)


def term_threshold(term_prob: float) -> int:
    return int(min(max(term_prob, 0.0), 1.0) * 4294967296.0) & 0xFFFFFFFF if term_prob < 1.0 else 0xFFFFFFFF


class SyntheticBoxEnv:
    """N synthetic Box envs, obs in R^O, action in [-1, 1]^A."""

    def __init__(self, num_envs, obs_dim=60, act_dim=8, max_episode_steps=64, seed=0,
                 term_prob=0.0, env_id_offset=0, cost_threshold=0.0):
        self.N, self.O, self.A = int(num_envs), int(obs_dim), int(act_dim)
        self.max_episode_steps = int(max_episode_steps)
        self.seed = int(seed) & 0xFFFFFFFF
        self.term_threshold = term_threshold(term_prob)
        self.cost_threshold = F32(cost_threshold)
        self.gid = (np.arange(self.N, dtype=np.int64) + int(env_id_offset)).astype(U32)
        self.bias = env_bias(self.O)
        self.s = np.zeros((self.N, self.O), F32)
        self.ep_step = np.zeros(self.N, np.int32)
        self.episode = np.zeros(self.N, U32)
        self.gstep = np.zeros(self.N, U32)

    def _reset_values(self, episode):
        j = np.arange(self.O, dtype=U32)[None, :]
        return u32_to_unit(hash4(U32(self.seed), self.gid[:, None], episode[:, None], j))

    def reset(self):
        with np.errstate(over='ignore'):
            self.episode = (self.episode + U32(1)).astype(U32)
        self.ep_step[:] = 0
        self.s = self._reset_values(self.episode)
        return self.s.copy()

    def step(self, action):
        """Returns (next_obs, reward, cost, terminated, truncated, final_obs, finished_mask)."""
        a = np.clip(np.asarray(action, F32), F32(-1), F32(1)).astype(F32)
        idx = np.arange(self.O) % self.A
        t1 = (F32(0.95) * self.s).astype(F32)
        t2 = (F32(0.1) * a[:, idx]).astype(F32)
        sn = ((t1 + t2).astype(F32) + self.bias[None, :]).astype(F32)
        sn = np.clip(sn, F32(-10), F32(10)).astype(F32)
        sq = (sn * sn).astype(F32)
        parts = []
        for q in range(8):
            p = np.zeros(self.N, F32)
            for j in range(q, self.O, 8):
                p = (p + sq[:, j]).astype(F32)
            parts.append(p)
        t0 = (parts[0] + parts[1]).astype(F32); t1_ = (parts[2] + parts[3]).astype(F32)
        t2_ = (parts[4] + parts[5]).astype(F32); t3 = (parts[6] + parts[7]).astype(F32)
        tot = ((t0 + t1_).astype(F32) + (t2_ + t3).astype(F32)).astype(F32)
        reward = (F32(1.0) - (tot / F32(self.O)).astype(F32)).astype(F32)
        cost = (sn[:, 0] > self.cost_threshold).astype(F32)
        truncated = (self.ep_step + 1) >= self.max_episode_steps
        if self.term_threshold:
            h = hash4(U32(self.seed ^ 0xA5A5A5A5), self.gid, self.gstep, U32(0xFFFF))
            terminated = h < U32(self.term_threshold)
        else:
            terminated = np.zeros(self.N, bool)
        fin = terminated | truncated
        final_obs = sn.copy()
        with np.errstate(over='ignore'):
            self.gstep = (self.gstep + U32(1)).astype(U32)
            new_episode = np.where(fin, self.episode + U32(1), self.episode).astype(U32)
        reset_vals = self._reset_values(new_episode)
        self.s = np.where(fin[:, None], reset_vals, sn).astype(F32)
        self.episode = new_episode
        self.ep_step = np.where(fin, 0, self.ep_step + 1).astype(np.int32)
        return self.s.copy(), reward, cost, terminated, truncated, final_obs, fin
