"""Vector on-policy buffer as one set of time-major slabs in HBM.

Mirrors omnisafe/common/buffer/vector_onpolicy_buffer.py:L26-138 (and the per-env
OnPolicyBuffer it wraps, onpolicy_buffer.py:L134-238): the reference keeps a Python list of N
single-env buffers; here one [T][N] slab per field is appended by the fused rollout kernel, all
paths are finished at once by `finish_paths()` (= every finish_path call of an epoch, dual GAE
kernel) and `get()` returns the same dict of tensors in the reference's env-major sample order.
"""
from __future__ import annotations

import torch

from omnisafe_b200._lib import current_stream, lib, ptr


class VectorOnPolicyBuffer:
    def __init__(self, obs_dim: int, act_dim: int, size: int, gamma: float, lam: float, lam_c: float,
                 advantage_estimator: str = 'gae', penalty_coefficient: float = 0.0,
                 standardized_adv_r: bool = False, standardized_adv_c: bool = False,
                 num_envs: int = 1, device='cuda', keep_discounted_ret: bool = True) -> None:
        if num_envs < 1:
            raise ValueError('num_envs must be greater than 0.')
        assert advantage_estimator in ('gae', 'gae-rtg', 'vtrace', 'plain')      # onpolicy_buffer.py:L122
        self._estimator = {'gae': 0, 'gae-rtg': 1, 'plain': 2, 'vtrace': 3}[advantage_estimator]
        assert penalty_coefficient >= 0, 'penalty_coefficient must be non-negative!'
        T, N, O, A = int(size), int(num_envs), int(obs_dim), int(act_dim)
        dev = torch.device(device)
        self.T, self.N, self.O, self.A, self.device = T, N, O, A, dev
        self._gamma, self._lam, self._lam_c = float(gamma), float(lam), float(lam_c)
        self._penalty = float(penalty_coefficient)
        self._standardized_adv_r, self._standardized_adv_c = bool(standardized_adv_r), bool(standardized_adv_c)
        f32 = dict(dtype=torch.float32, device=dev)
        self.data = {
            'obs': torch.zeros(T, N, O, **f32), 'act': torch.zeros(T, N, A, **f32),
            'logp': torch.zeros(T, N, **f32), 'reward': torch.zeros(T, N, **f32),
            'cost': torch.zeros(T, N, **f32), 'value_r': torch.zeros(T, N, **f32),
            'value_c': torch.zeros(T, N, **f32), 'boot_r': torch.zeros(T, N, **f32),
            'boot_c': torch.zeros(T, N, **f32),
            'flags': torch.zeros(T, N, dtype=torch.uint8, device=dev),
            'epfin': torch.zeros(3, T, N, **f32),
            'adv_r': torch.zeros(T, N, **f32), 'adv_c': torch.zeros(T, N, **f32),
            'target_value_r': torch.zeros(T, N, **f32), 'target_value_c': torch.zeros(T, N, **f32),
        }
        self.data['discounted_ret'] = torch.zeros(T, N, **f32) if keep_discounted_ret else None
        self._ws = torch.zeros(lib().osb_gae_workspace_doubles(N), dtype=torch.float64, device=dev)
        self.adv_sums = torch.zeros(4, dtype=torch.float64, device=dev)
        self.adv_moments = torch.zeros(4, dtype=torch.float32, device=dev)

    @property
    def num_buffers(self) -> int:
        return self.N

    @property
    def standardized_adv_r(self) -> bool:
        return self._standardized_adv_r

    @property
    def standardized_adv_c(self) -> bool:
        return self._standardized_adv_c

    def slab_ptrs(self) -> list:
        d = self.data
        return [ptr(d[k]) for k in ('obs', 'act', 'logp', 'reward', 'cost', 'value_r', 'value_c',
                                    'boot_r', 'boot_c', 'flags', 'epfin')]

    def finish_paths(self) -> None:
        """All finish_path calls of one epoch (onpolicy_buffer.py:L148-203) in one launch."""
        d = self.data
        # the reward-to-go estimators share their scan with discounted_ret: identical when penalty == 0
        ret = d['discounted_ret'] if (self._estimator in (0, 3) or self._penalty == 0.0) else None
        lib().osb_adv_estimate(ptr(d['reward']), ptr(d['cost']), ptr(d['value_r']), ptr(d['value_c']),
                               ptr(d['flags']), ptr(d['boot_r']), ptr(d['boot_c']), self.T, self.N,
                               self._gamma, self._lam, self._lam_c, self._penalty, self._estimator,
                               ptr(d['adv_r']), ptr(d['adv_c']), ptr(d['target_value_r']),
                               ptr(d['target_value_c']), ptr(ret), ptr(self._ws),
                               ptr(self.adv_sums), current_stream())

    def finalize_statistics(self) -> None:
        """Turn the (already all-reduced) fp64 sums into the moments the update kernels consume
        (vector_onpolicy_buffer.py:L131-136)."""
        lib().osb_adv_moments(ptr(self.adv_sums), int(self._standardized_adv_r),
                              int(self._standardized_adv_c), ptr(self.adv_moments), current_stream())

    def get(self) -> dict[str, torch.Tensor]:
        """Reference-shaped view of the epoch: env-major [N*T, ...] tensors, advantages
        standardised (vector_onpolicy_buffer.py:L113-138).  The training kernels do NOT use this
        (they read the slabs in place); it exists for API compatibility and tests."""
        d = self.data
        n = self.T * self.N
        out_r = torch.empty(self.T, self.N, dtype=torch.float32, device=self.device)
        out_c = torch.empty(self.T, self.N, dtype=torch.float32, device=self.device)
        lib().osb_adv_standardize(ptr(d['adv_r']), ptr(d['adv_c']), ptr(self.adv_moments), n,
                                  ptr(out_r), ptr(out_c), current_stream())

        def em(x):  # time-major [T, N, ...] -> env-major [N*T, ...]
            return x.transpose(0, 1).reshape(n, *x.shape[2:]).contiguous()

        out = {'obs': em(d['obs']), 'act': em(d['act']), 'logp': em(d['logp']),
               'target_value_r': em(d['target_value_r']), 'target_value_c': em(d['target_value_c']),
               'adv_r': em(out_r), 'adv_c': em(out_c)}
        if d['discounted_ret'] is not None:
            out['discounted_ret'] = em(d['discounted_ret'])
        return out
