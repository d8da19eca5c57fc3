"""Lagrange multiplier with its Adam state on the device.

Mirrors omnisafe/common/lagrange.py:L25-136 (constructor arguments, `lagrangian_multiplier`,
`update_lagrange_multiplier`).  The step itself is `osb_lagrange_update` (csrc/optim.cu): Adam on a
scalar with gradient -(Jc - cost_limit), then projection onto [0, upper_bound].
"""
from __future__ import annotations

import torch

from omnisafe_b200._lib import current_stream, lib, ptr


class Lagrange:
    def __init__(self, cost_limit: float, lagrangian_multiplier_init: float, lambda_lr: float,
                 lambda_optimizer: str = 'Adam', lagrangian_upper_bound: float | None = None,
                 device='cuda') -> None:
        assert lambda_optimizer == 'Adam', (
            f'Optimizer={lambda_optimizer}: only Adam is implemented on the device path')
        self.cost_limit = float(cost_limit)
        self.lambda_lr = float(lambda_lr)
        self.lagrangian_upper_bound = lagrangian_upper_bound
        init_value = max(float(lagrangian_multiplier_init), 0.0)
        # state = {lambda, adam m, adam v, adam t}
        self.state = torch.tensor([init_value, 0.0, 0.0, 0.0], dtype=torch.float32, device=device)
        self.nan_flag = torch.zeros(1, dtype=torch.int32, device=device)

    @property
    def lagrangian_multiplier(self) -> torch.Tensor:
        return self.state[0]

    def update_lagrange_multiplier(self, window_sums: torch.Tensor) -> None:
        """`window_sums` = device fp64 {sum EpRet, sum EpCost, sum EpLen, count} of the episode
        window (already all-reduced): Jc = sum EpCost / count (common/lagrange.py:L114-136)."""
        ub = -1.0 if self.lagrangian_upper_bound is None else float(self.lagrangian_upper_bound)
        lib().osb_lagrange_update(ptr(window_sums), self.cost_limit, self.lambda_lr, ub,
                                  ptr(self.state), ptr(self.nan_flag), current_stream())
