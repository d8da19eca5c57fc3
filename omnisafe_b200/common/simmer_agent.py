"""Safety-budget controller of the Simmer adapter (omnisafe/common/simmer_agent.py:L98-186): a PID on the polyak-blurred
error `budget - cost` with the integral over the last 10 errors; fp32 CPU tensors as in the reference (the controller
acts once per epoch on one scalar per env row, so it stays on the host exactly where the reference runs it)."""
from __future__ import annotations

from collections import deque

import torch


class SimmerPIDAgent:
    def __init__(self, cfgs, budget_bound: torch.Tensor, action_space: tuple[float, float] = (-1.0, 1.0)) -> None:
        self._cfgs = cfgs
        self._budget_bound = budget_bound
        self._action_space = action_space
        self._prev_action = torch.zeros(1)
        self._prev_error = torch.zeros(1)
        self._prev_raw_action = torch.zeros(1)
        self._integral_history: deque = deque([], maxlen=10)

    def act(self, safety_budget: torch.Tensor, observation: torch.Tensor) -> torch.Tensor:
        """simmer_agent.py:L132-186 (get_greedy_action + act)."""
        c = self._cfgs
        current_error = safety_budget - observation
        blured_error = c.polyak * self._prev_error + (1 - c.polyak) * current_error
        self._integral_history.append(blured_error)
        sum_history = torch.as_tensor(sum(self._integral_history))
        raw_action = c.kp * blured_error + c.ki * sum_history + c.kd * (self._prev_action - self._prev_raw_action)
        action = torch.clamp(raw_action, min=self._action_space[0], max=self._action_space[1])
        next_safety_budget = torch.clamp(safety_budget + action, 1e-6 * torch.ones_like(safety_budget), self._budget_bound)
        action = next_safety_budget - safety_budget
        self._prev_action, self._prev_raw_action, self._prev_error = action, raw_action, blured_error
        return next_safety_budget
