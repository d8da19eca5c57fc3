"""Oracle: the learner side -- Lagrange multiplier, PPO-Lag / FOCOPS minibatch update, KL early
stop, Fisher-vector product, conjugate gradients and the CPO step logic (torch-CPU autograd).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates
  common/lagrange.py:L67-136, algorithms/on_policy/base/policy_gradient.py:L308-588,
  base/ppo.py:L35-87, naive_lagrange/ppo_lag.py:L52-102, first_order/focops.py:L62-230,
  base/natural_pg.py:L74-230, base/trpo.py:L56-222, second_order/cpo.py:L57-462,
  utils/math.py:L86-132, utils/tools.py:L35-129.
torch.optim.Adam / clip_grad_norm_ / autograd are the reference's own (installed) dependencies and
are called directly.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.distributions import Normal, kl_divergence
from torch.nn.utils.clip_grad import clip_grad_norm_

from oracle import actor_critic as ac

NETS = ('actor', 'reward_critic', 'cost_critic')


class Lagrange:
    """common/lagrange.py:L67-136 with lambda_optimizer='Adam'."""

    def __init__(self, cost_limit, lagrangian_multiplier_init, lambda_lr, upper_bound=None):
        self.cost_limit = cost_limit
        self.upper_bound = upper_bound
        self.lam = torch.nn.Parameter(torch.as_tensor(max(lagrangian_multiplier_init, 0.0)))
        self.opt = torch.optim.Adam([self.lam], lr=lambda_lr)

    def update(self, Jc: float) -> float:
        self.opt.zero_grad()
        loss = -self.lam * (Jc - self.cost_limit)
        loss.backward()
        self.opt.step()
        self.lam.data.clamp_(0.0, self.upper_bound)
        return float(self.lam.item())


class Learner:
    """Parameters as per-tensor leaves in the reference's named_parameters order + 3 Adam optimisers."""

    def __init__(self, theta, O, A, lr_actor=3e-4, lr_critic=3e-4):
        self.O, self.A = O, A
        lay = ac.layout(O, A)
        theta = torch.as_tensor(np.asarray(theta, np.float32)).clone()
        self.params = {}
        for net in NETS:
            self.params[net] = {name: theta[o:o + int(np.prod(shape))].view(*shape).clone().requires_grad_(True)
                                for name, (o, shape) in lay[net]['entries'].items()}
        self.opt = {
            'actor': torch.optim.Adam(list(self.params['actor'].values()), lr=lr_actor) if lr_actor is not None else None,
            'reward_critic': torch.optim.Adam(list(self.params['reward_critic'].values()), lr=lr_critic),
            'cost_critic': torch.optim.Adam(list(self.params['cost_critic'].values()), lr=lr_critic),
        }

    # -- flat views (utils/tools.py:L35-129) --------------------------------------------------
    def flat(self, net=None) -> np.ndarray:
        nets = NETS if net is None else (net,)
        return torch.cat([p.detach().reshape(-1) for n in nets for p in self.params[n].values()]).numpy().copy()

    def flat_grad(self, net) -> torch.Tensor:
        return torch.cat([p.grad.reshape(-1) for p in self.params[net].values()])

    def set_flat(self, net, vec):
        vec = torch.as_tensor(vec, dtype=torch.float32)
        i = 0
        for p in self.params[net].values():
            n = p.numel()
            p.data.copy_(vec[i:i + n].view_as(p))
            i += n

    def zero_grad(self, net):
        for p in self.params[net].values():
            p.grad = None

    def dist(self, obs) -> Normal:
        return ac.actor_dist(self.params['actor'], obs)

    # -- losses -------------------------------------------------------------------------------
    def loss_pi_ppo(self, obs, act, logp, adv, clip, entropy_coef=0.0):
        d = self.dist(obs)
        ratio = torch.exp(d.log_prob(act).sum(-1) - logp)
        rc = torch.clamp(ratio, 1 - clip, 1 + clip)
        loss = -torch.min(ratio * adv, rc * adv).mean()
        loss = loss - entropy_coef * d.entropy().mean()
        return loss, ratio

    def loss_pi_plain(self, obs, act, logp, adv):
        d = self.dist(obs)
        ratio = torch.exp(d.log_prob(act).sum(-1) - logp)
        return -(ratio * adv).mean()

    def loss_pi_cost(self, obs, act, logp, adv_c):
        d = self.dist(obs)
        ratio = torch.exp(d.log_prob(act).sum(-1) - logp)
        return (ratio * adv_c).mean()

    def loss_pi_p3o(self, obs, act, logp, adv_r, adv_c, clip, kappa, jc_minus_limit, entropy_coef=0.0):
        """P3O._update_actor (penalty_function/p3o.py:L48-125): PPO clipped surrogate on adv_r +
        kappa * relu(mean(ratio * adv_c) + Jc - cost_limit)."""
        loss_reward, ratio = self.loss_pi_ppo(obs, act, logp, adv_r, clip, entropy_coef)
        surr_cadv = (ratio * adv_c).mean()
        return loss_reward + kappa * torch.relu(surr_cadv + jc_minus_limit), ratio

    def loss_pi_cup(self, obs, act, logp, adv_c, old_mean, old_std, lam, coef):
        """CUP._loss_pi_cost (first_order/cup.py:L93-147): lambda * coef * ratio * adv_c + KL(new || old), mean over
        the [b, 1] tensor (the [b] surrogate broadcasts against the [b, 1] KL exactly as in the reference)."""
        d = self.dist(obs)
        ratio = torch.exp(d.log_prob(act).sum(-1) - logp)
        kl = kl_divergence(d, Normal(old_mean, old_std)).sum(-1, keepdim=True)
        return (lam * coef * ratio * adv_c + kl).mean()

    def update_cup_stage2(self, data, perms, lam, *, batch_size, gamma=0.99, lam_gae=0.95, max_grad_norm=40.0,
                          target_kl=0.02, kl_early_stop=True):
        """Second stage of CUP._update (cup.py:L163-215): actor-only minibatch steps on the cost projection loss,
        KL early stop against the policy the stage started from.  Returns the number of passes executed."""
        t = {k: torch.as_tensor(v) for k, v in data.items()}
        obs_all = t['obs']
        with torch.no_grad():
            old = self.dist(obs_all)
            old_mean, old_std = old.loc.clone(), old.scale.clone()
        old = Normal(old_mean, old_std)
        coef = (1 - gamma * lam_gae) / (1 - gamma)
        done = 0
        for perm in perms:
            perm = torch.as_tensor(np.asarray(perm, np.int64))
            for s in range(0, len(perm), batch_size):
                idx = perm[s:s + batch_size]
                loss = self.loss_pi_cup(obs_all[idx], t['act'][idx], t['logp'][idx], t['adv_c'][idx], old_mean[idx],
                                        old_std[idx], lam, coef)
                self._step('actor', loss, max_grad_norm)
            done += 1
            with torch.no_grad():
                kl = kl_divergence(old, self.dist(obs_all)).sum(-1, keepdim=True).mean().item()
            if kl_early_stop and kl > target_kl:
                break
        return done

    def loss_pi_focops(self, obs, act, logp, adv, old_mean, old_std, lam_f, eta, entropy_coef=0.0):
        # NB: as in the reference, kl is [b, 1] while ratio * adv is [b]: the difference broadcasts
        # to [b, b] before the mean (first_order/focops.py:L85-89).  Kept verbatim for parity.
        d = self.dist(obs)
        ratio = torch.exp(d.log_prob(act).sum(-1) - logp)
        kl = kl_divergence(d, Normal(old_mean, old_std)).sum(-1, keepdim=True)
        loss = (kl - (1 / lam_f) * ratio * adv) * (kl.detach() <= eta).type(torch.float32)
        return loss.mean() - entropy_coef * d.entropy().mean(), ratio

    # -- one optimiser step per network (policy_gradient.py:L407-524) --------------------------
    def _step(self, net, loss, max_grad_norm):
        self.opt[net].zero_grad()
        loss.backward()
        if max_grad_norm is not None:
            clip_grad_norm_(list(self.params[net].values()), max_grad_norm)
        self.opt[net].step()

    def critic_step(self, net, obs, target, critic_norm_coef, max_grad_norm):
        v = ac.critic_value(self.params[net], obs)
        loss = torch.nn.functional.mse_loss(v, target)
        if critic_norm_coef:
            for p in self.params[net].values():
                loss = loss + p.pow(2).sum() * critic_norm_coef
        self._step(net, loss, max_grad_norm)
        return float(loss.item())

    def update_ppo(self, data, perms, lam, *, batch_size, clip=0.2, entropy_coef=0.0, use_cost=True,
                   critic_norm_coef=0.001, max_grad_norm=40.0, target_kl=0.02, kl_early_stop=True,
                   focops=None, p3o=None, plain=False):
        """PolicyGradient._update / FOCOPS._update.  `data` holds env-major tensors as returned by
        VectorOnPolicyBuffer.get(); `perms[i]` is the sample order of pass i (DataLoader shuffle)."""
        t = {k: torch.as_tensor(v) for k, v in data.items()}
        obs_all = t['obs']
        with torch.no_grad():
            old = self.dist(obs_all)
            old_mean, old_std = old.loc.clone(), old.scale.clone()
        old = Normal(old_mean, old_std)
        stats = {'loss_pi': [], 'loss_r': [], 'loss_c': [], 'kl': [], 'iters': 0}
        for perm in perms:
            perm = torch.as_tensor(np.asarray(perm, np.int64))
            for s in range(0, len(perm), batch_size):
                idx = perm[s:s + batch_size]
                obs = obs_all[idx]
                stats['loss_r'].append(self.critic_step('reward_critic', obs, t['target_value_r'][idx], critic_norm_coef, max_grad_norm))
                if use_cost:
                    stats['loss_c'].append(self.critic_step('cost_critic', obs, t['target_value_c'][idx], critic_norm_coef, max_grad_norm))
                adv = (t['adv_r'][idx] - lam * t['adv_c'][idx]) / (1 + lam)
                if p3o is not None:
                    loss, _ = self.loss_pi_p3o(obs, t['act'][idx], t['logp'][idx], t['adv_r'][idx], t['adv_c'][idx],
                                               clip, p3o['kappa'], p3o['jc_minus_limit'], entropy_coef)
                elif plain:      # PolicyGradient._loss_pi (policy_gradient.py:L483-524): no clipping
                    loss = self.loss_pi_plain(obs, t['act'][idx], t['logp'][idx], adv)
                    if entropy_coef:
                        loss = loss - entropy_coef * self.dist(obs).entropy().mean()
                elif focops is None:
                    loss, _ = self.loss_pi_ppo(obs, t['act'][idx], t['logp'][idx], adv, clip, entropy_coef)
                else:
                    loss, _ = self.loss_pi_focops(obs, t['act'][idx], t['logp'][idx], adv, old_mean[idx],
                                                  old_std[idx], focops['lam'], focops['eta'], entropy_coef)
                self._step('actor', loss, max_grad_norm)
                stats['loss_pi'].append(float(loss.item()))
            with torch.no_grad():
                kl = kl_divergence(old, self.dist(obs_all)).sum(-1, keepdim=True).mean().item()
            stats['kl'].append(kl)
            stats['iters'] += 1
            if kl_early_stop and kl > target_kl:
                break
        return stats

    # -- natural-gradient machinery ---------------------------------------------------------------
    def fvp(self, vec, obs, damping):
        """NaturalPG._fvp: Hessian of mean KL(p || q) at q == p via double backward."""
        params = list(self.params['actor'].values())
        q = self.dist(obs)
        with torch.no_grad():
            pd = self.dist(obs)
            pd = Normal(pd.loc.clone(), pd.scale.clone())
        kl = kl_divergence(pd, q).mean()
        grads = torch.autograd.grad(kl, params, create_graph=True)
        flat = torch.cat([g.reshape(-1) for g in grads])
        kl_p = (flat * torch.as_tensor(vec, dtype=torch.float32)).sum()
        grads2 = torch.autograd.grad(kl_p, params)
        out = torch.cat([g.contiguous().reshape(-1) for g in grads2])
        return out + torch.as_tensor(vec, dtype=torch.float32) * damping


def conjugate_gradients(fisher_product, b, num_steps=10, residual_tol=1e-10, eps=1e-6):
    """utils/math.py:L86-132."""
    b = torch.as_tensor(b, dtype=torch.float32)
    x = torch.zeros_like(b)
    r = b - fisher_product(x)
    p = r.clone()
    rdotr = torch.dot(r, r)
    for _ in range(num_steps):
        z = fisher_product(p)
        alpha = rdotr / (torch.dot(p, z) + eps)
        x = x + alpha * p
        r = r - alpha * z
        new_rdotr = torch.dot(r, r)
        if torch.sqrt(new_rdotr) < residual_tol:
            break
        mu = new_rdotr / (rdotr + eps)
        p = r + mu * p
        rdotr = new_rdotr
    return x


def cpo_determine_case(b_dot_b, ep_costs, q, r, s, target_kl):
    """CPO._determine_case (second_order/cpo.py:L215-268) on python floats."""
    if b_dot_b <= 1e-6 and ep_costs < 0:
        return 4, 0.0, 0.0
    A = q - r ** 2 / (s + 1e-8)
    B = 2 * target_kl - ep_costs ** 2 / (s + 1e-8)
    if ep_costs < 0 and B < 0:
        case = 3
    elif ep_costs < 0 <= B:
        case = 2
    elif ep_costs >= 0 and B >= 0:
        case = 1
    else:
        case = 0
    return case, A, B


def cpo_step_direction(case, xHx, x, A, B, q, p, r, s, ep_costs, target_kl):
    """CPO._step_direction (second_order/cpo.py:L271-337); x, p are torch vectors."""
    if case in (3, 4):
        alpha = float(np.sqrt(2 * target_kl / (xHx + 1e-8)))
        return alpha * x, 1 / (alpha + 1e-8), 0.0
    if case in (1, 2):
        lam_a = float(np.sqrt(A / B))
        lam_b = float(np.sqrt(q / (2 * target_kl)))
        bound = r / (ep_costs + 1e-8)
        if ep_costs < 0:
            lam_a_star = float(np.clip(lam_a, 0.0, bound))
            lam_b_star = float(np.clip(lam_b, bound, np.inf))
        else:
            lam_a_star = float(np.clip(lam_a, bound, np.inf))
            lam_b_star = float(np.clip(lam_b, 0.0, bound))
        f_a = lambda lam: -0.5 * (A / (lam + 1e-8) + B * lam) - r * ep_costs / (s + 1e-8)
        f_b = lambda lam: -0.5 * (q / (lam + 1e-8) + 2 * target_kl * lam)
        lam_star = lam_a_star if f_a(lam_a_star) >= f_b(lam_b_star) else lam_b_star
        nu_star = max(lam_star * ep_costs - r, 0.0) / (s + 1e-8)
        return 1.0 / (lam_star + 1e-8) * (x - nu_star * p), lam_star, nu_star
    nu_star = float(np.sqrt(2 * target_kl / (s + 1e-8)))
    return -nu_star * p, 0.0, nu_star


def pcpo_step_direction(x_hx, hx, p, r, s, ep_costs, target_kl):
    """PCPO projection step (second_order/pcpo.py:L99-106): sqrt(2 delta / (q + 1e-8)) * H x
    - max(0, (sqrt(2 delta / q) * r + c) / s) * p, with q = xHx; hx, p torch vectors, scalars fp32."""
    q = torch.tensor(x_hx, dtype=torch.float32)
    r_, s_ = torch.tensor(r, dtype=torch.float32), torch.tensor(s, dtype=torch.float32)
    return (torch.sqrt(2 * target_kl / (q + 1e-8)) * hx
            - torch.clamp_min((torch.sqrt(2 * target_kl / q) * r_ + ep_costs) / s_, torch.tensor(0.0)) * p)


class PIDLagrangian:
    """PID controller of the multiplier (common/pid_lagrange.py:L54-125), Python floats throughout."""

    def __init__(self, pid_kp, pid_ki, pid_kd, pid_d_delay, pid_delta_p_ema_alpha, pid_delta_d_ema_alpha,
                 sum_norm, diff_norm, penalty_max, lagrangian_multiplier_init, cost_limit):
        self.kp, self.ki, self.kd, self.delay = pid_kp, pid_ki, pid_kd, int(pid_d_delay)
        self.a_p, self.a_d = pid_delta_p_ema_alpha, pid_delta_d_ema_alpha
        self.sum_norm, self.diff_norm, self.penalty_max = bool(sum_norm), bool(diff_norm), penalty_max
        self.cost_limit = cost_limit
        self.pid_i = float(lagrangian_multiplier_init)
        self.cost_ds = [0.0]          # deque(maxlen=delay): oldest first
        self.delta_p, self.cost_d, self.cost_penalty = 0.0, 0.0, 0.0

    @property
    def lagrangian_multiplier(self) -> float:
        return self.cost_penalty

    def pid_update(self, ep_cost_avg: float) -> float:
        delta = float(ep_cost_avg - self.cost_limit)
        self.pid_i = max(0.0, self.pid_i + delta * self.ki)
        if self.diff_norm:
            self.pid_i = max(0.0, min(1.0, self.pid_i))
        self.delta_p = self.delta_p * self.a_p + (1 - self.a_p) * delta
        self.cost_d = self.cost_d * self.a_d + (1 - self.a_d) * float(ep_cost_avg)
        pid_d = max(0.0, self.cost_d - self.cost_ds[0])
        pid_o = self.kp * self.delta_p + self.pid_i + self.kd * pid_d
        self.cost_penalty = max(0.0, pid_o)
        if self.diff_norm:
            self.cost_penalty = min(1.0, self.cost_penalty)
        if not (self.diff_norm or self.sum_norm):
            self.cost_penalty = min(self.cost_penalty, self.penalty_max)
        self.cost_ds.append(self.cost_d)
        if len(self.cost_ds) > self.delay:
            self.cost_ds.pop(0)
        return self.cost_penalty


def trpo_actor_step(L: 'Learner', obs, act, logp, adv, *, cost_surrogate=False, damping=0.1, cg_iters=15,
                    target_kl=0.01, total_steps=15, decay=0.8, search=True):
    """NaturalPG._update_actor direction (natural_pg.py:L146-166) + TRPO._search_step_size
    (trpo.py:L56-138; `search=False`: plain NaturalPG step) on the oracle Learner.  `cost_surrogate`: the loss is mean(ratio * adv) instead of
    -mean(ratio * adv) (OnCRPO's -adv_c surrogate, crpo.py:L55-80, passes adv = adv_c with this flag).
    Sets the actor parameters to theta_old + accepted step; returns (accept_step, final_kl, step, x, xHx, alpha)."""
    loss_fn = (lambda: L.loss_pi_cost(obs, act, logp, adv)) if cost_surrogate else (lambda: L.loss_pi_plain(obs, act, logp, adv))
    theta_old = torch.as_tensor(L.flat('actor')).clone()
    with torch.no_grad():
        p = L.dist(obs)
        p_dist = Normal(p.loc.clone(), p.scale.clone())
    L.zero_grad('actor')
    loss_before = loss_fn()
    loss_before.backward()
    grads = -L.flat_grad('actor')
    fvp = lambda v: L.fvp(v, obs, damping)   # noqa: E731
    x = conjugate_gradients(fvp, grads.numpy(), cg_iters)
    x_hx = float(x.dot(fvp(x)))
    alpha = float(np.sqrt(2 * target_kl / (x_hx + 1e-8)))
    step_direction = alpha * x
    if not search:      # NaturalPG / RCPO: the natural step is taken as is (natural_pg.py:L166-170)
        L.set_flat('actor', theta_old + step_direction)
        return 0, 0.0, step_direction, x, x_hx, alpha
    step_frac, accept, final_kl = 1.0, 0, 0.0
    for step in range(total_steps):
        L.set_flat('actor', theta_old + step_frac * step_direction)
        with torch.no_grad():
            loss = loss_fn()
            kl = float(kl_divergence(p_dist, L.dist(obs)).mean())
        improve = float(loss_before.detach() - loss)
        if np.isfinite(float(loss)) and improve >= 0 and kl <= target_kl:
            accept, final_kl = step + 1, kl
            break
        step_frac *= decay
    else:
        step_direction = torch.zeros_like(step_direction)
    step = step_frac * step_direction
    L.set_flat('actor', theta_old + step)
    return accept, final_kl, step, x, x_hx, alpha


class SimmerPIDAgent:
    """Safety-budget controller of the Simmer adapter (common/simmer_agent.py:L91-170): a PID on the polyak-
    blurred error budget - cost, integral over the last 10 errors, fp32 tensors of shape [N, 1]."""

    def __init__(self, kp, ki, kd, polyak, budget_bound, action_space=(-1.0, 1.0)):
        self.kp, self.ki, self.kd, self.polyak = kp, ki, kd, polyak
        self.bound = torch.as_tensor(budget_bound, dtype=torch.float32)
        self.lo, self.hi = action_space
        self.prev_action = torch.zeros(1)
        self.prev_error = torch.zeros(1)
        self.prev_raw_action = torch.zeros(1)
        self.integral = []            # deque(maxlen=10)

    def act(self, safety_budget: torch.Tensor, observation: torch.Tensor) -> torch.Tensor:
        current_error = safety_budget - observation
        blurred = self.polyak * self.prev_error + (1 - self.polyak) * current_error
        self.integral.append(blurred)
        if len(self.integral) > 10:
            self.integral.pop(0)
        sum_history = torch.as_tensor(sum(self.integral))
        raw = self.kp * blurred + self.ki * sum_history + self.kd * (self.prev_action - self.prev_raw_action)
        action = torch.clamp(raw, min=self.lo, max=self.hi)
        nxt = torch.clamp(safety_budget + action, 1e-6 * torch.ones_like(safety_budget), self.bound)
        self.prev_action, self.prev_raw_action, self.prev_error = nxt - safety_budget, raw, blurred
        return nxt


def simmer_control_budget(agent: SimmerPIDAgent, safety_budget, ep_cost, saute_gamma, max_ep_len):
    """SimmerAdapter.control_budget (adapter/simmer_adapter.py:L113-131): the episode cost is put on the
    per-step discounted scale of the budgets before the controller acts."""
    scale = (1 - saute_gamma ** max_ep_len) / (1 - saute_gamma) / max_ep_len
    obs = torch.as_tensor(ep_cost, dtype=torch.float32) * scale
    return agent.act(torch.as_tensor(safety_budget, dtype=torch.float32), obs)


def cpo_actor_step(L: 'Learner', obs, act, logp, adv_r, adv_c, ep_costs, *, damping=0.1, cg_iters=15, target_kl=0.01,
                   total_steps=20, decay=0.8):
    """CPO._update_actor (second_order/cpo.py:L340-462) on the oracle Learner: reward / cost surrogate gradients,
    two CG solves, case analysis, closed-form step and the cost-aware backtracking line search (cpo.py:L57-180).
    `ep_costs` = mean episode cost - cost_limit.  Returns (optim_case, acceptance_step)."""
    theta_old = torch.as_tensor(L.flat('actor')).clone()
    with torch.no_grad():
        pd = L.dist(obs)
        p_dist = Normal(pd.loc.clone(), pd.scale.clone())
    L.zero_grad('actor')
    loss_reward_before = L.loss_pi_plain(obs, act, logp, adv_r)
    loss_reward_before.backward()
    grads = -L.flat_grad('actor')
    fvp = lambda v: L.fvp(v, obs, damping)   # noqa: E731
    x = conjugate_gradients(fvp, grads.numpy(), cg_iters)
    xHx = float(x.dot(fvp(x)))
    L.zero_grad('actor')
    loss_cost_before = L.loss_pi_cost(obs, act, logp, adv_c)
    loss_cost_before.backward()
    b_grads = L.flat_grad('actor').clone()
    p = conjugate_gradients(fvp, b_grads.numpy(), cg_iters)
    q, r, s = xHx, float(grads.dot(p)), float(b_grads.dot(p))
    case, A, B = cpo_determine_case(float(b_grads.dot(b_grads)), ep_costs, q, r, s, target_kl)
    step_direction, _, _ = cpo_step_direction(case, xHx, x, A, B, q, p, r, s, ep_costs, target_kl)
    step_direction = torch.as_tensor(step_direction, dtype=torch.float32)
    step_frac, accept = 1.0, 0
    for step in range(total_steps):
        L.set_flat('actor', theta_old + step_frac * step_direction)
        with torch.no_grad():
            loss_r = L.loss_pi_plain(obs, act, logp, adv_r)
            loss_c = L.loss_pi_cost(obs, act, logp, adv_c)
            kl = float(kl_divergence(p_dist, L.dist(obs)).mean())
        improve = float(loss_reward_before.detach() - loss_r)
        cost_diff = float(loss_c - loss_cost_before.detach())
        if not np.isfinite(kl):
            continue
        if (case > 1 and improve < 0) or cost_diff > max(-ep_costs, 0) or kl > target_kl:
            step_frac *= decay
            continue
        accept = step + 1
        break
    else:
        step_direction = torch.zeros_like(step_direction)
    L.set_flat('actor', theta_old + step_frac * step_direction)
    return case, accept


def update_ppo_multirank(L: 'Learner', datas, perms_per_rank, lam, *, batch_size, clip=0.2, entropy_coef=0.0,
                         critic_norm_coef=0.001, max_grad_norm=40.0, target_kl=0.02, kl_early_stop=True):
    """PolicyGradient._update under `parallel = P` (policy_gradient.py:L345-405, L437-443; distributed.py:L193-198),
    simulated in ONE process: every rank has its own env-major data and DataLoader orders, the parameters are
    replicated.  Per minibatch step and network: each rank's gradient is clipped locally (clip_grad_norm_), the clipped
    gradients are averaged over the ranks, then ONE Adam step.  The early-stop KL is the rank average."""
    P = len(datas)
    ts = [{k: torch.as_tensor(v) for k, v in d.items()} for d in datas]
    olds = []
    with torch.no_grad():
        for t in ts:
            o = L.dist(t['obs'])
            olds.append(Normal(o.loc.clone(), o.scale.clone()))
    n_iters = len(perms_per_rank[0])
    stats = {'kl': [], 'iters': 0}

    def avg_step(net, loss_fn):
        params = list(L.params[net].values())
        acc = [torch.zeros_like(p_) for p_ in params]
        for r in range(P):
            L.opt[net].zero_grad()
            loss_fn(r).backward()
            if max_grad_norm is not None:
                clip_grad_norm_(params, max_grad_norm)
            for a, p_ in zip(acc, params):
                a += p_.grad
        for a, p_ in zip(acc, params):
            p_.grad = a / P
        L.opt[net].step()

    for it in range(n_iters):
        perms = [torch.as_tensor(np.asarray(perms_per_rank[r][it], np.int64)) for r in range(P)]
        for s in range(0, len(perms[0]), batch_size):
            idx = [pm[s:s + batch_size] for pm in perms]

            def critic_loss(net, tgt):
                def f(r):
                    v = ac.critic_value(L.params[net], ts[r]['obs'][idx[r]])
                    loss = torch.nn.functional.mse_loss(v, ts[r][tgt][idx[r]])
                    for p_ in L.params[net].values():
                        loss = loss + p_.pow(2).sum() * critic_norm_coef
                    return loss
                return f

            avg_step('reward_critic', critic_loss('reward_critic', 'target_value_r'))
            avg_step('cost_critic', critic_loss('cost_critic', 'target_value_c'))

            def actor_loss(r):
                t = ts[r]
                adv = (t['adv_r'][idx[r]] - lam * t['adv_c'][idx[r]]) / (1 + lam)
                return L.loss_pi_ppo(t['obs'][idx[r]], t['act'][idx[r]], t['logp'][idx[r]], adv, clip, entropy_coef)[0]

            avg_step('actor', actor_loss)
        with torch.no_grad():
            kl = float(np.mean([kl_divergence(olds[r], L.dist(ts[r]['obs'])).sum(-1, keepdim=True).mean().item() for r in range(P)]))
        stats['kl'].append(kl)
        stats['iters'] += 1
        if kl_early_stop and kl > target_kl:
            break
    return stats
