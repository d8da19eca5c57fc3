"""Oracle: actor / critic MLPs and the Gaussian policy head (torch-CPU).  TEST INFRASTRUCTURE ONLY.

Follows omnisafe/utils/model.py:L73-111 (Linear-Tanh-Linear-Tanh-Linear), models/actor/
gaussian_learning_actor.py:L64-139 (Normal(mean, exp(log_std)), rsample, log_prob summed over the
action dim), models/critic/v_critic.py:L75-92, models/actor_critic/constraint_actor_critic.py:L84-109
and the flat parameter order of utils/tools.py:L35-129 (named_parameters order, log_std first).
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch.distributions import Normal

HID = 64


def layout(O: int, A: int, hid: int = HID):
    """Sizes/offsets of the flat vector theta = [actor | reward_critic | cost_critic]."""
    actor = [('log_std', (A,)), ('w1', (hid, O)), ('b1', (hid,)), ('w2', (hid, hid)), ('b2', (hid,)),
             ('w3', (A, hid)), ('b3', (A,))]
    critic = [('w1', (hid, O)), ('b1', (hid,)), ('w2', (hid, hid)), ('b2', (hid,)),
              ('w3', (1, hid)), ('b3', (1,))]
    out, off = {}, 0
    for net, spec in (('actor', actor), ('reward_critic', critic), ('cost_critic', critic)):
        start = off
        entries = {}
        for name, shape in spec:
            n = int(np.prod(shape))
            entries[name] = (off, shape)
            off += n
        out[net] = {'start': start, 'size': off - start, 'entries': entries}
    out['total'] = off
    return out


def unflatten(theta, O: int, A: int):
    """dict net -> dict name -> torch view (shares memory with theta if theta is a torch tensor)."""
    theta = torch.as_tensor(theta)
    lay = layout(O, A)
    return {net: {name: theta[o:o + int(np.prod(shape))].view(*shape)
                  for name, (o, shape) in lay[net]['entries'].items()}
            for net in ('actor', 'reward_critic', 'cost_critic')}


def init_theta(O: int, A: int, seed: int = 0) -> np.ndarray:
    """Random init with the reference's scheme: kaiming_uniform_(a=sqrt(5)) weights and torch's
    default Linear bias init (utils/model.py:L25-44), log_std = 0."""
    g = torch.Generator().manual_seed(seed)
    lay = layout(O, A)
    theta = torch.zeros(lay['total'])
    views = unflatten(theta, O, A)
    for net in ('actor', 'reward_critic', 'cost_critic'):
        for w, b in (('w1', 'b1'), ('w2', 'b2'), ('w3', 'b3')):
            W = views[net][w]
            fan_in = W.shape[1]
            bound_w = math.sqrt(6.0 / ((1 + 5.0) * fan_in))
            W.uniform_(-bound_w, bound_w, generator=g)
            bound_b = 1.0 / math.sqrt(fan_in)
            views[net][b].uniform_(-bound_b, bound_b, generator=g)
    return theta.numpy().copy()


def mlp(p, x):
    h = torch.tanh(torch.nn.functional.linear(x, p['w1'], p['b1']))
    h = torch.tanh(torch.nn.functional.linear(h, p['w2'], p['b2']))
    return torch.nn.functional.linear(h, p['w3'], p['b3'])


def actor_dist(p, obs) -> Normal:
    return Normal(mlp(p, obs), torch.exp(p['log_std']))


def critic_value(p, obs):
    return torch.squeeze(mlp(p, obs), -1)


def step(theta, obs, eps, O: int, A: int):
    """ConstraintActorCritic.step with the standard-normal draw `eps` supplied by the caller
    (Normal.rsample == loc + eps * scale).  Returns float32 numpy (act, v_r, v_c, logp)."""
    nets = unflatten(torch.as_tensor(np.asarray(theta, np.float32)), O, A)
    obs = torch.as_tensor(np.asarray(obs, np.float32))
    eps = torch.as_tensor(np.asarray(eps, np.float32))
    with torch.no_grad():
        v_r = critic_value(nets['reward_critic'], obs)
        v_c = critic_value(nets['cost_critic'], obs)
        dist = actor_dist(nets['actor'], obs)
        act = dist.loc + eps * dist.scale
        logp = dist.log_prob(act).sum(-1)
    return act.numpy(), v_r.numpy(), v_c.numpy(), logp.numpy()


def values(theta, obs, O: int, A: int):
    nets = unflatten(torch.as_tensor(np.asarray(theta, np.float32)), O, A)
    obs = torch.as_tensor(np.asarray(obs, np.float32))
    with torch.no_grad():
        return (critic_value(nets['reward_critic'], obs).numpy(),
                critic_value(nets['cost_critic'], obs).numpy())
