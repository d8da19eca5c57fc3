"""Quick device timings of the individual kernels (CUDA events); development aid, not bench.py."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace as NS
import numpy as np, torch
from omnisafe_b200.adapter.onpolicy_adapter import OnPolicyAdapter
from omnisafe_b200.common.buffer import VectorOnPolicyBuffer
from omnisafe_b200.models import ConstraintActorCritic

dev = torch.device('cuda:0')
N, O, A = 4096, 60, 8


def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


net = NS(hidden_sizes=[64, 64], activation='tanh', lr=3e-4)
mc = NS(actor=net, critic=net, actor_type='gaussian_learning', linear_lr_decay=True, weight_initialization_mode='kaiming_uniform')
for T in (128, 512, 2048):
    cfgs = NS(algo_cfgs=NS(obs_normalize=True, reward_normalize=False, cost_normalize=False),
              logger_cfgs=NS(window_lens=100), env_cfgs=dict(obs_dim=O, act_dim=A, max_episode_steps=64))
    ad = OnPolicyAdapter('SyntheticBox-v0', N, 0, cfgs, device=dev)
    agent = ConstraintActorCritic(O, A, mc, epochs=10, device=dev)
    buf = VectorOnPolicyBuffer(O, A, T, 0.99, 0.95, 0.95, 'gae', 0.0, True, True, num_envs=N, device=dev, keep_discounted_ret=False)
    if T == 128:
        ms = timeit(lambda: ad.rollout(T, agent, buf), iters=5)
        print(f'rollout N={N} T={T}: {ms:.3f} ms/epoch  ({ms*1e3/T:.2f} us/step)  {N*T/ms/1e3:.1f} M env-steps/s')
    else:
        buf.data['reward'].uniform_(); buf.data['value_r'].normal_(); buf.data['value_c'].normal_()
        buf.data['flags'][63::64] = 2
    ms = timeit(buf.finish_paths, iters=50)
    b = N * T * 33
    print(f'gae N={N} T={T}: {ms*1e3:.2f} us  {b/ms/1e6:.1f} GB/s algorithmic (33 B/sample)')
    del ad, agent, buf
