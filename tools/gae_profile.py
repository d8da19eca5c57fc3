"""A few launches of the training-path GAE kernel at a long horizon (for ncu captures): T = 2048 x 4096 envs = 277 MB
of algorithmic traffic, larger than L2."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnisafe_b200.common.buffer import VectorOnPolicyBuffer

T, N = int(os.environ.get('GAE_T', 2048)), int(os.environ.get('GAE_N', 4096))
buf = VectorOnPolicyBuffer(4, 2, T, 0.99, 0.95, 0.95, 'gae', 0.0, True, True, num_envs=N, device='cuda', keep_discounted_ret=False)
for k in ('reward', 'cost', 'value_r', 'value_c', 'boot_r', 'boot_c'):
    buf.data[k].normal_()
fl = torch.zeros(T, N, dtype=torch.uint8, device='cuda')
fl[63::64] = 2                                   # time-limit truncations every 64 steps (the bench env)
fl |= (torch.rand(T, N, device='cuda') < 0.01).to(torch.uint8)   # 1 % terminations
buf.data['flags'].copy_(fl)
for _ in range(8):
    buf.finish_paths()
torch.cuda.synchronize()
