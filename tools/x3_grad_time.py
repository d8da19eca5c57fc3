"""Per-launch time of the three minibatch gradient kernels at the bench minibatch (16384 rows, obs 60 / act 8)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from oracle import actor_critic as oac
from test_update_gpu import _rand_data, _setup
from omnisafe_b200._lib import current_stream, lib, ptr
dev = torch.device('cuda:0')
O, A, N, T = 60, 8, 4096, 16
rng = np.random.default_rng(0)
theta = oac.init_theta(O, A, seed=5)
data = _rand_data(rng, N, T, O, A, theta)
agent, buf, eng = _setup(dev, data, N, T, O, A, theta)
B = N * T
lag = torch.tensor([0.3], dtype=torch.float32, device=dev)
d = buf.data
def args(count):
    return (ptr(agent.theta), O, A, ptr(d['obs']), ptr(d['act']), ptr(d['logp']), ptr(d['adv_r']), ptr(d['adv_c']),
            ptr(d['target_value_r']), ptr(d['target_value_c']), ptr(eng.mu_old), ptr(buf.adv_moments), 0, B, 12345,
            0, count, 0, 0.2, 0.0, 1.0, 0.0, ptr(lag), ptr(eng.logstd_old), 7, ptr(eng.gpart), ptr(eng.stats_part), 0, current_stream())
for name in ('osb_minibatch_grad_x3', 'osb_minibatch_grad_tc', 'osb_minibatch_grad'):
    fn = getattr(lib(), name)
    for count in (16384, 49 * 128, 2 * 49 * 128):
        for _ in range(5): fn(*args(count))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): fn(*args(count))
        e1.record(); torch.cuda.synchronize()
        print(f'{name:26s} rows {count:6d}: {e0.elapsed_time(e1) / 50 * 1e3:8.1f} us / launch')
