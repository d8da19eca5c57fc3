import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, omnisafe_b200
from omnisafe_b200._lib import current_stream, lib, ptr
w = bench.WORKLOAD
algo = omnisafe_b200.Agent(w['algo'], w['env'], custom_cfgs=bench._custom_cfgs(1, tempfile.mkdtemp(), 50)).agent
algo.train_epoch(); torch.cuda.synchronize()
eng, buf, ac, d = algo._engine, algo._buf, algo._actor_critic, algo._buf.data
O, A, total = w['obs_dim'], w['act_dim'], buf.T * buf.N
for bs in (128, 6272, 16384, 6272 * 6):
    for rep in range(3):
        lib().osb_minibatch_grad_tc(ptr(ac.theta), O, A, ptr(d['obs']), ptr(d['act']), ptr(d['logp']), ptr(d['adv_r']), ptr(d['adv_c']), ptr(d['target_value_r']), ptr(d['target_value_c']), ptr(eng.mu_old), ptr(buf.adv_moments), 0, total, 1, 0, bs, 0, 0.2, 0.0, 1.0, 0.0, ptr(algo._lagrange.state), ptr(eng.logstd_old), 7, ptr(eng.gpart), ptr(eng.stats_part), 0, current_stream())
        torch.cuda.synchronize()
