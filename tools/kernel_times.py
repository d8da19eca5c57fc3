"""Per-kernel device times (CUDA events) at the bench workload; development aid."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, omnisafe_b200
from omnisafe_b200._lib import current_stream, lib, ptr

w = bench.WORKLOAD
agent = omnisafe_b200.Agent(w['algo'], w['env'], custom_cfgs=bench._custom_cfgs(1, tempfile.mkdtemp(), 50))
algo = agent.agent
for _ in range(2): algo.train_epoch()
torch.cuda.synchronize()
eng, buf, ac, d = algo._engine, algo._buf, algo._actor_critic, algo._buf.data
O, A, total, bs = w['obs_dim'], w['act_dim'], buf.T * buf.N, w['batch_size']
s = current_stream


def timeit(name, fn, iters=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    print(f'{name:28s} {e0.elapsed_time(e1) / iters * 1e3:10.2f} us')


nb = lib().osb_update_grid_blocks(bs)
timeit('minibatch_grad_tc', lambda: lib().osb_minibatch_grad_tc(ptr(ac.theta), O, A, ptr(d['obs']), ptr(d['act']), ptr(d['logp']), ptr(d['adv_r']), ptr(d['adv_c']), ptr(d['target_value_r']), ptr(d['target_value_c']), ptr(eng.mu_old), ptr(buf.adv_moments), 0, total, 1, 0, bs, 0, 0.2, 0.0, 1.0, 0.0, ptr(algo._lagrange.state), ptr(eng.logstd_old), 7, ptr(eng.gpart), ptr(eng.stats_part), 0, s()))
timeit('minibatch_grad (fp32)', lambda: lib().osb_minibatch_grad(ptr(ac.theta), O, A, ptr(d['obs']), ptr(d['act']), ptr(d['logp']), ptr(d['adv_r']), ptr(d['adv_c']), ptr(d['target_value_r']), ptr(d['target_value_c']), ptr(eng.mu_old), ptr(buf.adv_moments), 0, total, 1, 0, bs, 0, 0.2, 0.0, 1.0, 0.0, ptr(algo._lagrange.state), ptr(eng.logstd_old), 7, ptr(eng.gpart), ptr(eng.stats_part), 0, s()))
timeit('grad_reduce', lambda: lib().osb_grad_reduce(ptr(eng.gpart), ptr(eng.stats_part), nb, O, A, ptr(ac.theta), ptr(ac.grad), 0.001, 7, ptr(eng.sumsq_part), ptr(ac.adam_step), ptr(eng.train_stats), 0, s()))
timeit('clip_adam', lambda: lib().osb_clip_adam(ptr(ac.grad), ptr(ac.theta), ptr(ac.adam_m), ptr(ac.adam_v), ptr(ac.adam_step), ptr(eng.sumsq_part), O, A, 40.0, 0.0, 0.0, 0.0, 1.0, 0.001, ptr(eng.train_stats), 1, 1, 7, 0, s()))
timeit('actor_eval_tc (full batch)', lambda: lib().osb_actor_eval_tc(ptr(ac.theta), O, A, ptr(d['obs']), ptr(d['act']), ptr(d['logp']), ptr(d['adv_r']), ptr(d['adv_c']), ptr(eng.mu_old), ptr(eng.logstd_old), ptr(buf.adv_moments), ptr(algo._lagrange.state), total, 1, 0, ptr(eng.eval_ws), ptr(eng.eval_out), s()), iters=10)
timeit('optim_fused', lambda: lib().osb_optim_fused(ptr(eng.gpart), ptr(eng.stats_part), nb, O, A, ptr(ac.theta), ptr(ac.grad), ptr(ac.adam_m), ptr(ac.adam_v), ptr(ac.adam_step), 0.001, 40.0, 0.0, 0.0, 0.0, 7, ptr(eng.sumsq_part), ptr(eng.train_stats), 0, s()))
timeit('gae_dual', buf.finish_paths)
timeit('rollout epoch (T=128)', lambda: algo._env.rollout(buf.T, ac, buf), iters=5)
timeit('update epoch (8 passes)', algo._update, iters=3)
timeit('train_epoch', algo.train_epoch, iters=3)
