"""Device time of one PPOLag epoch (bench workload) per matmul_precision, split into rollout / GAE / update."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import omnisafe_b200
import bench

def timed(fn, k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k

for prec in (sys.argv[1:] or ['bf16x3', 'tf32', 'fp32']):
    cfg = bench._custom_cfgs(1, tempfile.mkdtemp(), 40)
    cfg['train_cfgs']['matmul_precision'] = prec
    agent = omnisafe_b200.Agent('PPOLag', 'SyntheticBox-v0', custom_cfgs=cfg)
    algo = agent.agent
    for _ in range(3): algo.train_epoch()
    ms = timed(algo.train_epoch, 10)
    ro = timed(lambda: algo._env.rollout(algo._steps_per_epoch, algo._actor_critic, algo._buf, algo._logger), 5)
    ga = timed(algo._buf.finish_paths, 20)
    up = timed(algo._update, 5)
    print(f'{prec:7s}: epoch {ms:7.3f} ms  ({4096*128/ms/1e3:6.2f} M env-steps/s)   rollout {ro:6.3f}  gae {ga:6.3f}  update {up:6.3f} ms', flush=True)
    del agent, algo
