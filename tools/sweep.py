"""BASELINE.json configs[2] and [4]: CPO at the headline size, and the TRPOLag / FOCOPS obs-dim sweep
17 -> 376 at 4096 envs (1 GPU): env-steps/s per epoch, GAE GB/s.  Development / reporting aid."""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import omnisafe_b200


def run(algo, O, epochs=4, N=4096, T=128):
    cfg = {'seed': 0,
           'train_cfgs': {'device': 'cuda', 'vector_env_nums': N, 'total_steps': N * T * (epochs + 4)},
           'algo_cfgs': {'steps_per_epoch': N * T, 'batch_size': 16384, 'update_iters': 8},
           'logger_cfgs': {'log_dir': tempfile.mkdtemp(), 'use_tensorboard': False, 'save_model_freq': 10 ** 9},
           'env_cfgs': {'obs_dim': O, 'act_dim': 8, 'max_episode_steps': 64}}
    algo_obj = omnisafe_b200.Agent(algo, 'SyntheticBox-v0', custom_cfgs=cfg).agent
    for _ in range(2):
        algo_obj.train_epoch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(epochs):
        algo_obj.train_epoch()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / epochs
    buf = algo_obj._buf
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(20):
        buf.finish_paths()
    g1.record(); torch.cuda.synchronize()
    gae_us = g0.elapsed_time(g1) / 20 * 1e3
    out = {'algo': algo, 'obs_dim': O, 'ms_per_epoch': round(ms, 2), 'env_steps_per_s': round(N * T / ms * 1e3),
           'gae_us': round(gae_us, 2), 'gae_GBps': round(33 * N * T / gae_us / 1e3, 1),
           'tensor_core_update': bool(algo_obj._engine.precision == 1 and O <= 512),
           'tensor_core_rollout': bool(algo_obj._engine.precision == 1 and O <= 64)}
    print(json.dumps(out), flush=True)
    del algo_obj
    torch.cuda.empty_cache()


if __name__ == '__main__':
    dims = (17, 60) if '--tc-only' in sys.argv else (17, 60, 111, 376)
    run('CPO', 60)
    for algo in ('TRPOLag', 'FOCOPS'):
        for O in dims:
            run(algo, O)
