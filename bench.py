"""bench.py -- env-steps/sec over full PPO-Lag epochs (rollout + dual GAE + update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one epoch of the BASELINE.json workload `configs[1]`: PPOLag on the synthetic Box env
(obs 60 / act 8), 4096 HBM-resident envs per GPU, T = 128 steps per env (524 288 samples per GPU),
update_iters 8, batch_size 16384 -- i.e. the reference's `Time/FPS = steps_per_epoch / epoch_time`
(omnisafe/algorithms/on_policy/base/policy_gradient.py:L280).  Prints ONE JSON line.

  value   : device-timed (CUDA events, barrier + synchronize on both sides, max over ranks),
            everything resident in HBM, in-kernel Philox noise.
  e2e     : the same metric through the public `omnisafe_b200.Agent(...)` training loop with HOST
            buffers: every epoch the standard-normal action-noise stream is copied from pinned host
            memory (parity-mode input of the rollout) and the epoch's logged metrics are read back.
  --impl reference : the CPU restatement of the reference path (oracle/, torch-CPU + numpy) timed
            on the host cores on a bounded sample of the same workload (fewer envs, same T / update
            schedule per sample); /root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(algo='PPOLag', env='SyntheticBox-v0', obs_dim=60, act_dim=8, envs_per_gpu=4096,
                steps_per_env=128, batch_size=16384, update_iters=8, max_episode_steps=64)


# ------------------------------------------------------------------------------------------------
def _peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {'hbm_gbs': p['hbm_gbs'], 'bf16_tflops': p['bf16_tflops'],
                'bf16_tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int) -> None:
        self.rows, self._stop, self.gpu = [], threading.Event(), gpu_index
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self) -> None:
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                      '-i', str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(',')])
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=3)

    def summary(self) -> dict:
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (('hw_slowdown', 4), ('hw_thermal_slowdown', 5), ('sw_thermal_slowdown', 6), ('sw_power_cap', 7)):
                if len(r) > col and r[col].lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx or None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
def _custom_cfgs(world: int, log_dir: str, epochs: int) -> dict:
    w = WORKLOAD
    spe = world * w['envs_per_gpu'] * w['steps_per_env']
    return {
        'seed': 0,
        'train_cfgs': {'device': 'cuda', 'vector_env_nums': w['envs_per_gpu'], 'parallel': world,
                       'total_steps': spe * epochs},
        'algo_cfgs': {'steps_per_epoch': spe, 'batch_size': w['batch_size'], 'update_iters': w['update_iters']},
        'logger_cfgs': {'log_dir': log_dir, 'use_tensorboard': False, 'save_model_freq': 10 ** 9},
        'env_cfgs': {'obs_dim': w['obs_dim'], 'act_dim': w['act_dim'], 'max_episode_steps': w['max_episode_steps']},
    }


def _launches_per_epoch() -> int:
    w = WORKLOAD
    T = w['steps_per_env']
    n_mb = -(-(w['envs_per_gpu'] * T) // w['batch_size'])
    rollout = 1 + (T + 1) + 2            # reset, T steps + bootstrap launch, episode window + sums
    gae = 2 + 1                          # scan + stats reduce, moments
    update = 1 + 1 + w['update_iters'] * (n_mb * 2 + 3)   # lagrange, old-policy snapshot, (fused fwd+bwd, fused reduce+clip+adam)*mb + (eval, reduce, kl)
    return rollout + gae + update


def run_b200(args) -> dict:
    import torch.distributed as dist

    import omnisafe_b200
    from omnisafe_b200._lib import current_stream, lib, ptr
    from omnisafe_b200.utils import distributed

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N > 1)'
    torch.cuda.set_device(local_rank)
    if world > 1:
        distributed.init_process_group('cuda')
    w = WORKLOAD
    T, N = w['steps_per_env'], w['envs_per_gpu']
    tmp = tempfile.mkdtemp(prefix='osb_bench_')
    agent = omnisafe_b200.Agent(w['algo'], w['env'], custom_cfgs=_custom_cfgs(world, tmp, args.steps + args.warmup + 8))
    algo = agent.agent

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        """K calls of fn bracketed by barrier + synchronize, device time (ms), max over ranks."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident number ("value") -------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        algo.train_epoch()
    with ClockSampler(local_rank) as clk:
        ms = timed(algo.train_epoch, args.steps)
    clocks = clk.summary()
    ms_per_step = ms / args.steps
    samples_global = world * N * T
    value = samples_global / (ms_per_step * 1e-3)

    # ---- end to end through the public loop with host buffers ----------------------------------
    A = w['act_dim']
    host_eps = torch.randn(T, N, A, dtype=torch.float32).pin_memory()
    dev_eps = torch.empty(T, N, A, dtype=torch.float32, device='cuda')
    d2h = {'bytes': 0}

    def e2e_epoch():
        dev_eps.copy_(host_eps, non_blocking=True)            # H2D of this epoch's noise stream
        row = algo.train_epoch(eps=dev_eps, log=True)          # logs => D2H read of the epoch metrics
        d2h['bytes'] = row['d2h_bytes']

    for _ in range(2):
        e2e_epoch()
    ms_e2e = timed(e2e_epoch, args.steps) / args.steps
    e2e = {'value': samples_global / (ms_e2e * 1e-3), 'unit': 'env-steps/s',
           'h2d_bytes_per_step': host_eps.numel() * 4, 'd2h_bytes_per_step': d2h['bytes'],
           'ms_per_step': ms_e2e}

    # ---- roofline of the dominant kernel (fused minibatch fwd+bwd), timed live -----------------
    peaks = _peaks()
    eng, buf, ac = algo._engine, algo._buf, algo._actor_critic
    d = buf.data
    total = T * N

    def grad_launch():
        lib().osb_minibatch_grad_tc(ptr(ac.theta), w['obs_dim'], A, ptr(d['obs']), ptr(d['act']), ptr(d['logp']),
                                    ptr(d['adv_r']), ptr(d['adv_c']), ptr(d['target_value_r']), ptr(d['target_value_c']),
                                    ptr(eng.mu_old), ptr(buf.adv_moments), 0, total, 12345, 0, w['batch_size'], 0, 0.2, 0.0,
                                    1.0, 0.0, ptr(algo._lagrange.state), ptr(eng.logstd_old), 7, ptr(eng.gpart), ptr(eng.stats_part), 0,
                                    current_stream())

    ms_grad = timed(grad_launch, 50) / 50
    flop_per_sample = _flops_per_sample(w['obs_dim'], A)
    ach_tf = flop_per_sample * w['batch_size'] / (ms_grad * 1e-3) / 1e12
    ms_gae = timed(buf.finish_paths, 50) / 50
    gae_gbs = 33.0 * total / (ms_gae * 1e-3) / 1e9
    roofline = {'kernel': 'minibatch_grad_tc_kernel', 'bound': 'tensor', 'achieved': ach_tf,
                'peak': peaks['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                'frac': ach_tf / peaks['bf16_tflops_sustained'],
                'traffic': 14.57e6,   # dram bytes read + written per launch, ncu --set full (profiles/r01_ncu_minibatch_grad_tc.md)
                'algorithmic_bytes': 292.0 * w['batch_size'] + 147 * 24850 * 4.0,   # sample rows read + per-CTA partial gradients written
                'peak_source': peaks['source'] + ' (cuBLAS bf16 sustained; kernel runs tcgen05 kind::tf32, nominal tf32 peak = half of bf16)',
                'us_per_launch': ms_grad * 1e3,
                'gae': {'kernel': 'gae_dual_kernel', 'bound': 'hbm', 'achieved': gae_gbs, 'peak': peaks['hbm_gbs'],
                        'unit': 'GB/s', 'frac': gae_gbs / peaks['hbm_gbs'], 'us_per_launch': ms_gae * 1e3,
                        'bytes_per_sample': 33}}

    out = {
        'metric': 'env-steps/sec (rollout+GAE+update) PPO-Lag', 'value': value, 'unit': 'env-steps/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'tf32 (fp32 storage/accumulate; GAE fp64 carry)', 'data': 'synthetic',
        'config': {'workload': 'PPOLag SyntheticBox-v0 obs60/act8, 4096 envs/GPU x T=128, batch 16384, update_iters 8 '
                               '(BASELINE.json configs[1])', 'envs_per_gpu': N, 'steps_per_env': T,
                   'global_samples_per_step': samples_global, 'parallelism': f'dp{world}',
                   'l2_policy': 'inputs larger than L2 (per-epoch slabs ~157 MB > 126 MB)', 'noise': 'in-kernel Philox'},
        'e2e': e2e, 'gpu_launches': _launches_per_epoch() * args.steps, 'clocks': clocks, 'roofline': roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out['cpu_baseline'] = cpu_baseline(budget_s=20.0)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out if rank == 0 else {}


def _flops_per_sample(O: int, A: int) -> int:
    """fwd + bwd multiply-adds x2 of the three trunks (actor O-64-64-A, two critics O-64-64-1);
    backward = 2x forward except that no dX is formed for layer 1."""
    def net(out):
        fwd = 2 * (O * 64 + 64 * 64 + 64 * out)
        bwd = 2 * (O * 64 + 2 * 64 * 64 + 2 * 64 * out)
        return fwd + bwd
    return net(A) + 2 * net(1)


# ------------------------------------------------------------------------------------------------
def _oracle_epoch(state, n_envs: int):
    """One epoch of the CPU restatement on n_envs envs (same T / update schedule per sample)."""
    from oracle import gae as ogae
    from oracle import learner as ol
    from oracle import rollout as orollout

    w = WORKLOAD
    T, O, A = w['steps_per_env'], w['obs_dim'], w['act_dim']
    eps = state['rng'].standard_normal((T, n_envs, A)).astype(np.float32)
    theta = state['learner'].flat()
    window = state['window']
    sl = orollout.rollout_epoch(state['env'], state['norm'], theta, T, eps, window=window)
    out = ogae.dual_gae_slab(sl['rew'], sl['cost'], sl['val_r'], sl['val_c'], sl['flags'], sl['boot_r'], sl['boot_c'],
                             0.99, 0.95, 0.95)
    sr, sc = ogae.standardize(out['adv_r'], out['adv_c'])
    B = T * n_envs
    em = lambda x: np.ascontiguousarray(np.swapaxes(x, 0, 1)).reshape(B, *x.shape[2:])   # noqa: E731
    data = {'obs': em(sl['obs']), 'act': em(sl['act']), 'logp': em(sl['logp']), 'adv_r': em(sr), 'adv_c': em(sc),
            'target_value_r': em(out['tv_r']), 'target_value_c': em(out['tv_c'])}
    jc = float(np.mean([c for _, c, _ in window[-100:]]))
    lam = state['lagrange'].update(jc)
    bs = max(64, w['batch_size'] * n_envs // w['envs_per_gpu'])
    perms = [state['rng'].permutation(B) for _ in range(w['update_iters'])]
    state['learner'].update_ppo(data, perms, lam, batch_size=bs)
    return B


def _oracle_state(n_envs: int):
    from oracle import actor_critic as oac
    from oracle import learner as ol
    from oracle.normalizer import Normalizer
    from oracle.synthetic_env import SyntheticBoxEnv

    w = WORKLOAD
    return {'env': SyntheticBoxEnv(n_envs, w['obs_dim'], w['act_dim'], max_episode_steps=w['max_episode_steps'], seed=0),
            'norm': Normalizer((w['obs_dim'],)), 'learner': ol.Learner(oac.init_theta(w['obs_dim'], w['act_dim'], 0), w['obs_dim'], w['act_dim']),
            'lagrange': ol.Lagrange(25.0, 0.001, 0.035), 'rng': np.random.default_rng(0), 'window': []}


def _host_threads() -> int:
    """Threads for the CPU arm.  torch-CPU on these small layers collapses when oversubscribed
    (128 threads on the GPU box made one 64-env epoch take minutes), so the arm uses up to 16."""
    cores = os.cpu_count() or 1
    n = min(cores, 16)
    torch.set_num_threads(n)
    return n


def _sized_sample(budget_s: float) -> int:
    """Largest env count (multiple of 64, <= 4096) whose epoch fits `budget_s`, from a 256-env probe."""
    t0 = time.time(); st = _oracle_state(256); _oracle_epoch(st, 256); probe = time.time() - t0
    per_env = probe / 256.0                      # pessimistic: per-env cost falls with n
    return int(min(WORKLOAD['envs_per_gpu'], max(64, (budget_s / max(per_env, 1e-9)) // 64 * 64)))


def cpu_baseline(budget_s: float = 20.0) -> dict:
    """The oracle port timed on the host cores on a bounded sample: one full epoch on n envs."""
    threads = _host_threads()
    n = _sized_sample(budget_s)
    st = _oracle_state(n)
    t0 = time.time(); B = _oracle_epoch(st, n); dt = time.time() - t0
    return {'value': B / dt, 'unit': 'env-steps/s', 'cores': threads, 'kind': 'port',
            'sample': f'1 PPOLag epoch (rollout+GAE+update, update_iters 8) on {n} envs x T=128 = {B} env-steps, '
                      f'oracle/ torch-CPU+numpy restatement, {threads} torch threads of {os.cpu_count()} cores, {dt:.1f} s'}


def run_reference(args) -> dict:
    """--impl reference: the CPU restatement of the reference path on the host cores, bounded sample
    per step.  Under torchrun only rank 0 works."""
    if int(os.environ.get('RANK', '0')) != 0:
        return {}
    threads = _host_threads()
    steps, warm = args.steps, max(args.warmup, 1)
    n = _sized_sample(120.0 / (steps + warm))            # whole run within a few minutes
    st = _oracle_state(n)
    for _ in range(warm):
        _oracle_epoch(st, n)
    t0 = time.time()
    done = 0
    for _ in range(steps):
        done += _oracle_epoch(st, n)
    dt = time.time() - t0
    v = done / dt
    sample = (f'each step = 1 PPOLag epoch (rollout+GAE+update, update_iters 8, batch scaled) on {n} envs x T=128; '
              f'oracle/ torch-CPU+numpy restatement of the reference path, {threads} torch threads of {os.cpu_count()} cores')
    return {'impl': 'reference', 'metric': 'env-steps/sec (rollout+GAE+update) PPO-Lag', 'value': v,
            'unit': 'env-steps/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': warm,
            'ms_per_step': dt / steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'PPOLag SyntheticBox-v0 obs60/act8, T=128, update_iters 8 (bounded sample of '
                                   'BASELINE.json configs[1])', 'sample_envs': n},
            'cpu_baseline': {'value': v, 'unit': 'env-steps/s', 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': v, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    out = run_reference(args) if args.impl == 'reference' else run_b200(args)
    if out:
        print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
