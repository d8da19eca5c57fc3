"""Golden fixture of the reference's DATA-PARALLEL semantics: two gloo ranks of the UNMODIFIED reference
(PPOLag, parallel = 2; /root/reference through oracle/ref_shim.py) roll out their own env shards and run ONE
`_update`; every rank records its data, its DataLoader orders and the (rank-identical) parameters before / after.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/golden/make_golden_parallel2.py

Pins: seed + 1000 * rank (base_algo.py:L34-53), sync_params (policy_gradient.py:L98-99), two-phase advantage
statistics over both ranks (distributed.py:L361-393), all-reduced window Jc for the multiplier, and per minibatch
clip locally -> average gradients -> Adam (policy_gradient.py:L437-443, distributed.py:L193-198).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

dist.init_process_group('gloo')            # before the reference is imported: its helpers see world_size() == 2
torch.set_num_threads(1)
import make_golden as mg  # noqa: E402  (installs the shim, registers the synthetic env with the reference)

rank, world = dist.get_rank(), dist.get_world_size()
assert world == 2
N, T, O, A, seed = 8, 24, 12, 3, 41
algo = mg._build_algo('PPOLag', N, T, O, A, seed, tmax=8, term_prob=0.05,
                      extra_algo={'steps_per_epoch': world * N * T})      # global steps: T per env on every rank (policy_gradient.py:L70-77)
assert algo._steps_per_epoch == T
theta0 = mg._flat_theta(algo._actor_critic)
algo._env.rollout(steps_per_epoch=algo._steps_per_epoch, agent=algo._actor_critic, buffer=algo._buf, logger=algo._logger)
data = {k: v.numpy().copy() for k, v in algo._buf.get().items()}       # advantages standardised over BOTH ranks
lam0 = float(algo._lagrange.lagrangian_multiplier.item())
Jc = algo._logger.get_stats('Metrics/EpCost')[0]                        # all-reduced window mean
_, perms = mg._record_randperm(algo._update)
B = data['obs'].shape[0]
perms = np.stack([p.numpy() for p in perms if p.numel() == B])
theta1 = mg._flat_theta(algo._actor_critic)
lg = algo._logger
rec = dict(theta0=theta0, theta1=theta1, lam0=lam0, lam1=float(algo._lagrange.lagrangian_multiplier.item()), Jc=Jc, perms=perms,
           kl=mg._last(lg, 'Train/KL'), stop_iter=mg._last(lg, 'Train/StopIter'),
           **{'data_' + k: v for k, v in data.items()})
gathered = [None, None]
dist.all_gather_object(gathered, rec)
if rank == 0:
    assert np.array_equal(gathered[0]['theta0'], gathered[1]['theta0']), 'sync_params'
    assert np.array_equal(gathered[0]['theta1'], gathered[1]['theta1']), 'ranks diverged'
    assert not np.array_equal(gathered[0]['data_obs'], gathered[1]['data_obs']), 'env shards must differ'
    out = dict(N=N, T=T, O=O, A=A, seed=seed, batch_size=32, update_iters=2, cost_limit=25.0, lambda_lr=0.035,
               theta0=gathered[0]['theta0'], theta1=gathered[0]['theta1'], lam0=gathered[0]['lam0'], lam1=gathered[0]['lam1'],
               Jc=gathered[0]['Jc'], kl=gathered[0]['kl'], stop_iter=gathered[0]['stop_iter'])
    for r in (0, 1):
        out[f'perms_r{r}'] = gathered[r]['perms']
        for k, v in gathered[r].items():
            if k.startswith('data_'):
                out[f'r{r}_{k}'] = v
    np.savez(os.path.join(mg.OUT, 'update_ppolag_parallel2.npz'), **out)
    print('wrote update_ppolag_parallel2.npz; |theta1 - theta0| =', float(np.linalg.norm(out['theta1'] - out['theta0'])),
          'kl', out['kl'], 'stop_iter', out['stop_iter'])
dist.barrier()
