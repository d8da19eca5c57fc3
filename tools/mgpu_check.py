"""2-rank sanity check (run under torchrun): parameters stay identical across ranks after training,
env shards differ, and the epoch statistics were all-reduced."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import omnisafe_b200
from omnisafe_b200.utils import distributed

torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
distributed.init_process_group('cuda')
w = dist.get_world_size()
cfg = {'seed': 1, 'train_cfgs': {'device': 'cuda', 'vector_env_nums': 256, 'parallel': w, 'total_steps': w * 256 * 32 * 3,
                                 'matmul_precision': os.environ.get('PREC', 'fp32')},
       'algo_cfgs': {'steps_per_epoch': w * 256 * 32, 'batch_size': 1024, 'update_iters': 2},
       'logger_cfgs': {'log_dir': tempfile.mkdtemp(), 'use_tensorboard': False},
       'env_cfgs': {'obs_dim': 60, 'act_dim': 8, 'max_episode_steps': 16}}
for algo in ('PPOLag', 'CPO'):
    agent = omnisafe_b200.Agent(algo, 'SyntheticBox-v0', custom_cfgs=cfg)
    out = agent.learn()
    th = agent.agent._actor_critic.theta
    ref = th.clone(); dist.broadcast(ref, 0)
    same = bool(torch.equal(ref, th))
    obs0 = agent.agent._buf.data['obs'][0, 0, :4].clone()
    g = [torch.zeros_like(obs0) for _ in range(w)]; dist.all_gather(g, obs0)
    if dist.get_rank() == 0:
        print(algo, 'p2p' if not os.environ.get('OSB_NO_P2P') else 'nccl', 'theta checksum %.9e' % float(th.double().sum()), 'world', w, 'params identical across ranks:', same, 'finite:', bool(torch.isfinite(th).all()),
              'env shards differ:', not torch.equal(g[0], g[1]), 'learn() ->', [round(x, 3) for x in out], flush=True)
    assert same
dist.barrier()
