"""Where one CTA of the streaming GAE kernel spends its time: clock64 stamps of CTA 0 (warp 0 and warp 15, lane 0)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnisafe_b200._lib import lib, ptr
from omnisafe_b200.common.buffer import VectorOnPolicyBuffer

T, N = int(os.environ.get('GAE_T', 1024)), int(os.environ.get('GAE_N', 4096))
buf = VectorOnPolicyBuffer(4, 2, T, 0.99, 0.95, 0.95, 'gae', 0.0, True, True, num_envs=N, device='cuda', keep_discounted_ret=False)
for k in ('reward', 'cost', 'value_r', 'value_c', 'boot_r', 'boot_c'):
    buf.data[k].normal_()
fl = torch.zeros(T, N, dtype=torch.uint8, device='cuda')
fl[63::64] = 2
fl |= (torch.rand(T, N, device='cuda') < 0.01).to(torch.uint8)
buf.data['flags'].copy_(fl)
for _ in range(3):
    buf.finish_paths()
dbg = torch.zeros(1024, dtype=torch.int64, device='cuda')
torch.cuda.synchronize()
lib().osb_gae_debug_buffer(ptr(dbg))
buf.finish_paths()
torch.cuda.synchronize()
lib().osb_gae_debug_buffer(0)
NAMES = {0: 'start', 1: 'tile0 landed', 2: 'tile start', 3: 'pass1 done', 4: 'barrier', 5: 'fold done (+boot prefetch)', 6: 'pass2 done (boot prefetched)', 7: 'pass2 done (boot not yet)', 8: 'values landed'}
for which, off in (('warp 0', 0), ('warp 15', 512)):
    d = dbg[off:off + 512].cpu().tolist()
    n = d[0]
    print(f'== {which}: {n} stamps, T={T} N={N}')
    prev = t0 = d[2]
    for i in range(min(n, 60)):
        id_, clk = d[1 + 2 * i], d[2 + 2 * i]
        print(f'  {NAMES.get(id_, id_):30s} +{clk - prev:7d} cyc   t={clk - t0:8d}')
        prev = clk
