import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omnisafe_b200._lib import lib, ptr
out = torch.zeros(2, dtype=torch.int64, device='cuda')
for M, N in ((128, 64), (64, 128), (64, 64), (128, 16), (64, 16), (128, 256)):
    for reps in (1, 8, 16, 64):
        for _ in range(2):
            lib().osb_umma_timing(M, N, reps, ptr(out), 0); torch.cuda.synchronize()
        o = out.tolist()
        print(f'M={M:3d} N={N:3d} reps={reps:3d}: total {o[0]:6d} cyc  issue {o[1]:6d} cyc  per-mma {o[0]/reps:7.1f}')
