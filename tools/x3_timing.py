import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omnisafe_b200._lib import lib, ptr
out = torch.zeros(16, dtype=torch.int64, device='cuda')
sink = torch.zeros(512, dtype=torch.float32, device='cuda')
print('== MMA issue: bf16 kind::f16, K=16 per MMA')
for style in (1, 2):
    for M, N in ((128, 64), (128, 16), (64, 64), (64, 16), (128, 128)):
        for reps in (48, 192):
            for _ in range(2):
                lib().osb_x3_timing(M, N, reps, style, ptr(out), 0); torch.cuda.synchronize()
            o = out.tolist()
            print(f'style={style} M={M:3d} N={N:3d} reps={reps:3d}: total {o[0]:6d} cyc  issue {o[1]:6d} cyc  per-mma {o[0]/reps:7.1f}')
sys.exit(0)
for mode in (0, 1, 2):
    for cols in (64, 192):
        for reps in (1, 8):
            for _ in range(2):
                lib().osb_x3_epilogue_probe(cols, reps, mode, ptr(out), ptr(sink), 0); torch.cuda.synchronize()
            o = out.tolist()
            print(f'mode={mode} cols={cols:3d} reps={reps}: max {max(o)/reps:8.1f} min {min(o)/reps:8.1f} cyc/pass')
