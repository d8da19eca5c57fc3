"""Scan descriptor conventions: which (LBO) makes each operand view fetch the right elements."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from omnisafe_b200._lib import lib, ptr, current_stream
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_x3_gpu import CASES
dev = torch.device('cuda')

def run(A, B, M, N, K, a_mn, b_mn, a_sw, b_sw, ones, albo, blbo):
    out = torch.full((128, N), float('nan'), dtype=torch.float32, device=dev)
    lib().osb_x3_selftest_dbg(ptr(torch.as_tensor(A).to(dev)), ptr(torch.as_tensor(B).to(dev)), M, N, K, a_mn, b_mn, a_sw, b_sw, ones,
                              albo, 1024 if a_sw == 128 else 256, blbo, 1024 if b_sw == 128 else 256, ptr(out), current_stream())
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    rows = list(range(128)) if M == 128 else [32 * (r // 16) + r % 16 for r in range(64)]
    return got[rows]

rng = np.random.default_rng(0)
for (M, N, K, a_mn, b_mn, a_sw, b_sw, ones) in CASES:
    print(f'=== M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} a_sw={a_sw} b_sw={b_sw} ones={ones}')
    A = rng.standard_normal((M, K)).astype(np.float32); B = rng.standard_normal((N, K)).astype(np.float32)
    if ones: B[:] = 1
    want = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    lbos_a = [16, 0, 1024, (K if a_mn else M) * a_sw]
    lbos_b = [16, 0, 1024, (K if b_mn else N) * b_sw]
    for albo, blbo in itertools.product(lbos_a, lbos_b):
        d = run(A, B, M, N, K, a_mn, b_mn, a_sw, b_sw, ones, albo, blbo)
        err = np.abs(d - want) / scale
        badr = np.where(err.max(1) > 1e-6)[0]
        print(f'  a_lbo={albo:6d} b_lbo={blbo:6d}: max err {np.nanmax(err):.3e}  bad rows {len(badr)}/{len(err)} {badr[:6].tolist()} bad cols {int((err.max(0) > 1e-6).sum())}/{N}')
