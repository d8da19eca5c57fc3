"""Math utilities mirroring omnisafe/utils/math.py on the device."""
from __future__ import annotations

import torch

from omnisafe_b200._lib import current_stream, lib, ptr


def discount_cumsum(vector_x: torch.Tensor, discount: float) -> torch.Tensor:
    """y_t = x_t + discount * y_{t+1} carried in float64 (omnisafe/utils/math.py:L59-82).

    `vector_x` is a CUDA tensor of shape [L] or [B, L] (fp32 or fp64); returns float64."""
    assert vector_x.is_cuda, 'omnisafe_b200 has no CPU fallback: pass a CUDA tensor'
    x = vector_x.contiguous()
    if x.dtype not in (torch.float32, torch.float64):
        x = x.to(torch.float32)
    rows, length = (1, x.shape[0]) if x.dim() == 1 else (x.shape[0], x.shape[1])
    out = torch.empty(x.shape, dtype=torch.float64, device=x.device)
    if x.numel() == 0:
        return out
    lib().osb_discount_cumsum(ptr(x), int(x.dtype == torch.float64), rows, length, float(discount),
                              ptr(out), current_stream())
    return out
