"""Oracle: running observation normaliser (numpy fp32).  TEST INFRASTRUCTURE ONLY.

Follows omnisafe/common/normalizer.py:L88-139 (Chan/Golub/LeVeque batched update, fp32 state,
std floor 1e-2, clip) and ObsNormalize's clip=5 (envs/wrapper.py:L202).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


class Normalizer:
    def __init__(self, shape, clip=5.0):
        self.mean = np.zeros(shape, F32)
        self.sumsq = np.zeros(shape, F32)
        self.std = np.zeros(shape, F32)
        self.count = 0
        self.clip = F32(clip)

    def push(self, raw):
        raw = np.asarray(raw, F32)
        if raw.ndim == self.mean.ndim:
            raw = raw[None]
        n = raw.shape[0]
        count = self.count + n
        mean_raw = raw.mean(axis=0, dtype=F32)
        delta = (mean_raw - self.mean).astype(F32)
        self.mean = (self.mean + (delta * F32(n)).astype(F32) / F32(count)).astype(F32)
        sumq_raw = ((raw - mean_raw) ** 2).sum(axis=0, dtype=F32)
        corr = (((delta ** 2).astype(F32) * F32(self.count)).astype(F32) * F32(n)).astype(F32) / F32(count)
        self.sumsq = (self.sumsq + (sumq_raw + corr.astype(F32)).astype(F32)).astype(F32)
        self.count = count
        with np.errstate(divide='ignore', invalid='ignore'):
            var = (self.sumsq / F32(self.count - 1)).astype(F32)
        self.std = np.maximum(np.sqrt(var), F32(1e-2)).astype(F32)

    def normalize(self, data):
        data = np.asarray(data, F32)
        self.push(data)
        if self.count <= 1:
            return data
        out = ((data - self.mean) / self.std).astype(F32)
        return np.clip(out, -self.clip, self.clip).astype(F32)
