"""Seeded, framework-independent weights for fixtures too large to commit (gomoku: 22 MB fp32)."""
import zlib

import numpy as np


def seeded_tensor(key, shape, seed=0):
    rs = np.random.RandomState((zlib.crc32(key.encode()) + seed) & 0x7FFFFFFF)
    shape = tuple(int(s) for s in shape)
    if key.endswith("num_batches_tracked"):
        return np.zeros(shape, dtype=np.int64)
    if key.endswith("running_var"):
        return rs.uniform(0.5, 1.5, shape).astype(np.float32)
    if key.endswith("running_mean"):
        return (rs.normal(0, 0.1, shape)).astype(np.float32)
    if ".bn" in key and key.endswith(".weight"):
        return rs.uniform(0.5, 1.5, shape).astype(np.float32)
    if key.endswith(".bias"):
        return rs.normal(0, 0.1, shape).astype(np.float32)
    fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else int(shape[0])
    return rs.normal(0, 1.0 / np.sqrt(fan_in), shape).astype(np.float32)


def seeded_state_dict(keys, shapes, seed=0):
    return {k: seeded_tensor(k, s, seed) for k, s in zip(keys, shapes)}
