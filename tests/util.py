"""Shared helpers for the parity tests."""
import numpy as np
import torch


def det_key(level, loc, cls):
    return (int(level), float(loc[0]), float(loc[1]), int(cls))


def match_by_key(keys_a, keys_b):
    """Returns index pairs (ia, ib) of detections present in both sets."""
    pos = {k: i for i, k in enumerate(keys_b)}
    pairs = [(i, pos[k]) for i, k in enumerate(keys_a) if k in pos]
    if not pairs:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    ia, ib = zip(*pairs)
    return np.array(ia), np.array(ib)


def quat_dist(qa, qb):
    """Distance up to global sign (q and -q are the same rotation; pytorch3d's sign is unpinned)."""
    qa, qb = torch.as_tensor(qa), torch.as_tensor(qb)
    return torch.minimum((qa - qb).norm(dim=-1), (qa + qb).norm(dim=-1))


def rel_err(a, b, floor=1.0):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).abs() / b.abs().clamp(min=floor)).max().item() if a.numel() else 0.0
