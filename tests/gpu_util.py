"""Helpers shared by the -m gpu parity tests (all product calls go through the C ABI)."""
import numpy as np
import pytest


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


_ctx = None


def ctx():
    global _ctx
    if _ctx is None:
        require_cuda()
        import mav_tube_trajectory_generation_b200 as m

        _ctx = m.Context(0)
    return _ctx


def random_problems(po, B, K, D, box=10.0, v_max=3.0, a_max=5.0, seed0=0):
    """createRandomVertices(4, K, +-box, seed0+b) + Nfabian times -> positions [B,K+1,D], times [B,K]."""
    pos = np.empty((B, K + 1, D))
    times = np.empty((B, K))
    for b in range(B):
        _, values = po.create_random_vertices(4, K, [-box] * D, [box] * D, seed0 + b)
        pos[b] = values[:, 0, :]
        times[b] = po.estimate_segment_times_nfabian(pos[b], v_max, a_max)
    return pos, times


def soa(x):
    """[B, ...] -> [..., B] contiguous."""
    return np.ascontiguousarray(np.moveaxis(x, 0, -1))


def aos(x):
    """[..., B] -> [B, ...] contiguous."""
    return np.ascontiguousarray(np.moveaxis(x, -1, 0))


def dev(x):
    import torch

    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def host(t):
    return t.cpu().numpy() if hasattr(t, "cpu") else t


def normwise(a, b):
    den = np.abs(b).max(axis=-1)
    num = np.abs(a - b).max(axis=-1)
    den = np.where(den == 0.0, 1.0, den)
    return num / den
