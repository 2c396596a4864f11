"""Synthetic workloads for the BASELINE.json configs.

Host-side (numpy) generators used by the parity tests -- the same arrays are fed
to the CPU oracle and to the CUDA library -- and device-side (torch) generators
used by bench.py to fill HBM with inputs of the full BASELINE sizes.  Ranges are
produced in the reference's own wire format: integer millimetres, floored
(PG.cpp:213), with <= 0 meaning "no ranging from this anchor" (PG.cpp:482,
TOA.cpp:50).
"""
from __future__ import annotations

import numpy as np

SEED = 20190816  # base seed, SURVEY.md §8d


def anchors_for(n_anchors: int) -> np.ndarray:
    """Anchor tables of the BASELINE configs, [n][3] float64 metres."""
    if n_anchors == 4:  # config 1: non-coplanar so the 3-D ML is well posed
        a = [(0, 0, 2.5), (10, 0, 0.5), (10, 10, 2.5), (0, 10, 0.5)]
    elif n_anchors == 8:  # config 2/3/5: perimeter of a 10 x 10 m room, z alternating
        xy = [(0, 0), (5, 0), (10, 0), (10, 5), (10, 10), (5, 10), (0, 10), (0, 5)]
        a = [(x, y, 0.5 if i % 2 else 2.5) for i, (x, y) in enumerate(xy)]
    elif n_anchors == 16:  # config 4: 4 x 4 grid, ceiling/floor alternating
        a = []
        for i in range(4):
            for j in range(4):
                a.append((10.0 * i / 3, 10.0 * j / 3, 2.5 if (i + j) % 2 == 0 else 0.5))
    else:
        rng = np.random.default_rng(SEED + n_anchors)
        a = np.column_stack([rng.uniform(0, 10, n_anchors), rng.uniform(0, 10, n_anchors),
                             rng.uniform(0.3, 2.7, n_anchors)])
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def truth_lissajous(n_filters: int, n_steps: int, dt: float, seed: int = SEED, z: float = 1.0):
    """Per-filter Lissajous truth, returns pos [T+1][3][N] (index 0 = initial position)."""
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi, size=(2, n_filters))
    t = (np.arange(n_steps + 1) * dt)[:, None]
    x = 5 + 3 * np.sin(0.20 * t + ph[0][None, :])
    y = 5 + 3 * np.sin(0.31 * t + ph[1][None, :])
    zz = np.full_like(x, z)
    return np.ascontiguousarray(np.stack([x, y, zz], axis=1))


def ranges_mm(truth: np.ndarray, anchors: np.ndarray, sigma: float = 0.10, seed: int = SEED + 1,
              p_missing: float = 0.0, p_nlos: float = 0.0, nlos_mean: float = 0.8,
              dtype=np.int32) -> np.ndarray:
    """Noisy floored-mm ranges [T][M][N] for truth [T][3][N] (or [3][N] -> [M][N])."""
    rng = np.random.default_rng(seed)
    single = truth.ndim == 2
    tr = truth[None] if single else truth
    d = np.sqrt(((tr[:, None, :, :] - anchors[None, :, :, None]) ** 2).sum(axis=2))  # [T][M][N]
    r = d + rng.normal(0.0, sigma, size=d.shape)
    if p_nlos > 0:
        nl = rng.random(size=d.shape) < p_nlos
        r = r + nl * rng.exponential(nlos_mean, size=d.shape)
    mm = np.floor(r * 1000.0)
    if p_missing > 0:
        mm[rng.random(size=d.shape) < p_missing] = 0
    mm = np.clip(mm, 0, np.iinfo(dtype).max).astype(dtype)
    return np.ascontiguousarray(mm[0] if single else mm)


def device_ranges_mm(n_filters: int, n_steps: int, anchors: np.ndarray, dt: float, device,
                     seed: int = SEED, sigma: float = 0.10, dtype=None, step0: int = 0):
    """torch/device version of truth_lissajous + ranges_mm for bench-sized inputs.

    Returns (ranges [T][M][N] int32 on `device`, x0 [3][N] float64, truth_end [3][N]).
    The inputs are i.i.d. per filter; only their shape/statistics matter to the bench.
    """
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    dtype = dtype or torch.int32
    ph = torch.rand((2, n_filters), generator=g, device=device, dtype=torch.float64) * (2 * np.pi)
    anc = torch.as_tensor(anchors, device=device, dtype=torch.float64)
    m = anc.shape[0]
    out = torch.empty((n_steps, m, n_filters), device=device, dtype=dtype)

    def pos(k):
        t = k * dt
        return torch.stack([5 + 3 * torch.sin(0.20 * t + ph[0]), 5 + 3 * torch.sin(0.31 * t + ph[1]),
                            torch.ones_like(ph[0])])

    x0 = pos(step0)
    p = x0
    for k in range(n_steps):
        p = pos(step0 + k + 1)
        d = torch.sqrt(((p[None, :, :] - anc[:, :, None]) ** 2).sum(dim=1))  # [M][N]
        d = d + sigma * torch.randn(d.shape, generator=g, device=device, dtype=torch.float64)
        out[k] = torch.floor(d * 1000.0).clamp_(min=0).to(dtype)
    return out, x0.contiguous(), p.contiguous()
