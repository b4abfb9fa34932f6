"""Synthetic inputs of the benchmark workloads; numpy only, shared by bench.py, the tests and the
CPU oracle's drivers (the oracle imports this module, never the other way round)."""
import numpy as np


def gen_rotated_boxes(n, seed=0, extent=1024.0, wmin=8.0, wmax=128.0):
    """centres U(0,extent)^2, w log-uniform [8,128], h = w*U(0.2,1), theta U(-pi/2,pi/2), unique scores
    U(0.05,1).  Returns float32 [n,9] (x1..y4, score)."""
    rng = np.random.RandomState(seed)
    cx = rng.uniform(0, extent, n)
    cy = rng.uniform(0, extent, n)
    w = np.exp(rng.uniform(np.log(wmin), np.log(wmax), n))
    h = w * rng.uniform(0.2, 1.0, n)
    th = rng.uniform(-np.pi / 2, np.pi / 2, n)
    sc = rng.uniform(0.05, 1.0, n)
    # make scores unique after the float32 cast (ties are implementation-defined in the reference)
    sc = np.sort(sc)[::-1].copy()
    sc32 = sc.astype(np.float32)
    for k in range(1, n):
        if sc32[k] >= sc32[k - 1]:
            sc32[k] = np.nextafter(sc32[k - 1], np.float32(-1))
    perm = rng.permutation(n)
    sc32 = sc32[perm]
    c, s = np.cos(th), np.sin(th)
    out = np.zeros((n, 9), np.float64)
    for k, (dx, dy) in enumerate(((w / 2, -h / 2), (w / 2, h / 2), (-w / 2, h / 2), (-w / 2, -h / 2))):
        out[:, 2 * k] = cx + c * dx - s * dy
        out[:, 2 * k + 1] = cy + s * dx + c * dy
    out = out.astype(np.float32)
    out[:, 8] = sc32
    return out


def const_density_extent(n):
    """extent that keeps the density of 1k boxes per 1024^2 (BASELINE.json configs[2], the sparse variant)"""
    return 1024.0 * float(np.sqrt(n / 1000.0))
