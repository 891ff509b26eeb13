"""Seeded point sets shared by the distCUDA2 tests and tests/golden/make_golden_knn.py."""
import numpy as np


def points(kind: str, P: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if kind == "uniform":          # EOGS++ random initialisation box (scene/dataset_readers/dataset_affine.py:273-281)
        lo, hi = np.array([-0.70, -0.70, -0.10]), np.array([0.70, 0.70, 0.25])
        p = lo + (hi - lo) * rng.random((P, 3))
    elif kind == "clustered":      # tight clusters + exact duplicates (distance 0 must count, simple_knn.cu:176)
        centres = rng.normal(0, 1, (max(P // 200, 1), 3))
        p = centres[rng.integers(0, len(centres), P)] + rng.normal(0, 1e-3, (P, 3))
        dup = rng.integers(0, P, P // 10)
        p[rng.integers(0, P, P // 10)] = p[dup]
    elif kind == "planar":         # degenerate extent in z: Morton normalisation divides by zero in the reference
        p = np.concatenate([rng.random((P, 2)), np.zeros((P, 1))], 1)
    elif kind == "offset":         # UTM-like magnitudes: fp32 cancellation in p_j - p_i
        p = np.array([4.3e5, 3.36e6, 10.0]) + rng.normal(0, 50.0, (P, 3))
    elif kind == "line":           # strongly anisotropic
        t = rng.random((P, 1))
        p = np.concatenate([t * 100.0, 1e-3 * rng.normal(0, 1, (P, 2))], 1)
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(p.astype(np.float32))


GOLDEN_CASES = {
    # name: (kind, P, seed)
    "p1": ("uniform", 1, 1), "p2": ("uniform", 2, 2), "p3": ("uniform", 3, 3), "p4": ("uniform", 4, 4),
    "p5": ("uniform", 5, 5), "p33": ("uniform", 33, 6), "p1025": ("uniform", 1025, 7),
    "uniform": ("uniform", 3000, 1337), "clustered": ("clustered", 2500, 11), "planar": ("planar", 2000, 5),
    "offset": ("offset", 2000, 23), "line": ("line", 1500, 9),
}
