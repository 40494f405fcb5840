"""Synthetic correspondence-set generators for the BASELINE.json configs.

All generators are numpy-only and seeded, so the GPU path, the oracle and the benchmark see the
same bytes.  Points are returned as (N, 3) float64 C-contiguous arrays, which is byte-identical to
the reference's column-major Eigen::Matrix<double,3,Dynamic> (xyzxyz...), registration.cc:568-570.
"""
from __future__ import annotations

import os
import numpy as np

NOISE_BOUND_SIGMA_001 = 0.033682  # test/benchmark/data/benchmark_6/parameters.txt (sigma=0.01)


def random_rotation(rng: np.random.Generator) -> np.ndarray:
    """Uniform random rotation from a uniform unit quaternion."""
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ]
    )


def _truncated_gaussian(rng, n, sigma, bound):
    """N(0, sigma^2 I) rejected to ||eps|| <= bound (keeps every inlier pair an edge)."""
    eps = rng.normal(scale=sigma, size=(n, 3))
    bad = np.linalg.norm(eps, axis=1) > bound
    while bad.any():
        eps[bad] = rng.normal(scale=sigma, size=(int(bad.sum()), 3))
        bad = np.linalg.norm(eps, axis=1) > bound
    return eps


def make_problem(n: int, outlier_ratio: float, seed: int, model: str = "ball", sigma: float = 0.01,
                 noise_bound: float | None = None, extent=(1.0, 1.0, 1.0)):
    """One registration problem.

    model: "ball"   — outlier dst uniform in a radius-5 ball (TEASER-paper style; primary for C2/C4)
           "incube" — outlier dst = R u + t with u ~ U(extent) (wrong matches inside the object; C3)
           "permute"— outlier dst = R src[pi(i)] + t + noise (wrong matches within the cloud; C5)
    Returns dict(src, dst, R, t, inliers(sorted idx), noise_bound).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    if noise_bound is None:
        noise_bound = NOISE_BOUND_SIGMA_001 * (sigma / 0.01)
    ext = np.asarray(extent, dtype=np.float64)
    src = rng.uniform(size=(n, 3)) * ext
    R = random_rotation(rng)
    t = rng.uniform(-1, 1, size=3)
    dst = src @ R.T + t + _truncated_gaussian(rng, n, sigma, noise_bound)
    n_out = int(round(outlier_ratio * n))
    perm = rng.permutation(n)
    out_idx = np.sort(perm[:n_out])
    if model == "ball":
        d = rng.normal(size=(n_out, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        r = 5.0 * rng.uniform(size=(n_out, 1)) ** (1.0 / 3.0)
        dst[out_idx] = d * r
    elif model == "incube":
        u = rng.uniform(size=(n_out, 3)) * ext
        dst[out_idx] = u @ R.T + t
    elif model == "permute":
        pick = rng.integers(0, n - 1, size=n_out)
        pick = pick + (pick >= out_idx)  # never the true match
        dst[out_idx] = src[pick] @ R.T + t + _truncated_gaussian(rng, n_out, sigma, noise_bound)
    else:
        raise ValueError(model)
    inl = np.sort(perm[n_out:])
    return dict(src=np.ascontiguousarray(src), dst=np.ascontiguousarray(dst), R=R, t=t,
                inliers=inl.astype(np.int32), noise_bound=float(noise_bound))


def config_problem(cfg: str, b: int = 0, n: int | None = None):
    """The five BASELINE.json configs. `n` overrides the size for scaled-down tests."""
    if cfg == "C2":
        n = n or 5000
        return make_problem(n, 0.95, 5000 * 1000 + b, "ball")
    if cfg == "C2scale":  # the unknown-scale line: see bench.py CONFIGS for why 80 %
        n = n or 5000
        return make_problem(n, 0.80, 5000 * 1000 + b, "ball")
    if cfg == "C2cube":
        n = n or 5000
        return make_problem(n, 0.95, 5000 * 1000 + b, "incube")
    if cfg == "C3":
        n = n or 10000
        return make_problem(n, 0.99, 10000 * 1000 + b, "incube")
    if cfg == "C3ball":
        n = n or 10000
        return make_problem(n, 0.99, 10000 * 1000 + b, "ball")
    if cfg == "C4":
        n = n or 2000
        return make_problem(n, 0.90, 2000 * 1000 + b, "ball")
    if cfg == "C5":
        n = n or 8000
        return make_problem(n, 0.97, 8000 * 1000 + b, "permute", sigma=0.015, noise_bound=0.05,
                            extent=(3.0, 3.0, 2.0))
    raise ValueError(cfg)


def read_ply_vertices(path: str) -> np.ndarray:
    """Minimal ASCII-PLY vertex reader (first three floats per vertex line), float32 like the
    reference's tinyply-based reader (teaser/src/ply_io.cc:28-112), widened to float64."""
    with open(path, "r") as f:
        nv = None
        line = f.readline()
        assert line.strip() == "ply", path
        while True:
            line = f.readline()
            if not line:
                raise ValueError("bad ply header")
            tok = line.split()
            if tok[:2] == ["element", "vertex"]:
                nv = int(tok[2])
            if tok and tok[0] == "end_header":
                break
        pts = np.empty((nv, 3), dtype=np.float32)
        for i in range(nv):
            tok = f.readline().split()
            pts[i] = (float(tok[0]), float(tok[1]), float(tok[2]))
    return pts.astype(np.float64)


def bunny_problem(ply_path: str, seed: int = 1889, n_outlier_draws: int = 1700, noise_bound: float = 0.001):
    """BASELINE config C1: examples/teaser_cpp_ply/teaser_cpp_ply.cc:21-40,62-75 with a fixed seed."""
    src = read_ply_vertices(ply_path)
    n = src.shape[0]
    T = np.array([[9.96926560e-01, 6.68735757e-02, -4.06664421e-02, -1.15576939e-01],
                  [-6.61289946e-02, 9.97617877e-01, 1.94008687e-02, -3.87705398e-02],
                  [4.18675510e-02, -1.66517807e-02, 9.98977765e-01, 1.14874890e-01],
                  [0, 0, 0, 1]])
    rng = np.random.Generator(np.random.PCG64(seed))
    dst = src @ T[:3, :3].T + T[:3, 3]
    dst = dst + rng.uniform(-1, 1, size=(n, 3)) * (noise_bound / 2)
    untouched = np.ones(n, dtype=bool)
    for _ in range(n_outlier_draws):
        c = int(rng.integers(0, n))
        dst[c] += float(rng.integers(5, 11))
        untouched[c] = False
    return dict(src=np.ascontiguousarray(src), dst=np.ascontiguousarray(dst), R=T[:3, :3].copy(), t=T[:3, 3].copy(),
                inliers=np.nonzero(untouched)[0].astype(np.int32), noise_bound=noise_bound)


def angular_error(Ra: np.ndarray, Rb: np.ndarray) -> float:
    """test/test-tools/test_utils.h:92-94"""
    c = (np.trace(Ra.T @ Rb) - 1) / 2
    return float(abs(np.arccos(min(max(c, -1.0), 1.0))))


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


# ---------------------------------------------------------------------------------------------
# Matcher inputs (Matcher::calculateCorrespondences, matcher.cc:21-53): points + per-point descriptors
# ---------------------------------------------------------------------------------------------
def random_fpfh(rng: np.random.Generator, n: int, dim: int = 33) -> np.ndarray:
    """FPFH-shaped descriptors: non-negative float32, each 11-bin sub-histogram sums to 100 (pcl::FPFHSignature33)."""
    f = rng.gamma(0.6, 1.0, size=(n, dim))
    for lo in range(0, dim, 11):
        blk = f[:, lo:lo + 11]
        blk *= 100.0 / np.maximum(blk.sum(axis=1, keepdims=True), 1e-12)
    return f.astype(np.float32)


def matcher_problem(ns: int, nd: int, n_common: int, seed: int, dim: int = 33, feat_noise: float = 0.5,
                    point_noise: float = 0.002):
    """Two clouds that share n_common physical points.  Returns dict(src_pts, dst_pts (float32 (n,3)), src_feat,
    dst_feat (float32 (n,dim)), R, t, true_pairs (n_common,2) int32 sorted by source index).  The shared points carry
    the same descriptor up to feat_noise; the rest get unrelated descriptors."""
    rng = np.random.default_rng(seed)
    R = random_rotation(rng)
    t = rng.uniform(-1, 1, size=3)
    src = rng.uniform(0, 1, size=(ns, 3))
    src_feat = random_fpfh(rng, ns, dim)
    dst = R @ rng.uniform(0, 1, size=(nd, 3)).T
    dst = dst.T + t
    dst_feat = random_fpfh(rng, nd, dim)
    si = np.sort(rng.permutation(ns)[:n_common])
    di = rng.permutation(nd)[:n_common]
    dst[di] = (R @ src[si].T).T + t + rng.normal(scale=point_noise, size=(n_common, 3))
    dst_feat[di] = np.maximum(src_feat[si] + rng.normal(scale=feat_noise, size=(n_common, dim)), 0).astype(np.float32)
    pairs = np.stack([si, di], axis=1).astype(np.int32)
    return dict(src_pts=src.astype(np.float32), dst_pts=dst.astype(np.float32), src_feat=src_feat, dst_feat=dst_feat,
                R=R, t=t, true_pairs=pairs)


def read_pcd_ascii(path: str) -> np.ndarray:
    """x y z rows of an ASCII .pcd (test/teaser/data/bunny.pcd)."""
    rows, data = [], False
    with open(path) as f:
        for line in f:
            if data:
                v = line.split()
                if len(v) >= 3:
                    rows.append([float(v[0]), float(v[1]), float(v[2])])
            elif line.startswith("DATA"):
                data = True
    return np.asarray(rows, dtype=np.float32)


def bunny_fpfh():
    """The reference's FPFH fixture: 397 bunny points and their 397x33 PCL descriptors (feature-test.cc:52-90)."""
    pts = read_pcd_ascii(os.path.join(GOLDEN_DIR, "bunny.pcd"))
    feat = np.loadtxt(os.path.join(GOLDEN_DIR, "bunny_fpfh.csv"), dtype=np.float32).reshape(-1, 33)
    assert pts.shape[0] == feat.shape[0]
    return pts, feat
