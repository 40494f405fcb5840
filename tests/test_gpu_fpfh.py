"""GPU parity tests of descriptor estimation (`pytest -m gpu`): tzr_compute_fpfh through the C-ABI against
oracle/fpfh_oracle.cc (reference teaser/src/fpfh.cc:15-43 = PCL normals + FPFH).  Both sides use the same float
operation sequences (including the elementary functions), so the comparison is bit-exact."""
import importlib

import numpy as np
import pytest

import oracle_lib as orc

capi = importlib.import_module("teaser-plusplus_b200.capi")
synth = importlib.import_module("teaser-plusplus_b200.synth")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


def test_bunny_fpfh_bit_exact_and_golden(ctx):
    pts, ref = synth.bunny_fpfh()
    got, nrm = ctx.compute_fpfh(pts, 0.03, 0.05, return_normals=True)
    want, wn = orc.compute_fpfh(pts, 0.03, 0.05)
    assert _same(nrm, wn)
    assert _same(got, want)
    d = np.abs(got - ref)   # the reference's golden vector (feature-test.cc:52-90), see tests/test_fpfh_cpu.py
    assert (d > 1e-4).mean() < 0.025 and np.median(d) < 1e-5


@pytest.mark.parametrize("n,rn,rf,seed", [(2000, 0.08, 0.12, 1), (5000, 0.05, 0.08, 2), (700, 0.3, 0.2, 3),
                                          (3, 0.5, 0.5, 4), (1, 0.1, 0.1, 5)])
def test_random_surface_bit_exact(ctx, n, rn, rf, seed):
    rng = np.random.default_rng(seed)
    uv = rng.uniform(-1, 1, size=(n, 2))
    pts = np.stack([uv[:, 0], uv[:, 1], 0.3 * np.sin(3 * uv[:, 0]) * np.cos(2 * uv[:, 1])], axis=1)
    pts = (pts + rng.normal(scale=0.003, size=pts.shape) + np.array([0.5, -0.2, 2.0])).astype(np.float32)
    got, nrm = ctx.compute_fpfh(pts, rn, rf, return_normals=True)
    want, wn = orc.compute_fpfh(pts, rn, rf)
    assert _same(nrm, wn)
    assert _same(got, want)


@pytest.mark.parametrize("n,rn,rf,seed", [(4096, 0.06, 0.09, 11), (20000, 0.02, 0.03, 12), (9000, 0.5, 0.7, 13)])
def test_grid_search_path_bit_exact(ctx, n, rn, rf, seed):
    """n >= 4096 takes the hashed-grid radius search: same neighbour sets, so still bit-exact vs the brute-force
    restatement — including a cloud with NaN / inf / far-away points and a radius comparable to the extent."""
    rng = np.random.default_rng(seed)
    uv = rng.uniform(-1, 1, size=(n, 2))
    pts = np.stack([uv[:, 0], uv[:, 1], 0.3 * np.sin(3 * uv[:, 0]) * np.cos(2 * uv[:, 1])], axis=1)
    pts = (pts + rng.normal(scale=0.003, size=pts.shape) + np.array([-0.5, 0.2, 2.0])).astype(np.float32)
    if seed == 13:
        pts = pts[:, [2, 0, 1]].copy()
        pts[:2000] *= np.float32(0.3)       # dense core: up to a few thousand neighbours, fewer than the 4096 cap
    pts[17] = [np.nan, 0, 0]
    pts[18] = [np.inf, 0, 0]
    pts[19] = [1e6, -1e6, 1e6]
    got, nrm = ctx.compute_fpfh(pts, rn, rf, return_normals=True)
    want, wn = orc.compute_fpfh(pts, rn, rf)
    assert _same(nrm, wn)
    assert _same(got, want)


def test_duplicates_isolated_points_and_nan_inputs(ctx):
    rng = np.random.default_rng(9)
    pts = rng.uniform(0, 1, size=(600, 3)).astype(np.float32)
    pts[10] = pts[11]                 # duplicate point: zero distance, skipped by the pair features and the weights
    pts[20] = [50, 50, 50]            # isolated: NaN normal, empty histogram
    pts[30] = [np.nan, 0, 0]          # NaN coordinate: not even its own neighbour
    got, nrm = ctx.compute_fpfh(pts, 0.15, 0.2, return_normals=True)
    want, wn = orc.compute_fpfh(pts, 0.15, 0.2)
    assert _same(nrm, wn) and _same(got, want)
    assert np.isnan(nrm[20]).all() and (got[20] == 0).all()


def test_too_many_neighbours_is_an_error(ctx):
    pts = np.random.default_rng(0).uniform(0, 0.01, size=(5000, 3)).astype(np.float32)
    with pytest.raises(capi.TzrError):
        ctx.compute_fpfh(pts, 1.0, 1.0)


def test_fpfh_matcher_solve_pipeline_on_bunny(ctx):
    """computeFPFHFeatures -> calculateCorrespondences -> solve, all three on the device
    (examples/teaser_cpp_fpfh/teaser_cpp_fpfh.cc:86-106)."""
    pts, _ = synth.bunny_fpfh()
    rng = np.random.default_rng(3)
    # the example's transform (teaser_cpp_fpfh.cc:64-69).  PCL orients normals towards the origin, so FPFH is only
    # pose-invariant while the cloud keeps its side of the viewpoint; a large motion flips normals and thins the
    # matches (35 of 397 for a random rotation with a 0.2 m shift).
    R = np.array([[9.96926560e-01, 6.68735757e-02, -4.06664421e-02],
                  [-6.61289946e-02, 9.97617877e-01, 1.94008687e-02],
                  [4.18675510e-02, -1.66517807e-02, 9.98977765e-01]])
    t = np.array([-1.15576939e-01, -3.87705398e-02, 1.14874890e-01])
    perm = rng.permutation(len(pts))
    dst = ((R @ pts[perm].astype(np.float64).T).T + t + rng.uniform(-1, 1, size=pts.shape) * 1e-4).astype(np.float32)
    fs = ctx.compute_fpfh(pts, 0.03, 0.05)
    fd = ctx.compute_fpfh(dst, 0.03, 0.05)
    pairs = ctx.match_correspondences(pts, dst, fs, fd, False, True, False, 0.95)
    right = (perm[pairs[:, 1]] == pairs[:, 0]).sum()
    assert len(pairs) > 300 and right > 0.9 * len(pairs)
    res = ctx.solve(pts[pairs[:, 0]].astype(np.float64), dst[pairs[:, 1]].astype(np.float64),
                    capi.default_params(noise_bound=0.001, cbar2=1.0, estimate_scaling=0,
                                        rotation_cost_threshold=0.005))
    assert res["valid"]
    assert synth.angular_error(res["R"], R) < 0.01 and np.linalg.norm(res["t"] - t) < 0.002


def test_cpp_fpfh_example_with_computed_descriptors():
    """host/examples/teaser_cpp_fpfh.cc without a descriptor file: teaser::FPFHEstimation -> teaser::Matcher ->
    RobustRegistrationSolver::solve, the reference example's flow (teaser_cpp_fpfh.cc:86-113) through the façade."""
    import os
    import subprocess
    host = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "teaser-plusplus_b200", "host")
    subprocess.check_call(["make", "-s", "-C", host])
    out = subprocess.run([os.path.join(host, "example_cpp_fpfh"), os.path.join(synth.GOLDEN_DIR, "bunny.pcd")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    vals = {l.split(":")[0]: float(l.split(":")[1]) for l in out.stdout.strip().splitlines()}
    assert vals["correct correspondences"] >= 30
    assert vals["rotation error (rad)"] < 0.03 and vals["translation error (m)"] < 0.01
