"""Pins the CPU oracle against the reference's own golden vectors / known-answer tests
(SURVEY.md §8c).  CPU-only: `pytest -m "not gpu"`.
"""
import importlib
import os

import numpy as np
import pytest

import oracle_lib as orc

pkg_synth = importlib.import_module("teaser-plusplus_b200.synth")
GOLD = pkg_synth.GOLDEN_DIR


def read_csv(path):
    return np.loadtxt(path, delimiter=",", ndmin=2)


# ---------------------------------------------------------------- tls-test.cc:21-86
@pytest.mark.parametrize("X,r,ref,ref_inl", [
    ([0.5, 1, 0.6, 0.7, 1.2], [0.9, 0.9, 0.4, 0.5, 0.4], 0.8383, [1, 1, 1, 1, 1]),
    ([0.5, 1, 0.6, 0.7, 1.2, 10], [0.9, 0.9, 0.4, 0.5, 0.4, 0.5], 0.8383, [1, 1, 1, 1, 1, 0]),
    ([0.5, 1, 0.6, 20, 16, 10], [0.9, 0.9, 0.4, 0.5, 0.4, 0.5], 0.6425, [1, 1, 1, 0, 0, 0]),
])
def test_scalar_tls_kat(X, r, ref, ref_inl):
    est, inl = orc.scalar_tls(X, r)
    assert abs(est - ref) < 1e-3  # tolerance of tls-test.cc:38
    assert inl.tolist() == [bool(v) for v in ref_inl]


# ---------------------------------------------------------------- translation-solver-test.cc:21-113
def test_translation_kat():
    v1 = read_csv(os.path.join(GOLD, "registration_test", "translation_test_v1_inliers.csv")).T
    v2 = read_csv(os.path.join(GOLD, "registration_test", "translation_test_v2_inliers.csv")).T
    assert v1.shape == (34, 3) and v2.shape == (34, 3)
    t, inl = orc.tls_translation(v1, v2, 0.00673642835, 1.0)
    exp = np.array([-0.098430131086161, 0.008679113091532, 0.197317864174211])
    assert np.linalg.norm(t - exp) < 1e-5
    # unit shifts (translation-solver-test.cc:45-91)
    for ax in range(3):
        d = v1.copy()
        d[:, ax] += 1
        t, _ = orc.tls_translation(v1, d, 0.01, 1.0)
        e = np.zeros(3)
        e[ax] = 1
        assert np.linalg.norm(t - e) < 1e-5
    t, _ = orc.tls_translation(v1, v1, 0.01, 1.0)
    assert np.linalg.norm(t) < 1e-5


# ---------------------------------------------------------------- rotation-solver-test.cc:137-251
EXPECTED_R = np.array([[0.997379773225804, -0.019905935977315, -0.069551000516966],
                       [0.013777311189888, 0.996068297974922, -0.087510750572249],
                       [0.071019530105605, 0.086323226782879, 0.993732623426126]])


def test_gnc_tls_rotation_kat():
    src = read_csv(os.path.join(GOLD, "registration_test", "rotation_only_src.csv"))
    assert src.shape == (200, 3)
    dst = src @ EXPECTED_R.T
    out = orc.gnc_tls(src, dst, 100, 1e-12, 1.4, 1e-3)
    assert pkg_synth.angular_error(EXPECTED_R, out["R"]) < 1e-5
    # identity and axis rotations
    out = orc.gnc_tls(src, src, 100, 1e-12, 1.4, 1e-3)
    assert pkg_synth.angular_error(np.eye(3), out["R"]) < 1e-5
    for ax in range(3):
        a = 0.7
        c, s = np.cos(a), np.sin(a)
        R = np.eye(3)
        i, j = [(1, 2), (0, 2), (0, 1)][ax]
        R[i, i] = c; R[j, j] = c; R[i, j] = -s; R[j, i] = s
        out = orc.gnc_tls(src, src @ R.T, 100, 1e-12, 1.4, 1e-3)
        assert pkg_synth.angular_error(R, out["R"]) < 1e-5


def test_fgr_rotation_kat():
    src = read_csv(os.path.join(GOLD, "registration_test", "rotation_only_src.csv"))
    dst = src @ EXPECTED_R.T
    out = orc.fgr(src, dst, 100, 1e-12, 1.4, 1e-3)  # rotation-solver-test.cc:23-135 uses the same Params
    assert pkg_synth.angular_error(EXPECTED_R, out["R"]) < 1e-5


def test_quatro_rotation_kat():  # registration-test.cc:183-218
    src = read_csv(os.path.join(GOLD, "registration_test", "rotation_only_src.csv"))
    Ryaw = np.array([[0.997379773225804, -0.072343541246221, 0.0],
                     [0.072343541246221, 0.997379773225804, 0.0],
                     [0.0, 0.0, 1.0]])
    out = orc.quatro(src, src @ Ryaw.T, 100, 0.005, 1.4, 0.0067364)
    assert pkg_synth.angular_error(Ryaw, out["R"]) < 1e-5
    # yaw + planted outliers: still recovers the yaw
    rng = np.random.default_rng(3)
    dst = src @ Ryaw.T
    dst[:40] += rng.normal(size=(40, 3))
    out = orc.quatro(src, dst, 100, 1e-9, 1.4, 0.01)
    assert pkg_synth.angular_error(Ryaw, out["R"]) < 1e-3 and out["iterations"] > 2


def test_svd3_matches_numpy():
    rng = np.random.default_rng(0)
    for k in range(200):
        H = rng.normal(size=(3, 3))
        if k % 5 == 0:
            H[:, 2] = H[:, 0] * 0.3 + H[:, 1]  # rank deficient
        if k % 7 == 0:
            H = H * 1e-9
        U, S, V = orc.svd3(H)
        assert np.allclose(U @ np.diag(S) @ V.T, H, atol=1e-13 * max(1, np.abs(H).max()) + 1e-22)
        assert np.allclose(U.T @ U, np.eye(3), atol=1e-13)
        assert np.allclose(V.T @ V, np.eye(3), atol=1e-13)
        assert S[0] >= S[1] >= S[2] >= 0
        assert np.allclose(S, np.linalg.svd(H, compute_uv=False), rtol=1e-12, atol=1e-13 * S[0])


# ---------------------------------------------------------------- scale-solver-test.cc:23-130
def test_scale_solvers_kat():
    obj = read_csv(os.path.join(GOLD, "registration_test", "objectIn.csv")).T
    assert obj.shape == (168, 3)
    tims, mp = orc.compute_tims(obj)
    K = 168 * 167 // 2
    assert tims.shape == (K, 3)
    # TIM ordering registration.cc:531: k = i*N - i(i+1)/2 + (j-i-1)
    assert mp[0].tolist() == [0, 1] and mp[166].tolist() == [0, 167] and mp[167].tolist() == [1, 2]
    assert np.array_equal(tims[167], obj[2] - obj[1])
    s, mask = orc.tls_scale_solver(tims, tims, 0.01)
    assert abs(s - 1) < 1e-5
    for k in (2.0, 0.5, 3.7):
        s, mask = orc.tls_scale_solver(tims, tims * k, 0.01)
        assert abs(s - k) < 1e-5
    s, mask = orc.scale_inliers_selector(tims, tims, 0.01)
    assert s == 1 and mask.all()
    s, mask = orc.scale_inliers_selector(tims, tims * 100, 1e-4)
    assert s == 1 and not mask.any()
    d = tims.copy()
    d[5] *= 50
    s, mask = orc.scale_inliers_selector(tims, d, 1e-4)
    assert (~mask).sum() == 1 and not mask[5]


# ---------------------------------------------------------------- registration-test.cc:107-142 / 256-392
def test_object_scene_scale():
    obj = read_csv(os.path.join(GOLD, "registration_test", "objectIn.csv")).T
    scene = read_csv(os.path.join(GOLD, "registration_test", "sceneIn.csv")).T
    st, _ = orc.compute_tims(obj)
    dt, _ = orc.compute_tims(scene)
    s, _ = orc.tls_scale_solver(st, dt, 0.0067364)
    assert abs(s - 0.955885) < 0.01  # registration-test.cc:139
    p = orc.default_params(noise_bound=0.0067364, estimate_scaling=1, rotation_estimation_algorithm=1,
                           rotation_cost_threshold=0.005)
    out = orc.solve(obj, scene, p)
    assert out["valid"]
    assert abs(out["scale"] - 0.955885) < 1e-4  # registration-test.cc:313
    Rexp = np.array([[0.9974, -0.0199, -0.0696], [0.0138, 0.9961, -0.0875], [0.0710, 0.0863, 0.9937]])
    texp = np.array([-0.1011, 0.0908, 0.1344])
    assert pkg_synth.angular_error(Rexp, out["R"]) < 0.25
    assert np.linalg.norm(out["t"] - texp) < 0.15
    # fixed scale (registration-test.cc:325-392)
    p = orc.default_params(noise_bound=0.0067364, estimate_scaling=0, rotation_estimation_algorithm=1,
                           rotation_cost_threshold=0.005)
    out = orc.solve(obj, scene, p)
    assert out["scale"] == 1
    assert pkg_synth.angular_error(Rexp, out["R"]) < 0.2
    assert np.linalg.norm(out["t"] - texp) < 0.1


# ---------------------------------------------------------------- registration-benchmark.cc:276-374
BENCH_TOL = {  # (s_ref, R_ref, t_ref, s_est, R_est, t_est)
    1: (1e-5,) * 6, 2: (1e-5,) * 6, 3: (1e-5,) * 6, 4: (1e-5,) * 6, 5: (1e-5,) * 6,
    6: (1e-2, 1e-2, 2e-2, 1e-5, 1e-3, 1e-3),
}
EXPECTED_CLIQUE = {1: 10, 2: 9, 3: 7, 4: 25, 5: 10, 6: 10}


def load_benchmark(i):
    d = os.path.join(GOLD, f"benchmark_{i}")
    src = pkg_synth.read_ply_vertices(os.path.join(d, "src.ply"))
    dst = pkg_synth.read_ply_vertices(os.path.join(d, "dst.ply"))
    nb = None
    for line in open(os.path.join(d, "parameters.txt")):
        if line.startswith("Noise Bound"):
            nb = float(line.split(":")[1])
    g = {k: read_csv(os.path.join(d, f"{k}.csv")) for k in ("R_ref", "R_est", "t_ref", "t_est", "s_ref", "s_est")}
    return src, dst, nb, g


@pytest.mark.parametrize("i", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("alg", [0, 1])
def test_benchmark_fixture(i, alg):
    src, dst, nb, g = load_benchmark(i)
    p = orc.default_params(noise_bound=nb, cbar2=1, estimate_scaling=1, rotation_max_iterations=100,
                           rotation_gnc_factor=1.4, rotation_estimation_algorithm=alg,
                           rotation_cost_threshold=1e-12 if alg == 0 else 0.005)
    out = orc.solve(src, dst, p)
    assert out["valid"]
    tol = BENCH_TOL[i]
    assert abs(out["scale"] - g["s_ref"].item()) <= tol[0]
    assert pkg_synth.angular_error(g["R_ref"], out["R"]) <= tol[1]
    assert np.linalg.norm(out["t"] - g["t_ref"].ravel()) <= tol[2]
    assert abs(out["scale"] - g["s_est"].item()) <= tol[3]
    assert pkg_synth.angular_error(g["R_est"], out["R"]) <= tol[4]
    assert np.linalg.norm(out["t"] - g["t_est"].ravel()) <= tol[5]
    assert len(out["clique"]) == EXPECTED_CLIQUE[i]
