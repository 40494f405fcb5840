"""The certifier restatement (oracle/certifier_oracle.py; reference teaser/src/certification.cc) against the
reference's own fixtures, at the reference test's tolerance (certification-test.cc:29 ACCEPTABLE_ERROR = 1e-7)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import certifier_oracle as co  # noqa: E402
import certifier_fixtures as cf  # noqa: E402

TOL = 1e-7
SMALL = cf.cases("small")
LARGE = cf.cases("large")


@pytest.mark.parametrize("c", SMALL, ids=[c["name"] for c in SMALL])
def test_building_blocks_match_reference_fixtures(c):
    N = c["v1"].shape[1]
    nb, cb = c["params"]["noise_bound"], c["params"]["cbar2"]
    q = c["q_est"]
    assert np.abs(co.get_omega1(q) - c["omega"]).max() < TOL                                  # :355-369 GetOmega1
    assert np.abs(co.get_block_diag_omega(4 * N + 4, q) - c["block_diag_omega"]).max() < TOL  # :371-388
    assert np.abs(co.get_q_cost(c["v1"], c["v2"], nb, cb) - c["Q_cost"]).max() < TOL          # :390-405 GetQCost
    lam = co.get_lambda_guess(c["R_est"], c["theta_est"], c["v1"], c["v2"], nb, cb)
    assert np.abs(lam - c["lambda_bar_init"]).max() < TOL                                     # :407-423
    thp = np.concatenate([[1.0], c["theta_est"]])
    A = co.get_linear_projection(thp)
    assert np.abs(A - c["A_inv"]).max() < TOL                                                 # :425-446
    Wd = co.get_optimal_dual_projection(c["W_1st_iter"], thp, A)
    assert np.abs(Wd - c["W_dual_1st_iter"]).max() < TOL                                      # :448-482
    Wd2 = co.get_optimal_dual_projection(c["W_1st_iter"], thp, None)                          # implicit A_inv
    assert np.abs(Wd2 - c["W_dual_1st_iter"]).max() < TOL
    gap = co.compute_suboptimality_gap(c["M_affine_1st_iter"], c["mu"], N)
    assert abs(gap - c["suboptimality_1st_iter"].ravel()[0]) < TOL                            # :484-497
    # the quaternion of R_est is the stored q_est up to sign
    qq = co.quaternion_from_rotation(c["R_est"])
    assert min(np.abs(qq - q).max(), np.abs(qq + q).max()) < 1e-6


@pytest.mark.parametrize("c", SMALL + LARGE[:1], ids=[c["name"] for c in SMALL + LARGE[:1]])
def test_certify_trajectory_matches_reference(c):
    """Certify / LargeInstance (certification-test.cc:499-527): same length, every value within 1e-7."""
    r = co.certify(c["R_est"], c["v1"], c["v2"], c["theta_est"], c["params"]["noise_bound"], c["params"]["cbar2"],
                   max_iterations=c["params"].get("max_iterations", 200), return_intermediates=True)
    want = c["suboptimality_traj"]
    assert len(r["suboptimality_traj"]) == len(want)
    assert np.abs(r["suboptimality_traj"] - want).max() < TOL
    assert abs(r["best_suboptimality"] - want.min()) < TOL
    assert r["is_optimal"] == bool(want.min() < 1e-3)
    if "mu" in c:
        assert abs(r["mu"] - c["mu"]) < TOL


def test_implicit_projection_equals_explicit_matrix():
    rng = np.random.default_rng(0)
    for N in (2, 3, 7, 12):
        th = np.concatenate([[1.0], rng.choice([-1.0, 1.0], size=N - 1)])
        A = co.get_linear_projection(th)
        b = rng.normal(size=(N * (N - 1) // 2, 3))
        assert np.abs(A @ b - co.apply_linear_projection(th, b)).max() < 1e-12
        assert np.abs(A - A.T).max() < 1e-15


def test_bool_theta_is_mapped_to_plus_minus_one():
    c = SMALL[0]
    r1 = co.certify(c["R_est"], c["v1"], c["v2"], c["theta_est"], c["params"]["noise_bound"], 1.0)
    r2 = co.certify(c["R_est"], c["v1"], c["v2"], c["theta_est"] > 0, c["params"]["noise_bound"], 1.0)
    assert np.array_equal(r1["suboptimality_traj"], r2["suboptimality_traj"])
