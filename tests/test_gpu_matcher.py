"""GPU parity tests of the correspondence-generation stage (`pytest -m gpu`): tzr_feature_nn / tzr_match_correspondences
through the C-ABI against oracle/matcher_oracle.cc (reference teaser/src/matcher.cc:21-337).  Index work: bit-exact."""
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


@pytest.mark.parametrize("nq,ndb,dim", [(257, 301, 33), (64, 64, 4), (1, 1, 33), (5, 1000, 1), (1000, 5, 3),
                                        (700, 1900, 7), (3000, 3000, 33), (130, 4097, 64), (65, 129, 128),
                                        (2, 70000, 33)])
def test_feature_nn_bit_exact(ctx, nq, ndb, dim):
    rng = np.random.default_rng(nq * 7 + ndb + dim)
    q = synth.random_fpfh(rng, nq, dim) if dim >= 11 else rng.normal(size=(nq, dim)).astype(np.float32)
    db = synth.random_fpfh(rng, ndb, dim) if dim >= 11 else rng.normal(size=(ndb, dim)).astype(np.float32)
    idx, dist = ctx.feature_nn(q, db)
    want = orc.nn1(q, db)
    assert np.array_equal(idx, want)
    # the reported distance is the flann::L2<float> value of the winner (float32 accumulation, groups of four)
    d64 = ((q.astype(np.float64) - db[idx].astype(np.float64)) ** 2).sum(1)
    assert np.allclose(dist, d64, rtol=1e-5, atol=1e-30)


def test_feature_nn_ties_pick_lowest_index(ctx):
    rng = np.random.default_rng(2)
    base = synth.random_fpfh(rng, 300, 33)
    db = np.concatenate([base, base[::-1], base])  # every descriptor three times, spread over several segments
    idx, dist = ctx.feature_nn(base, db)
    want = orc.nn1(base, db)
    assert np.array_equal(idx, want)
    assert (dist == 0).all()
    first = np.minimum(np.arange(300), 599 - np.arange(300))
    assert np.array_equal(idx, first)


def test_feature_nn_nan_rows_never_win(ctx):
    rng = np.random.default_rng(4)
    q = synth.random_fpfh(rng, 100, 33)
    db = synth.random_fpfh(rng, 500, 33)
    db[0, 3] = np.nan
    db[77, :] = np.nan
    idx, _ = ctx.feature_nn(q, db)
    assert np.array_equal(idx, orc.nn1(q, db))
    assert 0 not in idx and 77 not in idx


def test_self_matching_bunny_fpfh(ctx):
    """matcher-test.cc:17-39 (SelfMatching) on the reference's PCL descriptors of bunny.pcd."""
    pts, feat = synth.bunny_fpfh()
    pairs = ctx.match_correspondences(pts, pts, feat, feat, False, True, False, 0)
    assert pairs.shape == (pts.shape[0], 2)
    assert (pairs[:, 0] == pairs[:, 1]).all()
    assert np.array_equal(pairs, orc.match_correspondences(pts, pts, feat, feat, False, True, False, 0))


FLAG_CASES = [(abs_s, cc, tt, ts) for abs_s in (False, True) for cc in (False, True)
              for (tt, ts) in ((False, 0.0), (True, 0.95), (True, 0.0), (False, 0.95))]


@pytest.mark.parametrize("ns,nd,nc", [(900, 700, 300), (700, 900, 300), (64, 64, 64), (1000, 1000, 0), (3, 2, 2)])
@pytest.mark.parametrize("abs_scale,crosscheck,tuple_test,tuple_scale", FLAG_CASES)
def test_match_correspondences_equals_oracle(ctx, ns, nd, nc, abs_scale, crosscheck, tuple_test, tuple_scale):
    mp = synth.matcher_problem(ns, nd, nc, seed=ns + 3 * nd + nc)
    args = (mp["src_pts"], mp["dst_pts"], mp["src_feat"], mp["dst_feat"], abs_scale, crosscheck, tuple_test,
            tuple_scale)
    got, g = ctx.match_correspondences(*args, tuple_seed=42, return_scale=True)
    want, go = orc.match_correspondences(*args, tuple_seed=42, return_scale=True)
    assert g == go                      # Matcher::global_scale_ (float, sequential mean): bit-exact
    assert np.array_equal(got, want)    # sorted unique index pairs: bit-exact


@pytest.mark.parametrize("seed", [0, 1, 2**40 + 17])
def test_tuple_test_seeded_and_equal_to_oracle_full_size(ctx, seed):
    mp = synth.matcher_problem(5000, 4500, 1500, seed=77, feat_noise=0.3)
    args = (mp["src_pts"], mp["dst_pts"], mp["src_feat"], mp["dst_feat"], False, True, True, 0.95)
    got = ctx.match_correspondences(*args, tuple_seed=seed)
    want = orc.match_correspondences(*args, tuple_seed=seed)
    assert np.array_equal(got, want)
    truth = set(map(tuple, mp["true_pairs"]))
    assert truth <= set(map(tuple, got))


def test_matcher_then_solve_recovers_transform(ctx):
    """The examples' flow (examples/teaser_cpp_fpfh/teaser_cpp_fpfh.cc:86-106): correspondences from the matcher,
    gathered by index (registration.cc:553-566), then solve()."""
    mp = synth.matcher_problem(4000, 4000, 600, seed=5, feat_noise=0.3, point_noise=0.002)
    pairs = ctx.match_correspondences(mp["src_pts"], mp["dst_pts"], mp["src_feat"], mp["dst_feat"], False, True, False,
                                      0.95)
    assert np.array_equal(pairs, orc.match_correspondences(mp["src_pts"], mp["dst_pts"], mp["src_feat"],
                                                           mp["dst_feat"], False, True, False, 0.95))
    src = mp["src_pts"][pairs[:, 0]].astype(np.float64)
    dst = mp["dst_pts"][pairs[:, 1]].astype(np.float64)
    p = capi.default_params(noise_bound=0.01, cbar2=1.0, estimate_scaling=0, rotation_cost_threshold=1e-12)
    res = ctx.solve(src, dst, p)
    assert res["valid"]
    assert synth.angular_error(res["R"], mp["R"]) < 5e-3
    assert np.linalg.norm(res["t"] - mp["t"]) < 5e-3
    osol = orc.solve(src, dst, orc.default_params(noise_bound=0.01, cbar2=1.0, estimate_scaling=0,
                                                  rotation_cost_threshold=1e-12))
    assert np.array_equal(np.sort(res["clique"]), np.sort(osol["clique"]))
    assert synth.angular_error(res["R"], osol["R"]) < 1e-4 and np.linalg.norm(res["t"] - osol["t"]) < 1e-4


def test_match_rejects_bad_arguments(ctx):
    mp = synth.matcher_problem(10, 10, 5, seed=1)
    with pytest.raises(capi.TzrError):
        ctx.match_correspondences(mp["src_pts"], mp["dst_pts"], mp["src_feat"][:, :5], mp["dst_feat"], True, True)
    big = np.zeros((10, 129), dtype=np.float32)
    with pytest.raises(capi.TzrError):
        ctx.match_correspondences(mp["src_pts"], mp["dst_pts"], big, big, True, True)


def test_cpp_fpfh_example_through_facade():
    """host/examples/teaser_cpp_fpfh.cc: teaser::Matcher + solve(PointCloud, PointCloud, correspondences) via the
    C++ façade (reference flow: examples/teaser_cpp_fpfh/teaser_cpp_fpfh.cc:91-113)."""
    import os
    import subprocess
    host = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "teaser-plusplus_b200", "host")
    subprocess.check_call(["make", "-s", "-C", host])
    out = subprocess.run([os.path.join(host, "example_cpp_fpfh"), os.path.join(synth.GOLDEN_DIR, "bunny.pcd"),
                          os.path.join(synth.GOLDEN_DIR, "bunny_fpfh.csv")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    vals = {l.split(":")[0]: float(l.split(":")[1]) for l in out.stdout.strip().splitlines()}
    assert vals["correct correspondences"] >= 0.6 * 397
    assert vals["clique size"] >= 0.9 * vals["correct correspondences"]
    assert vals["rotation error (rad)"] < 0.01 and vals["translation error (m)"] < 0.005
