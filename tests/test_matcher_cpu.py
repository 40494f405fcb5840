"""Oracle-level tests of the Matcher restatement (oracle/matcher_oracle.cc; reference teaser/src/matcher.cc:21-337).

The reference's own matcher tests (test/teaser/matcher-test.cc:17-78) recompute FPFH with PCL, which this image does
not have; what can be carried over is their *semantics*: self-matching returns the identity (SelfMatching, :17-39),
run here on the reference's real PCL descriptors of bunny.pcd (test/teaser/data/bunny_fpfh.csv).
"""
import importlib

import numpy as np
import pytest

import oracle_lib as o

synth = importlib.import_module("teaser-plusplus_b200.synth")


def brute_nn(q, db):
    d = ((q[:, None, :].astype(np.float64) - db[None, :, :].astype(np.float64)) ** 2).sum(-1)
    return d.argmin(1), np.sort(d, axis=1)


def test_nn1_matches_numpy_bruteforce():
    rng = np.random.default_rng(0)
    for dim in (1, 3, 4, 7, 33, 64):
        q = rng.normal(size=(257, dim)).astype(np.float32)
        db = rng.normal(size=(301, dim)).astype(np.float32)
        got = o.nn1(q, db)
        want, srt = brute_nn(q, db)
        clear = (srt[:, 1] - srt[:, 0]) > 1e-4 * (1 + srt[:, 0])  # float64 argmin is decisive away from near-ties
        assert (got[clear] == want[clear]).all()
        assert clear.mean() > 0.5 or dim == 1


def test_nn1_lowest_index_among_duplicates():
    rng = np.random.default_rng(1)
    db = rng.normal(size=(50, 33)).astype(np.float32)
    db = np.concatenate([db, db, db])  # every row three times
    got = o.nn1(db[50:100] + 0, db)
    assert (got == np.arange(50)).all()


def test_self_matching_bunny_fpfh():
    """matcher-test.cc:17-39 (SelfMatching): (cloud, cloud, feat, feat, false, true, false, 0) -> identity pairs."""
    pts, feat = synth.bunny_fpfh()
    pairs = o.match_correspondences(pts, pts, feat, feat, False, True, False, 0)
    assert pairs.shape == (pts.shape[0], 2)
    assert (pairs[:, 0] == pairs[:, 1]).all()
    assert (pairs[:, 0] == np.arange(pts.shape[0])).all()  # sorted + unique (matcher.cc:295-296)


def test_permuted_cloud_recovers_permutation():
    pts, feat = synth.bunny_fpfh()
    perm = np.random.default_rng(3).permutation(pts.shape[0])
    pairs = o.match_correspondences(pts, pts[perm], feat, feat[perm], False, True, False, 0)
    assert pairs.shape[0] == pts.shape[0]
    assert (perm[pairs[:, 1]] == pairs[:, 0]).all()


@pytest.mark.parametrize("ns,nd", [(900, 700), (700, 900)])
def test_crosscheck_is_mutual_nn_and_union_without(ns, nd):
    mp = synth.matcher_problem(ns, nd, 300, seed=11)
    a2b = o.nn1(mp["src_feat"], mp["dst_feat"])
    b2a = o.nn1(mp["dst_feat"], mp["src_feat"])
    mutual = {(i, int(a2b[i])) for i in range(ns) if b2a[a2b[i]] == i}
    got = o.match_correspondences(mp["src_pts"], mp["dst_pts"], mp["src_feat"], mp["dst_feat"], False, True, False, 0)
    assert set(map(tuple, got)) == mutual
    assert set(map(tuple, mp["true_pairs"])) <= mutual
    # no cross check (matcher.cc:166-177): every point of the smaller cloud with its NN, plus the reverse NN of the
    # points of the larger cloud that were hit
    if ns >= nd:
        ji = {(int(b2a[j]), j) for j in range(nd)}
        ij = {(i, int(a2b[i])) for i in set(b2a.tolist())}
    else:
        ji = {(i, int(a2b[i])) for i in range(ns)}
        ij = {(int(b2a[j]), j) for j in set(a2b.tolist())}
    got2 = o.match_correspondences(mp["src_pts"], mp["dst_pts"], mp["src_feat"], mp["dst_feat"], True, False, False, 0)
    assert set(map(tuple, got2)) == ij | ji
    assert (np.lexsort((got2[:, 1], got2[:, 0])) == np.arange(len(got2))).all()


def test_tuple_test_filters_wrong_matches_and_is_seeded():
    mp = synth.matcher_problem(1500, 1500, 500, seed=5, feat_noise=0.3)
    args = (mp["src_pts"], mp["dst_pts"], mp["src_feat"], mp["dst_feat"])
    base = o.match_correspondences(*args, False, True, False, 0.95)
    t1 = o.match_correspondences(*args, False, True, True, 0.95, tuple_seed=1)
    t1b = o.match_correspondences(*args, False, True, True, 0.95, tuple_seed=1)
    t2 = o.match_correspondences(*args, False, True, True, 0.95, tuple_seed=2)
    assert (t1 == t1b).all()
    sb, s1, s2, truth = (set(map(tuple, x)) for x in (base, t1, t2, mp["true_pairs"]))
    assert s1 <= sb and s2 <= sb
    assert truth <= s1 and truth <= s2          # rigid triples always pass; each is drawn ~300 times
    assert len(s1 - truth) < len(sb - truth)    # wrong matches are thinned
    # tuple_scale == 0 disables the test (matcher.cc:223)
    t0 = o.match_correspondences(*args, False, True, True, 0.0, tuple_seed=1)
    assert (t0 == base).all()


def test_global_scale_and_absolute_scale_flag():
    mp = synth.matcher_problem(400, 300, 100, seed=9)
    _, g_abs = o.match_correspondences(mp["src_pts"], mp["dst_pts"], mp["src_feat"], mp["dst_feat"], True, True, False,
                                       0, return_scale=True)
    _, g_rel = o.match_correspondences(mp["src_pts"], mp["dst_pts"], mp["src_feat"], mp["dst_feat"], False, True,
                                       False, 0, return_scale=True)
    assert g_abs == 1.0
    ctr = [p - p.mean(0, dtype=np.float64) for p in (mp["src_pts"].astype(np.float64), mp["dst_pts"].astype(np.float64))]
    want = max(np.linalg.norm(c, axis=1).max() for c in ctr)
    assert abs(g_rel - want) < 1e-5 * want
