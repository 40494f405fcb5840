"""The reference's own matcher tests (test/teaser/matcher-test.cc) driven through the C-ABI on the GPU:
FPFHEstimation::computeFPFHFeatures -> Matcher::calculateCorrespondences on the reference's fixtures.
This pins the upstream stages to reference-held data instead of to the restatement only."""
import importlib
import os

import numpy as np
import pytest

capi = importlib.import_module("teaser-plusplus_b200.capi")
synth = importlib.import_module("teaser-plusplus_b200.synth")

pytestmark = pytest.mark.gpu
G = synth.GOLDEN_DIR


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def test_self_matching_canstick(ctx):
    """matcher-test.cc:17-39: a cloud matched against itself with cross-check gives the identity, one pair per point."""
    cloud = synth.read_ply_vertices(os.path.join(G, "canstick.ply")).astype(np.float32)
    d1 = ctx.compute_fpfh(cloud, 0.03, 0.05)
    d2 = ctx.compute_fpfh(cloud, 0.03, 0.05)
    assert np.array_equal(d1, d2, equal_nan=True)
    corr = ctx.match_correspondences(cloud, cloud, d1, d2, False, True, False, 0)
    assert corr.shape[0] == cloud.shape[0]            # EXPECT_EQ(correspondences.size(), cloud1.size())
    assert np.array_equal(corr[:, 0], corr[:, 1])     # EXPECT_EQ(pair.first, pair.second)


def test_match_case_1(ctx):
    """matcher-test.cc:41-77: object (1000 points) against scene (60865 points), FPFH radii 0.02 / 0.04,
    calculateCorrespondences(obj, scene, ..., false, true, false, 0.95); the i-th correspondence equals the i-th row of
    matcher-test-matches-1.csv (1-based MATLAB indices) for every row of the file."""
    obj = synth.read_ply_vertices(os.path.join(G, "matcher-test-object-1.ply")).astype(np.float32)
    scene = synth.read_ply_vertices(os.path.join(G, "matcher-test-scene-1.ply")).astype(np.float32)
    ref = np.loadtxt(os.path.join(G, "matcher-test-matches-1.csv"), delimiter=",", dtype=np.int64) - 1
    fo = ctx.compute_fpfh(obj, 0.02, 0.04)
    fs = ctx.compute_fpfh(scene, 0.02, 0.04)
    corr = ctx.match_correspondences(obj, scene, fo, fs, False, True, False, 0.95)
    assert corr.shape[0] >= ref.shape[0]
    same = (corr[:ref.shape[0]] == ref).all(axis=1)
    assert same.all(), f"{int((~same).sum())} of {ref.shape[0]} reference correspondences differ; first at row {int(np.argmin(same))}: " \
                       f"got {corr[int(np.argmin(same))]} want {ref[int(np.argmin(same))]}"
