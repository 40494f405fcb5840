"""Oracle max-clique checks: the reference's PMC toy graphs (test/teaser/graph-test.cc:131-409),
the N=20 outlier-detection identity test (test/teaser/registration-test.cc:394-467), and an
independent cross-check against networkx on random graphs."""
import importlib

import numpy as np
import pytest

import oracle_lib as orc

synth = importlib.import_module("teaser-plusplus_b200.synth")


def test_k5():  # graph-test.cc:131-180, 273-305
    adj = [[j for j in range(5) if j != i] for i in range(5)]
    for threads in (1, 12, 15):
        c, info = orc.max_clique_adj(adj, threads=threads)
        assert sorted(c.tolist()) == [0, 1, 2, 3, 4]
        assert info["lb"] == 5 and info["ub"] == 5 and info["max_core"] == 4


def test_four_vertex_graph():  # graph-test.cc:182-224
    adj = [[2, 3], [2], [0, 1, 3], [0, 2]]
    c, info = orc.max_clique_adj(adj)
    assert sorted(c.tolist()) == [0, 2, 3]


def test_isolated_vertices():  # graph-test.cc:226-271: PMC reports lb=0; solve() then declares invalid
    c, info = orc.max_clique_adj([[], [], [], []])
    assert len(c) == 0 and info["lb"] == 0


def test_kcore_heuristic_mode():  # graph.cc:66-81
    # K6 plus a pendant path: max_core=5 > 0.5*8 -> returns the max-core vertices
    adj = [[j for j in range(6) if j != i] for i in range(6)] + [[7], [6]]
    c, info = orc.max_clique_adj(adj, mode=2, kcore_thr=0.5)
    assert sorted(c.tolist()) == [0, 1, 2, 3, 4, 5]
    # threshold 1 short-circuits to the PMC heuristic path
    c, info = orc.max_clique_adj(adj, mode=2, kcore_thr=1.0)
    assert sorted(c.tolist()) == [0, 1, 2, 3, 4, 5]


@pytest.mark.parametrize("seed", range(6))
def test_random_graphs_vs_networkx(seed):
    nx = pytest.importorskip("networkx")
    rng = np.random.default_rng(seed)
    n = int(rng.integers(30, 140))
    p = float(rng.uniform(0.1, 0.6))
    A = rng.uniform(size=(n, n)) < p
    A = np.triu(A, 1)
    # plant a clique in half of the cases
    if seed % 2 == 0:
        k = int(rng.integers(8, 16))
        idx = rng.choice(n, size=k, replace=False)
        for a in idx:
            for b in idx:
                if a < b:
                    A[a, b] = True
    A = A | A.T
    adj = [np.nonzero(A[i])[0].tolist() for i in range(n)]
    c, info = orc.max_clique_adj(adj, threads=4)
    G = nx.from_numpy_array(A)
    ref, w = nx.max_weight_clique(G, weight=None)
    assert len(c) == w
    c = c.tolist()
    assert all(A[a, b] for a in c for b in c if a != b)


@pytest.mark.parametrize("n_out", [1, 2, 3, 4, 5])
def test_outlier_detection_identity(n_out):  # registration-test.cc:394-467
    rng = np.random.default_rng(100 + n_out)
    N = 20
    src = rng.uniform(-1, 1, size=(N, 3))
    R = synth.random_rotation(rng)
    t = rng.uniform(-1, 1, size=3)
    dst = src @ R.T + t
    out_idx = np.sort(rng.choice(N, size=n_out, replace=False))
    dst[out_idx] += rng.uniform(5, 10, size=(n_out, 3))
    p = orc.default_params(noise_bound=1e-3, cbar2=1, estimate_scaling=0, rotation_max_iterations=100,
                           rotation_gnc_factor=1.4, rotation_estimation_algorithm=0, rotation_cost_threshold=0.005)
    out = orc.solve(src, dst, p)
    assert synth.angular_error(R, out["R"]) <= 0.2
    assert np.linalg.norm(out["t"] - t) <= 0.1
    expect = sorted(set(range(N)) - set(out_idx.tolist()))
    assert out["clique"].tolist() == expect


@pytest.mark.parametrize("cfg,n", [("C2", 600), ("C2cube", 600), ("C3", 2500), ("C4", 400), ("C5", 800)])
def test_synthetic_configs_small(cfg, n):
    """Scaled-down BASELINE configs: the planted inlier set is THE maximum clique."""
    pr = synth.config_problem(cfg, 0, n=n)
    p = orc.default_params(noise_bound=pr["noise_bound"], estimate_scaling=0, rotation_cost_threshold=1e-12)
    out = orc.solve(pr["src"], pr["dst"], p)
    assert out["valid"]
    assert set(pr["inliers"].tolist()) <= set(out["clique"].tolist())
    assert synth.angular_error(pr["R"], out["R"]) < 0.05
    assert np.linalg.norm(out["t"] - pr["t"]) < 0.05


def test_canonical_tie_break():
    """Several maximum cliques: the oracle (and the GPU path) return the lexicographically smallest sorted
    index set; the reference (PMC, multi-threaded) returns an unspecified one of them."""
    # two disjoint K4
    adj = [[j for j in range(4) if j != i] for i in range(4)] + [[4 + j for j in range(4) if j != i] for i in range(4)]
    c, _ = orc.max_clique_adj(adj, threads=3)
    assert sorted(c.tolist()) == [0, 1, 2, 3]
    # overlapping K4s {2,3,4,5} and {0,3,4,5}
    import itertools
    n = 6
    A = np.zeros((n, n), dtype=bool)
    for K in ([2, 3, 4, 5], [0, 3, 4, 5]):
        for a, b in itertools.combinations(K, 2):
            A[a, b] = A[b, a] = True
    adj = [np.nonzero(A[i])[0].tolist() for i in range(n)]
    c, _ = orc.max_clique_adj(adj, threads=2)
    assert sorted(c.tolist()) == [0, 3, 4, 5]
