"""The restatement's two-pass exact search on a dense inlier graph (first pass, vertex-cover LP bound, Nemhauser-Trotter
reduction, beat-only second pass): the returned clique must be valid and of MAXIMUM size, checked independently with a
mixed-integer program (scipy / HiGHS) for the maximum independent set of the complement graph."""
import numpy as np
import pytest

import oracle_lib as orc


def _dense_problem(n=360, n_out=110, seed=5):
    rng = np.random.default_rng(seed)
    src = rng.uniform(0, 0.15, size=(n, 3))
    dst = src + (rng.random((n, 3)) - 0.5) * 0.1      # noise up to 0.087 > noise bound 0.05: inlier pairs may conflict
    dst[:n_out] += 7.0
    return src, dst


def test_dense_graph_maximum_clique_matches_milp():
    milp = pytest.importorskip("scipy.optimize").milp
    from scipy.optimize import Bounds, LinearConstraint
    from scipy.sparse import lil_matrix
    src, dst = _dense_problem()
    n = src.shape[0]
    bits, deg, ne = orc.build_graph_bits(src, dst, 0.05)
    A = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, :n].astype(bool)
    c, info = orc.max_clique_bits(bits, n, mode=0, time_limit=300.0)
    c = np.asarray(c)
    assert A[np.ix_(c, c)].sum() == len(c) * (len(c) - 1)          # a clique
    assert not info["timed_out"]
    # independent optimum: max sum x, x_u + x_v <= 1 for every non-adjacent pair
    non = np.argwhere(np.triu(~A, 1))
    con = lil_matrix((len(non), n))
    for k, (i, j) in enumerate(non):
        con[k, i] = 1
        con[k, j] = 1
    res = milp(c=-np.ones(n), constraints=LinearConstraint(con.tocsr(), ub=np.ones(len(non))), integrality=np.ones(n),
               bounds=Bounds(0, 1), options=dict(time_limit=120))
    assert res.status == 0
    assert len(c) == int(round(-res.fun))
