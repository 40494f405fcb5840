"""TEST INFRASTRUCTURE ONLY — numpy restatement of teaser::DRSCertifier (reference teaser/src/certification.cc:22-671,
teaser/include/teaser/linalg.h:20-99).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it.

PARITY STATUS: pinned.  Every function below is checked against the reference's own fixtures
(test/teaser/data/certification_small_instances/case_{1,2,3}: omega, block_diag_omega, Q_cost, lambda_bar_init,
A_inv, W_dual_1st_iter, suboptimality_1st_iter, suboptimality_traj; certification_large_instances/case_{1,2}:
suboptimality_traj) to the reference test's tolerance 1e-7 (certification-test.cc:29,109-129,355-520) in
tests/test_certifier_cpu.py.  The symmetric eigendecompositions (Eigen::SelfAdjointEigenSolver in the reference,
linalg.h:84-99 and certification.cc:195-206) are numpy.linalg.eigh (LAPACK): the quantities derived from them (nearest
PSD matrix, smallest eigenvalue) do not depend on the solver beyond rounding.
"""
from __future__ import annotations

import numpy as np


def quaternion_from_rotation(R: np.ndarray) -> np.ndarray:
    """Eigen::Quaterniond(R).normalized() as (x, y, z, w) (certification.cc:66-69); sign is immaterial downstream."""
    R = np.asarray(R, dtype=np.float64)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0)
        w = 0.5 * s
        s = 0.5 / s
        x, y, z = (R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q = np.zeros(3)
        q[i] = 0.5 * s
        s = 0.5 / s
        w = (R[k, j] - R[j, k]) * s
        q[j] = (R[j, i] + R[i, j]) * s
        q[k] = (R[k, i] + R[i, k]) * s
        x, y, z = q
    v = np.array([x, y, z, w])
    return v / np.linalg.norm(v)


def get_omega1(q_xyzw: np.ndarray) -> np.ndarray:
    """certification.cc:293-303."""
    x, y, z, w = q_xyzw
    return np.array([[w, -z, y, x], [z, w, -x, y], [-y, x, w, z], [-x, -y, -z, w]], dtype=np.float64)


def get_block_diag_omega(npm: int, q_xyzw: np.ndarray) -> np.ndarray:
    """certification.cc:305-314."""
    D = np.zeros((npm, npm))
    om = get_omega1(q_xyzw)
    for i in range(npm // 4):
        D[4 * i:4 * i + 4, 4 * i:4 * i + 4] = om
    return D


_P = np.array([
    [1, 0, 0, 0, 0, -1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1],
    [0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0],
    [0, 0, 1, 0, 0, 0, 0, -1, 1, 0, 0, 0, 0, -1, 0, 0],
    [0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 0, -1, 0, 0, -1, 0],
    [-1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1],
    [0, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 0],
    [0, 0, 1, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 1, 0, 0],
    [0, 0, 0, -1, 0, 0, 1, 0, 0, 1, 0, 0, -1, 0, 0, 0],
    [-1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1],
], dtype=np.float64)  # certification.cc:242-252: maps vec(q q^T) to vec(R)


def _pk(v1k: np.ndarray, v2k: np.ndarray) -> np.ndarray:
    """P_k = reshape(P' * vec(v2 v1'), [4,4]), column-major both ways (certification.cc:268-271)."""
    A = np.outer(v2k, v1k)
    return (_P.T @ A.reshape(9, order="F")).reshape(4, 4, order="F")


def get_q_cost(v1: np.ndarray, v2: np.ndarray, noise_bound: float, cbar2: float) -> np.ndarray:
    """certification.cc:233-291.  v1, v2: (3, N)."""
    N = v1.shape[1]
    npm = 4 + 4 * N
    nbs = cbar2 * noise_bound ** 2
    Q1 = np.zeros((npm, npm))
    Q2 = np.zeros((npm, npm))
    I4 = np.eye(4)
    for k in range(N):
        s = 4 * k + 4
        Pk = _pk(v1[:, k], v2[:, k])
        nn = v1[:, k] @ v1[:, k] + v2[:, k] @ v2[:, k]
        ck = 0.5 * (nn - nbs)
        Q1[0:4, s:s + 4] += -0.5 * Pk + ck / 2 * I4
        Q1[s:s + 4, 0:4] += -0.5 * Pk + ck / 2 * I4
        ck2 = 0.5 * (nn + nbs)
        Q2[s:s + 4, s:s + 4] += -Pk + ck2 * I4
    return Q1 + Q2


def hatmap(u: np.ndarray) -> np.ndarray:
    """linalg.h:20-29."""
    return np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]], dtype=np.float64)


def get_lambda_guess(R: np.ndarray, theta: np.ndarray, src: np.ndarray, dst: np.ndarray, noise_bound: float,
                     cbar2: float) -> np.ndarray:
    """certification.cc:448-529 (dense).  theta: (N,) of +-1."""
    K = theta.shape[0]
    npm = 4 * K + 4
    nbs = cbar2 * noise_bound ** 2
    L = np.zeros((npm, npm))
    top = np.zeros((4, 4))
    I3 = np.eye(3)
    for i in range(K):
        s = src[:, i]
        sh = hatmap(s)
        xi = R.T @ (dst[:, i] - R @ s)
        xh = hatmap(xi)
        blk = np.zeros((4, 4))
        if theta[i] > 0:
            blk[3, 3] = -0.75 * (xi @ xi) - 0.25 * nbs
            blk[:3, :3] = (sh @ sh - 0.5 * (s @ xi) * I3 + 0.5 * xh @ sh + 0.5 * np.outer(xi, s)
                           - 0.75 * (xi @ xi) * I3 - 0.25 * nbs * I3)
            blk[:3, 3] = -1.5 * xh @ s
        else:
            blk[3, 3] = -0.25 * (xi @ xi) - 0.75 * nbs
            blk[:3, :3] = (sh @ sh - 0.5 * (s @ xi) * I3 + 0.5 * xh @ sh + 0.5 * np.outer(xi, s)
                           - 0.25 * (xi @ xi) * I3 - 0.25 * nbs * I3)
            blk[:3, 3] = -0.5 * xh @ s
        blk[3, :3] = blk[:3, 3]
        L[4 * (i + 1):4 * (i + 1) + 4, 4 * (i + 1):4 * (i + 1) + 4] = -blk
        top += blk
    L[:4, :4] += top
    return L


def pair_index(N: int):
    """mat2vec of certification.cc:543-551: (i, j), i < j < N -> running index, row-major over the upper triangle."""
    iu = np.triu_indices(N, k=1)
    m = np.full((N, N), -1, dtype=np.int64)
    m[iu] = np.arange(iu[0].size)
    return m, iu


def get_linear_projection(theta_prepended: np.ndarray) -> np.ndarray:
    """certification.cc:531-655, dense (use only for small N: (N(N-1)/2)^2 entries)."""
    N = theta_prepended.shape[0]
    N0 = N - 1
    y = 1.0 / (2 * N0 + 6)
    x = (N0 + 1.0) * y
    m, _ = pair_index(N)
    nr = N * (N - 1) // 2
    A = np.zeros((nr, nr))
    th = theta_prepended
    for i in range(N - 1):
        for j in range(i + 1, N):
            c = m[i, j]
            for p in range(N):
                if p != j and p != i:
                    if p < i:
                        A[m[p, i], c] += y * th[j] * th[p]
                    else:
                        A[m[i, p], c] += -y * th[j] * th[p]
            for p in range(N):
                if p != i and p != j:
                    if p < j:
                        A[m[p, j], c] += -y * th[i] * th[p]
                    else:
                        A[m[j, p], c] += y * th[i] * th[p]
            A[c, c] += x
    return A


def apply_linear_projection(theta_prepended: np.ndarray, b: np.ndarray) -> np.ndarray:
    """A_inv @ b without forming A_inv (b: (N(N-1)/2, k)).  From the column pattern of certification.cc:577-641:
    out[a,b] = (x + 2y) B[a,b] + y (th_a (Rs[b] - Cs[b]) - th_b (Rs[a] - Cs[a])), Rs[v] = sum_{j>v} th_j B[v,j],
    Cs[v] = sum_{i<v} th_i B[i,v]."""
    N = theta_prepended.shape[0]
    N0 = N - 1
    y = 1.0 / (2 * N0 + 6)
    x = (N0 + 1.0) * y
    _, (ia, ib) = pair_index(N)
    th = theta_prepended
    k = b.shape[1]
    Rs = np.zeros((N, k))
    Cs = np.zeros((N, k))
    np.add.at(Rs, ia, th[ib, None] * b)
    np.add.at(Cs, ib, th[ia, None] * b)
    D = Rs - Cs
    return (x + 2 * y) * b + y * (th[ia, None] * D[ib] - th[ib, None] * D[ia])


def get_optimal_dual_projection(W: np.ndarray, theta_prepended: np.ndarray, A_inv=None) -> np.ndarray:
    """certification.cc:316-446.  A_inv: dense matrix, or None to apply it implicitly."""
    npm = W.shape[0]
    N = npm // 4 - 1
    th = theta_prepended
    _, (ia, ib) = pair_index(N + 1)
    # b_W rows (certification.cc:334-373): -th_ij * C + D - E + th_ij * F
    thij = th[ia] * th[ib]
    Cv = np.stack([W[4 * ia + 3, 4 * ia + c] for c in range(3)], axis=1)
    Dv = np.stack([W[4 * ib + 3, 4 * ia + c] for c in range(3)], axis=1)
    Ev = np.stack([W[4 * ia + 3, 4 * ib + c] for c in range(3)], axis=1)
    Fv = np.stack([W[4 * ib + 3, 4 * ib + c] for c in range(3)], axis=1)
    bW = -thij[:, None] * Cv + Dv - Ev + thij[:, None] * Fv
    bWd = apply_linear_projection(th, bW) if A_inv is None else A_inv @ bW
    Wd = np.zeros((npm, npm))
    for c in range(ia.size):
        i, j = ia[c], ib[c]
        Wij = W[4 * i:4 * i + 4, 4 * j:4 * j + 4]
        blk = (Wij - Wij.T) / 2
        blk[:3, 3] = bWd[c]
        blk[3, :3] = -bWd[c]
        Wd[4 * i:4 * i + 4, 4 * j:4 * j + 4] = blk
    Wd = Wd + Wd.T
    vec = np.zeros(npm)
    vec[3::4] = th
    diag_sum = np.zeros((3, 3))
    for i in range(N + 1):
        s = 4 * i
        rs = Wd[s:s + 4, :] @ vec          # getBlockRowSum (certification.cc:657-671); the diagonal block is still 0
        Wii = W[s:s + 4, s:s + 4].copy()
        Wii[:, 3] = -th[i] * rs
        Wii[3, :] = -th[i] * rs
        Wd[s:s + 4, s:s + 4] = Wii
        diag_sum += Wii[:3, :3]
    mean = diag_sum / (N + 1)
    for i in range(N + 1):
        Wd[4 * i:4 * i + 3, 4 * i:4 * i + 3] -= mean
    return Wd


def nearest_psd(A: np.ndarray) -> np.ndarray:
    """linalg.h:84-99."""
    B = (A + A.T) / 2
    w, V = np.linalg.eigh(B)
    return (V * np.maximum(w, 0)) @ V.T


def compute_suboptimality_gap(M: np.ndarray, mu: float, N: int) -> float:
    """certification.cc:192-231 (EIGEN solver branch)."""
    w = np.linalg.eigvalsh((M + M.T) / 2)
    mn = w.min()
    if mn > 0:
        return 0.0
    return (-mn * (N + 1)) / mu


def certify(R: np.ndarray, src: np.ndarray, dst: np.ndarray, theta: np.ndarray, noise_bound: float = 0.01,
            cbar2: float = 1.0, sub_optimality: float = 1e-3, max_iterations: float = 2e2,
            gamma_tau: float = 1.999999, return_intermediates: bool = False):
    """certification.cc:40-190.  src, dst: (3, N); theta: (N,) of +-1 (or bool mask -> +-1, :22-38)."""
    theta = np.asarray(theta)
    if theta.dtype == np.bool_:
        theta = np.where(theta, 1.0, -1.0)
    theta = theta.astype(np.float64).ravel()
    N = src.shape[1]
    npm = 4 + 4 * N
    thp = np.concatenate([[1.0], theta])
    Q = get_q_cost(src, dst, noise_bound, cbar2)
    q = quaternion_from_rotation(R)
    x = np.kron(thp, q)
    D = get_block_diag_omega(npm, q)
    Q_bar = D.T @ (Q @ D)
    mu = float(x @ (Q @ x))
    lam = get_lambda_guess(R, theta, src, dst, noise_bound, cbar2)
    M_init = Q_bar.copy()
    M_init[:4, :4] -= mu * np.eye(4)
    M_init -= lam
    M = M_init.copy()
    traj = []
    best = np.inf
    inter = {}
    it = 0
    while it < max_iterations:
        M_psd = nearest_psd(M)
        tW = 2 * M_psd - M - M_init
        Wd = get_optimal_dual_projection(tW, thp)
        M_aff = M_init + Wd
        gap = compute_suboptimality_gap(M_aff, mu, N)
        if it == 0 and return_intermediates:
            inter = dict(W=tW, W_dual=Wd, M_affine=M_aff, mu=mu, Q_cost=Q, lambda_guess=lam, M_init=M_init)
        traj.append(gap)
        best = min(best, gap)
        if gap < sub_optimality:
            break
        M = M + gamma_tau * (M_aff - M_psd)
        it += 1
    res = dict(is_optimal=bool(best < sub_optimality), best_suboptimality=float(best),
               suboptimality_traj=np.asarray(traj))
    if return_intermediates:
        res.update(inter)
    return res
