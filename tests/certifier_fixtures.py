"""Loader for the reference's certification fixtures (tests/golden/certification_*; the reference reads them in
test/teaser/certification-test.cc:78-330)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _mat(path):
    return np.atleast_2d(np.loadtxt(path, delimiter=",", dtype=np.float64))


def load_case(kind: str, name: str):
    d = os.path.join(GOLDEN, f"certification_{kind}_instances", name)
    params = {}
    for line in open(os.path.join(d, "parameters.txt")):
        if ":" in line:
            k, v = line.split(":")
            params[k.strip()] = float(v)
    c = dict(params=params, name=f"{kind}/{name}")
    for f in sorted(os.listdir(d)):
        if f.endswith(".csv"):
            c[f[:-4]] = _mat(os.path.join(d, f))
    c["theta_est"] = c["theta_est"].ravel()
    c["q_est"] = c["q_est"].ravel()          # stored x, y, z, w (certification-test.cc:160-163)
    c["suboptimality_traj"] = c["suboptimality_traj"].ravel()
    if "mu" in c:
        c["mu"] = float(c["mu"].ravel()[0])
    return c


def cases(kind: str):
    root = os.path.join(GOLDEN, f"certification_{kind}_instances")
    return [load_case(kind, n) for n in sorted(os.listdir(root))]
