"""One call of the matcher and of FPFH at fixed sizes (run under ncu to capture nn_kernel / normals / spfh / fpfh)."""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
capi = importlib.import_module("teaser-plusplus_b200.capi")
synth = importlib.import_module("teaser-plusplus_b200.synth")

ctx = capi.Context(0)
mp = synth.matcher_problem(20000, 20000, 5000, seed=1, feat_noise=0.3)
pairs = ctx.match_correspondences(mp["src_pts"], mp["dst_pts"], mp["src_feat"], mp["dst_feat"], False, True, False, 0.95)
print("pairs", len(pairs))
rng = np.random.default_rng(0)
n = 20000
uv = rng.uniform(-1, 1, size=(n, 2))
pts = np.stack([uv[:, 0], uv[:, 1], 0.3 * np.sin(3 * uv[:, 0]) * np.cos(2 * uv[:, 1])], 1).astype(np.float32) + np.float32(2.0)
f = ctx.compute_fpfh(pts, 0.025, 0.04)
print("fpfh", f.shape)
