"""Times tzr_match_correspondences (host buffers in, pairs out) against the CPU restatement on the same inputs.
Not part of bench.py's contract (the headline metric is registrations/s of solve()); used for DESIGN.md §3.6."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tests/ because it times the oracle too)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
capi = importlib.import_module("teaser-plusplus_b200.capi")
synth = importlib.import_module("teaser-plusplus_b200.synth")
import oracle_lib as orc  # noqa: E402  (CPU baseline leg only)

ctx = capi.Context(0)
rows = []
for ns, nd, nc, tuple_test in [(5000, 5000, 1500, False), (5000, 5000, 1500, True), (20000, 20000, 5000, False),
                               (100000, 100000, 20000, False)]:
    mp = synth.matcher_problem(ns, nd, nc, seed=1, feat_noise=0.3)
    args = (mp["src_pts"], mp["dst_pts"], mp["src_feat"], mp["dst_feat"], False, True, tuple_test, 0.95)
    got = ctx.match_correspondences(*args, tuple_seed=3)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        ctx.match_correspondences(*args, tuple_seed=3)
        ts.append(time.perf_counter() - t0)
    row = dict(ns=ns, nd=nd, tuple_test=tuple_test, gpu_ms=1e3 * min(ts), pairs=int(len(got)))
    if ns <= 20000:
        t0 = time.perf_counter()
        want = orc.match_correspondences(*args, tuple_seed=3)
        row["cpu_ms"] = 1e3 * (time.perf_counter() - t0)
        row["cpu_threads"] = orc.lib().orc_num_threads()
        row["equal"] = bool(np.array_equal(got, want))
    # distance evaluations per second (both directions)
    row["gpu_pair_dists_per_s"] = 2.0 * ns * nd / (row["gpu_ms"] * 1e-3)
    rows.append(row)
    print(json.dumps(row))
