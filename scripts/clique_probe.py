"""Per-root timing of the exact clique search (debug flag 4): which part of one problem's search is the critical path.
usage: python scripts/clique_probe.py [C3 C2cube C5]"""
import importlib, sys, time
import numpy as np
sys.path.insert(0, ".")
capi = importlib.import_module("teaser-plusplus_b200.capi")
synth = importlib.import_module("teaser-plusplus_b200.synth")

import os
EXTRA = int(os.environ.get("PROBE_FLAGS", "0"))   # 4096: no block bound, 8192: no singleton path
NPROB = int(os.environ.get("PROBE_PROBLEMS", "3"))
ctx = capi.Context()
for cfg in (sys.argv[1:] or ["C3", "C2cube", "C5"]):
    for b in range(NPROB):
        pr = synth.config_problem(cfg, b)
        p = capi.default_params(noise_bound=pr["noise_bound"], estimate_scaling=0, rotation_cost_threshold=1e-12)
        ctx.solve(pr["src"], pr["dst"], p)              # warm (allocations)
        ctx.set_flags(4 | EXTRA)
        t0 = time.perf_counter()
        r = ctx.solve(pr["src"], pr["dst"], p)
        dt = (time.perf_counter() - t0) * 1e3
        c = ctx.debug_counters()
        st = ctx.last_stage_ms()
        ctx.set_flags(0)
        inl = set(pr["inliers"].tolist())
        clq = np.asarray(r["clique"])
        rank = sorted(inl).index(c["slowest_root_vertex"]) if c["slowest_root_vertex"] in inl else -1
        print(cfg, b, "n", len(pr["src"]), "clique", len(clq), "wall %.2f ms" % dt, "stages", {k: round(v, 3) for k, v in st.items()},
              "| nodes", c["clique_nodes"], "colourings", c["colourings"], "coloured", c["coloured_vertices"],
              "reduce rounds", c["reduce_rounds"], "reduce vertices", c["reduce_vertices"],
              "| slowest root %.3f ms vertex %d (inlier rank %d)" % (c["slowest_root_ns"] / 1e6, c["slowest_root_vertex"], rank),
              "roots>1ms", c["roots_over_1ms"], "sum root ms %.1f" % (c["root_ns_total"] / 1e6),
              "Mcycles first/reduce/colour %.1f %.1f %.1f" % (c["root_colour_cycles"] / 1e6, c["reduce_cycles"] / 1e6, c["colour_cycles"] / 1e6),
              "block-bound prunes", c["block_bound_prunes"], "flags", EXTRA)
