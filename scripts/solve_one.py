"""Solve one synthetic BASELINE config on the GPU and print stage timings (profiling helper).
usage: python scripts/solve_one.py C3 [n] [repeat]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
capi = importlib.import_module("teaser-plusplus_b200.capi")
synth = importlib.import_module("teaser-plusplus_b200.synth")

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
n = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else None
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 2
pr = synth.config_problem(cfg, 0, n=n)
ctx = capi.Context(0)
ctx.set_flags(4 | int(os.environ.get('PROBE_FLAGS', '0')))
p = capi.default_params(noise_bound=pr["noise_bound"], estimate_scaling=0, rotation_cost_threshold=1e-12)
for _ in range(rep):
    g = ctx.solve(pr["src"], pr["dst"], p)
    print(cfg, "n", len(pr["src"]), "clique", len(g["clique"]), "proven", g["proven"], "edges", g["n_edges"],
          "stage_ms [prep, graph, clique, rot+trans]", [round(x, 3) for x in g["stage_ms"][:4]])
    print("   counters", ctx.debug_counters())
