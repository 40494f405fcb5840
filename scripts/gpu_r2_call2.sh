#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(make -s -j16 -C teaser-plusplus_b200/csrc && make -s -C teaser-plusplus_b200/host && make -s -C oracle) > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "graph" > gpurun_out/pytest_graph.log 2>&1; echo "graph rc=$?"; tail -3 gpurun_out/pytest_graph.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C2_tc2.json 2> gpurun_out/bench_C2_tc2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_C2_tc2.json').read().strip().splitlines()[-1])
print('value',d['value'],'stages',d['stage_ms_per_step'],'frac',d['roofline']['frac'],d['counters'])
PY
bash scripts/gpu_r2_profile.sh
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_matcher_reference.py::test_match_case_1 > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -25 gpurun_out/pytest_all.log
