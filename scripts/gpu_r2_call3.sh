#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(make -s -j16 -C teaser-plusplus_b200/csrc && make -s -C teaser-plusplus_b200/host && make -s -C oracle) > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -m gpu -k "graph or tc_kernel" > gpurun_out/pytest_graph.log 2>&1; echo "graph rc=$?"; tail -5 gpurun_out/pytest_graph.log
for fl in 0 2048 1024; do
  TZR_FLAGS=$fl timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C2_f$fl.json 2> gpurun_out/bench_C2_f$fl.err; echo "bench flags=$fl rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_C2_f$fl.json').read().strip().splitlines()[-1])
print('flags $fl value',round(d['value']),'e2e',round(d['e2e']['value']),'stages',{k:round(v,3) for k,v in d['stage_ms_per_step'].items()},'frac',round(d['roofline']['frac'],4),d['counters'])
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_C3_batch4.csv \
  python bench.py --config C3 --batch 4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1; echo "c3 launch list rc=$?"
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -8 gpurun_out/pytest_all.log
