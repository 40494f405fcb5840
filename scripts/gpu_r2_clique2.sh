#!/bin/bash
# round 2: clique-stage iteration: probe (per-root timing), GPU suite, the clique-bound configs
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(make -s -j16 -C teaser-plusplus_b200/csrc && make -s -C teaser-plusplus_b200/host && make -s -C oracle) > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 300 python scripts/clique_probe.py 2>&1 | cut -c1-700 | tail -9
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02c_pytest_gpu.log
for cfg in C3 C2cube C5; do
  timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_${cfg}.json 2> gpurun_out/r02c_bench_${cfg}.err; echo "bench $cfg rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02c_bench_*.json')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'ERR',e); continue
    print(f.split('/')[-1], 'value=%.1f'%d['value'], 'e2e=%.1f'%d['e2e']['value'], 'stages', {k:round(v,3) for k,v in d.get('stage_ms_per_step',{}).items()}, 'parity', d.get('parity',{}).get('vs_oracle'), 'lat', d.get('latency',{}).get('single_problem_ms_p50'))
PY
