#!/bin/bash
# round 2: full suite, every BASELINE config (default kernels) with parity + CPU baseline, unknown-scale line, sanitizer
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(make -s -j16 -C teaser-plusplus_b200/csrc && make -s -C teaser-plusplus_b200/host && make -s -C oracle) > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu.log
for cfg in C2 C3 C2cube C4 C5 C1; do
  timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/r02_bench_${cfg}.json 2> gpurun_out/r02_bench_${cfg}.err; echo "bench $cfg rc=$?"
done
timeout 900 python bench.py --config C2scale --steps 3 --warmup 3 > gpurun_out/r02_bench_C2scale.json 2> gpurun_out/r02_bench_C2scale.err; echo "bench C2scale rc=$?"
timeout 600 python bench.py --impl reference --config C2 --steps 6 > gpurun_out/r02_bench_C2_reference_arm.json 2>/dev/null; echo "ref C2 rc=$?"
timeout 600 python bench.py --impl reference --config C1 --steps 10 > gpurun_out/r02_bench_C1_reference_arm.json 2>/dev/null; echo "ref C1 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*.json')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'ERR',e); continue
    print(f.split('/')[-1], 'value=%.1f'%d['value'], 'e2e=%.1f'%d['e2e']['value'], 'pageable', d['e2e'].get('pageable',{}).get('value'), 'stages', {k:round(v,3) for k,v in d.get('stage_ms_per_step',{}).items()}, 'frac', round(d.get('roofline',{}).get('frac',0),4), 'parity', d.get('parity',{}).get('vs_oracle'), 'cpu', d.get('cpu_baseline',{}).get('value'), 'lat', d.get('latency',{}).get('single_problem_ms_p50'))
PY
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitizer_small.py --no-cert > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitizer_small.py --no-cert > gpurun_out/r02_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02_racecheck.log
