#!/bin/bash
# round 2: persistent exact-clique kernel (speculative colouring, colour-first roots): clique-stage tests + the configs
# whose step is the clique stage (C3, C2cube, C5), C2 as the no-regression check, launch list of C3
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(make -s -j16 -C teaser-plusplus_b200/csrc && make -s -C teaser-plusplus_b200/host && make -s -C oracle) > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02b_pytest_gpu.log
for cfg in C3 C2cube C5 C2; do
  timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 > gpurun_out/r02b_bench_${cfg}.json 2> gpurun_out/r02b_bench_${cfg}.err; echo "bench $cfg rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02b_launches_C3_batch4.csv python bench.py --config C3 --batch 4 --steps 1 --warmup 3 --parity-problems 0 > gpurun_out/ncu_c3.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02b_bench_*.json')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'ERR',e); continue
    print(f.split('/')[-1], 'value=%.1f'%d['value'], 'e2e=%.1f'%d['e2e']['value'], 'stages', {k:round(v,3) for k,v in d.get('stage_ms_per_step',{}).items()}, 'parity', d.get('parity',{}).get('vs_oracle'), 'lat', d.get('latency',{}).get('single_problem_ms_p50'), 'counters', d.get('counters'))
PY
grep -c "" gpurun_out/r02b_launches_C3_batch4.csv; grep "clique" gpurun_out/r02b_launches_C3_batch4.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -k1nr | head
