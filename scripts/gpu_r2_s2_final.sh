#!/bin/bash
# round 2, final evidence run (1 GPU): every BASELINE config with parity + CPU baseline, unknown-scale line, reference arm,
# launch list + full capture of the exact clique kernel on C3
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(make -s -j16 -C teaser-plusplus_b200/csrc && make -s -C teaser-plusplus_b200/host && make -s -C oracle) > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 400 python bench.py --config C2 --steps 20 --warmup 3 > gpurun_out/r02_bench_C2.json 2> gpurun_out/r02_bench_C2.err; echo "bench C2 rc=$?"
for cfg in C3 C2cube C5 C4 C1; do
  timeout 400 python bench.py --config $cfg --steps 5 --warmup 3 > gpurun_out/r02_bench_${cfg}.json 2> gpurun_out/r02_bench_${cfg}.err; echo "bench $cfg rc=$?"
done
timeout 400 python bench.py --config C2scale --steps 3 --warmup 3 > gpurun_out/r02_bench_C2scale.json 2> gpurun_out/r02_bench_C2scale.err; echo "bench C2scale rc=$?"
timeout 300 python bench.py --impl reference --config C2 --steps 6 > gpurun_out/r02_bench_C2_reference_arm.json 2>/dev/null; echo "ref C2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*.json')):
    try: d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    except Exception as e: print(f,'ERR',e); continue
    print(f.split('/')[-1], 'value=%.1f'%d['value'], 'e2e=%.1f'%d['e2e']['value'], 'pageable', d['e2e'].get('pageable',{}).get('value'), 'stages', {k:round(v,3) for k,v in d.get('stage_ms_per_step',{}).items()}, 'frac', round(d.get('roofline',{}).get('frac',0),4), 'parity', d.get('parity',{}).get('vs_oracle'), 'cpu', d.get('cpu_baseline',{}).get('value'), 'lat', d.get('latency',{}).get('single_problem_ms_p50'))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_C3_batch4.csv \
  python bench.py --config C3 --batch 4 --steps 2 --warmup 1 --no-cpu-baseline --parity-problems 0 > gpurun_out/ncu_c3.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:clique_exact -c 1 -o gpurun_out/r02_clique_exact_C3 -f \
  python bench.py --config C3 --batch 4 --steps 1 --warmup 1 --no-cpu-baseline --parity-problems 0 > gpurun_out/ncu_c3_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r02_clique_exact_C3.ncu-rep --page raw --csv > gpurun_out/r02_clique_exact_C3_ncu_raw.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_C2_batch1024.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-problems 0 > gpurun_out/ncu_c2.log 2>&1; echo "ncu C2 launches rc=$?"
ls gpurun_out | wc -l
