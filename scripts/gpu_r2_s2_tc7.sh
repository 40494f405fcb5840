#!/bin/bash
# round 2, session 2: tensor-core graph kernel v4 (3 CTAs/SM, one accumulator stage, TMA and MMA lanes in separate warps,
# shift-only tile schedule) + the exact clique kernel with the block bound compiled out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(make -s -j16 -C teaser-plusplus_b200/csrc && make -s -C teaser-plusplus_b200/host && make -s -C oracle) > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 300 python -m pytest tests/test_gpu_round2.py -q -x -k "tc_kernel or v7" > gpurun_out/r02_tc7_tests.log 2>&1; echo "tc tests rc=$?"; tail -5 gpurun_out/r02_tc7_tests.log
for fl in 1024 0; do
  TZR_FLAGS=$fl timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu-baseline --parity-problems 8 > gpurun_out/r02_tc7_bench_C2_flags$fl.json 2> gpurun_out/r02_tc7_bench_C2_flags$fl.err; echo "bench C2 flags=$fl rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_tc7_bench_*.json')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'ERR',e); continue
    print(f.split('/')[-1], 'value=%.1f'%d['value'], 'e2e=%.1f'%d['e2e']['value'], 'stages', {k:round(v,3) for k,v in d.get('stage_ms_per_step',{}).items()}, 'frac', round(d.get('roofline',{}).get('frac',0),4), 'parity', d.get('parity',{}).get('vs_oracle'))
PY
TZR_FLAGS=1024 timeout 600 ncu --set full --clock-control none --import-source on -k regex:graph_tc_kernel -s 3 -c 1 -o gpurun_out/r02_graph_tc7 -f \
  python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline --parity-problems 0 > gpurun_out/ncu_tc7.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r02_graph_tc7.ncu-rep --page raw --csv > gpurun_out/r02_graph_tc7_b128_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_graph_tc7.ncu-rep --page source --csv > gpurun_out/r02_graph_tc7_b128_ncu_source.csv 2>/dev/null
TZR_FLAGS=1024 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_tc7_batch1024.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-problems 0 > gpurun_out/ncu_tc7_launches.log 2>&1; echo "launch list rc=$?"
