#!/bin/bash
# round 2, session 2: tensor-core graph kernel v3 (square-root-free polynomial predicate): parity tests, bench A/B, ncu
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(make -s -j16 -C teaser-plusplus_b200/csrc && make -s -C teaser-plusplus_b200/host && make -s -C oracle) > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 600 python -m pytest tests/test_gpu_round2.py -q -x -k "tc_kernel or v7" > gpurun_out/r02_tc3_tests.log 2>&1; echo "tc tests rc=$?"; tail -5 gpurun_out/r02_tc3_tests.log
for fl in 1024 0; do
  TZR_FLAGS=$fl timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu-baseline --parity-problems 8 > gpurun_out/r02_tc3_bench_C2_flags$fl.json 2> gpurun_out/r02_tc3_bench_C2_flags$fl.err; echo "bench C2 flags=$fl rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_tc3_bench_*.json')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'ERR',e); continue
    print(f.split('/')[-1], 'value=%.1f'%d['value'], 'e2e=%.1f'%d['e2e']['value'], 'stages', {k:round(v,3) for k,v in d.get('stage_ms_per_step',{}).items()}, 'frac', round(d.get('roofline',{}).get('frac',0),4), 'parity', d.get('parity',{}).get('vs_oracle'), d.get('parity',{}).get('planted'))
PY
TZR_FLAGS=1024 timeout 600 ncu --set full --clock-control none --import-source on -k regex:graph_tc_kernel -s 3 -c 1 -o gpurun_out/r02_graph_tc3 -f \
  python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline --parity-problems 0 > gpurun_out/ncu_tc3.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r02_graph_tc3.ncu-rep --page raw --csv > gpurun_out/r02_graph_tc3_b128_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_graph_tc3.ncu-rep --page details --csv > gpurun_out/r02_graph_tc3_b128_ncu_details.csv 2>/dev/null
ncu -i gpurun_out/r02_graph_tc3.ncu-rep --page source --csv > gpurun_out/r02_graph_tc3_b128_ncu_source.csv 2>/dev/null
ls -la gpurun_out | tail -12
# --- illegal memory access of HEAD's exact clique kernel at full-size C3: locate it
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/solve_one.py C3 0 1 > gpurun_out/r02_memcheck_C3.log 2>&1; echo "memcheck C3 rc=$?"
grep -m 40 -E "Invalid|at |by thread|Address|ERROR SUMMARY|clique" gpurun_out/r02_memcheck_C3.log | head -60
