#!/bin/bash
# round 2, session 2: full GPU suite on the final kernels + device-batch lane experiment (TZR_DEV_CHUNKS) + stress configs
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(make -s -j16 -C teaser-plusplus_b200/csrc && make -s -C teaser-plusplus_b200/host && make -s -C oracle) > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu.log
for ch in 1 2 4 8 16; do
  TZR_DEV_CHUNKS=$ch timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu-baseline --parity-problems 4 > gpurun_out/r02_devchunks_${ch}.json 2> gpurun_out/r02_devchunks_${ch}.err; echo "bench C2 dev_chunks=$ch rc=$?"
done
for cfg in C3 C2cube C5; do
  timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --parity-problems 4 > gpurun_out/r02_quick_${cfg}.json 2> gpurun_out/r02_quick_${cfg}.err; echo "bench $cfg rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_devchunks_*.json'))+sorted(glob.glob('gpurun_out/r02_quick_*.json')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'ERR',e); continue
    print(f.split('/')[-1], 'value=%.1f'%d['value'], 'e2e=%.1f'%d['e2e']['value'], 'ms/step %.3f'%d['ms_per_step'], 'stages', {k:round(v,3) for k,v in d.get('stage_ms_per_step',{}).items()}, 'parity', d.get('parity',{}).get('vs_oracle',{}).get('clique_identical'), 'lat', d.get('latency',{}).get('single_problem_ms_p50'))
PY
