#!/bin/bash
# round 2: strong scaling of the fixed-batch configs C4 / C5 (sharded b mod G) and weak scaling of C2 on 8 GPUs of one box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(make -s -j16 -C teaser-plusplus_b200/csrc && make -s -C oracle) > gpurun_out/build.log 2>&1; echo "build rc=$?"
nvidia-smi -L | head -8
for cfg in C4 C5 C2; do
  timeout 300 python bench.py --gpus 1 --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --parity-problems 4 > gpurun_out/r02_scale_${cfg}_n1.json 2> gpurun_out/r02_scale_${cfg}_n1.err; echo "$cfg n=1 rc=$?"
  for N in 2 4 8; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --parity-problems 4 > gpurun_out/r02_scale_${cfg}_n$N.json 2> gpurun_out/r02_scale_${cfg}_n$N.err; echo "$cfg n=$N rc=$?"
  done
done
python - <<'PY'
import json,glob
base={}
for cfg in ['C4','C5','C2']:
    for N in [1,2,4,8]:
        f=f'gpurun_out/r02_scale_{cfg}_n{N}.json'
        try: d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        except Exception as e: print(f,'ERR',e); continue
        if N==1: base[cfg]=d['value']
        print(cfg,N,'value=%.1f'%d['value'],'e2e=%.1f'%d['e2e']['value'],'ms/step %.3f'%d['ms_per_step'],d['scaling'],'eff=%.3f'%(d['value']/(N*base.get(cfg,d['value']))))
PY
