#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(make -s -j16 -C teaser-plusplus_b200/csrc) > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:graph_strip3_kernel -s 3 -c 1 -o gpurun_out/r02_graph_v7 \
  python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline > gpurun_out/ncu_v7.log 2>&1; echo "v7 rc=$?"
TZR_FLAGS=2048 timeout 900 ncu --set full --clock-control none --import-source on -k regex:graph_strip2_kernel -s 3 -c 1 -o gpurun_out/r02_graph_v6 \
  python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline > gpurun_out/ncu_v6.log 2>&1; echo "v6 rc=$?"
for v in v7 v6; do ncu -i gpurun_out/r02_graph_$v.ncu-rep --page raw --csv > gpurun_out/r02_graph_${v}_ncu_raw.csv 2>/dev/null; done
ls -la gpurun_out | tail -6
