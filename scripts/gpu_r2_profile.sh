#!/bin/bash
# round 2: ncu captures of the tensor-core graph kernel (full set, batch 128) and the launch list of one bench step
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(make -s -j16 -C teaser-plusplus_b200/csrc && make -s -C teaser-plusplus_b200/host && make -s -C oracle) > gpurun_out/build.log 2>&1; echo "build rc=$?"
# launch list (cold-cache, serialised: compare SHARES): 1 warm-up + counters + 2 steps
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_batch1024.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
# full capture of the graph kernel at batch 128 (2 launches, ~40 replays each)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:graph_tc_kernel -s 3 -c 2 -o gpurun_out/r02_graph_tc \
  python bench.py --steps 2 --warmup 3 --batch 128 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
ncu -i gpurun_out/r02_graph_tc.ncu-rep --page raw --csv > gpurun_out/r02_graph_tc_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_graph_tc.ncu-rep --page details --csv > gpurun_out/r02_graph_tc_ncu_details.csv 2>/dev/null
# tail kernels (prep, tc_prep, heur, peel, rot_trans)
timeout 900 ncu --set full --clock-control none -k regex:'prep_kernel|clique_heur|clique_peel|rot_trans' -s 8 -c 5 -o gpurun_out/r02_tail \
  python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/ncu_tail.log 2>&1; echo "tail rc=$?"
ncu -i gpurun_out/r02_tail.ncu-rep --page raw --csv > gpurun_out/r02_tail_ncu_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
