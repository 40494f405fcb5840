#!/bin/bash
# round 2, GPU call 1: tensor-core probe, graph parity through the new kernel, first timings of every config
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
# incremental: rebuilds only what is older than its sources (guards against a snapshot taken mid-build)
(make -s -j16 -C teaser-plusplus_b200/csrc && make -s -C teaser-plusplus_b200/csrc tc_probe && make -s -C teaser-plusplus_b200/host && make -s -C oracle) > gpurun_out/build.log 2>&1; echo "build rc=$?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 240 teaser-plusplus_b200/csrc/tc_probe > gpurun_out/tc_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/tc_probe.log
tail -60 gpurun_out/tc_probe.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "graph" > gpurun_out/pytest_graph.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/pytest_graph.log
tail -30 gpurun_out/pytest_graph.log
if [ $rc -eq 0 ]; then
  timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_C2_tc.json 2> gpurun_out/bench_C2_tc.err; echo "bench tc rc=$?"
  tail -c 3000 gpurun_out/bench_C2_tc.json; tail -5 gpurun_out/bench_C2_tc.err
fi
TZR_FLAGS=512 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C2_legacy.json 2> gpurun_out/bench_C2_legacy.err; echo "bench legacy rc=$?"
tail -c 1500 gpurun_out/bench_C2_legacy.json; tail -5 gpurun_out/bench_C2_legacy.err
if [ $rc -eq 0 ]; then
  timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_all.log
  tail -15 gpurun_out/pytest_all.log
  for cfg in C3 C2cube C4 C5 C1; do
    timeout 900 python bench.py --config $cfg --steps 3 --warmup 3 > gpurun_out/bench_${cfg}.json 2> gpurun_out/bench_${cfg}.err; echo "bench $cfg rc=$?"
    tail -c 2500 gpurun_out/bench_${cfg}.json; tail -3 gpurun_out/bench_${cfg}.err
  done
fi
