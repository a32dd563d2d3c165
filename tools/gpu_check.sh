#!/bin/bash
# One gpurun call: microbenchmarks, GPU parity tests, smoke, a short bench, then the ncu launch list.
# Usage (from the repo root on the GPU box): bash tools/gpu_check.sh [stage ...]
set -u
mkdir -p gpurun_out
STAGES="${*:-micro tests smoke bench}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for s in $STAGES; do
  case $s in
    micro)
      (cd tools && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu && timeout 120 ./microbench) > gpurun_out/micro.log 2>&1
      echo "[micro] exit $?"; tail -25 gpurun_out/micro.log ;;
    tests)
      timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
      echo "[tests] exit $?"; tail -30 gpurun_out/pytest_gpu.log ;;
    smoke)
      timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
      echo "[smoke] exit $?"; tail -5 gpurun_out/smoke.log ;;
    bench)
      timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
      echo "[bench] exit $?"; tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err ;;
    benchdirect)
      timeout 900 python bench.py --steps 5 --warmup 3 --force-direct --no-e2e --no-cpu > gpurun_out/bench_direct.log 2> gpurun_out/bench_direct.err
      echo "[benchdirect] exit $?"; tail -3 gpurun_out/bench_direct.log; tail -5 gpurun_out/bench_direct.err ;;
    configs)
      timeout 1200 python tools/bench_configs.py 3 4 5 > gpurun_out/configs.log 2> gpurun_out/configs.err
      echo "[configs] exit $?"; cat gpurun_out/configs.log | cut -c1-700; tail -5 gpurun_out/configs.err ;;
    ncuk3)
      # full captures of one kLocal and one kGlobal launch of the A/B harness (launch order: see k3_bench.py)
      timeout 300 python tools/k3_bench.py plain > gpurun_out/k3_plain.log 2>&1 &&
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_steric_tma -s 8 -c 1 \
          -o gpurun_out/prof_klocal python tools/k3_bench.py ncu > gpurun_out/ncu_k3a.log 2>&1 &&
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_steric_tma -s 14 -c 1 \
          -o gpurun_out/prof_kglobal python tools/k3_bench.py ncu > gpurun_out/ncu_k3b.log 2>&1
      echo "[ncuk3] exit $?"; cat gpurun_out/k3_plain.log | tail -1 ;;
    e2esweep)
      # the host path by packing mode / threads / window width; "ring" = only the staging-ring modes against the default
      timeout 300 python tools/e2e_sweep.py 2 ${E2E_SWEEP_ARGS:-} > gpurun_out/e2e_sweep.log 2> gpurun_out/e2e_sweep.err
      echo "[e2esweep] exit $?"; cut -c1-260 gpurun_out/e2e_sweep.log; tail -3 gpurun_out/e2e_sweep.err ;;
    packbench)
      (cd tools && g++ -O3 -std=c++17 -pthread -o /tmp/packbench packbench.cpp ../momlevel_b200/csrc/ml_pack.cpp &&
        for n in 1 4 8 15; do /tmp/packbench $n; done) > gpurun_out/packbench.log 2>&1
      echo "[packbench] exit $?"; grep -E "pack step|memcpy" gpurun_out/packbench.log | tail -20 ;;
    benchref)
      timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
      echo "[benchref] exit $?"; tail -3 gpurun_out/bench_ref.log ;;
    nculist)
      timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_plain.log 2>&1 &&
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
          --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_list.log 2>&1
      echo "[nculist] exit $?"; tail -3 gpurun_out/ncu_list.log ;;
    ncufull)
      timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_plain2.log 2>&1 &&
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_steric -s 1 -c 1 \
          -o gpurun_out/prof_local python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_full.log 2>&1
      echo "[ncufull] exit $?"; tail -3 gpurun_out/ncu_full.log ;;
  esac
done
