#!/bin/bash
# One gpurun call = one pass over these steps; every step has its own timeout and log under gpurun_out/<tag>_*.
# usage: bash scripts/gpu_session.sh <tag> step1 step2 ...
tag=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
for step in "$@"; do
  echo "=== $step ($(date +%T))"
  case $step in
    ab_small)  timeout 300 python scripts/bench_ea_fwd_ab.py small > gpurun_out/${tag}_ab_small.jsonl 2> gpurun_out/${tag}_ab_small.err; echo "rc=$?"; cat gpurun_out/${tag}_ab_small.jsonl | cut -c1-260 ;;
    ab_large)  timeout 600 python scripts/bench_ea_fwd_ab.py large > gpurun_out/${tag}_ab_large.jsonl 2> gpurun_out/${tag}_ab_large.err; echo "rc=$?"; cat gpurun_out/${tag}_ab_large.jsonl | cut -c1-260 ;;
    t_kernels) timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/${tag}_t_kernels.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/${tag}_t_kernels.log ;;
    t_all)     timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${tag}_t_all.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/${tag}_t_all.log ;;
    gemm_acc)  for d in 0 1; do PFN_TC_DRAIN=$d timeout 300 python scripts/debug_gemm_acc.py > gpurun_out/${tag}_gemm_acc_drain$d.log 2>&1; echo "rc=$?"; cat gpurun_out/${tag}_gemm_acc_drain$d.log; done ;;
    parity_large) for d in 0 1; do PFN_TC_DRAIN=$d timeout 900 python scripts/debug_parity.py 2 6470rte 512 5 > gpurun_out/${tag}_parity_large_drain$d.log 2>&1; echo "rc=$?"; cat gpurun_out/${tag}_parity_large_drain$d.log | cut -c1-200; done ;;
    parity_std) timeout 600 python scripts/debug_parity.py 128 118v2 129 4 > gpurun_out/${tag}_parity_std.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/${tag}_parity_std.log | cut -c1-200 ;;
    bench)     timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "rc=$?"; cat gpurun_out/${tag}_bench.json | cut -c1-3000 ;;
    smoke)     timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${tag}_smoke.log ;;
    *) echo "unknown step $step" ;;
  esac
done
