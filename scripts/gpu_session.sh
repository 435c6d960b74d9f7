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
    gemm_acc)  for d in 1 2; do PFN_TC_DRAIN_TILES=$d timeout 300 python scripts/debug_gemm_acc.py > gpurun_out/${tag}_gemm_acc_dt$d.log 2>&1; echo "rc=$?"; grep abs gpurun_out/${tag}_gemm_acc_dt$d.log; done ;;
    parity_large) for d in 1 2; do PFN_TC_DRAIN_TILES=$d timeout 900 python scripts/debug_parity.py 2 6470rte 512 5 > gpurun_out/${tag}_parity_large_dt$d.log 2>&1; echo "rc=$?"; cat gpurun_out/${tag}_parity_large_dt$d.log | cut -c1-200; done ;;
    large_step) for d in 0 1 2; do PFN_TC_DRAIN=$([ $d = 0 ] && echo 0 || echo 1) PFN_TC_DRAIN_TILES=$([ $d = 0 ] && echo 2 || echo $d) timeout 900 python scripts/bench_large.py > gpurun_out/${tag}_large_step_dt$d.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/${tag}_large_step_dt$d.log | cut -c1-250; done ;;
    parity_std) timeout 600 python scripts/debug_parity.py 128 118v2 129 4 > gpurun_out/${tag}_parity_std.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/${tag}_parity_std.log | cut -c1-200 ;;
    bench)     timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "rc=$?"; cat gpurun_out/${tag}_bench.json | cut -c1-6000; tail -5 gpurun_out/${tag}_bench.err ;;
    smoke)     timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${tag}_smoke.log ;;
    ncu_pipe)  for v in tma cta; do AB_ITERS=24 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ea_fwd -s 30 -c 1 -o gpurun_out/${tag}_ncu_$v -f python scripts/bench_ea_fwd_ab.py small $v > gpurun_out/${tag}_ncu_$v.log 2>&1; echo "rc=$?"; done ;;
    ncu_pipe_large)  for v in tma; do timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ea_fwd -s 3 -c 1 -o gpurun_out/${tag}_ncu_large_$v -f python scripts/bench_ea_fwd_ab.py large $v > gpurun_out/${tag}_ncu_large_$v.log 2>&1; echo "rc=$?"; done ;;
    racecheck) timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_fused.py > gpurun_out/${tag}_racecheck.log 2>&1; echo "rc=$?"; grep -E "RACECHECK SUMMARY|fused|layer-wise|Error" gpurun_out/${tag}_racecheck.log | sort | uniq -c | head -30 ;;
    memcheck)  timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_fused.py > gpurun_out/${tag}_memcheck.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|fused|layer-wise|Invalid" gpurun_out/${tag}_memcheck.log | sort | uniq -c | head -30 ;;
    t_new)     timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_dataset.py tests/test_gpu_dropin_train.py tests/test_gpu_fused.py -q -m gpu > gpurun_out/${tag}_t_new.log 2>&1; echo "rc=$?"; tail -60 gpurun_out/${tag}_t_new.log | cut -c1-300 ;;
    bench_large) timeout 900 python bench.py --config large --steps 5 --warmup 3 > gpurun_out/${tag}_bench_large.json 2> gpurun_out/${tag}_bench_large.err; echo "rc=$?"; cat gpurun_out/${tag}_bench_large.json | cut -c1-3000; tail -5 gpurun_out/${tag}_bench_large.err ;;
    bench_mixed) timeout 900 python bench.py --config mixed --steps 5 --warmup 3 > gpurun_out/${tag}_bench_mixed.json 2> gpurun_out/${tag}_bench_mixed.err; echo "rc=$?"; cat gpurun_out/${tag}_bench_mixed.json | cut -c1-3000; tail -5 gpurun_out/${tag}_bench_mixed.err ;;
    bench_ref) timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "rc=$?"; cat gpurun_out/${tag}_bench_ref.json | cut -c1-1500 ;;
    parity_wide) for v in "" "PFN_TC_DRAIN=0" "PFN_GEMM=ffma" "PFN_WGRAD_GROUP=0" "PFN_EA_FWD=cta"; do env $v timeout 600 python scripts/debug_parity.py 16 118v2 129 6 6 > gpurun_out/${tag}_parity_wide_$(echo $v | tr '=' '_').log 2>&1; echo "rc=$? [$v]"; grep -E "env|<<<" gpurun_out/${tag}_parity_wide_$(echo $v | tr '=' '_').log | cut -c1-200 | head -12; done
               env timeout 600 python scripts/debug_parity.py 16 118v2 129 4 3 0 > gpurun_out/${tag}_parity_std_layerwise.log 2>&1; grep -E "env|<<<" gpurun_out/${tag}_parity_std_layerwise.log | cut -c1-200 | head ;;
    host_prof) timeout 600 python scripts/profile_host.py > gpurun_out/${tag}_host_prof.log 2>&1; echo "rc=$?"; head -60 gpurun_out/${tag}_host_prof.log | cut -c1-200 ;;
    debias)    for v in "PFN_TC_DEBIAS=0" "PFN_TC_DEBIAS=1" "PFN_TC_DEBIAS=2"; do env $v timeout 300 python scripts/debug_gemm_acc.py > gpurun_out/${tag}_gemm_acc_$(echo $v | tr '=' '_').log 2>&1; echo "rc=$? [$v]"; cut -c1-200 gpurun_out/${tag}_gemm_acc_$(echo $v | tr '=' '_').log | grep -v "^M=4096 K=32"; 
                 env $v timeout 600 python scripts/debug_parity.py 16 118v2 129 6 6 > gpurun_out/${tag}_parity_wide_$(echo $v | tr '=' '_').log 2>&1; python scripts/parity_worst.py gpurun_out/${tag}_parity_wide_$(echo $v | tr '=' '_').log
                 env $v timeout 600 python scripts/debug_parity.py 2 6470rte 512 5 > gpurun_out/${tag}_parity_large_$(echo $v | tr '=' '_').log 2>&1; python scripts/parity_worst.py gpurun_out/${tag}_parity_large_$(echo $v | tr '=' '_').log; done ;;
    ar_test)   N=$(nvidia-smi -L | wc -l); for m in one two; do PFN_AR_TIMING=1 AR_MODE=$m timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/test_oneshot_allreduce.py > gpurun_out/${tag}_ar_test_${N}_$m.log 2>&1; echo "rc=$?"; grep "^{\|mismatch\|Error" gpurun_out/${tag}_ar_test_${N}_$m.log | cut -c1-400; done ;;
    bench_mg)  N=$(nvidia-smi -L | wc -l); for impl in ours; do timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err; echo "rc=$?"; python scripts/bench_brief.py gpurun_out/${tag}_bench_${N}gpu.json; tail -3 gpurun_out/${tag}_bench_${N}gpu.err | cut -c1-300; done
               PFN_ONE_SHOT_ALLREDUCE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/${tag}_bench_${N}gpu_nccl.json 2> gpurun_out/${tag}_bench_${N}gpu_nccl.err; echo "rc=$?"; python scripts/bench_brief.py gpurun_out/${tag}_bench_${N}gpu_nccl.json ;;
    bench_mixed_mg) N=$(nvidia-smi -L | wc -l); timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --config mixed --steps 5 --warmup 3 > gpurun_out/${tag}_bench_mixed_${N}gpu.json 2> gpurun_out/${tag}_bench_mixed_${N}gpu.err; echo "rc=$?"; python scripts/bench_brief.py gpurun_out/${tag}_bench_mixed_${N}gpu.json; tail -3 gpurun_out/${tag}_bench_mixed_${N}gpu.err | cut -c1-300 ;;
    *) echo "unknown step $step" ;;
  esac
done
