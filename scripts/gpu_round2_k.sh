#!/bin/bash
# round-2 GPU session K: pipelined multi-warp FAC chains + dependence-fenced FCC mat-vec: parity, memcheck, timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_criterion.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/k_pytest_crit.log 2>&1; echo "pytest exit $?" >> gpurun_out/k_pytest_crit.log
tail -12 gpurun_out/k_pytest_crit.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_criterion.py -m gpu -q --tb=line -p no:cacheprovider -k "test_asg_parity and not baseline" > gpurun_out/k_memcheck.log 2>&1; tail -4 gpurun_out/k_memcheck.log
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_criterion.py -m gpu -q --tb=line -p no:cacheprovider -k "test_asg_parity and not baseline" > gpurun_out/k_racecheck.log 2>&1; tail -4 gpurun_out/k_racecheck.log
timeout 200 python bench.py --workload asg --steps 20 --warmup 3 --no-cpu > gpurun_out/k_bench_asg.json 2> gpurun_out/k_bench_asg.err; tail -c 1100 gpurun_out/k_bench_asg.json; tail -3 gpurun_out/k_bench_asg.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/k_asg_launches.csv python scripts/prof_asg.py asg 3 > gpurun_out/k_ncu1.log 2>&1
tail -6 gpurun_out/k_asg_launches.csv | awk -F'","' '{print $5, $NF}'
