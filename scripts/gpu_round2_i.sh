#!/bin/bash
# round-2 GPU session I: ASG v2 (stored FCC vectors, 8-frame FAC segments, packed adds, tree reduce): parity, memcheck, timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_criterion.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/i_pytest_crit.log 2>&1; echo "pytest exit $?" >> gpurun_out/i_pytest_crit.log
tail -25 gpurun_out/i_pytest_crit.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_criterion.py -m gpu -q --tb=line -p no:cacheprovider -k "test_asg_parity and not baseline" > gpurun_out/i_memcheck.log 2>&1; tail -4 gpurun_out/i_memcheck.log
timeout 300 python bench.py --workload asg --steps 20 --warmup 3 --no-cpu > gpurun_out/i_bench_asg.json 2> gpurun_out/i_bench_asg.err; tail -c 1300 gpurun_out/i_bench_asg.json; tail -3 gpurun_out/i_bench_asg.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/i_asg_launches.csv python scripts/prof_asg.py asg 3 > gpurun_out/i_ncu1.log 2>&1
tail -6 gpurun_out/i_asg_launches.csv | cut -d, -f5,9,15
