#!/bin/bash
# round-2 GPU session N: ASG v3 (single-warp chains with lagged two-float re-centring, batched FCC grad, prefetching FAC grad)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_criterion.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/n_pytest_crit.log 2>&1; echo "pytest exit $?" >> gpurun_out/n_pytest_crit.log
tail -6 gpurun_out/n_pytest_crit.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_criterion.py -m gpu -q --tb=line -p no:cacheprovider -k "test_asg_parity and not baseline" > gpurun_out/n_memcheck.log 2>&1; tail -3 gpurun_out/n_memcheck.log
timeout 200 python scripts/asg_roles.py > gpurun_out/n_roles.json 2>&1; cat gpurun_out/n_roles.json | tr -d '\n' | cut -c1-900; echo
timeout 200 python bench.py --workload asg --steps 20 --warmup 3 --no-cpu > gpurun_out/n_bench_asg.json 2> gpurun_out/n_bench_asg.err; grep -o '"asg_fwd_bwd_ms_per_batch": [0-9.]*\|"oracle_parity_rel": {[^}]*}\|"kernel_ms": [0-9.]*' gpurun_out/n_bench_asg.json; tail -3 gpurun_out/n_bench_asg.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/n_asg_launches.csv python scripts/prof_asg.py asg 3 > gpurun_out/n_ncu1.log 2>&1
tail -6 gpurun_out/n_asg_launches.csv | awk -F'","' '{print $5, $NF}'
