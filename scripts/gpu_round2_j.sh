#!/bin/bash
# round-2 GPU session J: ASG v2b (persistent FAC grad): parity, timing, full ncu of every ASG kernel with source pages
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_criterion.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/j_pytest_crit.log 2>&1; echo "pytest exit $?" >> gpurun_out/j_pytest_crit.log
tail -8 gpurun_out/j_pytest_crit.log
timeout 300 python bench.py --workload asg --steps 20 --warmup 3 --no-cpu > gpurun_out/j_bench_asg.json 2> gpurun_out/j_bench_asg.err; tail -c 1300 gpurun_out/j_bench_asg.json; tail -3 gpurun_out/j_bench_asg.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j_asg_launches.csv python scripts/prof_asg.py asg 3 > gpurun_out/j_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:asg_ -s 12 -c 6 -o /tmp/j_asg python scripts/prof_asg.py asg 3 > gpurun_out/j_ncu2.log 2>&1
ncu -i /tmp/j_asg.ncu-rep --page raw --csv > gpurun_out/j_asg_raw.csv 2>/dev/null
for k in chains fac_grad fcc_grad; do
  ncu -i /tmp/j_asg.ncu-rep --page source --csv -k regex:asg_$k > gpurun_out/j_asg_${k}_source.csv 2>/dev/null
done
python scripts/ncu_summary.py /tmp/j_asg.ncu-rep > gpurun_out/j_asg_summary.txt 2>&1
du -sh gpurun_out
