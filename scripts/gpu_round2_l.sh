#!/bin/bash
# round-2 GPU session L: source-level profile of the pipelined chains kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_criterion.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/l_pytest_crit.log 2>&1; echo "pytest exit $?" >> gpurun_out/l_pytest_crit.log
tail -4 gpurun_out/l_pytest_crit.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:asg_chains -s 2 -c 1 -o /tmp/l_asg python scripts/prof_asg.py asg 3 > gpurun_out/l_ncu2.log 2>&1
ncu -i /tmp/l_asg.ncu-rep --page source --csv > gpurun_out/l_asg_chains_source.csv 2>/dev/null
python scripts/ncu_summary.py /tmp/l_asg.ncu-rep > gpurun_out/l_asg_summary.txt 2>&1
du -sh gpurun_out
